// placeholder until the POA engine lands (hot path B)
#include "common.cuh"
void poa_state_free(rtl_ctx *) {}
int poa_msa(rtl_ctx *, const char *, const uint64_t *, uint32_t, int, int, int, int, char *, int64_t, int *, int64_t *,
            int32_t *, int64_t) {
    throw StateError("POA engine not built");
}
int correct_reads_impl(rtl_ctx *, const char *, const char *, const uint64_t *, uint32_t, const char *, const uint64_t *,
                       const int32_t *, const uint8_t *, const int32_t *, const int64_t *, const int32_t *,
                       const uint8_t *, const int32_t *, int, double, double, double, int, int, char *, int64_t *, char *,
                       int64_t *, char *, int64_t *) {
    throw StateError("POA engine not built");
}
