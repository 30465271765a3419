// placeholder, replaced by the POA restatement
#include "rattle_oracle.h"
extern "C" {
int orc_poa_msa(const char*, const uint64_t*, uint32_t, int, int, int, int, char*, int64_t, int*, int64_t*, int32_t*, int64_t) { return -1; }
int orc_correct_reads(const char*, const char*, const uint64_t*, uint32_t, const int32_t*, const uint8_t*, const int32_t*, const int64_t*, const int32_t*, const uint8_t*, const int32_t*, int, double, double, double, int, int, char*, int64_t*, char*, int64_t*, char*, int64_t*) { return -1; }
int64_t orc_poa_cells(void) { return 0; }
}
