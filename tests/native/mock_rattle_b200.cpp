// TEST INFRASTRUCTURE — a stand-in for librattle_b200.so that answers the clustering entry points of include/rattle_b200.h
// from the CPU oracle (oracle/liboracle.so), so that the drop-in CLI's HOST logic (integration/rattle_dropin.cpp: flattening,
// the recognition of main.cpp's per-gene --iso loop, the cache of batched results, device slots) runs in the CPU test suite
// behind the reference's unmodified main.cpp / fasta.cpp.  Never shipped, never linked by the product.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rattle_b200.h"
#include "../../oracle/rattle_oracle.h"

struct rtl_ctx {
    int device;
    std::string err;
};
static std::string g_err;
static long g_calls[3] = {0, 0, 0};  // cluster_reads, cluster_reads_batched, segments

extern "C" {
int rtl_init(int device, rtl_ctx **out) {
    *out = new rtl_ctx{device, ""};
    return RTL_OK;
}
void rtl_destroy(rtl_ctx *ctx) { delete ctx; }
const char *rtl_last_error(const rtl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int rtl_cluster_reads(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads, int kmer_size, double t_s,
                      double t_v, double bv_threshold, double min_bv_threshold, double bv_falloff, double repr_percentile,
                      int is_rna, int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id, uint8_t *mem_rev,
                      int32_t *n_clusters) {
    ++g_calls[0];
    const int nc = orc_cluster_reads(bases, offsets, n_reads, kmer_size, t_s, t_v, bv_threshold, min_bv_threshold, bv_falloff,
                                     repr_percentile, is_rna, 4, main_id, main_rev, cl_off, mem_id, mem_rev);
    if (nc < 0) {
        ctx->err = "oracle refused the input";
        return RTL_ERR_INPUT;
    }
    *n_clusters = nc;
    if (getenv("MOCK_RTL_TRACE")) fprintf(stderr, "mock: rtl_cluster_reads #%ld: %u reads -> %d clusters\n", g_calls[0], n_reads, nc);
    return RTL_OK;
}

// the contract of rtl_cluster_reads_batched, literally: one cluster_reads per segment, results concatenated
int rtl_cluster_reads_batched(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads, const uint32_t *seg_off,
                              uint32_t n_seg, int kmer_size, double t_s, double t_v, double bv_threshold,
                              double min_bv_threshold, double bv_falloff, double repr_percentile, int is_rna,
                              int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id, uint8_t *mem_rev,
                              int32_t *n_clusters, int64_t *seg_cl_off) {
    ++g_calls[1];
    g_calls[2] += n_seg;
    int64_t c_at = 0, m_at = 0;
    cl_off[0] = 0;
    for (uint32_t s = 0; s < n_seg; ++s) {
        seg_cl_off[s] = c_at;
        const uint32_t n = seg_off[s + 1] - seg_off[s];
        if (n == 0) continue;
        std::vector<uint64_t> off(n + 1);
        for (uint32_t i = 0; i <= n; ++i) off[i] = offsets[seg_off[s] + i] - offsets[seg_off[s]];
        std::vector<int32_t> mi(n), me(n);
        std::vector<uint8_t> mr(n), mer(n);
        std::vector<int64_t> co(n + 1);
        const int nc = orc_cluster_reads(bases + offsets[seg_off[s]], off.data(), n, kmer_size, t_s, t_v, bv_threshold,
                                         min_bv_threshold, bv_falloff, repr_percentile, is_rna, 4, mi.data(), mr.data(),
                                         co.data(), me.data(), mer.data());
        if (nc < 0) {
            ctx->err = "oracle refused a segment";
            return RTL_ERR_INPUT;
        }
        for (int c = 0; c < nc; ++c) {
            main_id[c_at + c] = mi[c];
            main_rev[c_at + c] = mr[c];
            cl_off[c_at + c + 1] = m_at + co[c + 1];
        }
        for (int64_t i = 0; i < co[nc]; ++i) {
            mem_id[m_at + i] = me[i];
            mem_rev[m_at + i] = mer[i];
        }
        c_at += nc;
        m_at += co[nc];
    }
    seg_cl_off[n_seg] = c_at;
    *n_clusters = (int32_t)c_at;
    (void)n_reads;
    if (getenv("MOCK_RTL_TRACE")) fprintf(stderr, "mock: rtl_cluster_reads_batched #%ld: %u segments -> %ld clusters\n", g_calls[1], n_seg, (long)c_at);
    return RTL_OK;
}

int rtl_set_labels(rtl_ctx *, const char *const *, int) { return RTL_OK; }
int rtl_set_cluster_ids(rtl_ctx *, const int32_t *, int) { return RTL_OK; }
int rtl_correct_reads(rtl_ctx *ctx, const char *, const char *, const uint64_t *, uint32_t, const char *, const uint64_t *,
                      const int32_t *, const uint8_t *, const int32_t *, const int64_t *, const int32_t *, const uint8_t *,
                      const int32_t *, int, double, double, double, int, int, char *, int64_t *, char *, int64_t *, char *,
                      int64_t *) {
    ctx->err = "the mock library only clusters";
    return RTL_ERR_STATE;
}
}
