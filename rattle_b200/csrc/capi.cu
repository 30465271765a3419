// extern "C" surface of librattle_b200 (include/rattle_b200.h).  Exceptions stop here and become status codes.
#include <cstring>

#include "common.cuh"

// cluster_engine.cu
void cluster_state_free(rtl_ctx *ctx);
void cluster_upload(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n);
void cluster_extract(rtl_ctx *ctx, int k, int both);
void cluster_sort_by_length(rtl_ctx *ctx, const uint64_t *offsets, uint32_t n, uint32_t *perm);
void cluster_run(rtl_ctx *ctx, int k, double t_s, double t_v, double bv_thr, double bv_min, double bv_falloff,
                 double repr_pct, int is_rna, int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id,
                 uint8_t *mem_rev, int32_t *n_clusters, const uint32_t *seg_off = nullptr, uint32_t n_seg = 0,
                 int64_t *seg_cl_off = nullptr);
void cluster_download_kmers(rtl_ctx *ctx, uint32_t *fh, int32_t *fp, uint32_t *rh, int32_t *rp, uint64_t *bf,
                            uint64_t *br);
void cluster_bv_scan_dense(rtl_ctx *ctx, int k, int is_rna, const int32_t *seed_reads, int n_seeds,
                           const int32_t *target_reads, int n_targets, double thr, uint32_t *common, uint8_t *pass);
void cluster_pair_similarity(rtl_ctx *ctx, int k, int is_rna, const int32_t *a_read, const int32_t *b_read,
                             const uint8_t *strand, int64_t n_tasks, double t_s, double t_v, int64_t *n_common,
                             int32_t *bases, int32_t *n_dist, double *var, uint8_t *accept);
// poa_engine.cu
void poa_state_free(rtl_ctx *ctx);
int poa_msa(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n, int m, int nn, int g, int e,
            char *msa_out, int64_t cap, int *msa_cols, int64_t *aln_off, int32_t *aln_pairs, int64_t aln_cap);
int correct_reads_impl(rtl_ctx *ctx, const char *bases, const char *quals, const uint64_t *offsets, uint32_t n_reads,
                       const char *headers, const uint64_t *header_off, const int32_t *main_id, const uint8_t *main_rev,
                       const int32_t *main_gene, const int64_t *cl_off, const int32_t *mem_id, const uint8_t *mem_rev,
                       const int32_t *mem_gene, int n_clusters, double min_occ, double gap_occ, double err_ratio,
                       int split, int min_reads, char *corrected, int64_t *corrected_len, char *uncorrected,
                       int64_t *uncorrected_len, char *consensi, int64_t *consensi_len);

static std::string g_init_error;

template <typename F>
static int guarded(rtl_ctx *ctx, F &&f) {
    if (!ctx) return RTL_ERR_STATE;
    try {
        cudaSetDevice(ctx->device);
        return f();
    } catch (const CudaError &e) {
        ctx->err = e.what();
        return RTL_ERR_CUDA;
    } catch (const InputError &e) {
        ctx->err = e.what();
        return RTL_ERR_INPUT;
    } catch (const CapacityError &e) {
        ctx->err = e.what();
        return RTL_ERR_CAPACITY;
    } catch (const StateError &e) {
        ctx->err = e.what();
        return RTL_ERR_STATE;
    } catch (const std::exception &e) {
        ctx->err = e.what();
        return RTL_ERR_CUDA;
    }
}

extern "C" {

int rtl_init(int device, rtl_ctx **out) {
    if (!out) return RTL_ERR_INPUT;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_init_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count=0") +
                       " (rattle_b200 has no CPU fallback)";
        return RTL_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        g_init_error = "device index out of range";
        return RTL_ERR_INPUT;
    }
    rtl_ctx *ctx = new rtl_ctx();
    ctx->device = device;
    try {
        CK(cudaSetDevice(device));
        cudaDeviceProp p;
        CK(cudaGetDeviceProperties(&p, device));
        if (p.major < 10) throw CudaError(std::string("device is sm_") + std::to_string(p.major * 10 + p.minor) + ", built for sm_100a");
        ctx->n_sm = p.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        ctx->stream = ctx->own_stream;
    } catch (const std::exception &ex) {
        g_init_error = ex.what();
        delete ctx;
        return RTL_ERR_CUDA;
    }
    *out = ctx;
    return RTL_OK;
}

void rtl_destroy(rtl_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cluster_state_free(ctx);
    poa_state_free(ctx);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *rtl_last_error(const rtl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int rtl_set_option(rtl_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) return RTL_ERR_STATE;
    std::string k(key);
    if (k == "wave") {
        if (value < 1 || value > 8192) return RTL_ERR_INPUT;
        ctx->wave = (int)value;
    } else if (k == "task_cap") {
        if (value < 4096 || value >= (1ll << 31)) return RTL_ERR_INPUT;
        ctx->task_cap = value;
    } else if (k == "scratch_mb") {
        if (value < 1) return RTL_ERR_INPUT;
        ctx->scratch_mb = value;
    } else if (k == "bv_kernel") {
        ctx->bv_kernel = (int)value;
    } else if (k == "poa_batch") {
        ctx->poa_batch = (int)value;
    } else if (k == "poa_units") {
        if (value < 0 || value > 16) return RTL_ERR_INPUT;
        ctx->poa_units = (int)value;
    } else if (k == "poa_mirror_pct") {
        if (value < 1 || value > 100) return RTL_ERR_INPUT;
        ctx->poa_mirror_pct = (int)value;
    } else if (k == "poa_gpu_sort") {
        ctx->poa_gpu_sort = (int)value;
    } else if (k == "poa_device_chain") {
        ctx->poa_device_chain = (int)value;
    } else if (k == "poa_device_vote") {
        ctx->poa_device_vote = (int)value;
    } else if (k == "poa_kernel") {
        ctx->poa_kernel = (int)value;
    } else if (k == "poa_arena_mb") {
        if (value < 0) return RTL_ERR_INPUT;
        ctx->poa_arena_mb = value;
    } else {
        ctx->err = "unknown option " + k;
        return RTL_ERR_INPUT;
    }
    return RTL_OK;
}

int rtl_set_labels(rtl_ctx *ctx, const char *const *labels, int n_labels) {
    if (!ctx || n_labels < 0 || (n_labels > 0 && !labels)) return RTL_ERR_STATE;
    ctx->labels.clear();
    for (int i = 0; i < n_labels; ++i) ctx->labels.emplace_back(labels[i] ? labels[i] : "");
    return RTL_OK;
}

int rtl_set_cluster_ids(rtl_ctx *ctx, const int32_t *ids, int n) {
    if (!ctx || n < 0 || (n > 0 && !ids)) return RTL_ERR_STATE;
    ctx->cluster_ids.assign(ids, ids + n);
    return RTL_OK;
}

int rtl_set_stream(rtl_ctx *ctx, void *cuda_stream) {
    if (!ctx) return RTL_ERR_STATE;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return RTL_OK;
}

int rtl_get_stats(const rtl_ctx *ctx, rtl_stats *out) {
    if (!ctx || !out) return RTL_ERR_STATE;
    *out = ctx->stats;
    return RTL_OK;
}

int rtl_set_shard(rtl_ctx *ctx, int rank, int world, rtl_allreduce_min_fn fn, void *user) {
    if (!ctx || world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) return RTL_ERR_INPUT;
    ctx->rank = rank;
    ctx->world = world;
    ctx->allreduce = fn;
    ctx->allreduce_user = user;
    return RTL_OK;
}

int rtl_sort_reads_by_length(rtl_ctx *ctx, const uint64_t *offsets, uint32_t n_reads, uint32_t *perm) {
    return guarded(ctx, [&]() {
        ctx->stats = rtl_stats{};
        cluster_sort_by_length(ctx, offsets, n_reads, perm);
        return RTL_OK;
    });
}

int rtl_set_broadcast(rtl_ctx *ctx, rtl_broadcast_fn fn, void *user) {
    if (!ctx) return RTL_ERR_STATE;
    ctx->broadcast = fn;
    ctx->broadcast_user = user;
    return RTL_OK;
}

int rtl_reads_upload(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads) {
    return guarded(ctx, [&]() {
        cluster_upload(ctx, bases, offsets, n_reads);
        return RTL_OK;
    });
}

int rtl_cluster_resident(rtl_ctx *ctx, int kmer_size, double t_s, double t_v, double bv_threshold,
                         double min_bv_threshold, double bv_falloff, double repr_percentile, int is_rna,
                         int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id, uint8_t *mem_rev,
                         int32_t *n_clusters) {
    return guarded(ctx, [&]() {
        ctx->stats = rtl_stats{};
        cluster_run(ctx, kmer_size, t_s, t_v, bv_threshold, min_bv_threshold, bv_falloff, repr_percentile, is_rna,
                    main_id, main_rev, cl_off, mem_id, mem_rev, n_clusters);
        return RTL_OK;
    });
}

int rtl_cluster_reads(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads, int kmer_size,
                      double t_s, double t_v, double bv_threshold, double min_bv_threshold, double bv_falloff,
                      double repr_percentile, int is_rna, int32_t *main_id, uint8_t *main_rev, int64_t *cl_off,
                      int32_t *mem_id, uint8_t *mem_rev, int32_t *n_clusters) {
    return guarded(ctx, [&]() {
        ctx->stats = rtl_stats{};
        if (n_reads == 0) {  // the reference returns an empty cluster set (cluster.cpp:93-259 with no reads)
            if (!n_clusters || !cl_off) throw InputError("null output buffers");
            *n_clusters = 0;
            cl_off[0] = 0;
            return RTL_OK;
        }
        const double t0 = now_ms();
        cluster_upload(ctx, bases, offsets, n_reads);
        cluster_run(ctx, kmer_size, t_s, t_v, bv_threshold, min_bv_threshold, bv_falloff, repr_percentile, is_rna,
                    main_id, main_rev, cl_off, mem_id, mem_rev, n_clusters);
        ctx->stats.total_ms = now_ms() - t0;
        return RTL_OK;
    });
}

int rtl_cluster_reads_batched(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads,
                              const uint32_t *seg_off, uint32_t n_seg, int kmer_size, double t_s, double t_v,
                              double bv_threshold, double min_bv_threshold, double bv_falloff, double repr_percentile,
                              int is_rna, int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id,
                              uint8_t *mem_rev, int32_t *n_clusters, int64_t *seg_cl_off) {
    return guarded(ctx, [&]() {
        ctx->stats = rtl_stats{};
        if (!seg_off || !seg_cl_off || !n_clusters || !cl_off) throw InputError("null segment / output buffers");
        if (n_reads == 0) {  // every segment is an empty read set: empty cluster sets (cluster.cpp:93-259 with no reads)
            *n_clusters = 0;
            cl_off[0] = 0;
            for (uint32_t s = 0; s <= n_seg; ++s) seg_cl_off[s] = 0;
            return RTL_OK;
        }
        const double t0 = now_ms();
        cluster_upload(ctx, bases, offsets, n_reads);
        cluster_run(ctx, kmer_size, t_s, t_v, bv_threshold, min_bv_threshold, bv_falloff, repr_percentile, is_rna,
                    main_id, main_rev, cl_off, mem_id, mem_rev, n_clusters, seg_off, n_seg, seg_cl_off);
        ctx->stats.total_ms = now_ms() - t0;
        return RTL_OK;
    });
}

int rtl_extract_kmers(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n_reads, int kmer_size,
                      int both_strands, uint32_t *fwd_hash, int32_t *fwd_pos, uint32_t *rev_hash, int32_t *rev_pos,
                      uint64_t *bv_fwd, uint64_t *bv_rev) {
    return guarded(ctx, [&]() {
        cluster_upload(ctx, bases, offsets, n_reads);
        cluster_extract(ctx, kmer_size, both_strands != 0);
        cluster_download_kmers(ctx, fwd_hash, fwd_pos, rev_hash, rev_pos, bv_fwd, bv_rev);
        return RTL_OK;
    });
}

int rtl_bv_scan(rtl_ctx *ctx, int kmer_size, int is_rna, const int32_t *seed_reads, int n_seeds,
                const int32_t *target_reads, int n_targets, double bv_threshold, uint32_t *common, uint8_t *pass) {
    return guarded(ctx, [&]() {
        cluster_bv_scan_dense(ctx, kmer_size, is_rna, seed_reads, n_seeds, target_reads, n_targets, bv_threshold, common,
                              pass);
        return RTL_OK;
    });
}

int rtl_pair_similarity(rtl_ctx *ctx, int kmer_size, int is_rna, const int32_t *a_read, const int32_t *b_read,
                        const uint8_t *strand, int64_t n_tasks, double t_s, double t_v, int64_t *n_common,
                        int32_t *bases, int32_t *n_dist, double *var, uint8_t *accept) {
    return guarded(ctx, [&]() {
        cluster_pair_similarity(ctx, kmer_size, is_rna, a_read, b_read, strand, n_tasks, t_s, t_v, n_common, bases,
                                n_dist, var, accept);
        return RTL_OK;
    });
}

int rtl_poa_msa(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n, int m, int nn, int g, int e,
                char *msa_out, int64_t cap, int *msa_cols, int64_t *aln_off, int32_t *aln_pairs, int64_t aln_cap) {
    int rows = 0;
    int rc = guarded(ctx, [&]() {
        rows = poa_msa(ctx, bases, offsets, n, m, nn, g, e, msa_out, cap, msa_cols, aln_off, aln_pairs, aln_cap);
        return RTL_OK;
    });
    return rc < 0 ? rc : rows;
}

int rtl_correct_reads(rtl_ctx *ctx, const char *bases, const char *quals, const uint64_t *offsets, uint32_t n_reads,
                      const char *headers, const uint64_t *header_off, const int32_t *main_id,
                      const uint8_t *main_rev, const int32_t *main_gene, const int64_t *cl_off,
                      const int32_t *mem_id, const uint8_t *mem_rev, const int32_t *mem_gene, int n_clusters,
                      double min_occ, double gap_occ, double err_ratio, int split, int min_reads, char *corrected,
                      int64_t *corrected_len, char *uncorrected, int64_t *uncorrected_len, char *consensi,
                      int64_t *consensi_len) {
    return guarded(ctx, [&]() {
        return correct_reads_impl(ctx, bases, quals, offsets, n_reads, headers, header_off, main_id, main_rev, main_gene,
                                  cl_off, mem_id, mem_rev, mem_gene, n_clusters, min_occ, gap_occ, err_ratio, split,
                                  min_reads, corrected, corrected_len, uncorrected, uncorrected_len, consensi,
                                  consensi_len);
    });
}

}  // extern "C"
