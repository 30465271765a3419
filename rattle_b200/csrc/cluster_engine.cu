// Host orchestration of hot path A on one GPU: resident read set, k-mer extraction, and the greedy passes of
// cluster.cpp:93-259 restated as device-side "waves" (DESIGN.md §2):
//
//   cluster_together(i,j,thr) is a pure function, so "read j joins the smallest earlier seed that matches it"
//   can be evaluated in batches: each wave takes the first W untaken items as candidate seeds, decides which of
//   them are seeds from their W x W match matrix (phase A), then scores the seeds against every later untaken
//   item (phase B).  All bookkeeping (selection, resolution, assignment) runs in small kernels; the host only
//   reads a 4-int status per wave.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "cluster_kernels.cuh"
#include "common.cuh"
#include "poa_engine.hpp"  // parallel_for / host_threads (shared worker pool)

using namespace rtl;

struct EvPool {
    std::vector<cudaEvent_t> free_, used_[4];
    double acc[4] = {0, 0, 0, 0};
    cudaEvent_t get() {
        cudaEvent_t e;
        if (!free_.empty()) {
            e = free_.back();
            free_.pop_back();
        } else
            CK(cudaEventCreate(&e));
        return e;
    }
    void begin(int cat, cudaStream_t s) {
        cudaEvent_t e = get();
        CK(cudaEventRecord(e, s));
        used_[cat].push_back(e);
    }
    void end(int cat, cudaStream_t s) { begin(cat, s); }
    void collect() {  // after a stream sync
        for (int c = 0; c < 4; ++c) {
            for (size_t i = 0; i + 1 < used_[c].size(); i += 2) {
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, used_[c][i], used_[c][i + 1]));
                acc[c] += ms;
            }
            for (auto e : used_[c]) free_.push_back(e);
            used_[c].clear();
        }
    }
    ~EvPool() {
        for (auto e : free_) cudaEventDestroy(e);
        for (int c = 0; c < 4; ++c)
            for (auto e : used_[c]) cudaEventDestroy(e);
    }
};
enum { EV_BV = 0, EV_JOIN = 1, EV_HEAVY = 2, EV_EXTRACT = 3 };

struct ClusterState {
    uint32_t n = 0;
    uint64_t total = 0;
    std::vector<uint64_t> h_off;
    std::vector<int32_t> h_len;
    DevBuf<uint8_t> d_bases;    // ASCII staging of the upload
    DevBuf<uint32_t> d_pk;      // the resident read set: 2-bit packed (k_pack_bases)
    DevBuf<int> pack_flag;      // != 0: the upload met a base outside A,C,G,T,U
    DevBuf<uint64_t> d_off;
    DevBuf<int32_t> d_len;
    int ex_k = -1, ex_both = -1;
    DevBuf<uint32_t> kh[2];
    DevBuf<int32_t> kp[2];
    DevBuf<uint64_t> bvbuf;   // bitvectors, [read][fwd 64 words | rev 64 words] when both strands are extracted
    uint64_t *bv[2] = {nullptr, nullptr};
    int bv_stride = 64;
    DevBuf<int32_t> pc;
    DevBuf<uint32_t> read_list;
    DevBuf<uint64_t> long_off, long_scratch;
    // wave state
    DevBuf<uint8_t> taken, owner_rev, is_seed;
    DevBuf<int32_t> owner, item_read, item_rid, item_seg, seg_first, seg_cur, cand, seed_item, wave, shard_list;
    DevBuf<uint32_t> memo;      // known k-mer-test failures between representatives (cluster_kernels.cuh: Memo)
    uint32_t rid_dim = 0;       // 0 = memo off
    DevBuf<uint32_t> best, acc;
    DevBuf<uint16_t> cut;
    DevBuf<uint64_t> tasks, surv, defer[2];
    DevBuf<unsigned long long> counters;  // [0]=n_tasks [1]=n_surv [2]=scratch_cur [3]=pairs
    DevBuf<int> flags;                    // [0]=input err [1]=overflow
    DevBuf<unsigned char> scratch;
    PinBuf<int32_t> h_wave;
    PinBuf<int> h_flags;
    PinBuf<unsigned long long> h_counters;
    EvPool ev;
    EventTimer up_timer;
    bool smem_attr_set = false;
};

static ClusterState &state(rtl_ctx *ctx) {
    if (!ctx->cl) ctx->cl = new ClusterState();
    return *ctx->cl;
}
void cluster_state_free(rtl_ctx *ctx) {
    delete ctx->cl;
    ctx->cl = nullptr;
}

static void set_smem_attrs(ClusterState &S) {
    if (S.smem_attr_set) return;
    const int max_smem = 227 * 1024;
    CK(cudaFuncSetAttribute(k_extract_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    // (k_bv_scan also holds a statically allocated mbarrier: its dynamic limit is what a full seed tile needs)
    CK(cudaFuncSetAttribute(k_bv_scan<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BVS_TS * (64 * 8 + 8)));
    CK(cudaFuncSetAttribute(k_bv_scan<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, BVS_TS * (64 * 8 + 8)));
    CK(cudaFuncSetAttribute(k_bv_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CK(cudaFuncSetAttribute(k_join_count, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CK(cudaFuncSetAttribute(k_pair_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    CK(cudaFuncSetAttribute(k_resolve_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    S.smem_attr_set = true;
}

// ------------------------------------------------------------------------------------------------ upload
void cluster_upload(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n) {
    ClusterState &S = state(ctx);
    if (n == 0) throw InputError("empty read set");
    if (!bases || !offsets) throw InputError("null read buffers");
    if (n >= (1u << 31)) throw InputError("more than 2^31 reads");
    for (uint32_t i = 0; i < n; ++i) {
        if (offsets[i + 1] < offsets[i]) throw InputError("read offsets are not monotonic");
        if (offsets[i + 1] - offsets[i] > 0x7fffffffull) throw InputError("read longer than 2^31 bases");
    }
    S.n = n;
    S.total = offsets[n] - offsets[0];
    S.h_off.resize(n + 1);
    const uint64_t o0 = offsets[0];
    for (uint32_t i = 0; i <= n; ++i) S.h_off[i] = offsets[i] - o0;
    S.h_len.resize(n);
    for (uint32_t i = 0; i < n; ++i) S.h_len[i] = (int32_t)(S.h_off[i + 1] - S.h_off[i]);
    S.d_bases.need(S.total + 16);
    S.d_off.need(n + 1);
    S.up_timer.init();
    CK(cudaEventRecord(S.up_timer.a, ctx->stream));
    CK(cudaMemcpyAsync(S.d_bases.p, bases + o0, S.total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(S.d_off.p, S.h_off.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(S.up_timer.b, ctx->stream));
    // K1: pack to 2 bits per base; everything downstream reads the packed copy
    S.d_pk.need((S.total >> 4) + n + 2);
    CK(cudaMemsetAsync(S.pack_flag.need(1), 0, sizeof(int), ctx->stream));
    k_pack_bases<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(S.d_bases.p, S.d_off.p, n, S.d_pk.p, S.pack_flag.p);
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    ctx->stats.h2d_bytes += (int64_t)S.total + (int64_t)(n + 1) * 8;
    S.ex_k = -1;
    CK(cudaStreamSynchronize(ctx->stream));
    float up_ms = 0;
    CK(cudaEventElapsedTime(&up_ms, S.up_timer.a, S.up_timer.b));
    ctx->stats.upload_ms += up_ms;
}

// ------------------------------------------------------------------------------------------------ visitation order
// sort_read_set (fasta.cpp:458-464; main.cpp:254 calls it before cluster_reads): perm[i] = index of the read that comes
// i-th, longest first, ties in input order
void cluster_sort_by_length(rtl_ctx *ctx, const uint64_t *offsets, uint32_t n, uint32_t *perm) {
    if (n == 0) return;
    if (!offsets || !perm) throw InputError("null buffers");
    for (uint32_t i = 0; i < n; ++i)
        if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] > 0x7fffffffull) throw InputError("bad read offsets");
    cudaStream_t st = ctx->stream;
    uint32_t n_pad = 2048;
    while (n_pad < n) n_pad <<= 1;
    DevBuf<uint64_t> d_off, d_keys;
    DevBuf<uint32_t> d_perm;
    CK(cudaMemcpyAsync(d_off.need(n + 1), offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    d_keys.need(n_pad);
    k_sort_keys_init<<<(n_pad + 255) / 256, 256, 0, st>>>(d_off.p, n, n_pad, d_keys.p);
    int launches = 2;
    for (uint32_t size = 2; size <= n_pad; size <<= 1) {
        uint32_t stride = size >> 1;
        for (; stride >= 1024; stride >>= 1, ++launches)
            k_bitonic_step<<<(n_pad / 2 + 255) / 256, 256, 0, st>>>(d_keys.p, n_pad, size, stride);
        k_bitonic_local<<<n_pad / 2048, 1024, 0, st>>>(d_keys.p, size, stride);
        ++launches;
    }
    k_sort_keys_perm<<<(n + 255) / 256, 256, 0, st>>>(d_keys.p, n, d_perm.need(n));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(perm, d_perm.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->stats.kernel_launches += launches;
    ctx->stats.h2d_bytes += (int64_t)(n + 1) * 8;
    ctx->stats.d2h_bytes += (int64_t)n * 4;
}

// ------------------------------------------------------------------------------------------------ extraction
static const int SORT_CLASSES[] = {1024, 2048, 4096, 8192, 16384};

void cluster_extract(rtl_ctx *ctx, int k, int both) {
    ClusterState &S = state(ctx);
    if (S.n == 0) throw StateError("no reads uploaded");
    if (k < 1 || k > 16) throw InputError("kmer_size must be in [1,16]");
    if (S.ex_k == k && S.ex_both == both) return;
    set_smem_attrs(S);
    const uint32_t n = S.n;
    const uint64_t total_k = S.total - (uint64_t)k * n;
    for (uint32_t i = 0; i < n; ++i)
        if (S.h_len[i] <= k || S.h_len[i] <= 6) throw InputError("read shorter than or equal to kmer size (kmer.cpp:9)");
    cudaStream_t st = ctx->stream;
    S.flags.need(4);
    k_clear_keep_pack_flag<<<1, 1, 0, st>>>(S.flags.p, S.pack_flag.p);  // what k_pack_bases saw at upload time -> flags[0]
    S.d_len.need(n);
    CK(cudaMemcpyAsync(S.d_len.p, S.h_len.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    ctx->stats.h2d_bytes += (int64_t)n * 4;
    for (int s = 0; s < 2; ++s) {
        const bool on = s == 0 || both;
        S.kh[s].need(on ? total_k + 4 : 4);
        S.kp[s].need(on ? total_k + 4 : 4);
    }
    // a read's forward and reverse bitvector are adjacent: the scan fetches both with one bulk copy
    S.bv_stride = both ? 128 : 64;
    S.bvbuf.need((size_t)n * S.bv_stride + 128);
    S.bv[0] = S.bvbuf.p;
    S.bv[1] = both ? S.bvbuf.p + 64 : S.bvbuf.p + (size_t)n * 64;  // (single strand: 64 words nobody reads)
    S.pc.need(n);
    // multi-GPU with an exchange callback: this rank extracts its block of the reads, the ranks exchange their slices
    // afterwards; otherwise every rank extracts everything
    const bool sharded = ctx->world > 1 && ctx->broadcast != nullptr && ctx->allreduce != nullptr;
    auto slice_lo = [&](int r) { return (uint32_t)((uint64_t)n * (uint64_t)r / (uint64_t)ctx->world); };
    const uint32_t my_lo = sharded ? slice_lo(ctx->rank) : 0, my_hi = sharded ? slice_lo(ctx->rank + 1) : n;
    // bucket reads by padded list size
    const int n_cls = sizeof(SORT_CLASSES) / sizeof(int);
    std::vector<std::vector<uint32_t>> bucket(n_cls + 1);
    for (uint32_t i = my_lo; i < my_hi; ++i) {
        const int nk = S.h_len[i] - k;
        int c = 0;
        while (c < n_cls && nk > SORT_CLASSES[c]) ++c;
        bucket[c].push_back(i);
    }
    std::vector<uint32_t> flat;
    flat.reserve(n);
    std::vector<size_t> start(n_cls + 2, 0);
    for (int c = 0; c <= n_cls; ++c) {
        start[c] = flat.size();
        flat.insert(flat.end(), bucket[c].begin(), bucket[c].end());
    }
    start[n_cls + 1] = flat.size();
    S.read_list.need(n);
    CK(cudaMemcpyAsync(S.read_list.p, flat.data(), flat.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    ctx->stats.h2d_bytes += (int64_t)flat.size() * 4;
    S.ev.begin(EV_EXTRACT, st);
    for (int c = 0; c < n_cls; ++c) {
        const size_t cnt = start[c + 1] - start[c];
        if (!cnt) continue;
        const int n_pad = SORT_CLASSES[c];
        const size_t smem = (size_t)n_pad * 8 + 512 + n_pad + 32;
        const int threads = n_pad <= 2048 ? 256 : (n_pad <= 4096 ? 512 : 1024);
        // gridDim.x limit is 2^31-1; y = strand
        dim3 grid((unsigned)cnt, both ? 2 : 1);
        k_extract_smem<<<grid, threads, smem, st>>>(S.d_pk.p, S.d_off.p, S.read_list.p + start[c], k, n_pad,
                                                    S.kh[0].p, S.kp[0].p, S.kh[1].p, S.kp[1].p, S.bv[0], S.bv[1],
                                                    S.bv_stride, S.pc.p, S.flags.p);
        CK(cudaGetLastError());
        ctx->stats.kernel_launches++;
    }
    {
        const size_t cnt = start[n_cls + 1] - start[n_cls];
        if (cnt) {
            std::vector<uint64_t> so(cnt * 2);
            uint64_t at = 0;
            for (size_t i = 0; i < cnt; ++i) {
                int nk = S.h_len[flat[start[n_cls] + i]] - k;
                uint64_t np = 1;
                while ((int64_t)np < nk) np <<= 1;
                so[2 * i] = at;
                at += np;
                so[2 * i + 1] = at;
                at += np;
            }
            S.long_off.need(cnt * 2);
            S.long_scratch.need(at);
            CK(cudaMemcpyAsync(S.long_off.p, so.data(), cnt * 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
            dim3 grid((unsigned)cnt, both ? 2 : 1);
            k_extract_long<<<grid, 1024, 0, st>>>(S.d_pk.p, S.d_off.p, S.read_list.p + start[n_cls], S.long_off.p,
                                                  S.long_scratch.p, k, S.kh[0].p, S.kp[0].p, S.kh[1].p, S.kp[1].p,
                                                  S.bv[0], S.bv[1], S.bv_stride, S.pc.p, S.flags.p);
            CK(cudaGetLastError());
            ctx->stats.kernel_launches++;
            CK(cudaStreamSynchronize(st));  // `so` must outlive the copy
        }
    }
    if (sharded) {
        // every rank's slice of the k-mer lists (read r's list starts at off[r] - k*r), bitvectors and popcounts goes to
        // all the others; a bad base seen by one rank fails the call on every rank
        for (int r = 0; r < ctx->world; ++r) {
            const uint32_t lo = slice_lo(r), hi = slice_lo(r + 1);
            if (hi == lo) continue;
            const uint64_t k0 = S.h_off[lo] - (uint64_t)k * lo, k1 = S.h_off[hi] - (uint64_t)k * hi;
            int rc = 0;
            for (int s = 0; s < (both ? 2 : 1) && !rc; ++s) {
                rc = ctx->broadcast(ctx->broadcast_user, S.kh[s].p + k0, (int64_t)((k1 - k0) * 4), r);
                if (!rc) rc = ctx->broadcast(ctx->broadcast_user, S.kp[s].p + k0, (int64_t)((k1 - k0) * 4), r);
            }
            if (!rc) rc = ctx->broadcast(ctx->broadcast_user, S.bvbuf.p + (size_t)lo * S.bv_stride,
                                         (int64_t)((size_t)(hi - lo) * S.bv_stride * 8), r);
            if (!rc) rc = ctx->broadcast(ctx->broadcast_user, S.pc.p + lo, (int64_t)(hi - lo) * 4, r);
            if (rc) throw CudaError("broadcast callback failed");
        }
        uint32_t *status = reinterpret_cast<uint32_t *>(S.flags.p + 3);
        k_fold_input_flag<<<1, 1, 0, st>>>(S.flags.p, status);
        if (ctx->allreduce(ctx->allreduce_user, status, 1) != 0) throw CudaError("allreduce callback failed");
        k_unfold_input_flag<<<1, 1, 0, st>>>(S.flags.p, status);
        ctx->stats.kernel_launches += 2;
    }
    S.ev.end(EV_EXTRACT, st);
    int *hf = S.h_flags.need(4);
    CK(cudaMemcpyAsync(hf, S.flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    S.ev.collect();
    if (hf[0]) throw InputError("base outside ACGTU in input (kmer.hpp:36)");
    S.ex_k = k;
    S.ex_both = both;
}

static ReadView view(ClusterState &S) {
    ReadView R;
    R.bases = S.d_bases.p;
    R.off = S.d_off.p;
    R.len = S.d_len.p;
    R.kh[0] = S.kh[0].p;
    R.kh[1] = S.kh[1].p;
    R.kp[0] = S.kp[0].p;
    R.kp[1] = S.kp[1].p;
    R.bv[0] = S.bv[0];
    R.bv[1] = S.bv[1];
    R.bv_stride = S.bv_stride;
    R.pc = S.pc.p;
    R.k = S.ex_k;
    R.n = S.n;
    return R;
}

// cluster.cpp:16-19: smallest common count c with double(c)/double(mmax) >= thr, per mmax (exact: same double ops).
static void make_cut_table(double thr, std::vector<uint16_t> &cut) {
    cut.assign(4097, 0);
    if (thr == 0) return;  // `thr == 0 ||` forward, c/mmax >= 0 reverse: everything passes
    int c = 0;
    for (int m = 1; m <= 4096; ++m) {
        while (c <= m && !((double)c / (double)m >= thr)) ++c;
        cut[m] = (uint16_t)std::min(c, 4097);
    }
    cut[0] = 4097;
}

// launches the bitvector scan over `n_targets` targets and seed tiles covering up to `max_seeds` seeds.  Option
// bv_kernel=2 sends scans with few seeds (the HBM-bound regime) to the bulk-copy ring kernel k_bv_stream; measured on B200
// it streams 0.54 of the HBM peak against 0.59 of the register-staged kernel, which therefore stays the default
static void launch_bv_scan(rtl_ctx *ctx, BvScanArgs a, int64_t n_targets, int max_seeds, cudaStream_t st) {
    if (ctx->bv_kernel != 2 || max_seeds > BVT_TS) {
        a.ts_cap = std::max(1, std::min(BVS_TS, max_seeds));
        int gx = (int)std::min<int64_t>((n_targets + 15) / 16, (int64_t)ctx->n_sm * 8);
        dim3 grid(std::max(gx, 1), (max_seeds + a.ts_cap - 1) / a.ts_cap);
        if (ctx->bv_kernel == 3) k_bv_scan<3><<<grid, BVS_THREADS, (size_t)a.ts_cap * (64 * 8 + 8), st>>>(a);
        else k_bv_scan<4><<<grid, BVS_THREADS, (size_t)a.ts_cap * (64 * 8 + 8), st>>>(a);
    } else {
        a.ts_cap = std::max(1, max_seeds);
        const size_t smem = bvt_smem_bytes(a.ts_cap);
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bv_stream, BVT_THREADS, smem));
        const int64_t groups = (n_targets + BVT_SLOT_TGT - 1) / BVT_SLOT_TGT;
        const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((groups + BVT_CONS - 1) / BVT_CONS, (int64_t)ctx->n_sm * std::max(1, occ)));
        k_bv_stream<<<dim3(gx, 1), BVT_THREADS, smem, st>>>(a);
    }
    CK(cudaGetLastError());
}

static const int JC_CAP_W = 1760;   // staged hashes (list B) per warp in k_join_count: 16 warps x 6.9 KB, 2 CTAs per SM
static const int PH_CAP_C = 1024;   // matches per warp kept in shared memory in k_pair_heavy (8 warps x 20 KB)

// join + heavy over the current task buffer
static void run_pair_kernels(rtl_ctx *ctx, ClusterState &S, const TaskView &tv, double t_s, double t_v, const Sink &sink,
                             int64_t *nmatch_out) {
    cudaStream_t st = ctx->stream;
    ReadView R = view(S);
    const int64_t surv_cap = (int64_t)S.surv.cap;
    S.ev.begin(EV_JOIN, st);
    k_join_count<<<ctx->n_sm * 2, JC_THREADS, (size_t)(JC_THREADS / 32) * JC_CAP_W * 4, st>>>(
        tv, S.tasks.p, S.counters.p + 0, R, t_s, JC_CAP_W, S.surv.p, S.counters.p + 1, surv_cap, nmatch_out,
        S.flags.p + 1, S.counters.p + 4);
    CK(cudaGetLastError());
    S.ev.end(EV_JOIN, st);
    S.ev.begin(EV_HEAVY, st);
    const size_t smem = (size_t)(PH_THREADS / 32) * heavy_bytes(PH_CAP_C, PH_CAP_C);
    // Survivors with more than PH_CAP_C matches sort in a global scratch arena (bump allocation).  Those that find it
    // full are deferred: two more launches follow, each with the arena empty again (they exit at once when nothing was
    // deferred, the usual case); only what is still left after them fails the call ("raise scratch_mb").
    const int64_t defer_cap = (int64_t)S.defer[0].cap;
    const unsigned long long scratch_bytes = std::min<size_t>(S.scratch.cap, (size_t)ctx->scratch_mb << 20);
    CK(cudaMemsetAsync(S.counters.p + 6, 0, 2 * sizeof(unsigned long long), st));
    k_pair_heavy<<<ctx->n_sm * 2, PH_THREADS, smem, st>>>(tv, S.tasks.p, S.surv.p, S.counters.p + 1, surv_cap, R, t_s, t_v,
                                                          PH_CAP_C, S.scratch.p, S.counters.p + 2,
                                                          scratch_bytes, sink, S.flags.p + 1,
                                                          S.counters.p + 5, S.defer[0].p, S.counters.p + 6, defer_cap);
    CK(cudaGetLastError());
    for (int r = 0; r < 2; ++r) {
        CK(cudaMemsetAsync(S.counters.p + 2, 0, sizeof(unsigned long long), st));
        k_pair_heavy<<<ctx->n_sm * 2, PH_THREADS, smem, st>>>(
            tv, S.tasks.p, S.defer[r].p, S.counters.p + 6 + r, defer_cap, R, t_s, t_v, PH_CAP_C, S.scratch.p, S.counters.p + 2,
            scratch_bytes, sink, S.flags.p + 1, nullptr, r == 0 ? S.defer[1].p : nullptr,
            S.counters.p + 7, defer_cap);
        CK(cudaGetLastError());
    }
    S.ev.end(EV_HEAVY, st);
    ctx->stats.kernel_launches += 4;
}

static void ensure_work_buffers(rtl_ctx *ctx, ClusterState &S, int64_t M) {
    const int W = ctx->wave;
    S.taken.need(M);
    S.owner.need(M);
    S.owner_rev.need(M);
    S.best.need(M + 1);
    S.item_read.need(M);
    S.cand.need(W);
    S.seed_item.need(W);
    S.is_seed.need(W);
    S.acc.need((size_t)W * W);
    S.wave.need(4);
    S.cut.need(4097);
    S.tasks.need(ctx->task_cap);
    S.surv.need(ctx->task_cap / 4 + 1024);
    S.defer[0].need(1 << 16);
    S.defer[1].need(1 << 16);
    S.counters.need(8);
    S.flags.need(4);
    S.scratch.need((size_t)ctx->scratch_mb << 20);
    S.h_wave.need(4);
    S.h_flags.need(4);
    S.h_counters.need(8);
}

// One greedy pass (cluster.cpp:124-166 with items = reads, :174-245 with items = cluster representatives).
// Result: owner[j] = item index of the seed that took item j (owner[j]==j for seeds), owner_rev[j] = rev flag.
// Batched clustering (h_item_seg != nullptr): the items are the concatenation of independent problems ("segments",
// contiguous item ranges; main.cpp:281-324 clusters every gene's reads on their own).  Pairs only exist inside a segment
// (the scan drops the others), so one pass over all items is every segment's greedy pass at once (see the windowed loop
// below).
static void greedy_pass(rtl_ctx *ctx, int M, const int32_t *h_item_read, const int32_t *h_item_rid, double thr, bool both,
                        double t_s, double t_v,
                        std::vector<int32_t> &owner, std::vector<uint8_t> &owner_rev,
                        const int32_t *h_item_seg = nullptr, const int32_t *h_item_seg_end = nullptr) {
    ClusterState &S = state(ctx);
    cudaStream_t st = ctx->stream;
    const int W = ctx->wave;
    ensure_work_buffers(ctx, S, M);
    set_smem_attrs(S);
    const int32_t *d_item_read = nullptr;
    if (h_item_read) {
        CK(cudaMemcpyAsync(S.item_read.p, h_item_read, (size_t)M * 4, cudaMemcpyHostToDevice, st));
        ctx->stats.h2d_bytes += (int64_t)M * 4;
        d_item_read = S.item_read.p;
    }
    Memo memo{};
    if (h_item_rid && S.rid_dim) {
        CK(cudaMemcpyAsync(S.item_rid.need(M), h_item_rid, (size_t)M * 4, cudaMemcpyHostToDevice, st));
        ctx->stats.h2d_bytes += (int64_t)M * 4;
        memo.item_rid = S.item_rid.p;
        memo.bits = S.memo.p;
        memo.rid_dim = S.rid_dim;
    }
    const int32_t *d_item_seg = nullptr;
    if (h_item_seg) {
        CK(cudaMemcpyAsync(S.item_seg.need(M), h_item_seg, (size_t)M * 4, cudaMemcpyHostToDevice, st));
        ctx->stats.h2d_bytes += (int64_t)M * 4;
        d_item_seg = S.item_seg.p;
    }
    // multi-GPU: this rank's targets as a compact ascending list (key = the representative's compact id in the merge
    // rounds, so that a rank meets the pairs it has memoised again; the item index otherwise) — phase B then walks only
    // its own share instead of filtering every item inside the scan kernel
    std::vector<int32_t> shard;
    if (ctx->world > 1) {
        shard.reserve((size_t)M / ctx->world + 1);
        const bool by_rid = h_item_rid && S.rid_dim;
        for (int i = 0; i < M; ++i)
            if ((by_rid ? h_item_rid[i] : i) % ctx->world == ctx->rank) shard.push_back(i);
        CK(cudaMemcpyAsync(S.shard_list.need(shard.size() + 1), shard.data(), shard.size() * 4, cudaMemcpyHostToDevice, st));
        ctx->stats.h2d_bytes += (int64_t)shard.size() * 4;
    }
    std::vector<uint16_t> cut;
    make_cut_table(thr, cut);
    CK(cudaMemcpyAsync(S.cut.p, cut.data(), 4097 * 2, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(S.taken.p, 0, M, st));
    CK(cudaMemsetAsync(S.best.p, 0xff, (size_t)M * 4, st));
    CK(cudaMemsetAsync(S.owner.p, 0xff, (size_t)M * 4, st));
    CK(cudaMemsetAsync(S.owner_rev.p, 0, M, st));
    CK(cudaMemsetAsync(S.wave.p, 0, 16, st));
    CK(cudaMemsetAsync(S.flags.p, 0, 16, st));
    CK(cudaStreamSynchronize(st));  // `cut` is pageable host memory

    if (h_item_seg) {
        // ---- batched clustering: windows of W consecutive segments; every wave takes the first untaken item of each
        // segment of the window — a seed by construction, so there is no candidate x candidate phase — and scores the
        // seeds against the window's untaken items (the scan drops pairs of different segments).  A window is done when
        // a wave finds no seed.  Work = seeds x segment sizes, as in the reference's per-gene loops (main.cpp:281-324).
        (void)h_item_seg_end;
        std::vector<int32_t> seg_first;
        for (int i = 0; i < M; ++i)
            if (i == 0 || h_item_seg[i] != h_item_seg[i - 1]) seg_first.push_back(i);
        const int n_segs = (int)seg_first.size();
        seg_first.push_back(M);
        CK(cudaMemcpyAsync(S.seg_first.need(n_segs + 1), seg_first.data(), (size_t)(n_segs + 1) * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.seg_cur.need(n_segs + 1), seg_first.data(), (size_t)n_segs * 4, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));  // seg_first is pageable host memory
        ctx->stats.h2d_bytes += (int64_t)(2 * n_segs + 1) * 4;
        const int64_t chunk = std::max<int64_t>(1024, ctx->task_cap / (2 * (int64_t)W));
        int32_t *hw = S.h_wave.p;
        int *hf = S.h_flags.p;
        for (int w0 = 0; w0 < n_segs; w0 += W) {
            const int w1 = std::min(n_segs, w0 + W);
            const int x_lo = seg_first[w0], x_hi = seg_first[w1];
            while (true) {
                CK(cudaMemsetAsync(S.wave.p, 0, 16, st));
                k_select_seg<<<(w1 - w0 + 255) / 256, 256, 0, st>>>(S.taken.p, S.seg_first.p + w0, S.seg_cur.p + w0, w1 - w0,
                                                                    S.seed_item.p, S.wave.p, S.owner.p, S.owner_rev.p);
                CK(cudaGetLastError());
                ctx->stats.kernel_launches++;
                for (int64_t c0 = x_lo; c0 < x_hi; c0 += chunk) {
                    const int64_t c1 = std::min<int64_t>(x_hi, c0 + chunk);
                    CK(cudaMemsetAsync(S.counters.p, 0, 3 * sizeof(unsigned long long), st));
                    BvScanArgs a{};
                    a.bv_f = S.bv[0];
                    a.bv_r = S.bv[1];
                    a.bv_stride = S.bv_stride;
                    a.pc = S.pc.p;
                    a.item_read = d_item_read;
                    a.item_seg = d_item_seg;
                    a.seed_item = S.seed_item.p;
                    a.n_seeds_p = S.wave.p + 2;
                    a.tgt_list = nullptr;
                    a.t0 = (int32_t)c0;
                    a.t1 = (int32_t)c1;
                    a.taken = S.taken.p;
                    a.cut = S.cut.p;
                    a.both = both;
                    a.order_check = 0;
                    a.rank = 0;
                    a.world = 1;
                    a.ts_cap = BVS_TS;
                    a.memo = memo;
                    a.tasks = S.tasks.p;
                    a.n_tasks = S.counters.p;
                    a.task_cap = (int64_t)S.tasks.cap;
                    a.ovf = S.flags.p + 1;
                    a.pair_counter = S.counters.p + 3;
                    S.ev.begin(EV_BV, st);
                    launch_bv_scan(ctx, a, c1 - c0, W, st);
                    S.ev.end(EV_BV, st);
                    ctx->stats.kernel_launches++;
                    ctx->stats.bv_launches++;
                    TaskView tv{S.seed_item.p, nullptr, d_item_read, memo};
                    Sink sink{};
                    sink.mode = 2;
                    sink.best = S.best.p;
                    run_pair_kernels(ctx, S, tv, t_s, t_v, sink, nullptr);
                }
                k_apply<<<ctx->n_sm, 256, 0, st>>>(S.best.p, x_lo, x_hi, S.seed_item.p, S.taken.p, S.owner.p, S.owner_rev.p);
                ctx->stats.kernel_launches++;
                CK(cudaMemcpyAsync(hw, S.wave.p, 12, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(hf, S.flags.p, 16, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                ctx->stats.d2h_bytes += 28;
                S.ev.collect();
                ctx->stats.waves++;
                if (hf[1]) {
                    if (hf[1] == 3) throw CapacityError("match scratch exhausted: raise option scratch_mb");
                    if (hf[1] == 4) throw CapacityError("a read pair has 2^31 or more common k-mer matches (not representable)");
                    throw CapacityError("candidate-pair buffer exhausted: raise option task_cap");
                }
                if (hw[2] == 0) break;  // no untaken item left in the window
            }
        }
        owner.resize(M);
        owner_rev.resize(M);
        CK(cudaMemcpyAsync(owner.data(), S.owner.p, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(owner_rev.data(), S.owner_rev.p, M, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(S.h_counters.p, S.counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->stats.d2h_bytes += (int64_t)M * 5;
        ctx->stats.rounds++;
        return;
    }

    ReadView R = view(S);
    int lo = 0;  // host-known lower bound of the device cursor
    const int64_t chunk = std::max<int64_t>(1024, ctx->task_cap / (2 * (int64_t)W));
    int32_t *hw = S.h_wave.p;
    int *hf = S.h_flags.p;
    hw[3] = -1;
    while (lo < M) {
        k_select<<<1, 1024, 0, st>>>(S.taken.p, M, W, S.cand.p, S.wave.p);
        k_mark_cand<<<4, 256, 0, st>>>(S.taken.p, S.cand.p, S.wave.p, S.owner.p, S.owner_rev.p);
        CK(cudaMemsetAsync(S.acc.p, 0xff, (size_t)W * W * 4, st));
        CK(cudaMemsetAsync(S.counters.p, 0, 3 * sizeof(unsigned long long), st));
        ctx->stats.kernel_launches += 2;
        const int b_hi = M;
        // ---- phase A: candidates x candidates
        {
            BvScanArgs a{};
            a.bv_f = S.bv[0];
            a.bv_r = S.bv[1];
            a.bv_stride = S.bv_stride;
            a.pc = S.pc.p;
            a.item_read = d_item_read;
            a.item_seg = d_item_seg;
            a.seed_item = S.cand.p;
            a.n_seeds_p = S.wave.p + 1;
            a.tgt_list = S.cand.p;
            a.n_tgt_p = S.wave.p + 1;
            a.t0 = 0;
            a.t1 = W;
            a.taken = nullptr;
            a.cut = S.cut.p;
            a.both = both;
            a.order_check = 1;
            a.rank = ctx->rank;
            a.world = ctx->world;
            a.ts_cap = BVS_TS;
            a.memo = memo;
            a.tasks = S.tasks.p;
            a.n_tasks = S.counters.p;
            a.task_cap = (int64_t)S.tasks.cap;
            a.ovf = S.flags.p + 1;
            a.pair_counter = S.counters.p + 3;
            S.ev.begin(EV_BV, st);
            launch_bv_scan(ctx, a, W, W, st);
            S.ev.end(EV_BV, st);
            ctx->stats.kernel_launches++;
            ctx->stats.bv_launches++;
            TaskView tv{S.cand.p, S.cand.p, d_item_read, memo};
            Sink sink{};
            sink.mode = 1;
            sink.acc = S.acc.p;
            sink.W = W;
            run_pair_kernels(ctx, S, tv, t_s, t_v, sink, nullptr);
            if (ctx->world > 1 && ctx->allreduce(ctx->allreduce_user, S.acc.p, (int64_t)W * W) != 0)
                throw CudaError("allreduce callback failed");
            if (W <= 1024 && (W & 31) == 0) {
                const size_t smem = ((size_t)W * (W / 32) + W / 32) * 4;
                k_resolve_cta<<<1, 1024, smem, st>>>(S.acc.p, W, S.cand.p, S.wave.p, S.seed_item.p, S.is_seed.p, S.owner.p,
                                                     S.owner_rev.p);
            } else {
                k_resolve<<<1, 32, 0, st>>>(S.acc.p, W, S.cand.p, S.wave.p, S.seed_item.p, S.is_seed.p, S.owner.p,
                                            S.owner_rev.p);
            }
            ctx->stats.kernel_launches++;
        }
        // ---- phase B: seeds x every later untaken item (items below the cursor are all taken)
        const int b_lo = lo;
        // target space: items [b_lo, M), or — sharded — the entries of this rank's list from the first one >= b_lo
        const bool sharded = ctx->world > 1;
        const int64_t x_lo = sharded ? (int64_t)(std::lower_bound(shard.begin(), shard.end(), b_lo) - shard.begin()) : b_lo;
        const int64_t x_hi = sharded ? (int64_t)shard.size() : b_hi;
        for (int64_t c0 = x_lo; c0 < x_hi; c0 += chunk) {
            const int64_t c1 = std::min<int64_t>(x_hi, c0 + chunk);
            CK(cudaMemsetAsync(S.counters.p, 0, 3 * sizeof(unsigned long long), st));
            BvScanArgs a{};
            a.bv_f = S.bv[0];
            a.bv_r = S.bv[1];
            a.bv_stride = S.bv_stride;
            a.pc = S.pc.p;
            a.item_read = d_item_read;
            a.item_seg = d_item_seg;
            a.seed_item = S.seed_item.p;
            a.n_seeds_p = S.wave.p + 2;
            a.tgt_list = sharded ? S.shard_list.p + c0 : nullptr;
            a.t0 = sharded ? 0 : (int32_t)c0;
            a.t1 = sharded ? (int32_t)(c1 - c0) : (int32_t)c1;
            a.taken = S.taken.p;
            a.cut = S.cut.p;
            a.both = both;
            a.order_check = 0;
            a.rank = ctx->rank;
            a.world = ctx->world;
            a.presharded = sharded ? 1 : 0;
            a.ts_cap = BVS_TS;
            a.memo = memo;
            a.tasks = S.tasks.p;
            a.n_tasks = S.counters.p;
            a.task_cap = (int64_t)S.tasks.cap;
            a.ovf = S.flags.p + 1;
            a.pair_counter = S.counters.p + 3;
            S.ev.begin(EV_BV, st);
            launch_bv_scan(ctx, a, c1 - c0, W, st);
            S.ev.end(EV_BV, st);
            ctx->stats.kernel_launches++;
            ctx->stats.bv_launches++;
            TaskView tv{S.seed_item.p, a.tgt_list, d_item_read, memo};
            Sink sink{};
            sink.mode = 2;
            sink.best = S.best.p;
            run_pair_kernels(ctx, S, tv, t_s, t_v, sink, nullptr);
        }
        if (b_lo < M) {
            if (ctx->world > 1) {
                k_fold_status<<<1, 1, 0, st>>>(S.flags.p, S.best.p + M);
                ctx->stats.kernel_launches++;
                if (ctx->allreduce(ctx->allreduce_user, S.best.p + b_lo, (int64_t)(M - b_lo) + 1) != 0)
                    throw CudaError("allreduce callback failed");
                CK(cudaMemcpyAsync(hw + 3, S.best.p + M, 4, cudaMemcpyDeviceToHost, st));  // wave[3] is unused by the host
            }
            k_apply<<<ctx->n_sm, 256, 0, st>>>(S.best.p, b_lo, sharded ? M : b_hi, S.seed_item.p, S.taken.p, S.owner.p,
                                               S.owner_rev.p);
            ctx->stats.kernel_launches++;
        }
        CK(cudaMemcpyAsync(hw, S.wave.p, 12, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hf, S.flags.p, 16, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->stats.d2h_bytes += 32;
        S.ev.collect();
        ctx->stats.waves++;
        if (ctx->world > 1 && b_lo < M && hw[3] == 0 && !hf[1])
            throw CapacityError("another rank ran out of candidate-pair or match scratch space: raise task_cap / scratch_mb");
        if (hf[1]) {
            if (hf[1] == 3) throw CapacityError("match scratch exhausted: raise option scratch_mb");
            if (hf[1] == 4) throw CapacityError("a read pair has 2^31 or more common k-mer matches (not representable)");
            throw CapacityError("candidate-pair buffer exhausted: raise option task_cap");
        }
        lo = hw[0];
    }
    owner.resize(M);
    owner_rev.resize(M);
    CK(cudaMemcpyAsync(owner.data(), S.owner.p, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(owner_rev.data(), S.owner_rev.p, M, cudaMemcpyDeviceToHost, st));
    unsigned long long *hc = S.h_counters.p;
    CK(cudaMemcpyAsync(hc, S.counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->stats.d2h_bytes += (int64_t)M * 5;
    ctx->stats.rounds++;
    (void)R;
}

// ------------------------------------------------------------------------------------------------ cluster_reads
struct Member {
    int32_t id;
    uint8_t rev;
};
struct Cluster {
    Member main;
    std::vector<Member> mem;
};

// cluster.cpp:67-91 (sorts the member vector in place; that order is what clusters.out stores)
static Member pick_main(std::vector<Member> &m, const std::vector<int32_t> &len, double pct) {
    const Member old = m[0];
    std::stable_sort(m.begin(), m.end(), [](const Member &a, const Member &b) { return a.id > b.id; });
    std::stable_sort(m.begin(), m.end(), [&len](const Member &a, const Member &b) { return len[a.id] > len[b.id]; });
    int nsid = (int)(m.size() * pct);
    Member ns = m[nsid];
    while (ns.rev != old.rev && (size_t)nsid < m.size() - 1) ns = m[++nsid];
    if ((size_t)nsid == m.size() - 1) return old;
    return ns;
}

void cluster_run(rtl_ctx *ctx, int k, double t_s, double t_v, double bv_thr, double bv_min, double bv_falloff,
                 double repr_pct, int is_rna, int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id,
                 uint8_t *mem_rev, int32_t *n_clusters, const uint32_t *seg_off, uint32_t n_seg, int64_t *seg_cl_off) {
    ClusterState &S = state(ctx);
    const double t_begin = now_ms();
    // batched clustering: reads [seg_off[s], seg_off[s+1]) are segment s, an independent cluster_reads problem whose
    // seq_ids count from the segment's first read; every segment's clusters come out together, in segment order
    const bool batched = seg_off != nullptr;
    std::vector<int32_t> read_seg, read_seg_end;
    if (batched) {
        if (n_seg == 0 || seg_off[0] != 0 || seg_off[n_seg] != S.n || !seg_cl_off) throw InputError("bad segment offsets");
        read_seg.resize(S.n);
        read_seg_end.resize(S.n);
        for (uint32_t s = 0; s < n_seg; ++s) {
            if (seg_off[s + 1] < seg_off[s] || seg_off[s + 1] > S.n) throw InputError("segment offsets are not monotonic");
            for (uint32_t i = seg_off[s]; i < seg_off[s + 1]; ++i) {
                read_seg[i] = (int32_t)s;
                read_seg_end[i] = (int32_t)seg_off[s + 1];
            }
        }
    }
    // (segments are bin-packed over GPUs by the caller, SURVEY.md 8e: no pair sharding inside a batched call)
    struct WorldGuard {
        rtl_ctx *c;
        int w;
        ~WorldGuard() { c->world = w; }
    } world_guard{ctx, ctx->world};
    if (batched) ctx->world = 1;
    const bool both = !is_rna;
    for (int c = 0; c < 4; ++c) S.ev.acc[c] = 0;
    S.ex_k = -1;  // one call = the whole hot path: extraction is never reused across cluster_reads calls
    cluster_extract(ctx, k, both);
    const int N = (int)S.n;
    ensure_work_buffers(ctx, S, N);
    CK(cudaMemsetAsync(S.counters.p, 0, 8 * sizeof(unsigned long long), ctx->stream));

    std::vector<int32_t> owner;
    std::vector<uint8_t> orev;
    greedy_pass(ctx, N, nullptr, nullptr, bv_thr, both, t_s, t_v, owner, orev, batched ? read_seg.data() : nullptr,
                batched ? read_seg_end.data() : nullptr);

    std::vector<Cluster> cl;
    {
        std::vector<int32_t> slot(N, -1);
        for (int i = 0; i < N; ++i)
            if (owner[i] == i) {
                slot[i] = (int32_t)cl.size();
                cl.emplace_back();
                cl.back().mem.push_back(Member{i, 0});
            }
        for (int i = 0; i < N; ++i)
            if (owner[i] != i) cl[slot[owner[i]]].mem.push_back(Member{i, orev[i]});
        parallel_for(host_threads(), cl.size(), [&](size_t c) { cl[c].main = pick_main(cl[c].mem, S.h_len, repr_pct); });
    }

    double thr = bv_thr - bv_falloff;
    bool last = false;
    std::vector<int32_t> rep, rep_rid, item_seg, item_seg_end;
    // Memo of failed k-mer tests between representatives (cluster_kernels.cuh: Memo).  A read gets a compact id when
    // it first becomes a representative; only clusters that absorbed others can change theirs, so there are fewer
    // than 2 x (clusters after the initial pass) ids in total.
    std::vector<int32_t> read_rid(N, -1);
    int32_t next_rid = 0;
    S.rid_dim = 0;
    {
        const uint64_t dim = 2 * (uint64_t)cl.size();
        const uint64_t words = (dim * dim * 2 + 31) / 32;
        if (dim > 0 && words * 4 <= (1ull << 30)) {
            S.rid_dim = (uint32_t)dim;
            CK(cudaMemsetAsync(S.memo.need(words), 0, words * 4, ctx->stream));
        }
    }
    while (thr >= bv_min || last) {
        const int M = (int)cl.size();
        rep.resize(M);
        rep_rid.resize(M);
        for (int i = 0; i < M; ++i) {
            rep[i] = cl[i].main.id;
            if (read_rid[rep[i]] < 0) read_rid[rep[i]] = next_rid++;
            rep_rid[i] = read_rid[rep[i]];
        }
        if (batched) {  // clusters stay grouped by segment: seeds are met in item order
            item_seg.resize(M);
            item_seg_end.resize(M);
            for (int i = 0; i < M; ++i) item_seg[i] = read_seg[rep[i]];
            for (int i = M - 1, end = M; i >= 0; --i) {
                if (i + 1 < M && item_seg[i + 1] != item_seg[i]) end = i + 1;
                if (i + 1 < M && item_seg[i + 1] < item_seg[i]) throw StateError("batched clustering: clusters left segment order");
                item_seg_end[i] = end;
            }
        }
        greedy_pass(ctx, M, rep.data(), rep_rid.data(), thr, both, t_s, t_v, owner, orev, batched ? item_seg.data() : nullptr,
                    batched ? item_seg_end.data() : nullptr);
        std::vector<Cluster> next;
        std::vector<int32_t> slot(M, -1);
        for (int i = 0; i < M; ++i)
            if (owner[i] == i) {
                slot[i] = (int32_t)next.size();
                next.emplace_back();
                next.back().mem = cl[i].mem;
            }
        for (int i = 0; i < M; ++i)
            if (owner[i] != i) {
                auto &dst = next[slot[owner[i]]].mem;
                for (Member s : cl[i].mem) {
                    if (orev[i]) s.rev = !s.rev;  // cluster.cpp:232-234
                    dst.push_back(s);
                }
            }
        parallel_for(host_threads(), next.size(), [&](size_t c) { next[c].main = pick_main(next[c].mem, S.h_len, repr_pct); });
        cl.swap(next);
        if (last) break;
        thr -= bv_falloff;
        if (thr < bv_min && !last) {
            last = true;
            thr = 0.0;
        }
    }

    int64_t o = 0;
    uint32_t seg_at = 0;  // batched: next segment whose first cluster has not been seen
    for (size_t c = 0; c < cl.size(); ++c) {
        int32_t base = 0;
        if (batched) {
            const uint32_t sg = (uint32_t)read_seg[cl[c].main.id];
            if (sg + 1 < seg_at) throw StateError("batched clustering: clusters left segment order");
            while (seg_at <= sg) seg_cl_off[seg_at++] = (int64_t)c;
            base = (int32_t)seg_off[sg];
        }
        main_id[c] = cl[c].main.id - base;
        main_rev[c] = cl[c].main.rev;
        cl_off[c] = o;
        for (auto &m : cl[c].mem) {
            mem_id[o] = m.id - base;
            mem_rev[o] = m.rev;
            ++o;
        }
    }
    cl_off[cl.size()] = o;
    if (batched)
        while (seg_at <= n_seg) seg_cl_off[seg_at++] = (int64_t)cl.size();
    *n_clusters = (int32_t)cl.size();
    ctx->stats.bv_pairs += (int64_t)S.h_counters.p[3];  // device counters are cumulative over the passes
    ctx->stats.full_pairs += (int64_t)S.h_counters.p[4];
    ctx->stats.heavy_pairs += (int64_t)S.h_counters.p[5];
    ctx->stats.bv_ms += S.ev.acc[EV_BV];
    ctx->stats.join_ms += S.ev.acc[EV_JOIN];
    ctx->stats.heavy_ms += S.ev.acc[EV_HEAVY];
    ctx->stats.extract_ms += S.ev.acc[EV_EXTRACT];
    ctx->stats.total_ms += now_ms() - t_begin;
}

// ------------------------------------------------------------------------------------------------ function-level entry points
void cluster_download_kmers(rtl_ctx *ctx, uint32_t *fh, int32_t *fp, uint32_t *rh, int32_t *rp, uint64_t *bf,
                            uint64_t *br) {
    ClusterState &S = state(ctx);
    const uint64_t total_k = S.total - (uint64_t)S.ex_k * S.n;
    cudaStream_t st = ctx->stream;
    if (fh) CK(cudaMemcpyAsync(fh, S.kh[0].p, total_k * 4, cudaMemcpyDeviceToHost, st));
    if (fp) CK(cudaMemcpyAsync(fp, S.kp[0].p, total_k * 4, cudaMemcpyDeviceToHost, st));
    if (S.ex_both) {
        if (rh) CK(cudaMemcpyAsync(rh, S.kh[1].p, total_k * 4, cudaMemcpyDeviceToHost, st));
        if (rp) CK(cudaMemcpyAsync(rp, S.kp[1].p, total_k * 4, cudaMemcpyDeviceToHost, st));
        if (br) CK(cudaMemcpy2DAsync(br, 512, S.bv[1], (size_t)S.bv_stride * 8, 512, S.n, cudaMemcpyDeviceToHost, st));
    } else if (br)
        memset(br, 0, (size_t)S.n * 512);
    if (bf) CK(cudaMemcpy2DAsync(bf, 512, S.bv[0], (size_t)S.bv_stride * 8, 512, S.n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
}

void cluster_bv_scan_dense(rtl_ctx *ctx, int k, int is_rna, const int32_t *seed_reads, int n_seeds,
                           const int32_t *target_reads, int n_targets, double thr, uint32_t *common, uint8_t *pass) {
    ClusterState &S = state(ctx);
    for (int i = 0; i < n_seeds; ++i)
        if (seed_reads[i] < 0 || (uint32_t)seed_reads[i] >= S.n) throw InputError("seed read index out of range");
    for (int i = 0; i < n_targets; ++i)
        if (target_reads[i] < 0 || (uint32_t)target_reads[i] >= S.n) throw InputError("target read index out of range");
    cluster_extract(ctx, k, !is_rna);
    set_smem_attrs(S);
    cudaStream_t st = ctx->stream;
    DevBuf<int32_t> d_seeds, d_tg, d_ns;
    DevBuf<uint32_t> d_common;
    DevBuf<uint8_t> d_pass;
    DevBuf<uint16_t> d_cut;
    DevBuf<unsigned long long> d_cnt;
    DevBuf<int> d_ovf;
    const size_t np = (size_t)n_seeds * n_targets;
    std::vector<uint16_t> cut;
    make_cut_table(thr, cut);
    CK(cudaMemcpyAsync(d_seeds.need(n_seeds), seed_reads, (size_t)n_seeds * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_tg.need(n_targets), target_reads, (size_t)n_targets * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_ns.need(1), &n_seeds, 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cut.need(4097), cut.data(), 4097 * 2, cudaMemcpyHostToDevice, st));
    if (common) d_common.need(np);
    if (pass) d_pass.need(np);
    CK(cudaMemsetAsync(d_cnt.need(2), 0, 16, st));
    CK(cudaMemsetAsync(d_ovf.need(1), 0, 4, st));
    BvScanArgs a{};
    a.bv_f = S.bv[0];
    a.bv_r = S.bv[1];
    a.bv_stride = S.bv_stride;
    a.pc = S.pc.p;
    a.item_read = nullptr;
    a.seed_item = d_seeds.p;
    a.n_seeds_p = d_ns.p;
    a.tgt_list = d_tg.p;
    a.n_tgt_p = nullptr;
    a.t0 = 0;
    a.t1 = n_targets;
    a.cut = d_cut.p;
    a.both = !is_rna;
    a.order_check = 0;
    a.rank = 0;
    a.world = 1;
    a.tasks = nullptr;
    a.n_tasks = d_cnt.p;
    a.ovf = d_ovf.p;
    a.dense_common = common ? d_common.p : nullptr;  // NULL outputs: scan only (timing runs)
    a.dense_pass = pass ? d_pass.p : nullptr;
    a.pair_counter = d_cnt.p + 1;
    S.ev.acc[EV_BV] = 0;
    S.ev.begin(EV_BV, st);
    launch_bv_scan(ctx, a, n_targets, n_seeds, st);
    S.ev.end(EV_BV, st);
    unsigned long long hcnt[2] = {0, 0};
    CK(cudaMemcpyAsync(hcnt, d_cnt.p, 16, cudaMemcpyDeviceToHost, st));
    if (common) CK(cudaMemcpyAsync(common, d_common.p, np * 4, cudaMemcpyDeviceToHost, st));
    if (pass) CK(cudaMemcpyAsync(pass, d_pass.p, np, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    S.ev.collect();
    // counters of this launch (tools/bv_stream_bench.py times the scan kernel alone through them)
    ctx->stats.bv_ms = S.ev.acc[EV_BV];
    ctx->stats.bv_pairs = (int64_t)hcnt[1];
    ctx->stats.bv_launches = 1;
    ctx->stats.kernel_launches = 1;
}

void cluster_pair_similarity(rtl_ctx *ctx, int k, int is_rna, const int32_t *a_read, const int32_t *b_read,
                             const uint8_t *strand, int64_t n_tasks, double t_s, double t_v, int64_t *n_common,
                             int32_t *bases, int32_t *n_dist, double *var, uint8_t *accept) {
    ClusterState &S = state(ctx);
    cluster_extract(ctx, k, !is_rna);
    set_smem_attrs(S);
    cudaStream_t st = ctx->stream;
    if (n_tasks >= (1ll << 31)) throw CapacityError("too many tasks");
    std::vector<uint64_t> tasks(n_tasks);
    for (int64_t i = 0; i < n_tasks; ++i) {
        if (a_read[i] < 0 || (uint32_t)a_read[i] >= S.n || b_read[i] < 0 || (uint32_t)b_read[i] >= S.n)
            throw InputError("task read index out of range");
        if (strand[i] && is_rna) throw InputError("reverse-strand task on an RNA (forward-only) extraction");
        tasks[i] = make_task((uint32_t)i, strand[i], (uint32_t)i);
    }
    DevBuf<int32_t> d_a, d_b, d_bases, d_nd;
    DevBuf<int64_t> d_nm;
    DevBuf<double> d_var;
    DevBuf<uint8_t> d_acc;
    S.tasks.need(std::max<int64_t>(n_tasks, 1));
    S.surv.need(std::max<int64_t>(n_tasks, 1));
    S.defer[0].need(1 << 16);
    S.defer[1].need(1 << 16);
    S.counters.need(8);
    S.flags.need(4);
    S.scratch.need((size_t)ctx->scratch_mb << 20);
    CK(cudaMemcpyAsync(S.tasks.p, tasks.data(), n_tasks * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_a.need(n_tasks), a_read, n_tasks * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_b.need(n_tasks), b_read, n_tasks * 4, cudaMemcpyHostToDevice, st));
    unsigned long long cnt[4] = {(unsigned long long)n_tasks, 0, 0, 0};
    CK(cudaMemcpyAsync(S.counters.p, cnt, sizeof(cnt), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(S.flags.p, 0, 16, st));
    CK(cudaMemsetAsync(d_bases.need(n_tasks), 0xff, n_tasks * 4, st));
    CK(cudaMemsetAsync(d_nd.need(n_tasks), 0, n_tasks * 4, st));
    CK(cudaMemsetAsync(d_var.need(n_tasks), 0, n_tasks * 8, st));
    CK(cudaMemsetAsync(d_acc.need(n_tasks), 0, n_tasks, st));
    d_nm.need(n_tasks);
    TaskView tv{d_a.p, d_b.p, nullptr};
    Sink sink{};
    sink.mode = 0;
    sink.bases = d_bases.p;
    sink.n_dist = d_nd.p;
    sink.var = d_var.p;
    sink.accept = d_acc.p;
    run_pair_kernels(ctx, S, tv, t_s, t_v, sink, d_nm.p);
    int hf[4];
    CK(cudaMemcpyAsync(hf, S.flags.p, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(n_common, d_nm.p, n_tasks * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(bases, d_bases.p, n_tasks * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(n_dist, d_nd.p, n_tasks * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(var, d_var.p, n_tasks * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(accept, d_acc.p, n_tasks, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    S.ev.collect();
    if (hf[1]) throw CapacityError("scratch or survivor buffer exhausted in pair_similarity");
}
