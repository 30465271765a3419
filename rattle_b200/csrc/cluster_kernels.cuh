// sm_100a kernels of hot path A (greedy k-mer/bitvector clustering).  No tensor-core use: every kernel here is
// integer/byte work bounded by HBM bandwidth, POPC issue or shared-memory latency.
//
//   k_extract_*     kmer.cpp:6-42       k-mer (hash,pos) lists sorted by (hash,pos) + 4096-bit 6-mer bitvectors
//   k_bv_scan       cluster.cpp:13-19,43 bitvector AND+popcount filter, seeds tiled in shared memory
//   k_join_count    kmer.cpp:45-67      size of the sorted-list multiset join (and the exact reject bound)
//   k_pair_heavy    kmer.cpp:45-67 + similarity.cpp:4-97 + utils.cpp:36-55 + cluster.cpp:24-37
//   k_select / k_resolve / k_apply      the greedy bookkeeping of cluster.cpp:124-166,171-245, batched in waves
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cluster_aux_kernels.cuh"

namespace rtl {

// device view of the resident read set
struct ReadView {
    const uint8_t *bases;
    const uint64_t *off;   // n+1
    const int32_t *len;    // n
    const uint32_t *kh[2]; // sorted k-mer hashes, forward / reverse-complement strand
    const int32_t *kp[2];  // positions
    const uint64_t *bv[2]; // read r's 64 words at bv[s] + r * bv_stride (both strands interleaved: stride 128)
    const int32_t *pc;     // popcount of bv[0]
    int k;
    uint32_t n;
    int bv_stride;
    __device__ __forceinline__ uint64_t koff(uint32_t r) const { return off[r] - (uint64_t)k * r; }
};

// ------------------------------------------------------------------------------------------------ K2: extraction
// One CTA per (read, strand).  Keys (hash<<32 | pos) are bitonic-sorted in shared memory.
// smem: keys[n_pad] u64 | bvw[128] u32 | codes[n_pad+32] u8
__global__ void k_extract_smem(const uint32_t *__restrict__ pk, const uint64_t *__restrict__ off,
                               const uint32_t *__restrict__ read_list, int k, int n_pad, uint32_t *kh_f, int32_t *kp_f,
                               uint32_t *kh_r, int32_t *kp_r, uint64_t *bv_f, uint64_t *bv_r, int bv_stride, int32_t *pc, int *err) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t *keys = (uint64_t *)sm_raw;
    uint32_t *bvw = (uint32_t *)(keys + n_pad);
    uint8_t *codes = (uint8_t *)(bvw + 128);
    const uint32_t r = read_list[blockIdx.x];
    const int strand = blockIdx.y;
    const uint64_t o = off[r];
    const int len = (int)(off[r + 1] - o);
    const int n = len - k;
    const int tid = threadIdx.x, nt = blockDim.x;

    (void)err;
    const uint64_t w0 = pk_start(off, r);
    for (int w = tid; w * 16 < len; w += nt) {  // one packed word = 16 bases per thread and step
        const uint32_t word = pk[w0 + w];
        const int m = min(16, len - w * 16);
        for (int i = 0; i < m; ++i) {
            const int p = w * 16 + i;
            const int c = (int)((word >> (2 * i)) & 3u);
            if (strand == 0) codes[p] = (uint8_t)c;
            else codes[len - 1 - p] = (uint8_t)(c ^ 2);  // reverse complement: A<->T, C<->G  (utils.cpp:15-24)
        }
    }
    if (tid < 128) bvw[tid] = 0;
    __syncthreads();
    for (int p = tid; p < n_pad; p += nt) {
        uint64_t key = ~0ull;
        if (p < n) {
            uint32_t h = 0;
            for (int i = 0; i < k; ++i) h = (h << 2) | codes[p + i];
            key = ((uint64_t)h << 32) | (uint32_t)p;
        }
        keys[p] = key;
    }
    for (int p = tid; p < len - 6; p += nt) {  // 6-mers at pos 0..len-7 (kmer.cpp:28)
        uint32_t h = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) h = (h << 2) | codes[p + i];
        atomicOr(&bvw[h >> 5], 1u << (h & 31));
    }
    // bitonic sort, ascending
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (n_pad >> 1); t += nt) {
                int i = 2 * t - (t & (stride - 1));
                int j = i + stride;
                uint64_t a = keys[i], b = keys[j];
                bool asc = (i & size) == 0;
                if ((a > b) == asc) {
                    keys[i] = b;
                    keys[j] = a;
                }
            }
        }
    }
    __syncthreads();
    const uint64_t ko = o - (uint64_t)k * r;
    uint32_t *kh = strand ? kh_r : kh_f;
    int32_t *kp = strand ? kp_r : kp_f;
    for (int p = tid; p < n; p += nt) {
        uint64_t key = keys[p];
        kh[ko + p] = (uint32_t)(key >> 32);
        kp[ko + p] = (int32_t)(uint32_t)key;
    }
    uint64_t *bv = strand ? bv_r : bv_f;
    if (tid < 64) bv[(uint64_t)r * bv_stride + tid] = (uint64_t)bvw[2 * tid] | ((uint64_t)bvw[2 * tid + 1] << 32);
    if (strand == 0 && tid < 32) {
        int c = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) c += __popc(bvw[tid * 4 + w]);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
        if (tid == 0) pc[r] = c;
    }
}

// Same, for reads whose key list does not fit shared memory: keys live in a global scratch segment.
__global__ void k_extract_long(const uint32_t *__restrict__ pk, const uint64_t *__restrict__ off,
                               const uint32_t *__restrict__ read_list, const uint64_t *__restrict__ scratch_off,
                               uint64_t *scratch, int k, uint32_t *kh_f, int32_t *kp_f, uint32_t *kh_r, int32_t *kp_r,
                               uint64_t *bv_f, uint64_t *bv_r, int bv_stride, int32_t *pc, int *err) {
    __shared__ uint32_t bvw[128];
    const uint32_t r = read_list[blockIdx.x];
    const int strand = blockIdx.y;
    const uint64_t o = off[r];
    const int len = (int)(off[r + 1] - o);
    const int n = len - k;
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    uint64_t *keys = scratch + scratch_off[blockIdx.x * 2 + strand];
    const int tid = threadIdx.x, nt = blockDim.x;
    (void)err;
    const uint64_t w0 = pk_start(off, r);
    auto code_at = [&](int p) -> int { return strand == 0 ? pk_code(pk, w0, p) : (pk_code(pk, w0, len - 1 - p) ^ 2); };
    if (tid < 128) bvw[tid] = 0;
    __syncthreads();
    for (int p = tid; p < n_pad; p += nt) {
        uint64_t key = ~0ull;
        if (p < n) {
            uint32_t h = 0;
            for (int i = 0; i < k; ++i) h = (h << 2) | (uint32_t)code_at(p + i);
            key = ((uint64_t)h << 32) | (uint32_t)p;
        }
        keys[p] = key;
    }
    for (int p = tid; p < len - 6; p += nt) {
        uint32_t h = 0;
        for (int i = 0; i < 6; ++i) h = (h << 2) | (uint32_t)code_at(p + i);
        atomicOr(&bvw[h >> 5], 1u << (h & 31));
    }
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (n_pad >> 1); t += nt) {
                int i = 2 * t - (t & (stride - 1));
                int j = i + stride;
                uint64_t a = keys[i], b = keys[j];
                bool asc = (i & size) == 0;
                if ((a > b) == asc) {
                    keys[i] = b;
                    keys[j] = a;
                }
            }
        }
    }
    __syncthreads();
    const uint64_t ko = o - (uint64_t)k * r;
    uint32_t *kh = strand ? kh_r : kh_f;
    int32_t *kp = strand ? kp_r : kp_f;
    for (int p = tid; p < n; p += nt) {
        uint64_t key = keys[p];
        kh[ko + p] = (uint32_t)(key >> 32);
        kp[ko + p] = (int32_t)(uint32_t)key;
    }
    uint64_t *bv = strand ? bv_r : bv_f;
    if (tid < 64) bv[(uint64_t)r * bv_stride + tid] = (uint64_t)bvw[2 * tid] | ((uint64_t)bvw[2 * tid + 1] << 32);
    if (strand == 0 && tid < 32) {
        int c = 0;
        for (int w = 0; w < 4; ++w) c += __popc(bvw[tid * 4 + w]);
        for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
        if (tid == 0) pc[r] = c;
    }
}

// extraction prologue: clear the flags, carrying over "the upload met a base outside ACGTU" (k_pack_bases)
__global__ void k_clear_keep_pack_flag(int *flags, const int *pack_flag) {
    flags[0] = *pack_flag ? 2 : 0;
    flags[1] = flags[2] = flags[3] = 0;
}

__global__ void k_lengths(const uint64_t *__restrict__ off, int32_t *len, uint32_t n, int k, int *err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int l = (int)(off[i + 1] - off[i]);
        len[i] = l;
        if (l <= k || l <= 6) atomicExch(err, 1);
    }
}

// ------------------------------------------------------------------------------------------------ tasks
// task = t (bits 0..31) | strand (bit 32) | s (bits 33..63)
//   s indexes `seed_item` (or is the item itself), t indexes `tgt_list` (or is the item itself);
//   items map to reads through `item_read` (or are reads themselves).
// Merge rounds (items = clusters, compared through their representative reads): cluster_together(a, b, thr) is a pure
// function of the two reads whose k-mer test does not depend on thr (cluster.cpp:24-37,48-61), and the merge rounds
// compare the same surviving representatives again at every threshold (cluster.cpp:171-255).  `memo` remembers the
// (seed representative, target representative, strand) triples whose k-mer test failed: bit
// ((rid_a * rid_dim + rid_b) * 2 + strand), rid = compact id of a read that is or was a representative.  The scan drops
// such pairs instead of emitting a task — the outcome ("no match") is known.
struct Memo {
    const int32_t *item_rid;  // nullable: no memo (initial pass over reads, function-level entry points)
    uint32_t *bits;
    uint32_t rid_dim;
    __device__ __forceinline__ bool known_failure(uint32_t a_item, uint32_t b_item, int strand) const {
        if (!item_rid) return false;
        const uint32_t ra = (uint32_t)item_rid[a_item], rb = (uint32_t)item_rid[b_item];
        if (ra >= rid_dim || rb >= rid_dim) return false;
        const uint64_t bit = ((uint64_t)ra * rid_dim + rb) * 2 + (uint32_t)strand;
        return (bits[bit >> 5] >> (bit & 31)) & 1u;
    }
    __device__ __forceinline__ void record_failure(uint32_t a_item, uint32_t b_item, int strand) const {
        if (!item_rid) return;
        const uint32_t ra = (uint32_t)item_rid[a_item], rb = (uint32_t)item_rid[b_item];
        if (ra >= rid_dim || rb >= rid_dim) return;
        const uint64_t bit = ((uint64_t)ra * rid_dim + rb) * 2 + (uint32_t)strand;
        atomicOr(&bits[bit >> 5], 1u << (bit & 31));
    }
};

struct TaskView {
    const int32_t *seed_item;
    const int32_t *tgt_list;
    const int32_t *item_read;
    Memo memo;
    __device__ __forceinline__ void items(uint64_t task, uint32_t &a_item, uint32_t &b_item) const {
        const uint32_t t = (uint32_t)task, s = (uint32_t)(task >> 33);
        a_item = seed_item ? (uint32_t)seed_item[s] : s;
        b_item = tgt_list ? (uint32_t)tgt_list[t] : t;
    }
    __device__ __forceinline__ void decode(uint64_t task, uint32_t &a_read, uint32_t &b_read, int &strand) const {
        uint32_t t = (uint32_t)task;
        strand = (int)((task >> 32) & 1);
        uint32_t s = (uint32_t)(task >> 33);
        uint32_t ai = seed_item ? (uint32_t)seed_item[s] : s;
        uint32_t bi = tgt_list ? (uint32_t)tgt_list[t] : t;
        a_read = item_read ? (uint32_t)item_read[ai] : ai;
        b_read = item_read ? (uint32_t)item_read[bi] : bi;
    }
};
__host__ __device__ __forceinline__ uint64_t make_task(uint32_t s, int strand, uint32_t t) {
    return ((uint64_t)s << 33) | ((uint64_t)(strand & 1) << 32) | t;
}

// ------------------------------------------------------------------------------------------------ K3: bitvector scan
// grid.x strides over targets (one warp per target), grid.y = seed tiles of `TS` seeds staged in shared memory.
// Each lane owns two 64-bit words of the target's forward and reverse bitvector (one coalesced uint4 load each);
// per seed: LDS.128 of the seed words, AND, POPC, 5-step butterfly; 32 seeds are evaluated per threshold/ballot step.
// Emits (seed slot, target, strand) tasks for pairs that pass cluster.cpp:19 / :43, or dense results for tests.
constexpr int BVS_TS = 128;       // seeds per shared-memory tile (64 KB)
constexpr int BVS_THREADS = 256;  // 8 warps
struct BvScanArgs {
    const uint64_t *bv_f, *bv_r;
    int bv_stride;             // uint64 words between consecutive reads' bitvectors (128: [fwd 64 | rev 64] per read)
    const int32_t *pc;
    const int32_t *item_read;  // nullable
    const int32_t *item_seg;   // nullable: segment of every item (batched clustering: only pairs inside a segment exist)
    const int32_t *seed_item;  // n_seeds (device count in *n_seeds_p)
    const int32_t *n_seeds_p;
    const int32_t *tgt_list;   // nullable: explicit target items
    const int32_t *n_tgt_p;    // nullable: device count for tgt_list
    int32_t t0, t1;            // target range when tgt_list == nullptr (t1 also caps the list)
    const uint8_t *taken;      // nullable
    const uint16_t *cut;       // 4097 entries: min common count that passes at mmax
    int both;                  // reverse strand evaluated
    int order_check;           // require seed item < target item
    int rank, world;           // target sharding
    int presharded;            // tgt_list already holds only this rank's targets (no filtering in the kernel)
    int ts_cap;                // seeds per tile the shared memory is sized for (<= BVS_TS; fewer seeds -> more CTAs/SM)
    Memo memo;                 // known k-mer-test failures between representatives (merge rounds)
    uint64_t *tasks;
    unsigned long long *n_tasks;
    int64_t task_cap;
    int *ovf;
    uint32_t *dense_common;    // nullable (tests): [n_seeds x n_targets]
    uint8_t *dense_pass;
    unsigned long long *pair_counter;  // evaluated pairs
};

// mbarrier + bulk-copy (TMA) primitives
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned); completes on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// MINB = CTAs per SM the register allocation aims at: 4 (64 registers, 32 warps per SM) keeps a third more loads in flight
// in the streaming regime than the 80 registers / 3 CTAs the compiler picks on its own (option bv_kernel=3)
template <int MINB>
__global__ void __launch_bounds__(BVS_THREADS, MINB) k_bv_scan(BvScanArgs A) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int TS = A.ts_cap;
    uint64_t *sseed = (uint64_t *)sm_raw;             // [TS][64]
    int32_t *sitem = (int32_t *)(sseed + TS * 64);    // [TS]
    int32_t *spc = sitem + TS;                        // [TS]
    const int n_seeds = *A.n_seeds_p;
    const int s0 = blockIdx.y * TS;
    if (s0 >= n_seeds) return;
    const int ts = min(TS, n_seeds - s0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the seed tile comes in through the bulk-copy engine (TMA): one 512-byte copy per seed, completion on an mbarrier
    __shared__ unsigned long long s_seedbar;
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&s_seedbar);
    if (tid == 0) {
        mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) mbar_arrive_expect_tx(bar_s, (uint32_t)ts * 512u);
    __syncthreads();
    {
        const uint32_t seed_s = (uint32_t)__cvta_generic_to_shared(sseed);
        for (int s = tid; s < ts; s += BVS_THREADS) {
            const int it = A.seed_item[s0 + s];
            const uint32_t rd = A.item_read ? (uint32_t)A.item_read[it] : (uint32_t)it;
            sitem[s] = it;
            spc[s] = A.pc[rd];
            bulk_g2s(seed_s + (uint32_t)s * 512u, A.bv_f + (uint64_t)rd * A.bv_stride, 512u, bar_s);
        }
    }
    __syncthreads();
    mbar_wait(bar_s, 0);
    int n_t = A.tgt_list ? (A.n_tgt_p ? *A.n_tgt_p : A.t1) : (A.t1 - A.t0);
    unsigned long long my_pairs = 0;
    // Half a warp per target (two targets per warp iteration): every lane holds 32 bytes of the target's forward
    // and reverse bitvector, so one warp instruction scores a seed against two reads and the cross-lane sum is a
    // 4-step butterfly over 16 lanes.  The loop is software-pipelined two iterations deep: stage A resolves WHICH
    // read the target two iterations ahead is (target list, taken flag, cluster -> representative read: one
    // dependent load), stage B requests the bitvectors of the next iteration's target, and the current one is scored.
    const int half = lane >> 4, sub = lane & 15;
    const int stride = gridDim.x * (BVS_THREADS / 32) * 2;
    struct Pre {  // stage A
        int tslot, item;
        uint32_t rd;
        bool live;
    };
    struct Tgt {  // stage B
        ulonglong2 f0, f1, r0, r1;
        int tslot, item, pcj;
        bool live;
    };
    auto stage_a = [&](int x) -> Pre {
        Pre p;
        p.tslot = p.item = 0;
        p.rd = 0;
        p.live = false;
        if (x >= n_t) return p;
        p.tslot = A.tgt_list ? x : A.t0 + x;  // value stored in the task
        p.item = A.tgt_list ? A.tgt_list[x] : p.tslot;
        // multi-GPU: targets are sharded by a key that is stable over the merge rounds (the representative's compact
        // id), so that a rank meets the pairs it has memoised again
        if (A.world > 1 && !A.presharded &&
            ((A.memo.item_rid ? A.memo.item_rid[p.item] : p.item) % A.world) != A.rank)
            return p;
        const bool taken = A.taken ? (A.taken[p.item] != 0) : false;
        p.rd = A.item_read ? (uint32_t)A.item_read[p.item] : (uint32_t)p.item;
        p.live = !taken;
        return p;
    };
    auto stage_b = [&](const Pre &p) -> Tgt {
        Tgt t;
        t.f0 = t.f1 = t.r0 = t.r1 = make_ulonglong2(0, 0);
        t.tslot = p.tslot;
        t.item = p.item;
        t.pcj = 0;
        t.live = p.live;
        if (!p.live) return t;
        const ulonglong2 *pf = reinterpret_cast<const ulonglong2 *>(A.bv_f + (uint64_t)p.rd * A.bv_stride + 4 * sub);
        t.f0 = pf[0];
        t.f1 = pf[1];
        if (A.both) {
            const ulonglong2 *pr = reinterpret_cast<const ulonglong2 *>(A.bv_r + (uint64_t)p.rd * A.bv_stride + 4 * sub);
            t.r0 = pr[0];
            t.r1 = pr[1];
        }
        t.pcj = A.pc[p.rd];
        return t;
    };
    int x = (blockIdx.x * (BVS_THREADS / 32) + warp) * 2 + half;  // my half-warp's target
    Tgt nxt = stage_b(stage_a(x));
    Pre pre = stage_a(x + stride);
    for (int xw = x - half; xw < n_t; xw += stride, x += stride) {  // xw: warp-uniform loop variable
        const Tgt cur = nxt;
        nxt = stage_b(pre);
        pre = stage_a(x + 2 * stride);
        if (!__any_sync(0xffffffffu, cur.live)) continue;
        const int tslot = cur.tslot, item = cur.item, pcj = cur.pcj;
        for (int g = 0; g < ts; g += 16) {
            uint32_t mine = 0;
            const int lim = min(16, ts - g);
            for (int q = 0; q < lim; ++q) {
                const ulonglong2 *sp = reinterpret_cast<const ulonglong2 *>(sseed + (g + q) * 64 + 4 * sub);
                const ulonglong2 s0v = sp[0], s1v = sp[1];
                uint32_t c = (uint32_t)(__popcll(s0v.x & cur.f0.x) + __popcll(s0v.y & cur.f0.y) + __popcll(s1v.x & cur.f1.x) +
                                        __popcll(s1v.y & cur.f1.y));
                c |= (uint32_t)(__popcll(s0v.x & cur.r0.x) + __popcll(s0v.y & cur.r0.y) + __popcll(s1v.x & cur.r1.x) +
                                __popcll(s1v.y & cur.r1.y))
                     << 16;
#pragma unroll
                for (int sft = 8; sft > 0; sft >>= 1) c += __shfl_xor_sync(0xffffffffu, c, sft);
                if (sub == q) mine = c;
            }
            const int s = g + sub;
            bool valid = cur.live && sub < lim;
            if (valid && A.order_check) valid = sitem[s] < item;
            if (valid && A.item_seg) valid = A.item_seg[sitem[s]] == A.item_seg[item];
            const uint32_t cf = mine & 0xffffu, cr = mine >> 16;
            bool pf = false, pr = false;
            if (valid) {
                const int mmax = max(spc[s], pcj);
                const uint32_t cutv = A.cut[mmax];
                pf = cf >= cutv;
                pr = A.both && cr >= cutv;
                if (pf && A.memo.known_failure((uint32_t)sitem[s], (uint32_t)item, 0)) pf = false;
                if (pr && A.memo.known_failure((uint32_t)sitem[s], (uint32_t)item, 1)) pr = false;
            }
            my_pairs += (unsigned long long)__popc(__ballot_sync(0xffffffffu, valid));
            if (A.dense_common) {
                if (cur.live && sub < lim) {
                    size_t idx = (size_t)(s0 + s) * (size_t)n_t + (size_t)x;
                    A.dense_common[idx] = mine;
                    A.dense_pass[idx] = (uint8_t)((pf ? 1 : 0) | (pr ? 2 : 0));
                }
            }
            if (A.tasks) {
                const uint32_t mf = __ballot_sync(0xffffffffu, pf), mr = __ballot_sync(0xffffffffu, pr);
                const int tot = __popc(mf) + __popc(mr);
                if (tot) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(A.n_tasks, (unsigned long long)tot);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base + tot > (unsigned long long)A.task_cap) {
                        if (lane == 0) atomicExch(A.ovf, 1);
                    } else {
                        const uint32_t below = (1u << lane) - 1u;
                        if (pf) A.tasks[base + __popc(mf & below)] = make_task((uint32_t)(s0 + s), 0, (uint32_t)tslot);
                        if (pr)
                            A.tasks[base + __popc(mf) + __popc(mr & below)] =
                                make_task((uint32_t)(s0 + s), 1, (uint32_t)tslot);
                    }
                }
            }
        }
    }
    if (lane == 0 && my_pairs && A.pair_counter) atomicAdd(A.pair_counter, my_pairs);
}


// ------------------------------------------------------------------------------------------------ K3, streaming regime
// The same scan for FEW seeds (at most BVT_TS), where the kernel is bound by HBM bandwidth: the reads' bitvectors are
// streamed through a shared-memory ring by the bulk-copy engine (TMA, cp.async.bulk + mbarrier).  Two producer warps
// resolve which reads the next targets are (target list, taken flag, cluster -> representative) and issue one 512-byte
// bulk copy per strand and target into the ring, far ahead of the six consumer warps, so that the bytes in flight do
// not depend on registers and the consumers spend ~40 warp instructions per 1-KB target (the register-staged kernel
// above needs ~240 and is issue-bound in this regime).  The seeds' bitvectors come in by bulk copies too.
// A consumer warp takes one ring slot = 4 targets, a quarter-warp per target: a lane holds 64 B of the target's forward
// and reverse bitvector and scores a seed with 4 LDS.128 (the seed; broadcast across the quarters) + 16 AND + 16 POPC.64.
// Every slot has ONE producer and ONE consumer (slot (c, d) = consumer c, iteration parity d = producer d), so that
// each mbarrier is waited on in phase order by a single warp.
constexpr int BVT_THREADS = 256;   // warps 0-1 produce, warps 2-7 consume
constexpr int BVT_PROD = 2, BVT_CONS = 6;
constexpr int BVT_SLOT_TGT = 4;    // targets per ring slot
constexpr int BVT_SLOT_BYTES = BVT_SLOT_TGT * 1024;
constexpr int BVT_RING = BVT_CONS * BVT_PROD;  // 12 slots = 48 KB of bitvectors in flight per CTA
constexpr int BVT_TS = 16;         // seeds at most

__host__ __device__ __forceinline__ size_t bvt_smem_bytes(int ts) {
    return (size_t)ts * 512 + (size_t)ts * 8 + (size_t)BVT_RING * BVT_SLOT_BYTES + (size_t)BVT_RING * BVT_SLOT_TGT * 16 +
           (size_t)(2 * BVT_RING + 1) * 8 + 128;
}

__global__ void __launch_bounds__(BVT_THREADS, 3) k_bv_stream(BvScanArgs A) {
    extern __shared__ __align__(16) unsigned char sm_raw[];  // (dynamic shared memory starts 1024-byte aligned)
    const int TS = A.ts_cap;
    const int n_seeds = *A.n_seeds_p;
    const int s0 = blockIdx.y * TS;
    if (s0 >= n_seeds) return;
    const int ts = min(TS, n_seeds - s0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // layout: seeds [TS][512 B] | ring [12][4][fwd 512 | rev 512] | sitem [TS] | spc [TS] | meta [12][4] int4 | barriers
    unsigned char *sseed = sm_raw;
    unsigned char *ring = sseed + (size_t)TS * 512;
    int32_t *sitem = (int32_t *)(ring + (size_t)BVT_RING * BVT_SLOT_BYTES);
    int32_t *spc = sitem + TS;
    int4 *meta = (int4 *)(((uintptr_t)(spc + TS) + 15) & ~(uintptr_t)15);
    unsigned long long *bars = (unsigned long long *)(meta + BVT_RING * BVT_SLOT_TGT);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t seed_s = (uint32_t)__cvta_generic_to_shared(sseed);
    const uint32_t full_s = (uint32_t)__cvta_generic_to_shared(bars);  // [12]
    const uint32_t empty_s = full_s + BVT_RING * 8;                     // [12]
    const uint32_t seedbar_s = empty_s + BVT_RING * 8;
    if (tid == 0) {
        for (int i = 0; i < BVT_RING; ++i) {
            mbar_init(full_s + i * 8, 1);
            mbar_init(empty_s + i * 8, 1);
        }
        mbar_init(seedbar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // ---- seeds: one bulk copy each
    if (tid == 0) mbar_arrive_expect_tx(seedbar_s, (uint32_t)ts * 512u);
    __syncthreads();
    for (int s = tid; s < ts; s += BVT_THREADS) {
        const int it = A.seed_item[s0 + s];
        const uint32_t rd = A.item_read ? (uint32_t)A.item_read[it] : (uint32_t)it;
        sitem[s] = it;
        spc[s] = A.pc[rd];
        bulk_g2s(seed_s + (uint32_t)s * 512u, A.bv_f + (uint64_t)rd * A.bv_stride, 512u, seedbar_s);
    }
    __syncthreads();
    const int n_t = A.tgt_list ? (A.n_tgt_p ? *A.n_tgt_p : A.t1) : (A.t1 - A.t0);
    const long long n_groups = ((long long)n_t + BVT_SLOT_TGT - 1) / BVT_SLOT_TGT;
    const uint32_t per_tgt = A.both ? 1024u : 512u;

    if (warp < BVT_PROD) {
        // ---------------- producers: warp p fills the slots of the iterations it = p, p+2, ... (6 slots = 24 targets each)
        const int sub = lane & 3, c = lane >> 2;  // quad c feeds consumer c
        for (int it = warp;; it += BVT_PROD) {
            if ((long long)blockIdx.x + (long long)it * BVT_CONS * gridDim.x >= n_groups) break;  // warp-uniform
            const long long q = (long long)blockIdx.x + ((long long)it * BVT_CONS + c) * gridDim.x;
            const bool slot_valid = c < BVT_CONS && q < n_groups;
            const int slot = c * BVT_PROD + warp, use = it / BVT_PROD;
            const long long x = q * BVT_SLOT_TGT + sub;
            int tslot = 0, item = 0, pcj = 0;
            uint32_t rd = 0;
            bool live = false;
            if (slot_valid && x < n_t) {
                tslot = A.tgt_list ? (int)x : A.t0 + (int)x;
                item = A.tgt_list ? A.tgt_list[x] : tslot;
                bool mine = true;
                if (A.world > 1 && !A.presharded &&
                    ((A.memo.item_rid ? A.memo.item_rid[item] : item) % A.world) != A.rank)
                    mine = false;
                if (mine && !(A.taken && A.taken[item] != 0)) {
                    rd = A.item_read ? (uint32_t)A.item_read[item] : (uint32_t)item;
                    pcj = A.pc[rd];
                    live = true;
                }
            }
            if (slot_valid && use > 0) mbar_wait(empty_s + slot * 8, (uint32_t)((use - 1) & 1));
            if (slot_valid) meta[slot * BVT_SLOT_TGT + sub] = make_int4(tslot, item, pcj, live ? 1 : 0);
            const unsigned lm = __ballot_sync(0xffffffffu, live);
            const int nlive = __popc((lm >> (c * 4)) & 0xfu);
            // a read's forward and reverse bitvectors are adjacent in memory (one copy per target), and the four reads
            // of a slot are often consecutive (unclustered stretches of the initial pass): then ONE 4-KB copy fills the slot
            const uint32_t rd0 = __shfl_sync(0xffffffffu, rd, lane & ~3);
            const bool seq = A.both && live && rd == rd0 + (uint32_t)sub;
            const unsigned sm = __ballot_sync(0xffffffffu, seq);
            const bool whole = ((sm >> (c * 4)) & 0xfu) == 0xfu;
            __syncwarp();
            if (slot_valid && sub == 0) {
                if (nlive) mbar_arrive_expect_tx(full_s + slot * 8, (uint32_t)nlive * per_tgt);
                else mbar_arrive(full_s + slot * 8);
            }
            __syncwarp();
            const uint32_t dst = ring_s + (uint32_t)slot * BVT_SLOT_BYTES + (uint32_t)sub * 1024u;
            if (whole) {
                if (sub == 0) bulk_g2s(dst, A.bv_f + (uint64_t)rd * A.bv_stride, 4096u, full_s + slot * 8);
            } else if (live) {
                if (A.both && A.bv_r == A.bv_f + 64) {
                    bulk_g2s(dst, A.bv_f + (uint64_t)rd * A.bv_stride, 1024u, full_s + slot * 8);
                } else {
                    bulk_g2s(dst, A.bv_f + (uint64_t)rd * A.bv_stride, 512u, full_s + slot * 8);
                    if (A.both) bulk_g2s(dst + 512u, A.bv_r + (uint64_t)rd * A.bv_stride, 512u, full_s + slot * 8);
                }
            }
        }
        return;
    }
    // ---------------- consumers: warp 2 + c takes slot (c, it & 1) of every iteration
    mbar_wait(seedbar_s, 0);
    const int c = warp - BVT_PROD;
    const int qt = lane >> 3, j = lane & 7;  // quarter-warp qt owns target qt of the slot; lane j holds bytes [64j, 64j+64)
    unsigned long long my_pairs = 0;
    for (int it = 0;; ++it) {
        const long long q = (long long)blockIdx.x + ((long long)it * BVT_CONS + c) * gridDim.x;
        if (q >= n_groups) break;
        const int slot = c * BVT_PROD + (it & 1);
        mbar_wait(full_s + slot * 8, (uint32_t)((it >> 1) & 1));
        const int4 m = meta[slot * BVT_SLOT_TGT + qt];
        const bool live = m.w != 0;
        const int tslot = m.x, item = m.y, pcj = m.z;
        const long long x = q * BVT_SLOT_TGT + qt;
        if (__any_sync(0xffffffffu, live)) {
            ulonglong2 f[4], r[4];
            const ulonglong2 *tp =
                reinterpret_cast<const ulonglong2 *>(ring + (size_t)slot * BVT_SLOT_BYTES + (size_t)qt * 1024 + (size_t)j * 64);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                f[w] = live ? tp[w] : make_ulonglong2(0, 0);
                r[w] = (live && A.both) ? tp[32 + w] : make_ulonglong2(0, 0);
            }
            for (int g = 0; g < ts; g += 8) {
                uint32_t mine = 0;
                const int lim = min(8, ts - g);
                for (int k = 0; k < lim; ++k) {
                    const ulonglong2 *sp = reinterpret_cast<const ulonglong2 *>(sseed + (size_t)(g + k) * 512 + (size_t)j * 64);
                    uint32_t cf = 0, cr = 0;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const ulonglong2 sv = sp[w];
                        cf += (uint32_t)(__popcll(sv.x & f[w].x) + __popcll(sv.y & f[w].y));
                        cr += (uint32_t)(__popcll(sv.x & r[w].x) + __popcll(sv.y & r[w].y));
                    }
                    uint32_t cc = cf | (cr << 16);
#pragma unroll
                    for (int sft = 4; sft > 0; sft >>= 1) cc += __shfl_xor_sync(0xffffffffu, cc, sft);
                    if (j == k) mine = cc;
                }
                const int s = g + j;
                bool valid = live && j < lim;
                if (valid && A.order_check) valid = sitem[s] < item;
                if (valid && A.item_seg) valid = A.item_seg[sitem[s]] == A.item_seg[item];
                const uint32_t cf = mine & 0xffffu, cr = mine >> 16;
                bool pf = false, pr = false;
                if (valid) {
                    const int mmax = max(spc[s], pcj);
                    const uint32_t cutv = A.cut[mmax];
                    pf = cf >= cutv;
                    pr = A.both && cr >= cutv;
                    if (pf && A.memo.known_failure((uint32_t)sitem[s], (uint32_t)item, 0)) pf = false;
                    if (pr && A.memo.known_failure((uint32_t)sitem[s], (uint32_t)item, 1)) pr = false;
                }
                my_pairs += (unsigned long long)__popc(__ballot_sync(0xffffffffu, valid));
                if (A.dense_common) {
                    if (live && j < lim) {
                        const size_t idx = (size_t)(s0 + s) * (size_t)n_t + (size_t)x;
                        A.dense_common[idx] = mine;
                        A.dense_pass[idx] = (uint8_t)((pf ? 1 : 0) | (pr ? 2 : 0));
                    }
                }
                if (A.tasks) {
                    const uint32_t mf = __ballot_sync(0xffffffffu, pf), mr = __ballot_sync(0xffffffffu, pr);
                    const int tot = __popc(mf) + __popc(mr);
                    if (tot) {
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(A.n_tasks, (unsigned long long)tot);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (base + tot > (unsigned long long)A.task_cap) {
                            if (lane == 0) atomicExch(A.ovf, 1);
                        } else {
                            const uint32_t below = (1u << lane) - 1u;
                            if (pf) A.tasks[base + __popc(mf & below)] = make_task((uint32_t)(s0 + s), 0, (uint32_t)tslot);
                            if (pr)
                                A.tasks[base + __popc(mf) + __popc(mr & below)] = make_task((uint32_t)(s0 + s), 1, (uint32_t)tslot);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_s + slot * 8);
    }
    if (lane == 0 && my_pairs && A.pair_counter) atomicAdd(A.pair_counter, my_pairs);
}

// ------------------------------------------------------------------------------------------------ K4: list join
__device__ __forceinline__ int lower_bound_u32(const uint32_t *p, int n, uint32_t v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (p[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Lane-local part of get_common_kmers (kmer.cpp:45-67): A positions [a0,a1) against all of B.
// EMIT=false: returns the number of cross pairs.  EMIT=true: also writes (posA<<32|posB) keys from out[0].
template <bool EMIT>
__device__ __forceinline__ long long join_range(const uint32_t *pa, const uint32_t *pb, int a0, int a1, int n2,
                                                const int32_t *posa, const int32_t *posb, uint64_t *out) {
    long long cnt = 0;
    if (a0 >= a1) return 0;
    int y = lower_bound_u32(pb, n2, pa[a0]);
    int ye = y;
    uint32_t hprev = 0;
    bool have = false;
    for (int x = a0; x < a1; ++x) {
        const uint32_t h = pa[x];
        if (!(have && h == hprev)) {
            y = ye;
            while (y < n2 && pb[y] < h) ++y;
            ye = y;
            while (ye < n2 && pb[ye] == h) ++ye;
            hprev = h;
            have = true;
        }
        if (EMIT) {
            const uint64_t hi = (uint64_t)(uint32_t)posa[x] << 32;
            for (int z = y; z < ye; ++z) out[cnt + (z - y)] = hi | (uint32_t)posb[z];
        }
        cnt += ye - y;
    }
    return cnt;
}

// warp per task; survivors of the exact bound  k*n_common/min_len >= t_s  are appended to `surv`
// as (task index | min(n_common, 2^32-1) << 32).
constexpr int JC_THREADS = 512;  // 16 warps per CTA, 2 CTAs per SM, 6.9 KB of staged hashes (list B) per warp
__global__ void __launch_bounds__(JC_THREADS) k_join_count(TaskView tv, const uint64_t *__restrict__ tasks,
                                                           const unsigned long long *n_tasks_p, ReadView R, double t_s,
                                                           int cap_w, uint64_t *surv, unsigned long long *n_surv,
                                                           int64_t surv_cap, int64_t *nmatch_out, int *ovf,
                                                           unsigned long long *stat_full) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *sw = (uint32_t *)sm_raw + (size_t)warp * cap_w;
    const unsigned long long n_tasks = *n_tasks_p;
    if (blockIdx.x == 0 && threadIdx.x == 0 && stat_full) atomicAdd(stat_full, n_tasks);
    const unsigned long long wstride = (unsigned long long)gridDim.x * (JC_THREADS / 32);
    for (unsigned long long ti = (unsigned long long)blockIdx.x * (JC_THREADS / 32) + warp; ti < n_tasks; ti += wstride) {
        uint32_t ar, br;
        int strand;
        tv.decode(tasks[ti], ar, br, strand);
        const int la = R.len[ar], lb = R.len[br];
        const int n1 = la - R.k, n2 = lb - R.k;
        const uint32_t *A = R.kh[0] + R.koff(ar);
        const uint32_t *B = (strand ? R.kh[1] : R.kh[0]) + R.koff(br);
        // Only list B is staged in shared memory: every lane walks a contiguous 1/32 of A (two or three cache lines,
        // read once, L1-resident) but probes all over B.  Half the staging traffic and twice the warps per SM of
        // staging both lists.
        const uint32_t *pa = A, *pb = B;
        if (n2 <= cap_w) {
            for (int i = lane; i < n2; i += 32) sw[i] = B[i];
            __syncwarp();
            pb = sw;
        }
        const int per = (n1 + 31) >> 5;
        const int a0 = min(n1, lane * per), a1 = min(n1, a0 + per);
        long long cnt = join_range<false>(pa, pb, a0, a1, n2, nullptr, nullptr, nullptr);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
        __syncwarp();
        if (lane == 0) {
            if (nmatch_out) nmatch_out[ti] = cnt;
            const double mn = (double)min(la, lb);
            const double bound = (double)((long long)R.k * cnt) / mn;
            if (!(bound >= t_s)) {
                uint32_t ai, bi;
                tv.items(tasks[ti], ai, bi);
                tv.memo.record_failure(ai, bi, strand);
            }
            if (bound >= t_s) {
                unsigned long long idx = atomicAdd(n_surv, 1ull);
                if ((long long)idx < surv_cap)
                    surv[idx] = (uint64_t)(uint32_t)ti | ((uint64_t)(cnt > 0xffffffffll ? 0xffffffffu : (uint32_t)cnt) << 32);
                else
                    atomicExch(ovf, 2);
            }
        }
    }
}

// utils.cpp:26-55 in the reference's exact operation order, no FMA contraction.
__device__ __forceinline__ double var_exact(const int32_t *d, int n) {
    if (n == 0) return 0.0;
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum = __dadd_rn(sum, (double)d[i]);
    const double dn = (double)n;
    const double m = __ddiv_rn(sum, dn);
    double ss = 0.0, comp = 0.0;
    for (int i = 0; i < n; ++i) {
        const double x = __dsub_rn((double)d[i], m);
        ss = __dadd_rn(ss, __dmul_rn(x, x));
        comp = __dadd_rn(comp, x);
    }
    // (ss - comp*comp/n) / (n-1): n-1 is computed in size_t, n==1 -> 0 -> x/0
    return __ddiv_rn(__dsub_rn(ss, __ddiv_rn(__dmul_rn(comp, comp), dn)), (double)(n - 1));
}

struct Sink {
    int mode;            // 0: none, 1: wave phase A (atomicMin acc[s*W+t] <- strand), 2: wave phase B (atomicMin best[t] <- 2s+strand)
    uint32_t *acc;
    int W;
    uint32_t *best;
    // test outputs, indexed by task (nullable)
    int32_t *bases, *n_dist;
    double *var;
    uint8_t *accept;
};

// warp per survivor: emit matches, sort by (posA,posB), LIS on posB, chain filter, variance, accept test.
// per-warp smem for cap C matches: keys u64[C] | prev i32[C] | tail i32[C+1] | tval i32[C+1]
constexpr int PH_THREADS = 256;
__host__ __device__ __forceinline__ size_t heavy_bytes(size_t n_pad, size_t n) {
    return 8 * n_pad + 4 * n + 8 * (n + 1) + 16;
}
__global__ void __launch_bounds__(PH_THREADS) k_pair_heavy(TaskView tv, const uint64_t *__restrict__ tasks,
                                                           const uint64_t *__restrict__ surv,
                                                           const unsigned long long *n_surv_p, int64_t surv_cap,
                                                           ReadView R, double t_s, double t_v, int cap_c,
                                                           unsigned char *scratch, unsigned long long *scratch_cur,
                                                           unsigned long long scratch_bytes, Sink S, int *ovf,
                                                           unsigned long long *stat_heavy, uint64_t *defer,
                                                           unsigned long long *n_defer, int64_t defer_cap) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wbase = sm_raw + (size_t)warp * heavy_bytes(cap_c, cap_c);
    unsigned long long n_surv = *n_surv_p;
    if ((long long)n_surv > surv_cap) n_surv = surv_cap;
    if (blockIdx.x == 0 && threadIdx.x == 0 && stat_heavy) atomicAdd(stat_heavy, n_surv);
    const unsigned long long wstride = (unsigned long long)gridDim.x * (PH_THREADS / 32);
    for (unsigned long long si = (unsigned long long)blockIdx.x * (PH_THREADS / 32) + warp; si < n_surv; si += wstride) {
        const uint64_t rec = surv[si];
        const uint32_t ti = (uint32_t)rec;
        const uint32_t n32 = (uint32_t)(rec >> 32);
        const uint64_t task = tasks[ti];
        uint32_t ar, br;
        int strand;
        tv.decode(task, ar, br, strand);
        const int la = R.len[ar], lb = R.len[br];
        const int n1 = la - R.k, n2 = lb - R.k;
        const uint32_t *A = R.kh[0] + R.koff(ar);
        const uint32_t *B = (strand ? R.kh[1] : R.kh[0]) + R.koff(br);
        const int32_t *PA = R.kp[0] + R.koff(ar);
        const int32_t *PB = (strand ? R.kp[1] : R.kp[0]) + R.koff(br);
        int bases = 0, nd = 0;
        double v = 0.0;
        bool fail = false;
        if (n32 >= 0x7fffffffu) {  // 2^31 or more matches (saturated at 2^32-1 by k_join_count): not representable here
            if (lane == 0) atomicExch(ovf, 4);
            continue;
        }
        const int n = (int)n32;
        int n_pad = 1;
        while (n_pad < n) n_pad <<= 1;
        uint64_t *keys = nullptr;
        if (!fail) {
            if (n_pad <= cap_c) {
                keys = (uint64_t *)wbase;
            } else {
                unsigned long long need = (heavy_bytes(n_pad, n) + 15) & ~15ull, at = 0;
                if (lane == 0) at = atomicAdd(scratch_cur, need);
                at = __shfl_sync(0xffffffffu, at, 0);
                if (at + need > scratch_bytes) fail = true;
                else keys = (uint64_t *)(scratch + at);
            }
        }
        if (fail) {
            // the match scratch is full: the survivor is deferred to the next launch (which starts with an empty
            // scratch); only when there is no room to remember it does the call fail
            if (lane == 0) {
                const unsigned long long d = defer ? atomicAdd(n_defer, 1ull) : ~0ull;
                if (defer && (long long)d < defer_cap) defer[d] = rec;
                else atomicExch(ovf, 3);
            }
            continue;
        }
        int32_t *prev = (int32_t *)(keys + n_pad);
        int32_t *tail = prev + n;
        int32_t *tval = tail + (n + 1);
        if (n > 0) {
            // emit
            const int per = (n1 + 31) >> 5;
            const int a0 = min(n1, lane * per), a1 = min(n1, a0 + per);
            long long cnt = join_range<false>(A, B, a0, a1, n2, nullptr, nullptr, nullptr);
            long long incl = cnt;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                long long o = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= s) incl += o;
            }
            join_range<true>(A, B, a0, a1, n2, PA, PB, keys + (incl - cnt));
            for (int i = n + lane; i < n_pad; i += 32) keys[i] = ~0ull;
            // warp bitonic sort
            for (int size = 2; size <= n_pad; size <<= 1)
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    __syncwarp();
                    for (int t = lane; t < (n_pad >> 1); t += 32) {
                        int i = 2 * t - (t & (stride - 1));
                        int j = i + stride;
                        uint64_t a = keys[i], b = keys[j];
                        bool asc = (i & size) == 0;
                        if ((a > b) == asc) {
                            keys[i] = b;
                            keys[j] = a;
                        }
                    }
                }
            __syncwarp();
            if (lane == 0) {
                // similarity.cpp:10-31: patience LIS, strict on .second
                int l = 0;
                tail[0] = 0;
                for (int i = 0; i < n; ++i) {
                    const int vi = (int)(uint32_t)keys[i];
                    int lo;
                    if (l > 0 && tval[l] < vi) lo = l + 1;  // all tails smaller (tails strictly increase)
                    else {
                        lo = 1;
                        int hi = l;
                        while (lo <= hi) {
                            int mid = (lo + hi + 1) >> 1;
                            if (tval[mid] < vi) lo = mid + 1;
                            else hi = mid - 1;
                        }
                    }
                    prev[i] = tail[lo - 1];
                    tail[lo] = i;
                    tval[lo] = vi;
                    if (lo > l) l = lo;
                }
                // similarity.cpp:36-44: backtrack; LIS indices overwrite tail[0..l)
                int at = tail[l];
                for (int i = l - 1; i >= 0; --i) {
                    int nx = prev[at];
                    tail[i] = at;
                    at = nx;
                }
                // similarity.cpp:52-91: chain filter, covered bases, gap differences (into prev[])
                const int k = R.k;
                bases = k;
                uint64_t lastk = keys[tail[0]];
                int sprev = (int)(uint32_t)lastk;
                for (int i = 1; i < l; ++i) {
                    const uint64_t cur = keys[tail[i]];
                    const int cf = (int)(cur >> 32), cs = (int)(uint32_t)cur;
                    const int df = cf - (int)(lastk >> 32), ds = cs - (int)(uint32_t)lastk;
                    if ((df < k && ds < k) || (df >= k && ds >= k)) {
                        bases += k;
                        const int ex = k - (cs - sprev);
                        if (ex > 0) bases -= ex;
                        prev[nd++] = ds - df;
                        lastk = cur;
                    }
                    sprev = cs;
                }
                v = var_exact(prev, nd);
            }
        }
        if (lane == 0) {
            const double mn = (double)min(la, lb);
            const bool ok = (__ddiv_rn((double)bases, mn) >= t_s) && (v < t_v);
            if (S.bases) {
                S.bases[ti] = bases;
                S.n_dist[ti] = nd;
                S.var[ti] = v;
                S.accept[ti] = ok ? 1 : 0;
            }
            if (!ok) {
                uint32_t ai, bi;
                tv.items(task, ai, bi);
                tv.memo.record_failure(ai, bi, strand);
            }
            if (ok) {
                const uint32_t t = (uint32_t)task;
                const uint32_t s = (uint32_t)(task >> 33);
                if (S.mode == 1) atomicMin(&S.acc[(size_t)s * S.W + t], (uint32_t)strand);
                else if (S.mode == 2) {  // best[] is indexed by item (t indexes the target list when there is one)
                    uint32_t ai, bi;
                    tv.items(task, ai, bi);
                    atomicMin(&S.best[bi], s * 2 + (uint32_t)strand);
                }
            }
        }
        __syncwarp();
    }
}

__global__ void k_fold_status(const int *flags, uint32_t *status) { *status = flags[1] ? 0u : 0xffffffffu; }
// sharded extraction: "some rank met a base outside ACGTU" travels as a min-reduced word (0 = error)
__global__ void k_fold_input_flag(const int *flags, uint32_t *status) { *status = flags[0] ? 0u : 0xffffffffu; }
__global__ void k_unfold_input_flag(int *flags, const uint32_t *status) { if (*status == 0u) flags[0] = 1; }

__global__ void k_fill_u32(uint32_t *p, uint32_t v, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace rtl
