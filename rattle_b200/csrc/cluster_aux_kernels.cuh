// Small kernels of hot path A that need nothing but standard CUDA (warp votes, shared-memory atomics, block barriers):
//   K1  read ingest: 2-bit packing of the uploaded read set (kmer.hpp:25-31) and the visitation order of sort_read_set
//       (fasta.cpp:458-464) as a bitonic sort of unique (length, index) keys;
//   the bookkeeping of the greedy waves (cluster.cpp:124-166,171-245 batched): candidate selection, resolution of the
//       candidate x candidate decision matrix, seed selection per segment (batched clustering), owner assignment.
// Kept apart from cluster_kernels.cuh (which needs sm_100a PTX: mbarrier, bulk copies) so that tests/native/cuda_emu.h can
// run THIS source on the CPU (tests/test_cluster_aux_emulated_cpu.py).
#pragma once
#ifndef CUDA_EMU
#include <cuda_runtime.h>
#define RTL_DYNAMIC_SHARED(type, name) extern __shared__ type name[]
#else
#define RTL_DYNAMIC_SHARED(type, name) type *name = reinterpret_cast<type *>(emu::g_smem)
#endif
#include <stdint.h>

namespace rtl {

// A/C/T(U)/G -> 0/1/2/3 (kmer.hpp:25-31): bits 1..2 of the ASCII code give exactly that order.
__device__ __forceinline__ int base_code(uint8_t c) {
    bool ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T') | (c == 'U');
    return ok ? ((c >> 1) & 3) : -1;
}

// ------------------------------------------------------------------------------------------------ K1: 2-bit packing
// The read set on the device is 2-bit packed (kmer.hpp:25-31 codes, 16 bases per 32-bit word, base i of a word in bits
// 2i..2i+1); every read starts on a word: read r's words start at pk_start(off, r).  The ASCII copy is only staging for this
// kernel; k-mer extraction reads a quarter of the bytes, coalesced.  A base outside A,C,G,T,U raises the input flag.
__host__ __device__ __forceinline__ uint64_t pk_start(const uint64_t *off, uint32_t r) { return (off[r] >> 4) + r; }
__global__ void __launch_bounds__(256) k_pack_bases(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ off, uint32_t n,
                                                    uint32_t *__restrict__ pk, int *err) {
    // one warp per read at a time; a lane packs one word (16 bases read as one 16-byte load when aligned)
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (uint32_t r = warp; r < n; r += n_warps) {
        const uint64_t o = off[r];
        const int len = (int)(off[r + 1] - o);
        const uint64_t w0 = pk_start(off, r);
        for (int w = lane; w * 16 < len; w += 32) {
            uint32_t word = 0;
            const int m = min(16, len - w * 16);
            for (int i = 0; i < m; ++i) {
                const int c = base_code(bases[o + (uint64_t)w * 16 + i]);
                if (c < 0) bad = true;
                word |= (uint32_t)(c & 3) << (2 * i);
            }
            pk[w0 + w] = word;
        }
    }
    if (bad) atomicExch(err, 2);
}
__device__ __forceinline__ int pk_code(const uint32_t *__restrict__ pk, uint64_t w0, int p) {
    return (int)((pk[w0 + (uint32_t)(p >> 4)] >> (2 * (p & 15))) & 3u);
}

// K1, visitation order: sort_read_set (fasta.cpp:458-464) is a stable sort by length, longest first.  The keys
// (2^32-1 - length) << 32 | index are unique, so a plain bitonic sort of them IS that stable order.
__global__ void k_sort_keys_init(const uint64_t *__restrict__ off, uint32_t n, uint32_t n_pad, uint64_t *keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    keys[i] = i < n ? (((uint64_t)(0xffffffffu - (uint32_t)(off[i + 1] - off[i])) << 32) | i) : ~0ull;
}
__global__ void k_bitonic_step(uint64_t *keys, uint32_t n_pad, uint32_t size, uint32_t stride) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n_pad >> 1)) return;
    const uint32_t i = 2 * t - (t & (stride - 1)), j = i + stride;
    const uint64_t a = keys[i], b = keys[j];
    const bool asc = (i & size) == 0;
    if ((a > b) == asc) {
        keys[i] = b;
        keys[j] = a;
    }
}
// all steps of the network with stride < 1024 of one merge stage, inside shared memory (one CTA per 2048 keys)
__global__ void __launch_bounds__(1024) k_bitonic_local(uint64_t *keys, uint32_t size, uint32_t first_stride) {
    __shared__ uint64_t s[2048];
    const uint32_t base = blockIdx.x * 2048u, tid = threadIdx.x;
    s[tid] = keys[base + tid];
    s[tid + 1024] = keys[base + tid + 1024];
    for (uint32_t stride = first_stride; stride > 0; stride >>= 1) {
        __syncthreads();
        const uint32_t i = 2 * tid - (tid & (stride - 1)), j = i + stride;
        const uint64_t a = s[i], b = s[j];
        const bool asc = ((base + i) & size) == 0;
        if ((a > b) == asc) {
            s[i] = b;
            s[j] = a;
        }
    }
    __syncthreads();
    keys[base + tid] = s[tid];
    keys[base + tid + 1024] = s[tid + 1024];
}
__global__ void k_sort_keys_perm(const uint64_t *__restrict__ keys, uint32_t n, uint32_t *perm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = (uint32_t)keys[i];
}

// ------------------------------------------------------------------------------------------------ greedy wave bookkeeping
// wave state (device): [0]=cursor, [1]=n_cand, [2]=n_seeds, [3]=done flag
// k_select: one CTA scans items from the cursor and takes the first W untaken ones as this wave's candidates.
__global__ void k_select(uint8_t *taken, int M, int W, int32_t *cand, int32_t *wave) {
    __shared__ int s_count;
    const int tid = threadIdx.x, nt = blockDim.x;
    int cursor = wave[0];
    if (tid == 0) s_count = 0;
    __syncthreads();
    int count = 0;
    while (cursor < M && count < W) {
        const int j = cursor + tid;
        const bool un = j < M && !taken[j];
        // block-level ordered compaction
        const unsigned m = __ballot_sync(0xffffffffu, un);
        __shared__ int wcnt[32];
        const int lane = tid & 31, w = tid >> 5;
        if (lane == 0) wcnt[w] = __popc(m);
        __syncthreads();
        if (tid == 0) {
            int run = s_count;
            for (int i = 0; i < (nt >> 5); ++i) {
                int c = wcnt[i];
                wcnt[i] = run;
                run += c;
            }
            s_count = run;
        }
        __syncthreads();
        if (un) {
            int slot = wcnt[w] + __popc(m & ((1u << lane) - 1u));
            if (slot < W) cand[slot] = j;
        }
        __syncthreads();
        count = s_count;
        if (count >= W) {
            // cursor must stop right after the W-th candidate: find it
            break;
        }
        cursor += nt;
    }
    __syncthreads();
    if (tid == 0) {
        int nc = min(s_count, W);
        wave[1] = nc;
        wave[2] = 0;
        int newcur = nc ? cand[nc - 1] + 1 : M;
        if (nc < W) newcur = M;  // scanned to the end
        wave[0] = newcur;
        wave[3] = (nc == 0) ? 1 : 0;
    }
}

// mark candidates taken (after k_select so that the scan above reads a consistent state)
// The same resolution for W <= 1024 by a whole CTA: the W x W decision matrix is read once, coalesced, by all warps and
// turned into per-candidate bit masks "earlier candidates that match me" in shared memory (the matrix is sparse: one
// shared atomicOr per match); warp 0 then walks the candidates in order with the mask of seeds so far in shared memory —
// one AND + ballot per candidate instead of a strided global-memory scan (0.7 ms -> a few tens of microseconds per wave).
// dynamic shared memory: W * (W / 32) + W / 32 words.
__global__ void __launch_bounds__(1024) k_resolve_cta(const uint32_t *__restrict__ acc, int W, const int32_t *__restrict__ cand,
                                                     int32_t *wave, int32_t *seed_item, uint8_t *is_seed, int32_t *owner,
                                                     uint8_t *owner_rev) {
    RTL_DYNAMIC_SHARED(uint32_t, sm_res);
    const int WW = W >> 5;               // mask words per candidate (W is a multiple of 32 here)
    uint32_t *hit = sm_res;              // [W][WW]: bit a of hit[b] = candidate a < b matches b
    uint32_t *seeds = sm_res + (size_t)W * WW;  // [WW]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    const int nc = wave[1];
    for (int i = tid; i < W * WW + WW; i += blockDim.x) sm_res[i] = 0u;
    __syncthreads();
    for (int a = warp; a < nc; a += n_warps)
        for (int b = (a + 1 - ((a + 1) & 31)) + lane; b < nc; b += 32)  // b > a, row a read coalesced
            if (b > a && acc[(size_t)a * W + b] != 0xffffffffu) atomicOr(&hit[(size_t)b * WW + (a >> 5)], 1u << (a & 31));
    __syncthreads();
    if (warp != 0) return;
    int ns = 0;
    for (int b = 0; b < nc; ++b) {
        uint32_t m = 0u;
        if (lane < WW) m = hit[(size_t)b * WW + lane] & seeds[lane];
        const unsigned any = __ballot_sync(0xffffffffu, m != 0u);
        int found = -1;
        if (any) {
            const int wl = __ffs(any) - 1;
            const uint32_t mw = __shfl_sync(0xffffffffu, m, wl);
            found = wl * 32 + __ffs(mw) - 1;
        }
        if (lane == 0) {
            if (found >= 0) {
                is_seed[b] = 0;
                owner[cand[b]] = cand[found];
                owner_rev[cand[b]] = (uint8_t)acc[(size_t)found * W + b];
            } else {
                is_seed[b] = 1;
                seed_item[ns] = cand[b];
                seeds[b >> 5] |= 1u << (b & 31);
            }
        }
        if (found < 0) ++ns;
        __syncwarp();
    }
    if (lane == 0) wave[2] = ns;
}
// Batched clustering: the first untaken item of a segment has no earlier untaken item it could join, so it IS a seed.  One
// thread per segment of the window picks it (seg_cur remembers where the segment's search stands), marks it taken and
// appends it to the wave's seeds; their order does not matter (an item only ever matches the seed of its own segment).
__global__ void k_select_seg(uint8_t *taken, const int32_t *__restrict__ seg_first, int32_t *seg_cur, int n_seg,
                             int32_t *seed_item, int32_t *wave, int32_t *owner, uint8_t *owner_rev) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    int i = seg_cur[s];
    const int end = seg_first[s + 1];
    while (i < end && taken[i]) ++i;
    if (i < end) {
        taken[i] = 1;
        owner[i] = i;
        owner_rev[i] = 0;
        seed_item[atomicAdd(&wave[2], 1)] = i;
        ++i;
    }
    seg_cur[s] = i;
}
__global__ void k_mark_cand(uint8_t *taken, const int32_t *cand, const int32_t *wave, int32_t *owner, uint8_t *owner_rev) {
    const int nc = wave[1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) {
        taken[cand[i]] = 1;
        owner[cand[i]] = cand[i];
        owner_rev[cand[i]] = 0;
    }
}

// k_resolve (one warp): greedy inside the wave (cluster.cpp:125-166 restricted to the candidates).
// candidate b joins the smallest earlier candidate a that is a seed and matches it; otherwise it is a seed.
__global__ void k_resolve(const uint32_t *acc, int W, const int32_t *cand, int32_t *wave, int32_t *seed_item,
                          uint8_t *is_seed /*[W]*/, int32_t *owner, uint8_t *owner_rev) {
    const int lane = threadIdx.x;
    const int nc = wave[1];
    int ns = 0;
    for (int b = 0; b < nc; ++b) {
        int found = -1;
        for (int a0 = 0; a0 < b && found < 0; a0 += 32) {
            const int a = a0 + lane;
            bool hit = a < b && is_seed[a] && acc[(size_t)a * W + b] != 0xffffffffu;
            unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) found = a0 + __ffs(m) - 1;
        }
        __syncwarp();
        if (lane == 0) {
            if (found >= 0) {
                is_seed[b] = 0;
                owner[cand[b]] = cand[found];
                owner_rev[cand[b]] = (uint8_t)acc[(size_t)found * W + b];
            } else {
                is_seed[b] = 1;
                seed_item[ns] = cand[b];
            }
        }
        if (found < 0) ++ns;
        __syncwarp();
    }
    if (lane == 0) wave[2] = ns;
}

// k_apply: targets that found a seed in phase B join it.
__global__ void k_apply(uint32_t *best, int t0, int M, const int32_t *seed_item, uint8_t *taken, int32_t *owner,
                        uint8_t *owner_rev) {
    for (int j = t0 + blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const uint32_t b = best[j];
        if (b != 0xffffffffu) {
            owner[j] = seed_item[b >> 1];
            owner_rev[j] = (uint8_t)(b & 1u);
            taken[j] = 1;
            best[j] = 0xffffffffu;
        }
    }
}

// multi-GPU: a rank-local overflow must stop EVERY rank (the others would wait in the next exchange for ever): the flag
// travels with the decision array as one more uint32 (0 = some rank overflowed; smaller wins in the min-reduction)

}  // namespace rtl
