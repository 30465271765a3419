// sm_100a kernel of hot path B, second generation: the same sequence-to-graph local alignment as poa_kernels.cuh
// (spoa kSW affine, correct.cpp:400,431,526 -> simd_alignment_engine.cpp:1086-1188 forward, :1210-1458 traceback)
// laid out for the Blackwell integer pipes:
//
//   * scores are int16, two DP cells per 32-bit register, computed with the DPX instructions
//     (VIADD.16x2 / VIMNMX.S16x2 / VIADDMNMX.S16x2 / VIMNMX3.S16x2) — half the issue slots of the int32 kernel;
//   * one WARP owns a strip of 256 query columns (8 consecutive columns per lane, register k holds columns
//     k and k+4 so that the "column - 1" operand of the diagonal is a register rename plus one PRMT);
//   * the strips of one alignment run as a wavefront inside the CTA: warp w starts row r as soon as warp w-1
//     hands over the row's gap carry E and its last H through a shared-memory mailbox.  There is no block
//     barrier inside the DP; the row-wise recurrence E[j] = max(H[j-1]+g, E[j-1]+e) is a max-plus scan
//     (3 packed steps inside the lane, 5 shuffle steps across the warp, one scalar across strips);
//   * the H/F rows of the last PS_K graph rows live in a per-warp shared-memory ring: in spoa's topological order
//     99 % of all predecessors are at most 6 rows back (aligned siblings and short branches sit between a node and
//     its predecessor), so predecessor rows are two LDS.128 away.  Only rows that some later row needs from
//     further back ("spilled" rows, marked by the host) are also written to HBM — the DP's DRAM traffic is the
//     2-byte traceback code per cell and little else;
//   * the traceback decisions are stored as one 16-bit code per cell (as in the first kernel), built from
//     0/1 "not equal" flags (XOR + unsigned min) that the FMA pipe packs with IMADs; the traceback itself is a
//     second kernel with one warp per alignment (k_poa_strip_traceback).
//
// Code of cell (r, j):
//   bit 0   H != 0                      (0 -> traceback stops here)
//   bit 1   H != Hdiag                  (0 -> diagonal move to predecessor `dp`)
//   bit 2   H != F                      (0 -> vertical move to predecessor `fp`; extend_up iff F-from-F attains F)
//   bit 3   E[j+1] != E[j]+e            (0 -> the E of the NEXT column is a gap extension)
//   bits 4-5  0: F[fp]+e > H[fp]+g   1: equal   2: H[fp]+g > F[fp]+e        (for the winning F predecessor)
//   bits 6-10  fp  first predecessor (in_edges order) attaining F
//   bits 11-15 dp  first predecessor attaining Hdiag
// Eligibility (checked by the host, otherwise the int32 kernel runs): scores 5/-4/-8/-6 that fit int16, in-degree
// <= 32, letters within {A,C,G,T,U}, fewer than 65535 spilled rows.
#pragma once
#ifndef CUDA_EMU  // tests/native/cuda_emu.h runs this source on the host (SIMT emulation, CPU tests)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include <type_traits>

namespace rtl {

struct PoaSJob {
    uint64_t hf_off;     // u32 words into the arena: (n_spill+1) x n_strips x 256 words (H then F per strip) of the
                         // spilled rows (slot 0 = virtual start row), then halo[(n_spill+1) x n_strips] (H of the
                         // column left of the strip), then passb[n+1] (hand-over between passes), then 8 words per
                         // thread: H of the lane's best row and its best cell over all passes (value, row, column)
    uint64_t code_off;   // u32 words into the arena: ceil(n/8) x n_strips x 1024 words (tiles of 8 rows, ps_code_words)
    uint32_t q_off;      // bytes into the query buffer: n_strips*256 letter codes (0..4, pad 255)
    uint32_t row_off;    // into rec: n+1 entries, entry r describes row r (1-based)
    uint32_t pred_base;  // into preds: predecessor words of rows with more than 3 predecessors
    uint32_t aln_off;    // pairs, into the alignment output
    uint32_t spill_off;  // into spill_rows: row of every spill slot (slot 0 -> row 0)
    int32_t L, n, n_strips, n_spill, pad;
    uint64_t order_off;  // int32 words into the graph-mirror pool: rank -> node id of this graph (poa_devgraph.cuh), or
                         // ~0: the traceback reports rows and the host maps them
};
// Row record (16 B): x = letter code | n_pred << 8 | own spill slot << 16 (0 = row is not spilled),
// y, z, w = predecessor words (n_pred <= 3) or y, z = the first two and w = offset of the full list in preds.
// Predecessor word: d (1..K) = the row d ranks back (in the ring), or 0x80000000 | spill slot.

constexpr int PS_STRIP = 256;  // columns per warp
constexpr int PS_D = 16;       // mailbox ring depth (rows a strip may run ahead of its right neighbour)
constexpr int PS_MAXW = 8;     // warps per CTA (strips per pass)
constexpr int PS_NLET = 5;     // A C G T U
constexpr int PS_K = 6;        // graph rows kept in the shared-memory ring (upper bound; the launch picks K <= PS_K)
constexpr int PS_NEGF = -1000; // F of the virtual start row: below every H+g, and NEGF+e stays far from int16 limits
constexpr uint32_t PS_FAR = 0x80000000u;

__host__ __device__ __forceinline__ size_t ps_hf_words(int n, int n_strips, int n_spill) {
    return (size_t)(n_spill + 1) * n_strips * 256 + (size_t)(n_spill + 1) * n_strips + (size_t)((n + 1 + 3) & ~3) +
           (size_t)PS_MAXW * 32 * 8;
}
// traceback codes are stored in tiles of 8 graph rows x 8 columns (one lane's columns): 128 contiguous bytes, so that the
// traceback fetches the 64 cells around its position with ONE coalesced load and walks inside the tile with shuffles
__host__ __device__ __forceinline__ size_t ps_code_words(int n, int n_strips) {
    return (size_t)((n + 7) / 8) * n_strips * 1024;
}
__host__ __device__ __forceinline__ size_t ps_smem_bytes(int n_warps, int K) {
    return (size_t)n_warps * (PS_NLET * 32 * 16 + K * 64 * 16 + 32);
}

__device__ __forceinline__ int ps_lo(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int ps_hi(uint32_t v) { return ((int)v) >> 16; }
__host__ __device__ constexpr uint32_t ps_pk(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__host__ __device__ constexpr uint32_t ps_pk2(int v) { return ps_pk(v, v); }

#ifndef CUDA_EMU
using ps_saddr = uint32_t;  // 32-bit shared-space address
#define PS_DYNAMIC_SHARED(type, name) extern __shared__ type name[]
// volatile shared-memory accesses by 32-bit shared-space address (keeps the mailbox addresses in one register each)
__device__ __forceinline__ unsigned long long ps_lds64(uint32_t addr) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void ps_sts64(uint32_t addr, unsigned long long v) {
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ int ps_lds32(uint32_t addr) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void ps_sts32(uint32_t addr, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// H (4 words), F (4 words, 512 B further) and the halo word of one predecessor row from the shared-memory ring ...
__device__ __forceinline__ void ps_fetch_ring(uint32_t s_row, uint32_t s_halo, uint32_t (&cH)[4], uint32_t (&cF)[4], int &hl) {
    asm volatile(
        "ld.shared.v4.u32 {%0,%1,%2,%3}, [%9];\n\t"
        "ld.shared.v4.u32 {%4,%5,%6,%7}, [%9+512];\n\t"
        "ld.shared.s32 %8, [%10];"
        : "=r"(cH[0]), "=r"(cH[1]), "=r"(cH[2]), "=r"(cH[3]), "=r"(cF[0]), "=r"(cF[1]), "=r"(cF[2]), "=r"(cF[3]), "=r"(hl)
        : "r"(s_row), "r"(s_halo)
        : "memory");
}
// ... overwritten in place, for a spilled row, from HBM ("+r": the same registers, so the rare branch around this
// block needs no register moves at its join)
__device__ __forceinline__ void ps_fetch_spilled(const uint32_t *g_row, const uint32_t *g_halo, uint32_t (&cH)[4],
                                                 uint32_t (&cF)[4], int &hl) {
    asm volatile(
        "ld.global.v4.u32 {%0,%1,%2,%3}, [%9];\n\t"
        "ld.global.v4.u32 {%4,%5,%6,%7}, [%9+512];\n\t"
        "ld.global.s32 %8, [%10];"
        : "+r"(cH[0]), "+r"(cH[1]), "+r"(cH[2]), "+r"(cH[3]), "+r"(cF[0]), "+r"(cF[1]), "+r"(cF[2]), "+r"(cF[3]), "+r"(hl)
        : "l"(g_row), "l"(g_halo)
        : "memory");
}
__device__ __forceinline__ void ps_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ps_lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ void ps_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#else  // CUDA_EMU: "shared-space addresses" are plain pointers
inline void ps_prefetch_l2(const void *) {}
using ps_saddr = uintptr_t;
#define PS_DYNAMIC_SHARED(type, name) type *name = reinterpret_cast<type *>(emu::g_smem)
inline unsigned long long ps_lds64(ps_saddr a) { return __atomic_load_n(reinterpret_cast<unsigned long long *>(a), __ATOMIC_SEQ_CST); }
inline void ps_sts64(ps_saddr a, unsigned long long v) { __atomic_store_n(reinterpret_cast<unsigned long long *>(a), v, __ATOMIC_SEQ_CST); }
inline int ps_lds32(ps_saddr a) { return __atomic_load_n(reinterpret_cast<int *>(a), __ATOMIC_SEQ_CST); }
inline void ps_sts32(ps_saddr a, int v) { __atomic_store_n(reinterpret_cast<int *>(a), v, __ATOMIC_SEQ_CST); }
inline void ps_fetch_ring(ps_saddr s_row, ps_saddr s_halo, uint32_t (&cH)[4], uint32_t (&cF)[4], int &hl) {
    memcpy(cH, reinterpret_cast<const void *>(s_row), 16);
    memcpy(cF, reinterpret_cast<const void *>(s_row + 512), 16);
    hl = ps_lds32(s_halo);
}
inline void ps_fetch_spilled(const uint32_t *g_row, const uint32_t *g_halo, uint32_t (&cH)[4], uint32_t (&cF)[4], int &hl) {
    memcpy(cH, g_row, 16);
    memcpy(cF, g_row + 128, 16);
    hl = (int)*g_halo;
}
inline void ps_sts128(ps_saddr a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    const uint32_t v[4] = {x, y, z, w};
    memcpy(reinterpret_cast<void *>(a), v, 16);
}
inline uint4 ps_lds128(ps_saddr a) {
    uint4 v;
    memcpy(&v, reinterpret_cast<const void *>(a), 16);
    return v;
}
#endif

// One predecessor row (cH, cF = its H and F in my 8 columns, hl = its H left of the strip, lane 0 only) folded into
// the running Hdiag / F of the current row.  FIRST: plain assignment, predecessor index 0.  Otherwise the strictly
// better candidate replaces the running value and its index (first arg-max = the reference's in_edges order).
template <int SG, int SE, bool FIRST>
__device__ __forceinline__ void ps_fold_pred(const uint32_t (&cH)[4], const uint32_t (&cF)[4], int hl, bool lane0, int p,
                                             const uint32_t (&sc)[4], uint32_t (&Hd)[4], uint32_t (&Fv)[4],
                                             uint32_t (&fpk)[4], uint32_t (&dpk)[4]) {
    constexpr uint32_t g2 = ps_pk2(SG), e2 = ps_pk2(SE), one2 = 0x00010001u, two2 = 0x00020002u;
    uint32_t left = __shfl_up_sync(0xffffffffu, cH[3], 1);
    if (lane0) left = (uint32_t)hl << 16;
    uint32_t hprev[4];
    hprev[0] = __byte_perm(left, cH[3], 0x5432);  // (column -1, column 3)
    hprev[1] = cH[0];
    hprev[2] = cH[1];
    hprev[3] = cH[2];
    if (FIRST) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            Hd[k] = __vadd2(hprev[k], sc[k]);
            const uint32_t fh = __vadd2(cH[k], g2), fe = __vadd2(cF[k], e2);
            Fv[k] = __vmaxs2(fh, fe);
            // clamp(fh - fe, -1, 1) + 1 without a packed subtract: fh + ~fe = fh - fe - 1
            fpk[k] = __viaddmin_s16x2_relu(__vadd2(fh, ~fe), two2, two2);
        }
    } else {
        const uint32_t pp = ((uint32_t)p << 16) | (uint32_t)p, pp4 = pp << 2;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t nh = __viaddmax_s16x2(hprev[k], sc[k], Hd[k]);
            const uint32_t mh = __vminu2(nh ^ Hd[k], one2) * 0xffffu;  // strictly better -> 0xffff
            dpk[k] = (dpk[k] & ~mh) | (pp & mh);
            Hd[k] = nh;
            const uint32_t fh = __vadd2(cH[k], g2), fe = __vadd2(cF[k], e2);
            const uint32_t sf = __viaddmin_s16x2_relu(__vadd2(fh, ~fe), two2, two2);
            const uint32_t nf = __vimax3_s16x2(fh, fe, Fv[k]);
            const uint32_t mf = __vminu2(nf ^ Fv[k], one2) * 0xffffu;
            fpk[k] = (fpk[k] & ~mf) | ((sf + pp4) & mf);
            Fv[k] = nf;
        }
    }
}

// The DP of ONE alignment by the whole CTA (all threads call it; it starts and ends with block barriers).  Returns the
// best cell (value, row, column) through *best_cell (shared memory), valid for every thread after the call.
template <int SM, int SN, int SG, int SE>
__device__ __forceinline__ void ps_align_job(const PoaSJob &J, const uint8_t *__restrict__ qcodes, const uint4 *__restrict__ rec,
                                             const int32_t *__restrict__ preds, uint32_t *arena, int K, uint4 *s_dyn,
                                             int4 *best_cell) {
    // dynamic: [warp][letter][lane] packed match/mismatch scores of the warp's strip, then [warp][PS_K][H 32 | F 32]
    // ring of recent rows, then [warp][8] ring of H left of the strip
    __shared__ unsigned long long s_mb[PS_MAXW][PS_D];
    __shared__ int s_done[PS_MAXW];
    __shared__ int s_best[PS_MAXW][3];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NW = blockDim.x >> 5;
    const bool lane0 = lane == 0;
    constexpr uint32_t g2 = ps_pk2(SG), e2 = ps_pk2(SE);
    constexpr uint32_t one2 = 0x00010001u;
    constexpr int e8 = 8 * SE;
    // 32-bit shared-space addresses (one register each, no generic-address arithmetic in the row loop)
    const ps_saddr dyn_s = (ps_saddr)__cvta_generic_to_shared(s_dyn);
    ps_saddr prof_s = dyn_s + (uint32_t)((wid * PS_NLET * 32 + lane) * 16);             // + letter*512
    ps_saddr ring_s = dyn_s + (uint32_t)((NW * PS_NLET * 32 + wid * K * 64 + lane) * 16);  // + idx*1024 (F +512)
    ps_saddr hring_s = dyn_s + (uint32_t)(NW * (PS_NLET * 32 + K * 64) * 16 + wid * 32);   // + idx*4
#ifndef CUDA_EMU
    // keep the three addresses in registers: left to itself the compiler re-derives them from %tid / %ntid / K in every
    // row and for every predecessor (some thirty instructions per row)
    asm volatile("" : "+r"(prof_s), "+r"(ring_s), "+r"(hring_s));
#endif
    {
        const int n = J.n, nst = J.n_strips;
        uint32_t *hf = arena + J.hf_off;  // word offsets below are relative to hf (a job's region is far below 16 GB)
        const uint32_t halo_o = (uint32_t)(J.n_spill + 1) * (uint32_t)nst * 256u;
        const uint32_t passb_o = (halo_o + (uint32_t)(J.n_spill + 1) * (uint32_t)nst + 3u) & ~3u;
        uint32_t *const mine = hf + ((passb_o + (uint32_t)((n + 1 + 3) & ~3)) + (uint32_t)tid * 8u);  // [0..3] best-row H, [4..6] best cell
        uint32_t *cd = arena + J.code_off;
        const uint8_t *q = qcodes + J.q_off;
        const uint4 *recs = rec + J.row_off;
        const uint32_t pred_base = J.pred_base;
        const int n_pass = (nst + NW - 1) / NW;
        const uint32_t hf_stride = (uint32_t)nst * 256u;
        const uint32_t cd_tile_step = (uint32_t)nst * 1024u - 28u;  // from the last row of a tile to the first of the next

        *reinterpret_cast<uint4 *>(mine + 4) = make_uint4(0u, 0u, 0u, 0u);

        for (int pass = 0; pass < n_pass; ++pass) {
            for (int i = tid; i < PS_MAXW * PS_D; i += blockDim.x) (&s_mb[0][0])[i] = ~0ull;
            if (tid < PS_MAXW) s_done[tid] = 0;
            __syncthreads();  // also orders passb of the previous pass before its readers
            // Strip position inside the pass: the LEFTMOST strip goes to the HIGHEST warp id.  The issue arbiter
            // prefers high warp ids, and a strip can only wait for the strip on its left: with this order the
            // producers run ahead (up to PS_D rows) instead of being starved by consumers that spin on them.
            const int pos = NW - 1 - wid;
            const int t = pass * NW + pos;
            if (t < nst) {
                const int j0 = t * PS_STRIP + lane * 8;  // 0-based index of my first column (column j0+1)
                {
                    const uint2 q8 = *reinterpret_cast<const uint2 *>(q + j0);
#pragma unroll
                    for (int c = 0; c < PS_NLET; ++c) {
                        uint4 v;
                        v.x = ps_pk(((q8.x) & 0xff) == (uint32_t)c ? SM : SN, ((q8.y) & 0xff) == (uint32_t)c ? SM : SN);
                        v.y = ps_pk(((q8.x >> 8) & 0xff) == (uint32_t)c ? SM : SN, ((q8.y >> 8) & 0xff) == (uint32_t)c ? SM : SN);
                        v.z = ps_pk(((q8.x >> 16) & 0xff) == (uint32_t)c ? SM : SN, ((q8.y >> 16) & 0xff) == (uint32_t)c ? SM : SN);
                        v.w = ps_pk(((q8.x >> 24) & 0xff) == (uint32_t)c ? SM : SN, ((q8.y >> 24) & 0xff) == (uint32_t)c ? SM : SN);
                        ps_sts128(prof_s + c * 512, v.x, v.y, v.z, v.w);
                    }
                }
                // row 0 (virtual start): H = 0, F = -inf (sisd_alignment_engine.cpp:137-141,159-165); it is spill
                // slot 0 and ring entry 0
                const uint32_t hfs_o = (uint32_t)t * 256u + (uint32_t)lane * 4u;  // my 16 bytes of H in slot 0 (F 512 B further)
                const uint32_t halo_so = halo_o + (uint32_t)t;
                {
                    const uint4 h0 = make_uint4(0u, 0u, 0u, 0u);
                    const uint4 f0 = make_uint4(ps_pk2(PS_NEGF), ps_pk2(PS_NEGF), ps_pk2(PS_NEGF), ps_pk2(PS_NEGF));
                    *reinterpret_cast<uint4 *>(hf + hfs_o) = h0;
                    *reinterpret_cast<uint4 *>(hf + hfs_o + 128) = f0;
                    ps_sts128(ring_s, h0.x, h0.y, h0.z, h0.w);
                    ps_sts128(ring_s + 512, f0.x, f0.y, f0.z, f0.w);
                    if (lane0) {
                        hf[halo_so] = 0u;
                        ps_sts32(hring_s, 0);
                    }
                }
                __syncwarp();
                // codes of row r: tile (r-1)/8, inside it [strip][lane][row in tile][4 words]
                uint32_t *cd_w = cd + ((size_t)t * 32 + lane) * 32;
                // where this strip stands (bit 0: a strip on its left, bit 1: that one runs in this pass, bit 2: a strip on
                // its right, bit 3: that one runs in this pass) and its mailbox addresses — opaque to the compiler, which
                // would otherwise re-derive all of them from %tid / %ntid in every row
                uint32_t where = (t > 0 ? 1u : 0u) | (pos > 0 ? 2u : 0u) | (t + 1 < nst ? 4u : 0u) | (pos + 1 < NW ? 8u : 0u);
                ps_saddr mb_in = (ps_saddr)__cvta_generic_to_shared(&s_mb[pos > 0 ? pos - 1 : 0][0]);
                ps_saddr mb_out = (ps_saddr)__cvta_generic_to_shared(&s_mb[pos][0]);
                ps_saddr done_in = (ps_saddr)__cvta_generic_to_shared(&s_done[pos]);
                ps_saddr done_out = (ps_saddr)__cvta_generic_to_shared(&s_done[pos + 1 < NW ? pos + 1 : pos]);
#ifndef CUDA_EMU
                asm volatile("" : "+r"(where), "+r"(mb_in), "+r"(mb_out), "+r"(done_in), "+r"(done_out));
#endif
                const bool has_left = (where & 1u) != 0u, left_smem = (where & 2u) != 0u;
                const bool has_right = (where & 4u) != 0u, right_smem = (where & 8u) != 0u;
                const int lane_e8 = lane * e8;
                int cdone = 0;  // rows the right neighbour is known to have consumed
                int bestv = 0, bestr = 0;
                int idx = 0;    // ring entry of row r-1 (row r goes to idx+1 mod PS_K)
                uint4 rc = make_uint4(0, 0, 0, 0);
                if (n >= 1) rc = recs[1];

                for (int r = 1; r <= n; ++r) {
                    const uint4 cur = rc;
                    if (r < n) rc = recs[r + 1];
                    const int np = (int)((cur.x >> 8) & 0xffu);
                    const uint4 sc4 = ps_lds128(prof_s + (cur.x & 0xffu) * 512u);
                    const uint32_t sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
                    uint32_t Hd[4], Fv[4], fpk[4], dpk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) dpk[k] = 0u;
                    // ---- predecessors in in_edges order: ring (near) or HBM (spilled rows), predicated loads
                    auto fold = [&](uint32_t pw, int p, auto first) {
                        uint32_t cH[4], cF[4];
                        int hl;
                        const bool far = (pw & PS_FAR) != 0u;
                        int pi = idx + 1 - (int)(pw & 0xffu);  // ring entry of row r - pw
                        pi += (pi < 0) ? K : 0;
                        pi = far ? 0 : pi;  // (a valid entry; its values are replaced below)
                        ps_fetch_ring(ring_s + (uint32_t)pi * 1024u, hring_s + (uint32_t)pi * 4u, cH, cF, hl);
                        if (far) {
                            const uint32_t slot = pw & 0xffffu;
                            ps_fetch_spilled(hf + (hfs_o + slot * hf_stride), hf + (halo_so + slot * (uint32_t)nst), cH, cF, hl);
                        }
                        ps_fold_pred<SG, SE, decltype(first)::value>(cH, cF, hl, lane0, p, sc, Hd, Fv, fpk, dpk);
                    };
                    fold(cur.y, 0, std::true_type());
                    if (np > 1) {  // straight-line code for the common in-degrees, a loop beyond
                        fold(cur.z, 1, std::false_type());
                        if (np > 2) {
                            fold((np <= 3) ? cur.w : (uint32_t)preds[pred_base + cur.w + 2], 2, std::false_type());
#pragma unroll 1
                            for (int p = 3; p < np; ++p) fold((uint32_t)preds[pred_base + cur.w + p], p, std::false_type());
                        }
                    }

                    // ---- E: contribution of my own columns, warp scan, carry from the strip on the left
                    uint32_t X[4], Xg[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        X[k] = __vimax_s16x2_relu(Hd[k], Fv[k]);
                        Xg[k] = __vadd2(X[k], g2);
                    }
                    uint32_t c = Xg[0];
                    c = __viaddmax_s16x2(c, e2, Xg[1]);
                    c = __viaddmax_s16x2(c, e2, Xg[2]);
                    c = __viaddmax_s16x2(c, e2, Xg[3]);
                    const int tlo = ps_lo(c), thi = ps_hi(c);  // columns 0-3 -> E[4], columns 4-7 -> E[8]
                    int w = __viaddmax_s32(tlo, 4 * SE, thi);
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, w, d);  // lanes < d get their own w: w + d*e8 < w
                        w = __viaddmax_s32(o, d * e8, w);
                    }
                    const int excl = __shfl_up_sync(0xffffffffu, w, 1);
                    int cin = SG, msgH = 0;  // strip 0: E[1] = H[r][0] + g = g
                    if (has_left) {
                        // every lane polls the same word (broadcast read): no divergence, no shuffle
                        uint32_t payload;
                        if (left_smem) {
                            const ps_saddr slot = mb_in + (uint32_t)(r & (PS_D - 1)) * 8u;
                            unsigned long long v = ps_lds64(slot);
                            while ((uint32_t)(v >> 32) != (uint32_t)r) {
                                __nanosleep(40);  // leave the issue slots to the warps that produce the row
                                v = ps_lds64(slot);
                            }
                            payload = (uint32_t)v;
                            __syncwarp();  // all lanes have read the slot before the producer may reuse it
                            if (lane0) ps_sts32(done_in, r);
                        } else {
                            payload = hf[passb_o + (uint32_t)r];
                        }
                        cin = ps_lo(payload);
                        msgH = ps_hi(payload);
                    }
                    const int Ein = lane0 ? cin : max(excl, cin + lane_e8);  // E at my first column
                    const int E4 = __viaddmax_s32(Ein, 4 * SE, tlo);
                    uint32_t E[4], Ee[4], H[4];
                    E[0] = __byte_perm((uint32_t)Ein, (uint32_t)E4, 0x5410);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        Ee[k] = __vadd2(E[k], e2);
                        E[k + 1] = __vmaxs2(Ee[k], Xg[k]);
                    }
                    Ee[3] = __vadd2(E[3], e2);
                    const uint32_t En = __vmaxs2(Ee[3], Xg[3]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) H[k] = __vmaxs2(X[k], E[k]);

                    // ---- hand the row over to the strip on the right: E at its first column, my last H
                    if (has_right) {
                        const int cout = __viaddmax_s32(cin, 32 * e8, w);  // meaningful in lane 31 (inclusive scan)
                        const uint32_t payload = __byte_perm((uint32_t)cout, H[3], 0x7610);
                        if (right_smem) {
                            while (cdone < r - PS_D) {  // all lanes track the consumer's progress (broadcast read)
                                cdone = ps_lds32(done_out);
                                if (cdone < r - PS_D) __nanosleep(100);
                            }
                            if (lane == 31)
                                ps_sts64(mb_out + (uint32_t)(r & (PS_D - 1)) * 8u, ((unsigned long long)(uint32_t)r << 32) | payload);
                        } else if (lane == 31) {
                            hf[passb_o + (uint32_t)r] = payload;
                        }
                    }
                    // ---- the row enters the ring (entry of row r-PS_K, which no later row reads from the ring) and,
                    // if a row further than PS_K ahead needs it, its spill slot in HBM
                    idx = (idx == K - 1) ? 0 : idx + 1;
                    ps_sts128(ring_s + (uint32_t)idx * 1024u, H[0], H[1], H[2], H[3]);
                    ps_sts128(ring_s + (uint32_t)idx * 1024u + 512u, Fv[0], Fv[1], Fv[2], Fv[3]);
                    if (lane0) ps_sts32(hring_s + (uint32_t)idx * 4u, msgH);
                    const uint32_t myslot = cur.x >> 16;
                    if (myslot) {
                        uint32_t *dst = hf + (hfs_o + myslot * hf_stride);
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(H[0], H[1], H[2], H[3]);
                        *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(Fv[0], Fv[1], Fv[2], Fv[3]);
                        if (lane0) hf[halo_so + myslot * (uint32_t)nst] = (uint32_t)msgH;
                    }

                    // ---- traceback codes
                    uint32_t cw[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t nz = __vminu2(H[k], one2);
                        const uint32_t nd = __vminu2(H[k] ^ Hd[k], one2);
                        const uint32_t nf = __vminu2(H[k] ^ Fv[k], one2);
                        const uint32_t nee = __vminu2(((k < 3) ? E[(k + 1) & 3] : En) ^ Ee[k], one2);
                        uint32_t x = nd * 2u + nz;
                        x = nf * 4u + x;
                        x = nee * 8u + x;
                        x = fpk[k] * 16u + x;
                        cw[k] = x;
                    }
                    if (np > 1) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) cw[k] = dpk[k] * 2048u + cw[k];
                    }
                    *reinterpret_cast<uint4 *>(cd_w) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
                    cd_w += (r & 7) ? 4u : cd_tile_step;
                    // ---- best cell of this lane: first row with a strictly larger H (padding columns never win:
                    // they only see mismatches and gaps, so they stay below a real cell's H)
                    uint32_t hm = __vimax3_s16x2(H[0], H[1], H[2]);
                    hm = __vmaxs2(hm, H[3]);
                    const int m = max((int)(hm & 0xffffu), (int)(hm >> 16));
                    if (m > bestv) {
                        bestv = m;
                        bestr = r;
                        *reinterpret_cast<uint4 *>(mine) = make_uint4(H[0], H[1], H[2], H[3]);
                    }
                    __syncwarp();  // ring entry of row r complete before the next row's reads
                }
                // fold this pass' best into the lane's running best: larger H, then smaller row, then smaller column
                if (bestv > 0) {
                    const uint4 b4 = *reinterpret_cast<const uint4 *>(mine);
                    const uint32_t bh[4] = {b4.x, b4.y, b4.z, b4.w};
                    int u = 0;
#pragma unroll
                    for (int x = 7; x >= 0; --x) {
                        const int val = (x < 4) ? ps_lo(bh[x & 3]) : ps_hi(bh[x & 3]);
                        if (val == bestv) u = x;
                    }
                    const int col = j0 + u + 1;
                    const int gbv = (int)mine[4], gbr = (int)mine[5], gbc = (int)mine[6];
                    if (bestv > gbv || (bestv == gbv && (bestr < gbr || (bestr == gbr && col < gbc)))) {
                        mine[4] = (uint32_t)bestv;
                        mine[5] = (uint32_t)bestr;
                        mine[6] = (uint32_t)col;
                    }
                }
            }
            __syncthreads();  // pass complete: mailboxes may be reset, passb and all codes are written
        }

        // global maximum: largest H, then first row in rank order, then first column
        // (simd_alignment_engine.cpp:1162-1167,1194-1196); the traceback runs in k_poa_strip_traceback
        int best = (int)mine[4], bi = (int)mine[5], bj = (int)mine[6];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, d);
            if (ob > best || (ob == best && (oi < bi || (oi == bi && oj < bj)))) {
                best = ob;
                bi = oi;
                bj = oj;
            }
        }
        if (lane == 0) {
            s_best[wid][0] = best;
            s_best[wid][1] = bi;
            s_best[wid][2] = bj;
        }
        __syncthreads();
        if (tid == 0) {
            for (int wv = 1; wv < NW; ++wv) {
                const int ob = s_best[wv][0], oi = s_best[wv][1], oj = s_best[wv][2];
                if (ob > best || (ob == best && (oi < bi || (oi == bi && oj < bj)))) {
                    best = ob;
                    bi = oi;
                    bj = oj;
                }
            }
            *best_cell = make_int4(best, bi, bj, 0);
        }
        __syncthreads();  // *best_cell is visible; s_best and the mailboxes may be reused by the next job
    }
}

template <int SM, int SN, int SG, int SE>
__global__ void __launch_bounds__(PS_MAXW * 32, 3)
k_poa_strip(const PoaSJob *__restrict__ jobs, int n_jobs, const uint8_t *__restrict__ qcodes,
            const uint4 *__restrict__ rec, const int32_t *__restrict__ preds, uint32_t *arena, int4 *best_out,
            unsigned int *job_counter, int K) {
    PS_DYNAMIC_SHARED(uint4, s_dyn);
    __shared__ int s_job;
    __shared__ int4 s_cell;
    while (true) {
        if (threadIdx.x == 0) s_job = (int)atomicAdd(job_counter, 1u);
        __syncthreads();
        const int jb = s_job;
        __syncthreads();
        if (jb >= n_jobs) break;
        const PoaSJob J = jobs[jb];
        ps_align_job<SM, SN, SG, SE>(J, qcodes, rec, preds, arena, K, s_dyn, &s_cell);
        if (threadIdx.x == 0) best_out[jb] = s_cell;
    }
}

// Traceback of the strip kernel's codes (sisd_alignment_engine.cpp:527-656): one WARP per alignment, so that the
// DP kernel's CTAs are not held by this latency-bound walk.  Pairs are (row or -1, query pos or -1), emitted
// end-to-start; the host reverses them and maps rows to node ids.  Lane k speculatively fetches the code of cell
// (i-k, j-k); leading lanes whose move is "diagonal to row i-k-1" are committed 32 at a time, anything else takes
// the general single step.
// (one warp; every lane returns the number of pairs written).  The walk is serial; what it costs is the latency of the
// dependent loads, so the warp keeps two 128-byte tiles in registers — the codes of 8 rows x 8 columns around the
// current cell and the row records of those 8 rows, one word per lane, each fetched with one coalesced load — and reads
// them with shuffles: about five steps of the path per pair of loads instead of two loads per step.
__device__ __forceinline__ int ps_traceback_warp(const PoaSJob &J, const int4 b, const uint4 *__restrict__ rec,
                                                 const int32_t *__restrict__ preds, const int32_t *__restrict__ spill_rows,
                                                 const uint32_t *__restrict__ arena, const int32_t *__restrict__ pool,
                                                 int32_t *aln_out) {
    const int lane = threadIdx.x & 31;
    const int nst = J.n_strips, n = J.n;
    const uint32_t *cd = arena + J.code_off;
    const uint32_t *recs32 = reinterpret_cast<const uint32_t *>(rec + J.row_off);
    const int32_t *pr = preds + J.pred_base;
    const int32_t *sp = spill_rows + J.spill_off;
    const int best = b.x;
    int32_t *out = aln_out + 2 * (size_t)J.aln_off;
    int cnt = 0;
    int i = b.y, j = b.z;
    // graphs with a device mirror: report node ids (what Graph::add_alignment consumes) instead of rows
    const int32_t *order = (J.order_off != ~0ull) ? pool + J.order_off : nullptr;
    int c_tile = -1, c_grp = -1, r_tile = -1;
    uint32_t cw = 0u, rw = 0u;
    auto code_at = [&](int row, int col) -> uint32_t {  // row >= 1, col >= 1
        const int tile = (row - 1) >> 3, jj = col - 1, grp = jj >> 3;
        if (tile != c_tile || grp != c_grp) {
            cw = cd[((size_t)tile * nst + (grp >> 5)) * 1024 + (size_t)(grp & 31) * 32 + lane];
            c_tile = tile;
            c_grp = grp;
            // The walk goes up and to the left, about two to three graph rows per query column, and every new tile costs
            // a DRAM latency (the codes were streamed out by the DP and are long gone from L2): 24 lanes ask L2 for the
            // tiles the path is likely to enter next — eight tile rows upwards in this and the next two tile columns,
            // shifted by the path's slope — so that the dependent loads of the next steps hit L2.
            if (lane < 24) {
                const int dc = lane >> 3, pt = tile - (lane & 7) - 2 * dc, pg = grp - dc;
                if (pt >= 0 && pg >= 0 && (dc | (lane & 7)) != 0)
                    ps_prefetch_l2(cd + ((size_t)pt * nst + (pg >> 5)) * 1024 + (size_t)(pg & 31) * 32);
            }
        }
        const uint32_t wv = __shfl_sync(0xffffffffu, cw, ((row - 1) & 7) * 4 + (jj & 3));
        return (jj & 4) ? (wv >> 16) : (wv & 0xffffu);
    };
    auto rec_word = [&](int row, int k) -> uint32_t {  // word k of the record of row >= 1
        const int tile = (row - 1) >> 3;
        if (tile != r_tile) {
            const int rr = tile * 8 + 1 + (lane >> 2);
            rw = rr <= n ? recs32[4 * (size_t)rr + (lane & 3)] : 0u;
            r_tile = tile;
            if (lane >= 1 && lane <= 4 && tile - lane >= 0) ps_prefetch_l2(recs32 + 4 * ((size_t)(tile - lane) * 8 + 1));
        }
        return __shfl_sync(0xffffffffu, rw, ((row - 1) & 7) * 4 + k);
    };
    auto pred_row = [&](int row, int idx) -> int {
        uint32_t w;
        if (idx == 0) w = rec_word(row, 1);
        else if (idx == 1) w = rec_word(row, 2);
        else {
            const int np = (int)((rec_word(row, 0) >> 8) & 0xffu);
            const uint32_t w3 = rec_word(row, 3);
            w = np <= 3 ? w3 : (uint32_t)pr[w3 + idx];
        }
        return (w & PS_FAR) ? sp[w & 0xffffu] : row - (int)w;
    };
    // pairs are written with ROWS; the rows are mapped to node ids at the end, by all lanes at once (a dependent
    // load -> store per step would stall the walk for a memory latency each time)
    auto emit = [&](int row, int pos) {
        if (lane == 0) {
            out[2 * cnt] = row;
            out[2 * cnt + 1] = pos;
        }
        ++cnt;
    };
    if (best > 0) {
        while (i > 0 && j > 0) {
            // ---- fast path: consecutive "diagonal to the previous row" steps, as far as the cached tile reaches
            // (lane k looks at cell (i-k, j-k); in round 2 of the correction nearly the whole path is such a run)
            {
                const int ri = (i - 1) & 7, cj = (j - 1) & 7;
                (void)code_at(i, j);       // tiles of (i, j) cached
                (void)rec_word(i, 0);
                const int m = min(ri, cj) + 1;
                const int rk = max(ri - lane, 0), ck = max(cj - lane, 0);
                const uint32_t wv = __shfl_sync(0xffffffffu, cw, rk * 4 + (ck & 3));
                const uint32_t ck_code = (ck & 4) ? (wv >> 16) : (wv & 0xffffu);
                const uint32_t dp = ck_code >> 11;
                const uint32_t pw = __shfl_sync(0xffffffffu, rw, rk * 4 + 1 + (int)min(dp, 1u));
                const bool chain = lane < m && (ck_code & 3u) == 1u && dp <= 1u && pw == 1u;
                const unsigned mk = __ballot_sync(0xffffffffu, chain);
                const int run = __ffs(~mk) - 1;
                if (run > 0) {
                    if (lane < run) {
                        out[2 * (cnt + lane)] = i - lane;
                        out[2 * (cnt + lane) + 1] = j - lane - 1;
                    }
                    cnt += run;
                    i -= run;
                    j -= run;
                    continue;
                }
            }
            const uint32_t c0 = code_at(i, j);
            if (!(c0 & 1u)) break;  // H == 0
            if (!(c0 & 2u)) {       // diagonal
                emit(i, j - 1);
                i = pred_row(i, (int)(c0 >> 11));
                j = j - 1;
            } else if (!(c0 & 4u)) {  // vertical; extend_up iff H == F[p][j]+e
                emit(i, -1);
                const bool ext = ((c0 >> 4) & 3u) <= 1u;
                i = pred_row(i, (int)((c0 >> 6) & 31u));
                if (ext) {
                    while (true) {  // extend_up walk (simd_alignment_engine.cpp:1388-1425)
                        const uint32_t c2 = code_at(i, j);
                        const bool stop = ((c2 >> 4) & 3u) >= 1u;  // F == H[p][j]+g
                        emit(i, -1);
                        i = pred_row(i, (int)((c2 >> 6) & 31u));
                        if (stop || i == 0) break;
                    }
                }
            } else {  // horizontal; extend_left iff H == E[j-1]+e, i.e. E[j] is an extension of E[j-1]
                const bool ext = (j >= 2) && !(code_at(i, j - 1) & 8u);
                emit(-1, j - 1);
                j = j - 1;
                if (ext) {
                    while (true) {  // extend_left walk (simd_alignment_engine.cpp:1364-1387)
                        emit(-1, j - 1);
                        --j;
                        if (j < 1) break;
                        if (code_at(i, j) & 8u) break;  // E[j+1] != E[j]+e
                    }
                }
            }
        }
    }
    __syncwarp();
    if (order)
        for (int x = lane; x < cnt; x += 32) {
            const int row = out[2 * x];
            if (row > 0) out[2 * x] = order[row - 1];
        }
    __syncwarp();
    return cnt;
}

__global__ void __launch_bounds__(128) k_poa_strip_traceback(const PoaSJob *__restrict__ jobs, int n_jobs,
                                                             const uint4 *__restrict__ rec,
                                                             const int32_t *__restrict__ preds,
                                                             const int32_t *__restrict__ spill_rows,
                                                             const uint32_t *__restrict__ arena,
                                                             const int4 *__restrict__ best_in,
                                                             const int32_t *__restrict__ pool, int32_t *aln_out,
                                                             int32_t *aln_len) {
    const int jb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (jb >= n_jobs) return;
    const PoaSJob J = jobs[jb];
    const int cnt = ps_traceback_warp(J, best_in[jb], rec, preds, spill_rows, arena, pool, aln_out);
    if ((threadIdx.x & 31) == 0) aln_len[jb] = cnt;
}

}  // namespace rtl
