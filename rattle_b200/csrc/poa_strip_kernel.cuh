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
//   * previous-row values stay in registers; other predecessor rows are read back by the lane that wrote them
//     (two coalesced LDG.128 per predecessor), so H/F traffic never crosses threads;
//   * the traceback decisions are stored as one 16-bit code per cell (as in the first kernel), built from
//     0/1 "not equal" flags (XOR + unsigned min) that the FMA pipe packs with IMADs.
//
// Code of cell (r, j):
//   bit 0   H != 0                      (0 -> traceback stops here)
//   bit 1   H != Hdiag                  (0 -> diagonal move to predecessor `dp`)
//   bit 2   H != F                      (0 -> vertical move to predecessor `fp`; extend_up iff F-from-F attains F)
//   bit 3   E[j+1] != E[j]+e            (0 -> the E of the NEXT column is a gap extension)
//   bits 4-5  0: F[fp]+e > H[fp]+g   1: equal   2: H[fp]+g > F[fp]+e        (for the winning F predecessor)
//   bits 6-10  fp  first predecessor (in_edges order) attaining F
//   bits 11-15 dp  first predecessor attaining Hdiag
// Eligibility (checked by the host, otherwise the int32 kernel runs): scores fit int16, in-degree <= 32,
// letters within {A,C,G,T,U}, m > 0 > n,g,e.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtl {

struct PoaSJob {
    uint64_t hf_off;     // u32 words into the HF arena: (n+1) x n_strips x 256 words (H then F per strip), then
                         // halo[(n+1) x n_strips] (H of the column left of the strip), then passb[n+1]
    uint64_t code_off;   // u32 words into the code arena: n x n_strips x 128 words
    uint32_t q_off;      // bytes into the query buffer: n_strips*256 letter codes (0..4, pad 255)
    uint32_t row_off;    // into rec: n+1 entries, entry r describes row r (1-based)
    uint32_t pred_base;  // into preds: CSR of predecessor rows (read for rows with more than 3 predecessors)
    uint32_t aln_off;    // pairs, into the alignment output
    int32_t L, n, n_strips, pad;
};

constexpr int PS_STRIP = 256;  // columns per warp
constexpr int PS_D = 8;        // mailbox ring depth (rows a strip may run ahead of its right neighbour)
constexpr int PS_MAXW = 8;     // warps per CTA (strips per pass)
constexpr int PS_NLET = 5;     // A C G T U
constexpr int PS_NEGF = -1000; // F of the virtual start row: below every H+g, and NEGF+e stays far from int16 limits

__host__ __device__ __forceinline__ size_t ps_hf_words(int n, int n_strips) {
    return (size_t)(n + 1) * n_strips * 256 + (size_t)(n + 1) * n_strips + (size_t)(n + 1);
}
__host__ __device__ __forceinline__ size_t ps_code_words(int n, int n_strips) { return (size_t)n * n_strips * 128; }

__device__ __forceinline__ int ps_lo(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int ps_hi(uint32_t v) { return ((int)v) >> 16; }
__device__ __forceinline__ uint32_t ps_pk(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ uint32_t ps_pk2(int v) { return ps_pk(v, v); }

// predecessor row of (row record, index) — records keep up to three predecessors inline
__device__ __forceinline__ int ps_pred_row(const uint4 &rc, int idx, const int32_t *pr) {
    const int np = (int)(rc.x >> 8);
    if (idx == 0) return (int)rc.y;
    if (idx == 1) return (int)rc.z;
    if (np <= 3) return (int)rc.w;
    return pr[rc.w + idx];
}

__global__ void __launch_bounds__(PS_MAXW * 32, 3)
k_poa_strip(const PoaSJob *__restrict__ jobs, int n_jobs, const uint8_t *__restrict__ qcodes,
            const uint4 *__restrict__ rec, const int32_t *__restrict__ preds, uint32_t *HF, uint32_t *codes,
            int32_t *aln_out, int32_t *aln_len, int sm, int sn, int sg, int se, unsigned int *job_counter) {
    extern __shared__ uint4 s_prof[];  // [warp][letter][lane] packed match/mismatch scores of the warp's strip
    __shared__ unsigned long long s_mb[PS_MAXW][PS_D];
    __shared__ int s_done[PS_MAXW];
    __shared__ int s_job;
    __shared__ int s_best[PS_MAXW][3];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NW = blockDim.x >> 5;
    const uint32_t g2 = ps_pk2(sg), e2 = ps_pk2(se);
    const uint32_t one2 = 0x00010001u, two2 = 0x00020002u;
    const int e8 = 8 * se;

    while (true) {
        if (tid == 0) s_job = (int)atomicAdd(job_counter, 1u);
        __syncthreads();
        const int jb = s_job;
        __syncthreads();
        if (jb >= n_jobs) break;
        const PoaSJob J = jobs[jb];
        const int n = J.n, nst = J.n_strips;
        uint32_t *hf = HF + J.hf_off;
        int *halo = reinterpret_cast<int *>(hf + (size_t)(n + 1) * nst * 256);
        uint32_t *passb = reinterpret_cast<uint32_t *>(halo + (size_t)(n + 1) * nst);
        uint32_t *cd = codes + J.code_off;
        const uint8_t *q = qcodes + J.q_off;
        const uint4 *recs = rec + J.row_off;
        const int32_t *pr = preds + J.pred_base;
        const int n_pass = (nst + NW - 1) / NW;

        int gbv = 0, gbr = 0, gbc = 0;  // best cell of this lane over all passes: value, row, column (1-based)

        for (int pass = 0; pass < n_pass; ++pass) {
            for (int i = tid; i < PS_MAXW * PS_D; i += blockDim.x) (&s_mb[0][0])[i] = ~0ull;
            if (tid < PS_MAXW) s_done[tid] = 0;
            __syncthreads();  // also orders passb of the previous pass before its readers
            const int t = pass * NW + wid;
            if (t < nst) {
                const int j0 = t * PS_STRIP + lane * 8;  // 0-based index of my first column (column j0+1)
                uint4 *prof = s_prof + (size_t)wid * PS_NLET * 32;
                {
                    const uint2 q8 = *reinterpret_cast<const uint2 *>(q + j0);
#pragma unroll
                    for (int c = 0; c < PS_NLET; ++c) {
                        uint4 v;
                        v.x = ps_pk(((q8.x) & 0xff) == (uint32_t)c ? sm : sn, ((q8.y) & 0xff) == (uint32_t)c ? sm : sn);
                        v.y = ps_pk(((q8.x >> 8) & 0xff) == (uint32_t)c ? sm : sn, ((q8.y >> 8) & 0xff) == (uint32_t)c ? sm : sn);
                        v.z = ps_pk(((q8.x >> 16) & 0xff) == (uint32_t)c ? sm : sn, ((q8.y >> 16) & 0xff) == (uint32_t)c ? sm : sn);
                        v.w = ps_pk(((q8.x >> 24) & 0xff) == (uint32_t)c ? sm : sn, ((q8.y >> 24) & 0xff) == (uint32_t)c ? sm : sn);
                        prof[c * 32 + lane] = v;
                    }
                }
                // row 0 (virtual start): H = 0, F = -inf (sisd_alignment_engine.cpp:137-141,159-165)
                uint32_t pH[4], pF[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    pH[k] = 0u;
                    pF[k] = ps_pk2(PS_NEGF);
                }
                {
                    uint32_t *dst = hf + (size_t)t * 256 + lane * 4;
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(pH[0], pH[1], pH[2], pH[3]);
                    *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(pF[0], pF[1], pF[2], pF[3]);
                    if (lane == 0) halo[t] = 0;
                }
                __syncwarp();
                int prevHalo = 0;  // lane 0: H[r-1][column left of the strip]
                int cdone = 0;     // lane 31: rows the right neighbour is known to have consumed
                int bestv = 0, bestr = 0;
                uint32_t bh[4] = {0u, 0u, 0u, 0u};
                uint4 rc = make_uint4(0, 0, 0, 0);
                if (n >= 1) rc = recs[1];

                for (int r = 1; r <= n; ++r) {
                    const uint4 cur = rc;
                    if (r < n) rc = recs[r + 1];
                    const int letter = (int)(cur.x & 0xffu);
                    const int np = (int)(cur.x >> 8);
                    const uint4 sc4 = prof[letter * 32 + lane];
                    const uint32_t sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
                    uint32_t Hd[4], Fv[4], fpk[4], dpk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) dpk[k] = 0u;

                    for (int p = 0; p < np; ++p) {
                        const int prow = (p == 0) ? (int)cur.y
                                                  : (p == 1) ? (int)cur.z : ((np <= 3) ? (int)cur.w : pr[cur.w + p]);
                        uint32_t cH[4], cF[4];
                        int hl;
                        if (prow == r - 1) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                cH[k] = pH[k];
                                cF[k] = pF[k];
                            }
                            hl = prevHalo;
                        } else {
                            const uint32_t *src = hf + ((size_t)prow * nst + t) * 256 + lane * 4;
                            const uint4 h4 = *reinterpret_cast<const uint4 *>(src);
                            const uint4 f4 = *reinterpret_cast<const uint4 *>(src + 128);
                            hl = (lane == 0 && t > 0) ? halo[(size_t)prow * nst + t] : 0;  // column 0 of every row is H = 0
                            cH[0] = h4.x; cH[1] = h4.y; cH[2] = h4.z; cH[3] = h4.w;
                            cF[0] = f4.x; cF[1] = f4.y; cF[2] = f4.z; cF[3] = f4.w;
                        }
                        uint32_t left = __shfl_up_sync(0xffffffffu, cH[3], 1);
                        if (lane == 0) left = (uint32_t)hl << 16;
                        uint32_t hprev[4];
                        hprev[0] = __byte_perm(left, cH[3], 0x5432);  // (column -1, column 3)
                        hprev[1] = cH[0];
                        hprev[2] = cH[1];
                        hprev[3] = cH[2];
                        if (p == 0) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                Hd[k] = __vadd2(hprev[k], sc[k]);
                                const uint32_t fh = __vadd2(cH[k], g2), fe = __vadd2(cF[k], e2);
                                Fv[k] = __vmaxs2(fh, fe);
                                // clamp(fh - fe, -1, 1) + 1 without a packed subtract: fh + ~fe = fh - fe - 1
                                fpk[k] = __viaddmin_s16x2_relu(__vadd2(fh, ~fe), two2, two2);
                            }
                        } else {
                            const uint32_t pp = ps_pk2(p), pp4 = ps_pk2(4 * p);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t d = __vadd2(hprev[k], sc[k]);
                                const uint32_t nh = __vmaxs2(Hd[k], d);
                                const uint32_t mh = __vminu2(nh ^ Hd[k], one2) * 0xffffu;  // strictly better -> 0xffff
                                dpk[k] = (dpk[k] & ~mh) | (pp & mh);
                                Hd[k] = nh;
                                const uint32_t fh = __vadd2(cH[k], g2), fe = __vadd2(cF[k], e2);
                                const uint32_t fm = __vmaxs2(fh, fe);
                                const uint32_t sf = __viaddmin_s16x2_relu(__vadd2(fh, ~fe), two2, two2);
                                const uint32_t nf = __vmaxs2(Fv[k], fm);
                                const uint32_t mf = __vminu2(nf ^ Fv[k], one2) * 0xffffu;
                                fpk[k] = (fpk[k] & ~mf) | ((sf + pp4) & mf);
                                Fv[k] = nf;
                            }
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
                    int w = max(thi, tlo + 4 * se);
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, w, d);  // lanes < d get their own w: w + d*e8 < w
                        w = max(w, o + d * e8);
                    }
                    const int excl = __shfl_up_sync(0xffffffffu, w, 1);
                    int cin = sg, msgH = 0;  // strip 0: E[1] = H[r][0] + g = g
                    if (t > 0) {
                        uint32_t payload = 0;
                        if (lane == 0) {
                            if (wid > 0) {
                                const volatile unsigned long long *slot = &s_mb[wid - 1][r & (PS_D - 1)];
                                unsigned long long v;
                                do {
                                    v = *slot;
                                } while ((uint32_t)(v >> 32) != (uint32_t)r);
                                payload = (uint32_t)v;
                                *reinterpret_cast<volatile int *>(&s_done[wid]) = r;
                            } else {
                                payload = passb[r];
                            }
                        }
                        payload = __shfl_sync(0xffffffffu, payload, 0);
                        cin = ps_lo(payload);
                        msgH = ps_hi(payload);
                    }
                    const int Ein = (lane == 0) ? cin : max(excl, cin + lane * e8);  // E at my first column
                    const int E4 = max(tlo, Ein + 4 * se);
                    uint32_t E[4], Ee[4], H[4];
                    E[0] = ps_pk(Ein, E4);
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
                    if (t + 1 < nst && lane == 31) {
                        const int cout = max(w, cin + 32 * e8);
                        const uint32_t payload = ps_pk(cout, ps_hi(H[3]));
                        if (wid + 1 < NW) {
                            while (cdone < r - PS_D) cdone = *reinterpret_cast<volatile int *>(&s_done[wid + 1]);
                            *reinterpret_cast<volatile unsigned long long *>(&s_mb[wid][r & (PS_D - 1)]) =
                                ((unsigned long long)(uint32_t)r << 32) | payload;
                        } else {
                            passb[r] = payload;
                        }
                    }
                    if (t > 0 && lane == 0) {
                        halo[(size_t)r * nst + t] = msgH;
                        prevHalo = msgH;
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
                        x = dpk[k] * 2048u + x;
                        cw[k] = x;
                    }
                    {
                        uint32_t *dst = hf + ((size_t)r * nst + t) * 256 + lane * 4;
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(H[0], H[1], H[2], H[3]);
                        *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(Fv[0], Fv[1], Fv[2], Fv[3]);
                        *reinterpret_cast<uint4 *>(cd + ((size_t)(r - 1) * nst + t) * 128 + lane * 4) =
                            make_uint4(cw[0], cw[1], cw[2], cw[3]);
                    }
                    // ---- best cell of this lane: first row with a strictly larger H (padding columns never win:
                    // they only see mismatches and gaps, so they stay below a real cell's H)
                    uint32_t hm = __vimax3_s16x2(H[0], H[1], H[2]);
                    hm = __vmaxs2(hm, H[3]);
                    const int m = max((int)(hm & 0xffffu), (int)(hm >> 16));
                    if (m > bestv) {
                        bestv = m;
                        bestr = r;
#pragma unroll
                        for (int k = 0; k < 4; ++k) bh[k] = H[k];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        pH[k] = H[k];
                        pF[k] = Fv[k];
                    }
                }
                // fold this pass' best into the lane's running best: larger H, then smaller row, then smaller column
                if (bestv > 0) {
                    int u = 0;
#pragma unroll
                    for (int x = 7; x >= 0; --x) {
                        const int val = (x < 4) ? ps_lo(bh[x & 3]) : ps_hi(bh[x & 3]);
                        if (val == bestv) u = x;
                    }
                    const int col = j0 + u + 1;
                    if (bestv > gbv || (bestv == gbv && (bestr < gbr || (bestr == gbr && col < gbc)))) {
                        gbv = bestv;
                        gbr = bestr;
                        gbc = col;
                    }
                }
            }
            __syncthreads();  // pass complete: mailboxes may be reset, passb and all codes are written
        }

        // global maximum: largest H, then first row in rank order, then first column
        // (simd_alignment_engine.cpp:1162-1167,1194-1196)
        int best = gbv, bi = gbr, bj = gbc;
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
        if (wid == 0) {
            for (int wv = 1; wv < NW; ++wv) {
                const int ob = s_best[wv][0], oi = s_best[wv][1], oj = s_best[wv][2];
                if (ob > best || (ob == best && (oi < bi || (oi == bi && oj < bj)))) {
                    best = ob;
                    bi = oi;
                    bj = oj;
                }
            }
            // ---- traceback (warp 0; sisd_alignment_engine.cpp:527-656).  Pairs are (row or -1, query pos or -1),
            // emitted end-to-start; the host reverses them and maps rows to node ids.  Lane k speculatively
            // fetches the code of cell (i-k, j-k); leading lanes whose move is "diagonal to row i-k-1" are
            // committed 32 at a time, anything else takes the general single step.
            int32_t *out = aln_out + 2 * (size_t)J.aln_off;
            int cnt = 0;
            int i = bi, j = bj;
            auto code_at = [&](int row, int col) -> uint32_t {  // row >= 1, col >= 1
                const int jj = col - 1;
                const uint32_t wv = cd[((size_t)(row - 1) * nst + (jj >> 8)) * 128 + ((jj >> 3) & 31) * 4 + (jj & 3)];
                return (jj & 4) ? (wv >> 16) : (wv & 0xffffu);
            };
            if (best > 0) {
                while (i > 0 && j > 0) {
                    const int ik = i - lane, jk = j - lane;
                    const bool valid = ik >= 1 && jk >= 1;
                    uint32_t c = 0;
                    int prow = -1;
                    if (valid) {
                        c = code_at(ik, jk);
                        const uint4 rcd = recs[ik];
                        if ((c & 3u) == 1u) prow = ps_pred_row(rcd, (int)(c >> 11), pr);
                    }
                    const bool chain = valid && (c & 3u) == 1u && prow == ik - 1;
                    const unsigned mk = __ballot_sync(0xffffffffu, chain);
                    const int run = (mk == 0xffffffffu) ? 32 : (__ffs(~mk) - 1);
                    if (lane < run) {
                        out[2 * (cnt + lane)] = ik;
                        out[2 * (cnt + lane) + 1] = jk - 1;
                    }
                    cnt += run;
                    i -= run;
                    j -= run;
                    if (run == 32) continue;
                    if (i <= 0 || j <= 0) break;
                    const uint32_t c0 = __shfl_sync(0xffffffffu, c, run);
                    const int prow0 = __shfl_sync(0xffffffffu, prow, run);
                    if (!(c0 & 1u)) break;  // H == 0
                    if (!(c0 & 2u)) {       // diagonal
                        if (lane == 0) {
                            out[2 * cnt] = i;
                            out[2 * cnt + 1] = j - 1;
                        }
                        ++cnt;
                        i = prow0;
                        j = j - 1;
                    } else if (!(c0 & 4u)) {  // vertical; extend_up iff H == F[p][j]+e
                        if (lane == 0) {
                            out[2 * cnt] = i;
                            out[2 * cnt + 1] = -1;
                        }
                        ++cnt;
                        const bool ext = ((c0 >> 4) & 3u) <= 1u;
                        i = ps_pred_row(recs[i], (int)((c0 >> 6) & 31u), pr);
                        if (ext) {
                            while (true) {  // extend_up walk (simd_alignment_engine.cpp:1388-1425)
                                const uint32_t c2 = code_at(i, j);
                                const bool stop = ((c2 >> 4) & 3u) >= 1u;  // F == H[p][j]+g
                                if (lane == 0) {
                                    out[2 * cnt] = i;
                                    out[2 * cnt + 1] = -1;
                                }
                                ++cnt;
                                i = ps_pred_row(recs[i], (int)((c2 >> 6) & 31u), pr);
                                if (stop || i == 0) break;
                            }
                        }
                    } else {  // horizontal; extend_left iff H == E[j-1]+e, i.e. E[j] is an extension of E[j-1]
                        const bool ext = (j >= 2) && !(code_at(i, j - 1) & 8u);
                        if (lane == 0) {
                            out[2 * cnt] = -1;
                            out[2 * cnt + 1] = j - 1;
                        }
                        ++cnt;
                        j = j - 1;
                        if (ext) {
                            while (true) {  // extend_left walk (simd_alignment_engine.cpp:1364-1387)
                                if (lane == 0) {
                                    out[2 * cnt] = -1;
                                    out[2 * cnt + 1] = j - 1;
                                }
                                ++cnt;
                                --j;
                                if (j < 1) break;
                                if (code_at(i, j) & 8u) break;  // E[j+1] != E[j]+e
                            }
                        }
                    }
                }
            }
            if (lane == 0) aln_len[jb] = cnt;
        }
        __syncthreads();
    }
}

}  // namespace rtl
