// sm_100a kernel of hot path B: sequence-to-graph local alignment with affine gaps (spoa kSW, m/n/g/e), i.e. what
// correct.cpp:400,431,526 calls through spoa::AlignmentEngine::align
//   forward DP   spoa/src/simd_alignment_engine.cpp:1086-1188  (scalar twin sisd_alignment_engine.cpp:465-525)
//   traceback    spoa/src/simd_alignment_engine.cpp:1210-1458  (scalar twin sisd_alignment_engine.cpp:527-656)
//
// One CTA per alignment job.  Rows = graph nodes in topological rank order (row 0 = virtual start), processed
// sequentially; threads are spread over query columns, four consecutive columns per thread.  The row-wise gap
// recurrence E[j] = max(H[j-1]+g, E[j-1]+e) is a max-plus prefix scan: with Z[t] = max(Hdiag,F,0)[t] + g - t*e,
// E[j] = (j-1)*e + max_{t<j} Z[t], evaluated with one block-wide max-scan per row.
//
// Instead of materialising H/F/E for the traceback (12 B/cell in the reference) the forward pass stores one
// 16-bit traceback code per cell holding every decision the reference's traceback would make from H/F/E:
//   bits 0-1  move of the main step: 0 stop (H==0), 1 diagonal, 2 vertical, 3 horizontal
//   bits 2-6  predecessor index (in in_edges order) of that move: first p with H==H[p][j-1]+s, resp. first p with
//             H==F[p][j]+e or H==H[p][j]+g
//   bit  7    extend flag: vertical -> H==F[p][j]+e (extend_up), horizontal -> H==E[j-1]+e (extend_left)
//   bits 8-12 F predecessor: first p with F==H[p][j]+g or F==F[p][j]+e      (extend_up walk)
//   bit  13   F stop flag: F==H[p][j]+g for that p
//   bit  14   E extend flag: E[j]==E[j-1]+e                                 (extend_left walk)
// (WIDE variant: 32-bit codes with 13-bit predecessor indices and int32 scores, for in-degree > 32 or reads
// whose scores do not fit int16.)  H and F rows are kept (4 B/cell) because later rows read them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtl {

struct PoaJob {
    uint64_t hf_off;    // cells, into the HF arena
    uint64_t code_off;  // codes, into the code arena
    uint32_t q_off;     // bytes, into the query buffer (padded to a multiple of 4, pad = 0)
    uint32_t row_off;   // into row_info / row_poff (n+1 entries, entry r describes row r, 1-based)
    uint32_t pred_base; // into preds
    uint32_t aln_off;   // pairs, into the alignment output
    int32_t L;          // query length
    int32_t n;          // graph nodes (rows)
};

constexpr int POA_T = 512;        // threads per CTA
constexpr int POA_CPT = 4;        // columns per thread
constexpr int POA_NEG16 = -16000; // "-inf" that survives +e in int16
constexpr int POA_NEGBIG = -(1 << 28);

template <bool WIDE>
struct PoaTypes {
    using cell_t = short2;
    using code_t = uint16_t;
    static constexpr int PB = 5;
    static constexpr int NEG = POA_NEG16;
};
template <>
struct PoaTypes<true> {
    using cell_t = int2;
    using code_t = uint32_t;
    static constexpr int PB = 13;
    static constexpr int NEG = POA_NEGBIG;
};

__host__ __device__ __forceinline__ int poa_lp(int L) { return (L + 3) & ~3; }          // padded columns
__host__ __device__ __forceinline__ int poa_ws(int L) { return poa_lp(L) + 4; }         // HF row stride (cells)

template <bool WIDE>
__global__ void __launch_bounds__(POA_T, 2) k_poa_align(const PoaJob *__restrict__ jobs, int n_jobs,
                                                     const uint8_t *__restrict__ qbytes,
                                                     const uint32_t *__restrict__ row_info,  // letter | npred<<8
                                                     const uint32_t *__restrict__ row_poff,
                                                     const int32_t *__restrict__ preds,      // rows (0 = virtual start)
                                                     typename PoaTypes<WIDE>::cell_t *HF,
                                                     typename PoaTypes<WIDE>::code_t *codes, int32_t *aln_out,
                                                     int32_t *aln_len, int sm, int sn, int sg, int se,
                                                     unsigned int *job_counter) {
    using cell_t = typename PoaTypes<WIDE>::cell_t;
    using code_t = typename PoaTypes<WIDE>::code_t;
    constexpr int PB = PoaTypes<WIDE>::PB;
    constexpr int NEG = PoaTypes<WIDE>::NEG;
    constexpr int MAXP_SMEM = 32;

    __shared__ int s_job;
    __shared__ int s_warp[POA_T / 32], s_warp_ex[POA_T / 32];
    __shared__ int s_pl[POA_T], s_zl[POA_T];
    __shared__ int s_hl[2][POA_T];  // H of every thread's last column, by row parity (left neighbour of the next thread)
    __shared__ int s_carry[3];  // running prefix max, P_last, Z_last of the previous chunk (or of column 0)
    __shared__ int s_pred[2][MAXP_SMEM];
    __shared__ uint32_t s_info[2], s_poff[2];
    __shared__ int s_best[POA_T / 32][3];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = POA_T / 32;

    while (true) {
        if (tid == 0) s_job = (int)atomicAdd(job_counter, 1u);
        __syncthreads();
        const int jb = s_job;
        __syncthreads();
        if (jb >= n_jobs) break;
        const PoaJob J = jobs[jb];
        const int L = J.L, n = J.n;
        const int Ws = poa_ws(L), Wc = poa_lp(L);
        cell_t *hf = HF + J.hf_off;
        code_t *cd = codes + J.code_off;
        const uint8_t *q = qbytes + J.q_off;
        const uint32_t *rinfo = row_info + J.row_off;
        const uint32_t *rpoff = row_poff + J.row_off;
        const int32_t *pr = preds + J.pred_base;
        const int n_chunks = (L + POA_T * POA_CPT - 1) / (POA_T * POA_CPT);

        // row 0: H = 0, F = -inf (sisd_alignment_engine.cpp:137-141,159-165)
        for (int c = tid; c < Ws; c += POA_T) {
            cell_t z;
            z.x = 0;
            z.y = NEG;
            hf[c] = z;
        }
        // metadata of row 1
        if (tid == 0 && n >= 1) {
            s_info[1] = rinfo[1];
            s_poff[1] = rpoff[1];
        }
        __syncthreads();
        if (n >= 1) {
            const int np1 = (int)(s_info[1] >> 8);
            if (tid < np1 && tid < MAXP_SMEM) s_pred[1][tid] = pr[s_poff[1] + tid];
        }
        __syncthreads();

        int best = 0, bi = 0, bj = 0;
        // Single-chunk jobs (L <= 2048) carry the previous row in registers: when a predecessor is row r-1 (the common
        // case inside linear stretches of the graph) its H/F are never re-read from memory, and the row needs three
        // block barriers instead of four.
        const bool single = n_chunks == 1;
        int pH[4], pF[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            pH[i] = 0;
            pF[i] = NEG;
        }
        s_hl[0][tid] = 0;
        __syncthreads();

        for (int r = 1; r <= n; ++r) {
            const int buf = r & 1;
            const uint32_t info = s_info[buf];
            const uint32_t poff = s_poff[buf];
            const int letter = (int)(info & 0xff);
            const int np = (int)(info >> 8);
            // prefetch metadata of the next row into registers (written to smem at the end of the row)
            uint32_t nx_info = 0, nx_poff = 0;
            int nx_pred = 0;
            if (r < n) {
                nx_info = rinfo[r + 1];
                nx_poff = rpoff[r + 1];
                if (tid < (int)(nx_info >> 8) && tid < MAXP_SMEM) nx_pred = pr[nx_poff + tid];
            }
            cell_t *hrow = hf + (size_t)r * Ws;
            code_t *crow = cd + (size_t)(r - 1) * Wc;
            if (tid == 0) {
                cell_t z;
                z.x = 0;
                z.y = NEG;
                hrow[3] = z;  // column 0
                if (!single) {
                    s_carry[0] = sg;          // prefix max over column 0: Z[0] = H[r][0] + g - 0*e
                    s_carry[1] = POA_NEGBIG;  // P of column 0 (E[r][0] = -inf)
                    s_carry[2] = sg;          // Z[0]
                }
            }
            if (!single) __syncthreads();

            for (int ch = 0; ch < n_chunks; ++ch) {
                const int j0 = ch * POA_T * POA_CPT + tid * POA_CPT + 1;  // first of my 4 columns (1-based)
                const bool act = j0 <= L;
                int Hd[4], Fv[4], dp[4], fp[4];
                bool ffh[4], ffe[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    Hd[i] = POA_NEGBIG;
                    Fv[i] = POA_NEGBIG;
                    dp[i] = 0;
                    fp[i] = 0;
                    ffh[i] = false;
                    ffe[i] = false;
                }
                if (act) {
                    const uchar4 q4 = *reinterpret_cast<const uchar4 *>(q + (j0 - 1));
                    int sc[4];
                    sc[0] = (q4.x == letter) ? sm : sn;
                    sc[1] = (q4.y == letter) ? sm : sn;
                    sc[2] = (q4.z == letter) ? sm : sn;
                    sc[3] = (q4.w == letter) ? sm : sn;
                    for (int p = 0; p < np; ++p) {
                        const int prow = (p < MAXP_SMEM) ? s_pred[buf][p] : pr[poff + p];
                        int cH[4], cF[4], leftH;
                        if (single && prow == r - 1) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                cH[i] = pH[i];
                                cF[i] = pF[i];
                            }
                            leftH = (tid == 0) ? 0 : s_hl[(r - 1) & 1][tid - 1];
                        } else {
                            const cell_t *src = hf + (size_t)prow * Ws + (j0 + 3);
                            cell_t c4[4];
                            if (!WIDE) {
                                const uint4 v = *reinterpret_cast<const uint4 *>(src);
                                *reinterpret_cast<uint4 *>(c4) = v;
                            } else {
                                const uint4 v0 = *reinterpret_cast<const uint4 *>(src);
                                const uint4 v1 = *reinterpret_cast<const uint4 *>(src + 2);
                                reinterpret_cast<uint4 *>(c4)[0] = v0;
                                reinterpret_cast<uint4 *>(c4)[1] = v1;
                            }
                            leftH = (int)src[-1].x;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                cH[i] = (int)c4[i].x;
                                cF[i] = (int)c4[i].y;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int hprev = (i == 0) ? leftH : cH[i - 1];
                            const int d = hprev + sc[i];
                            if (d > Hd[i]) {
                                Hd[i] = d;
                                dp[i] = p;
                            }
                            const int fh = cH[i] + sg, fe = cF[i] + se;
                            const int fm = max(fh, fe);
                            if (fm > Fv[i]) {
                                Fv[i] = fm;
                                fp[i] = p;
                                ffh[i] = fh == fm;
                                ffe[i] = fe == fm;
                            }
                        }
                    }
                }
                int X[4], Z[4];
                int zloc = POA_NEGBIG;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    X[i] = max(max(Hd[i], Fv[i]), 0);
                    Z[i] = act ? (X[i] + sg - (j0 + i) * se) : POA_NEGBIG;
                    zloc = max(zloc, Z[i]);
                }
                // block-wide exclusive max-scan of zloc
                int w = zloc;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, w, d);
                    if (lane >= d) w = max(w, o);
                }
                if (lane == 31) s_warp[wid] = w;
                __syncthreads();
                if (wid == 0) {
                    int x = (lane < NW) ? s_warp[lane] : POA_NEGBIG;
#pragma unroll
                    for (int d = 1; d < NW; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, x, d);
                        if (lane >= d) x = max(x, o);
                    }
                    const int ex = __shfl_up_sync(0xffffffffu, x, 1);
                    if (lane < NW) s_warp_ex[lane] = (lane == 0) ? POA_NEGBIG : ex;
                }
                __syncthreads();
                int excl = __shfl_up_sync(0xffffffffu, w, 1);
                if (lane == 0) excl = POA_NEGBIG;
                // column 0 (or the previous chunk): running prefix max, and P / Z of the column left of this chunk
                const int carryP = single ? sg : s_carry[0];
                const int carryPl = single ? POA_NEGBIG : s_carry[1], carryZl = single ? sg : s_carry[2];
                const int Pin = max(carryP, max(s_warp_ex[wid], excl));  // max Z over all columns < j0
                // my columns
                int E[4], H[4], P[4];
                int run = Pin;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    P[i] = run;  // prefix over columns <= j0+i-1
                    E[i] = (j0 + i - 1) * se + run;
                    H[i] = max(X[i], E[i]);
                    run = max(run, Z[i]);
                }
                s_pl[tid] = P[3];
                s_zl[tid] = Z[3];
                s_hl[r & 1][tid] = H[3];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    pH[i] = H[i];
                    pF[i] = max(Fv[i], NEG);
                }
                if (ch + 1 == n_chunks && r < n) {  // next row's metadata (prefetched at the start of this row)
                    if (tid == 0) {
                        s_info[buf ^ 1] = nx_info;
                        s_poff[buf ^ 1] = nx_poff;
                    }
                    if (tid < (int)(nx_info >> 8) && tid < MAXP_SMEM) s_pred[buf ^ 1][tid] = nx_pred;
                }
                __syncthreads();
                int Pl, Zl;  // left neighbour column j0-1: its prefix P and its Z
                if (tid == 0) {
                    Pl = carryPl;
                    Zl = carryZl;
                } else {
                    Pl = s_pl[tid - 1];
                    Zl = s_zl[tid - 1];
                }
                if (act) {
                    cell_t o4[4];
                    code_t k4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = j0 + i;
                        const int Pprev = (i == 0) ? Pl : P[i - 1];
                        const int Zprev = (i == 0) ? Zl : Z[i - 1];
                        const int Eprev = (j - 2) * se + Pprev;  // E[j-1]
                        const bool eext = Zprev <= Pprev;        // E[j] == E[j-1] + e
                        uint32_t code;
                        const int h = H[i];
                        if (h == 0) code = 0;
                        else if (Hd[i] == h) code = 1u | ((uint32_t)dp[i] << 2);
                        else if (Fv[i] == h) code = 2u | ((uint32_t)fp[i] << 2) | ((ffe[i] ? 1u : 0u) << (2 + PB));
                        else code = 3u | (((h == Eprev + se) ? 1u : 0u) << (2 + PB));
                        code |= ((uint32_t)fp[i] << (3 + PB)) | ((ffh[i] ? 1u : 0u) << (3 + 2 * PB)) |
                                ((eext ? 1u : 0u) << (4 + 2 * PB));
                        k4[i] = (code_t)code;
                        cell_t o;
                        o.x = h;
                        o.y = max(Fv[i], NEG);
                        o4[i] = o;
                        if (j <= L && h > best) {
                            best = h;
                            bi = r;
                            bj = j;
                        }
                    }
                    if (!WIDE) {
                        *reinterpret_cast<uint4 *>(hrow + (j0 + 3)) = *reinterpret_cast<const uint4 *>(o4);
                        *reinterpret_cast<uint2 *>(crow + (j0 - 1)) = *reinterpret_cast<const uint2 *>(k4);
                    } else {
                        reinterpret_cast<uint4 *>(hrow + (j0 + 3))[0] = reinterpret_cast<const uint4 *>(o4)[0];
                        reinterpret_cast<uint4 *>(hrow + (j0 + 3))[1] = reinterpret_cast<const uint4 *>(o4)[1];
                        *reinterpret_cast<uint4 *>(crow + (j0 - 1)) = *reinterpret_cast<const uint4 *>(k4);
                    }
                }
                if (!single && tid == POA_T - 1) {
                    s_carry[0] = max(carryP, max(s_warp_ex[NW - 1], w));  // prefix incl. this chunk (w = inclusive scan)
                    s_carry[1] = P[3];
                    s_carry[2] = Z[3];
                }
                if (!single) __syncthreads();  // multi-chunk rows: s_carry hand-over and global H/F rows of row r-1
            }
        }

        // global maximum: largest H, then first row in rank order, then first column
        // (simd_alignment_engine.cpp:1162-1167,1194-1196)
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
            if (lane == 0) {
                for (int wv = 1; wv < NW; ++wv) {
                    const int ob = s_best[wv][0], oi = s_best[wv][1], oj = s_best[wv][2];
                    if (ob > best || (ob == best && (oi < bi || (oi == bi && oj < bj)))) {
                        best = ob;
                        bi = oi;
                        bj = oj;
                    }
                }
            }
            best = __shfl_sync(0xffffffffu, best, 0);
            bi = __shfl_sync(0xffffffffu, bi, 0);
            bj = __shfl_sync(0xffffffffu, bj, 0);
            // ---- traceback (warp 0; sisd_alignment_engine.cpp:527-656).  Pairs are (row or -1, query pos or -1),
            // emitted end-to-start; the host reverses them and maps rows to node ids.  The walk itself is serial, but
            // most of it is a run of diagonal moves through consecutive rows: lane k speculatively fetches the code
            // of cell (i-k, j-k) and its diagonal predecessor, and the leading lanes whose move is "diagonal to row
            // i-k-1" are committed 32 at a time; anything else takes the general single step.
            int32_t *out = aln_out + 2 * (size_t)J.aln_off;
            int cnt = 0;
            int i = bi, j = bj;
            constexpr uint32_t PM = (1u << PB) - 1u;
            if (best > 0) {
                while (i > 0 && j > 0) {
                    const int ik = i - lane, jk = j - lane;
                    const bool valid = ik >= 1 && jk >= 1;
                    uint32_t c = 0;
                    int prow = -1;
                    if (valid) {
                        c = cd[(size_t)(ik - 1) * Wc + (jk - 1)];
                        if ((c & 3u) == 1u) prow = pr[rpoff[ik] + ((c >> 2) & PM)];
                    }
                    const bool chain = valid && (c & 3u) == 1u && prow == ik - 1;
                    const unsigned m = __ballot_sync(0xffffffffu, chain);
                    const int run = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
                    if (lane < run) {
                        out[2 * (cnt + lane)] = ik;
                        out[2 * (cnt + lane) + 1] = jk - 1;
                    }
                    cnt += run;
                    i -= run;
                    j -= run;
                    if (run == 32) continue;
                    if (i <= 0 || j <= 0) break;
                    // general step at (i,j): its code (and diagonal predecessor) sit in lane `run`
                    const uint32_t c0 = __shfl_sync(0xffffffffu, c, run);
                    const int prow0 = __shfl_sync(0xffffffffu, prow, run);
                    const uint32_t type = c0 & 3u;
                    if (type == 0) break;
                    const uint32_t pidx = (c0 >> 2) & PM;
                    const bool ext = (c0 >> (2 + PB)) & 1u;
                    if (type == 1) {
                        if (lane == 0) {
                            out[2 * cnt] = i;
                            out[2 * cnt + 1] = j - 1;
                        }
                        ++cnt;
                        i = prow0;
                        j = j - 1;
                    } else if (type == 2) {
                        if (lane == 0) {
                            out[2 * cnt] = i;
                            out[2 * cnt + 1] = -1;
                        }
                        ++cnt;
                        i = pr[rpoff[i] + pidx];
                        if (ext) {  // extend_up walk
                            while (true) {
                                const uint32_t c2 = cd[(size_t)(i - 1) * Wc + (j - 1)];
                                const uint32_t fpi = (c2 >> (3 + PB)) & PM;
                                const bool stop = (c2 >> (3 + 2 * PB)) & 1u;
                                if (lane == 0) {
                                    out[2 * cnt] = i;
                                    out[2 * cnt + 1] = -1;
                                }
                                ++cnt;
                                i = pr[rpoff[i] + fpi];
                                if (stop || i == 0) break;
                            }
                        }
                    } else {
                        if (lane == 0) {
                            out[2 * cnt] = -1;
                            out[2 * cnt + 1] = j - 1;
                        }
                        ++cnt;
                        j = j - 1;
                        if (ext) {  // extend_left walk
                            while (true) {
                                if (lane == 0) {
                                    out[2 * cnt] = -1;
                                    out[2 * cnt + 1] = j - 1;
                                }
                                ++cnt;
                                --j;
                                if (j < 1) break;
                                const uint32_t c2 = cd[(size_t)(i - 1) * Wc + j];  // cell (i, j+1)
                                if (!((c2 >> (4 + 2 * PB)) & 1u)) break;
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
