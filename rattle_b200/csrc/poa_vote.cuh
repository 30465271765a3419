// MSA post-processing of hot path B on the GPU: what correct.cpp does with a pack's multiple sequence
// alignment between and after the POA rounds —
//
//   fix_msa_ends               correct.cpp:32-92     pv_fix_row      (thread per MSA row)
//   generate_consensus_vector  correct.cpp:94-193    pv_column       (thread per MSA column; rows visited in read order,
//                                                                     so every double sum has the reference's order)
//   correct_read_pack          correct.cpp:196-309   pv_apply_row    (thread per MSA row)
//   consensus = vote minus '-' correct.cpp:436-445   pv_consensus
//
// — so that a pack's MSA never leaves the device: round 1's corrected reads become round 2's queries in place, and only
// the corrected reads / the consensus travel to the host.  The routines are plain functions of raw arrays and compile for
// the host as well (tests/native/vote_check.cpp runs them against the host restatement of correct_engine.cu).
//
// Exactness.  Counts are integers.  The error sums are doubles added in read order; 10^(-q/10) comes from a 256-entry
// table the HOST fills with libm's pow (utils.cpp:11-13), means are IEEE divisions.  The one libm call that cannot be
// tabulated is phred_symbol(mean error) = (char)(-10*log10(p)+33) (utils.cpp:6-9): the device evaluates it with its own
// log10 and accepts the result only when the value is further than 1e-6 from an integer (both libraries are accurate to a
// few ulp, so they truncate alike); a value that sits on an integer because p IS a table entry takes the host's symbol
// for that entry; everything else is flagged and the host evaluates those few columns before the correction kernel runs.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PV_HD __host__ __device__ __forceinline__
#else
#define PV_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define PV_UNROLL _Pragma("unroll")
#else
#define PV_UNROLL
#endif

namespace rtl {

struct DVPack {          // one per pack of a vote launch
    uint64_t msa_off;    // bytes into the MSA buffer: n_seq rows of ncol chars
    uint64_t out_off;    // bytes into the corrected-read buffers (sequence / quality): n_seq rows of ncol chars
    uint64_t col_off;    // entries into the column records
    uint32_t row_base;   // entries into the row records / corrected lengths
    uint32_t seq_base;   // first entry of the pack in the read table
    int32_t n_seq, ncol;
    int32_t status;      // DVStatus, written by the kernels
    int32_t cons_len;    // round 2: length of the consensus
};
struct DVRead {          // one per MSA row: the read the row belongs to
    uint64_t qual_off;   // bytes into the quality buffer (round 1)
    int32_t L;
    int32_t pad;
};
struct DVRow {           // per MSA row, written by k_vote_rows
    int32_t tf, tb;      // bases fix_msa_ends removed from the front / the back of the read
    int32_t first_col, last_col;  // columns of the row's first and last base (first_col > last_col: no base left)
};
struct DVCol {           // per MSA column, written by k_vote_cols
    double cerr;         // mean error of the consensus symbol
    int32_t occ, tot;    // occurrences of the consensus symbol / of all symbols (0 / 0 for a column nobody votes in)
    char cons;           // consensus symbol
    char psym;           // phred_symbol(cerr)
    uint8_t flag;        // 1: psym has to come from the host's libm
    uint8_t pad[5];
};
enum DVStatus { DV_OK = 0, DV_DEGENERATE = 1, DV_BAD_LETTER = 2 };

// symbols in the iteration order of the reference's unordered_map (correct_engine.cu: SYM)
PV_HD int pv_sym_index(char c) {
    switch (c) {
        case 'U': return 0;
        case '-': return 1;
        case 'G': return 2;
        case 'T': return 3;
        case 'C': return 4;
        case 'A': return 5;
        default: return -1;
    }
}
PV_HD char pv_sym(int s) { return "U-GTCA"[s]; }

// One side of fix_msa_ends (msa_ends.hpp: trim_row_front) on a row of n columns, read from the front (REV = false) or
// from the back; `trimmed` += bases removed.  false: the trimming reached the end of the row.
template <bool REV>
PV_HD bool pv_trim_front(char *row, int n, int &trimmed) {
    auto at_ = [&](int i) -> char & { return row[REV ? n - 1 - i : i]; };
    int at = 0;
    while (at < n) {
        while (at < n && at_(at) == '-') ++at;
        int stop = at, gap_run = 0, block = 0;
        for (; gap_run < 4 && stop < n; ++stop) {
            if (at_(stop) == '-') {
                ++gap_run;
            } else {
                ++block;
                gap_run = 0;
            }
        }
        if (block >= 10) return true;
        for (; stop < n && at_(stop) == '-'; ++stop) ++gap_run;
        if (gap_run < 20) return true;
        for (int i = at; i < stop; ++i) at_(i) = '-';
        trimmed += block;
        at = stop;
    }
    return false;
}

// fix_msa_ends of one row: blanks the trimmed blocks in place, fills R.  A row whose trimming runs to the end of the row
// stays reversed in the reference (msa_ends.hpp): that case is not handled here — DV_DEGENERATE sends the pack to the host.
PV_HD int pv_fix_row(char *row, int n, DVRow &R) {
    R.tf = R.tb = 0;
    int st = DV_OK;
    if (!pv_trim_front<false>(row, n, R.tf)) st = DV_DEGENERATE;
    else if (!pv_trim_front<true>(row, n, R.tb)) st = DV_DEGENERATE;
    int a = 0, b = n - 1;
    while (a < n && row[a] == '-') ++a;
    while (b >= 0 && row[b] == '-') --b;
    R.first_col = a;
    R.last_col = b;
    return st;
}

// quality of the base in every cell of the row (0 = gap); qual = the read's qualities behind the trimmed front
PV_HD void pv_fill_qrow(const char *row, int n, const char *qual, char *qm) {
    int pos = 0;
    for (int k = 0; k < n; ++k) qm[k] = row[k] != '-' ? qual[pos++] : (char)0;
}

// generate_consensus_vector for column k.  msa / qm: n_seq rows of ncol chars; tab[c] = 10^(-(c-33)/10) and symtab[c] =
// phred_symbol(tab[c]) from the host's libm; with_err = false (round 2): only the vote.  Returns DV_OK or DV_BAD_LETTER.
PV_HD int pv_column(const char *msa, const char *qm, const DVRow *rows, int n_seq, int ncol, int k, const double *tab,
                    const unsigned char *symtab, bool with_err, DVCol &C) {
    int occ[6] = {0, 0, 0, 0, 0, 0};
    double err[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int st = DV_OK;
    for (int i = 0; i < n_seq; ++i) {
        if (k < rows[i].first_col || k > rows[i].last_col) continue;  // the row votes between its first and last base
        const char nt = msa[(size_t)i * ncol + k];
        const int s = pv_sym_index(nt);
        if (s < 0) {
            st = DV_BAD_LETTER;
            continue;
        }
        double e = 0.0;
        if (with_err && nt != '-') e = tab[(unsigned char)qm[(size_t)i * ncol + k]];
PV_UNROLL
        for (int x = 0; x < 6; ++x)
            if (x == s) {
                occ[x] += 1;
                err[x] += e;
            }
    }
    int tot = 0;
PV_UNROLL
    for (int x = 0; x < 6; ++x) tot += occ[x];
    int best = -1, max_occ = 0;
PV_UNROLL
    for (int x = 0; x < 6; ++x)
        if (occ[x] > max_occ) {
            max_occ = occ[x];
            best = x;
        }
    C.flag = 0;
    C.psym = 0;
    if (best < 0) {  // nobody votes: '-' with 0 / 0 occurrences (correct.cpp:186-188)
        C.cons = '-';
        C.occ = 0;
        C.tot = 0;
        C.cerr = 0.0;
        return st;
    }
    double cerr = 0.0;
PV_UNROLL
    for (int x = 0; x < 6; ++x)
        if (x == best) cerr = err[x] / (double)occ[x];
    C.cons = pv_sym(best);
    C.occ = max_occ;
    C.tot = tot;
    C.cerr = cerr;
    if (with_err && C.cons != '-') {
        const double x = -10.0 * log10(cerr) + 33.0;
        const double near = rint(x);
        if (fabs(x - near) >= 1e-6 && x > 0.0 && x < 127.0) {  // (a char beyond 127 is the host compiler's business)
            C.psym = (char)x;
        } else {
            const int cand = (int)near;
            if (cand >= 0 && cand < 256 && cerr == tab[cand]) C.psym = (char)symtab[cand];
            else C.flag = 1;
        }
    }
    return st;
}

// correct_read_pack for one row; oseq / oqual hold ncol chars.  Returns the corrected read's length.
PV_HD int pv_apply_row(const char *row, const char *qm, const DVRow &R, const DVCol *cols, const double *tab, double min_occ,
                       double gap_occ, char *oseq, char *oqual) {
    int m = 0;
    for (int k = R.first_col; k <= R.last_col; ++k) {
        const char nt = row[k];
        const DVCol &C = cols[k];
        const char cnt = C.cons;
        const double occ_ratio = (double)C.occ / (double)C.tot;
        if (cnt == '-') {
            if (nt != '-' && !(occ_ratio >= gap_occ)) {
                oseq[m] = nt;
                oqual[m++] = qm[k];
            }
        } else if (nt == '-') {
            if (occ_ratio >= gap_occ) {
                oseq[m] = cnt;
                oqual[m++] = C.psym;
            }
        } else if (nt == cnt) {
            oseq[m] = nt;
            oqual[m++] = qm[k];
        } else if (occ_ratio >= min_occ && 30.0 * tab[(unsigned char)qm[k]] > C.cerr) {  // correct.cpp:409 passes 30.0
            oseq[m] = cnt;
            oqual[m++] = C.psym;
        } else {
            oseq[m] = nt;
            oqual[m++] = qm[k];
        }
    }
    return m;
}

// the pack's consensus: the vote without '-'
PV_HD int pv_consensus(const DVCol *cols, int ncol, char *out) {
    int m = 0;
    for (int k = 0; k < ncol; ++k)
        if (cols[k].cons != '-') out[m++] = cols[k].cons;
    return m;
}

#if defined(__CUDACC__)

// rows: fix_msa_ends and (round 1) the per-cell qualities.  One CTA per pack, one thread per row.
__global__ void __launch_bounds__(64) k_vote_rows(DVPack *packs, const DVRead *__restrict__ reads, char *msa, char *qm,
                                                  const char *__restrict__ quals, DVRow *rows, int with_err) {
    DVPack &P = packs[blockIdx.x];
    if (P.status != DV_OK) return;
    int st = DV_OK;
    for (int i = threadIdx.x; i < P.n_seq; i += blockDim.x) {
        char *row = msa + P.msa_off + (size_t)i * P.ncol;
        DVRow R;
        st = max(st, pv_fix_row(row, P.ncol, R));
        rows[P.row_base + i] = R;
        if (with_err && st == DV_OK) {
            const DVRead rd = reads[P.seq_base + i];
            pv_fill_qrow(row, P.ncol, quals + rd.qual_off + R.tf, qm + P.msa_off + (size_t)i * P.ncol);
        }
    }
    if (st != DV_OK) atomicMax(&P.status, st);
}

// columns: one thread per column; grid.y = pack.  Columns whose quality symbol needs the host are appended to `flagged`
// (column record index as two words, mean error as two words) up to flag_cap entries; n_flagged counts all of them.
__global__ void __launch_bounds__(128) k_vote_cols(DVPack *packs, const char *__restrict__ msa, const char *__restrict__ qm,
                                                   const DVRow *__restrict__ rows, const double *__restrict__ tab,
                                                   const unsigned char *__restrict__ symtab, DVCol *cols, int with_err,
                                                   int4 *flagged, unsigned int *n_flagged, unsigned int flag_cap) {
    DVPack &P = packs[blockIdx.y];
    if (P.status != DV_OK) return;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.ncol; k += gridDim.x * blockDim.x) {
        DVCol C;
        const int st = pv_column(msa + P.msa_off, qm + P.msa_off, rows + P.row_base, P.n_seq, P.ncol, k, tab, symtab,
                                 with_err != 0, C);
        cols[P.col_off + k] = C;
        if (st != DV_OK) atomicMax(&P.status, st);
        if (C.flag) {
            const unsigned int at = atomicAdd(n_flagged, 1u);
            const unsigned long long ci = P.col_off + (unsigned long long)k;
            if (at < flag_cap)
                flagged[at] = make_int4((int)(ci & 0xffffffffull), (int)(ci >> 32), __double2loint(C.cerr), __double2hiint(C.cerr));
        }
    }
}

// the host's answers for the flagged columns: patch[i] = (column record index lo, hi, symbol, -)
__global__ void k_vote_patch(const int4 *__restrict__ patch, unsigned int n, DVCol *cols) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 e = patch[i];
    const unsigned long long ci = (unsigned long long)(unsigned int)e.x | ((unsigned long long)(unsigned int)e.y << 32);
    cols[ci].psym = (char)e.z;
    cols[ci].flag = 0;
}

// read correction: one CTA per pack, one thread per row; out_len[row] = length of the corrected read
__global__ void __launch_bounds__(64) k_vote_apply(const DVPack *__restrict__ packs, const char *__restrict__ msa,
                                                   const char *__restrict__ qm, const DVRow *__restrict__ rows,
                                                   const DVCol *__restrict__ cols, const double *__restrict__ tab,
                                                   double min_occ, double gap_occ, char *out_seq, char *out_qual,
                                                   int32_t *out_len) {
    const DVPack &P = packs[blockIdx.x];
    if (P.status != DV_OK) return;
    for (int i = threadIdx.x; i < P.n_seq; i += blockDim.x) {
        const size_t ro = P.msa_off + (size_t)i * P.ncol, oo = P.out_off + (size_t)i * P.ncol;
        out_len[P.row_base + i] = pv_apply_row(msa + ro, qm + ro, rows[P.row_base + i], cols + P.col_off, tab, min_occ, gap_occ,
                                               out_seq + oo, out_qual + oo);
    }
}

// round 2: the pack's consensus (one thread per pack is plenty: a few thousand columns)
__global__ void __launch_bounds__(32) k_vote_consensus(DVPack *packs, const DVCol *__restrict__ cols, char *out) {
    DVPack &P = packs[blockIdx.x];
    if (P.status != DV_OK || threadIdx.x != 0) return;
    P.cons_len = pv_consensus(cols + P.col_off, P.ncol, out + P.out_off);
}

// round 2's queries from round 1's corrected reads: letter codes (0..4), padded with 255 to whole strips of `strip` columns
struct DVStage {
    uint64_t src_off;  // bytes into the corrected-read buffer
    uint32_t q_off;    // bytes into the query buffer
    int32_t L;
};
__global__ void __launch_bounds__(256) k_vote_stage_queries(const DVStage *__restrict__ st, const char *__restrict__ src,
                                                            uint8_t *q, int strip) {
    const DVStage S = st[blockIdx.x];
    const int padded = (S.L + strip - 1) / strip * strip;
    for (int i = threadIdx.x; i < padded; i += blockDim.x) {
        uint8_t c = 255;
        if (i < S.L) {
            switch (src[S.src_off + i]) {
                case 'A': c = 0; break;
                case 'C': c = 1; break;
                case 'G': c = 2; break;
                case 'T': c = 3; break;
                case 'U': c = 4; break;
                default: c = 255; break;
            }
        }
        q[S.q_off + i] = c;
    }
}

#endif  // __CUDACC__

}  // namespace rtl
