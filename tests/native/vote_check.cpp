// The device-side MSA post-processing (rattle_b200/csrc/poa_vote.cuh: pv_fix_row, pv_fill_qrow, pv_column, pv_apply_row,
// pv_consensus — the bodies of k_vote_rows / k_vote_cols / k_vote_apply / k_vote_consensus) on the CPU, against the host
// restatement of correct.cpp:32-309 (msa_ends.hpp, vote_host.hpp), on cases read from stdin:
//   n ncol min_occ gap_occ, then n lines "row seq qual".
// Prints one line per case: "ok" / "degenerate" (the device path hands the pack to the host) or the first difference.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>
struct InputError : std::runtime_error {
    explicit InputError(const std::string &m) : std::runtime_error(m) {}
};
#include "../../rattle_b200/csrc/vote_host.hpp"
#include "../../rattle_b200/csrc/poa_vote.cuh"
using namespace rtl;

int main() {
    double tab[256];
    unsigned char symtab[256];
    for (int i = 0; i < 256; ++i) {  // as poa_engine.cu: vote_tables
        const double q = (char)i - 33;
        tab[i] = pow(10.0, -q / 10.0);
        symtab[i] = (unsigned char)(char)(-10 * log10(tab[i]) + 33);
    }
    int n, ncol, n_cases = 0, n_bad = 0, n_deg = 0;
    long flagged = 0, table_hits = 0, cols_total = 0;
    double min_occ, gap_occ;
    while (std::cin >> n >> ncol >> min_occ >> gap_occ) {
        std::vector<Read> reads(n);
        std::vector<std::string> rows(n);
        for (int i = 0; i < n; ++i) {
            std::cin >> rows[i] >> reads[i].seq >> reads[i].quality;
            reads[i].header = "@r" + std::to_string(i);
            reads[i].ann = "+";
        }
        ++n_cases;
        // ---- device routines
        std::vector<char> msa((size_t)n * ncol), qm((size_t)n * ncol), oseq((size_t)n * ncol), oqual((size_t)n * ncol);
        std::vector<DVRow> R(n);
        std::vector<DVCol> C(ncol);
        std::vector<int> olen(n);
        int st = DV_OK;
        for (int i = 0; i < n; ++i) {
            memcpy(&msa[(size_t)i * ncol], rows[i].data(), ncol);
            st = std::max(st, pv_fix_row(&msa[(size_t)i * ncol], ncol, R[i]));
        }
        // ---- host restatement
        std::vector<Read> hreads = reads;
        std::vector<std::string> hrows = rows;
        fix_msa_ends(hreads, hrows);
        if (st != DV_OK) {
            ++n_deg;
            printf("degenerate\n");
            continue;
        }
        std::vector<Read> corrected, uncorrected;
        correct_pack(hreads, hrows, min_occ, gap_occ, 30.0, corrected, uncorrected);
        ColStats cs;
        consensus_vector(hreads, hrows, cs);
        const std::string hcons = strip_gaps(cs.consensus);
        std::string what;
        for (int i = 0; i < n; ++i) {
            pv_fill_qrow(&msa[(size_t)i * ncol], ncol, reads[i].quality.data() + R[i].tf, &qm[(size_t)i * ncol]);
            if (std::string(&msa[(size_t)i * ncol], ncol) != hrows[i]) what = "row " + std::to_string(i) + " after fix_msa_ends";
            const std::string trimmed = reads[i].seq.substr(R[i].tf, reads[i].seq.size() - R[i].tf - R[i].tb);
            if (trimmed != hreads[i].seq) what = "read " + std::to_string(i) + " after fix_msa_ends";
        }
        for (int k = 0; k < ncol; ++k) {
            if (pv_column(msa.data(), qm.data(), R.data(), n, ncol, k, tab, symtab, true, C[k]) != DV_OK) what = "bad letter";
            ++cols_total;
            if (C[k].flag) {
                ++flagged;
                C[k].psym = (char)(-10 * log10(C[k].cerr) + 33);  // what the host does with a flagged column
            } else if (C[k].cons != '-' && C[k].psym != (char)(-10 * log10(C[k].cerr) + 33)) {
                what = "quality symbol of column " + std::to_string(k);
            }
            if (C[k].cons != '-' && !C[k].flag) {
                const double x = -10.0 * log10(C[k].cerr) + 33.0;
                if (fabs(x - rint(x)) < 1e-6) ++table_hits;
            }
            if (C[k].cons != cs.consensus[k]) what = "consensus of column " + std::to_string(k);
        }
        // the vote alone (round 2) gives the same consensus
        for (int k = 0; k < ncol; ++k) {
            DVCol c2;
            pv_column(msa.data(), nullptr, R.data(), n, ncol, k, tab, symtab, false, c2);
            if (c2.cons != C[k].cons || c2.occ != C[k].occ || c2.tot != C[k].tot) what = "round-2 vote of column " + std::to_string(k);
        }
        std::string dcons((size_t)ncol, ' ');
        dcons.resize((size_t)pv_consensus(C.data(), ncol, &dcons[0]));
        if (dcons != hcons) what = "consensus string";
        size_t ci = 0, ui = 0;
        for (int i = 0; i < n; ++i) {
            olen[i] = pv_apply_row(&msa[(size_t)i * ncol], &qm[(size_t)i * ncol], R[i], C.data(), tab, min_occ, gap_occ,
                                   &oseq[(size_t)i * ncol], &oqual[(size_t)i * ncol]);
            if (olen[i] > 0) {
                if (ci >= corrected.size() || corrected[ci].header != reads[i].header ||
                    corrected[ci].seq != std::string(&oseq[(size_t)i * ncol], olen[i]) ||
                    corrected[ci].quality != std::string(&oqual[(size_t)i * ncol], olen[i]))
                    what = "corrected read " + std::to_string(i);
                ++ci;
            } else {
                if (ui >= uncorrected.size() || uncorrected[ui].header != reads[i].header) what = "uncorrected read " + std::to_string(i);
                ++ui;
            }
        }
        if (ci != corrected.size() || ui != uncorrected.size()) what = "corrected / uncorrected counts";
        if (!what.empty()) {
            ++n_bad;
            printf("DIFF: %s\n", what.c_str());
        } else {
            printf("ok\n");
        }
    }
    fprintf(stderr, "cases %d, mismatches %d, degenerate %d, columns %ld, flagged for the host's log10 %ld, table symbols %ld\n", n_cases,
            n_bad, n_deg, cols_total, flagged, table_hits);
    return n_bad ? 1 : 0;
}
