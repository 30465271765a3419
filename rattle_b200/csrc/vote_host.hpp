// Host restatement of the MSA post-processing of hot path B (correct.cpp:94-309): column vote and read correction.
// It is the path of packs the device pipeline (poa_vote.cuh) does not take, and what tests/native/vote_check.cpp compares
// the device routines with.  Plain C++ (no CUDA); InputError comes from the includer (common.cuh) or is defined by the test.
#pragma once
#include <math.h>

#include <string>
#include <vector>

#include "msa_ends.hpp"

namespace {

// utils.cpp:6-13
inline char phred_symbol(double p) { return (char)(-10 * log10(p) + 33); }
inline double phred_err_exact(char c) {
    double q = c - 33;
    return pow(10.0, -q / 10.0);
}
// same libm pow() values, tabulated once per byte value (the column vote calls this for every base twice)
struct PhredTable {
    double v[256];
    PhredTable() {
        for (int i = 0; i < 256; ++i) v[i] = phred_err_exact((char)i);
    }
};
inline double phred_err(char c) {
    static const PhredTable T;
    return T.v[(unsigned char)c];
}

// Column statistics of generate_consensus_vector (correct.cpp:94-193).  Symbols are indexed in the iteration order
// of the reference's std::unordered_map<char,pos_info_t> (insertion A,C,T,U,G,'-' -> iteration U,'-',G,T,C,A with
// libstdc++), because the strict '>' vote keeps the first symbol in that order on ties.
constexpr int NSYM = 6;
const char SYM[NSYM] = {'U', '-', 'G', 'T', 'C', 'A'};
inline int sym_index(char c) {
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
struct ColStats {
    std::vector<int> occ, total_occ;  // [col*6 + sym]
    std::vector<double> err;
    std::vector<char> consensus;
};

void consensus_vector(const std::vector<Read> &reads, const std::vector<std::string> &aln, ColStats &cs) {
    cs.occ.clear();
    cs.total_occ.clear();
    cs.err.clear();
    cs.consensus.clear();
    if (reads.empty() || aln.empty()) return;
    const size_t ncol = aln[0].size();
    cs.occ.assign(ncol * NSYM, 0);
    cs.total_occ.assign(ncol * NSYM, 0);
    cs.err.assign(ncol * NSYM, 0.0);
    for (size_t i = 0; i < reads.size(); ++i) {
        const std::string &row = aln[i];
        const std::string &qual = reads[i].quality;
        int seq_pos = -1;
        for (size_t k = 0; k < row.size(); ++k) {
            const char nt = row[k];
            double err_p = 0.0;
            if (nt != '-') {
                ++seq_pos;
                err_p = phred_err(qual[seq_pos]);
            }
            if (seq_pos >= 0 && (size_t)seq_pos < qual.size()) {
                const int s = sym_index(nt);
                if (s < 0) throw InputError("base outside ACGTU in a read to correct");
                cs.occ[k * NSYM + s]++;
                cs.err[k * NSYM + s] += err_p;
                if ((size_t)seq_pos == qual.size() - 1) ++seq_pos;  // end of read: trailing gaps do not vote
            }
        }
    }
    cs.consensus.resize(ncol);
    for (size_t k = 0; k < ncol; ++k) {
        int max_occ = 0;
        char max_nt = 0;
        int tot = 0;
        for (int s = 0; s < NSYM; ++s) tot += cs.occ[k * NSYM + s];
        for (int s = 0; s < NSYM; ++s) {
            const int o = cs.occ[k * NSYM + s];
            if (o > 0) {
                cs.total_occ[k * NSYM + s] += tot;
                cs.err[k * NSYM + s] /= double(o);
            }
            if (o > max_occ) {
                max_occ = o;
                max_nt = SYM[s];
            }
        }
        if (max_nt == 0) max_nt = '-';
        cs.consensus[k] = max_nt;
    }
}

std::string strip_gaps(const std::vector<char> &c) {
    std::string s;
    s.reserve(c.size());
    for (char x : c)
        if (x != '-') s.push_back(x);
    return s;
}

// correct.cpp:196-309 (err_ratio is the literal 30.0 the reference passes at correct.cpp:409)
void correct_pack(const std::vector<Read> &reads, const std::vector<std::string> &aln, double min_occ, double gap_occ,
                  double err_ratio, std::vector<Read> &corrected, std::vector<Read> &uncorrected) {
    ColStats cs;
    consensus_vector(reads, aln, cs);
    for (size_t i = 0; i < reads.size(); ++i) {
        const std::string &row = aln[i];
        const std::string &qual = reads[i].quality;
        int seq_pos = -1;
        std::string res_read, res_qt;
        for (size_t k = 0; k < row.size(); ++k) {
            const char nt = row[k];
            double err_p = 0.0;
            if (nt != '-') {
                ++seq_pos;
                err_p = phred_err(qual[seq_pos]);
            }
            if (seq_pos >= 0 && (size_t)seq_pos < qual.size()) {
                const char cnt = cs.consensus[k];
                const int ci = sym_index(cnt);
                const double c_err = cs.err[k * NSYM + ci];
                const double occ_ratio = double(cs.occ[k * NSYM + ci]) / double(cs.total_occ[k * NSYM + ci]);
                if (cnt == '-') {
                    if (nt != '-') {
                        if (!(occ_ratio >= gap_occ)) {
                            res_read += nt;
                            res_qt += qual[seq_pos];
                        }
                    }
                } else {
                    if (nt == '-') {
                        if (occ_ratio >= gap_occ) {
                            res_read += cnt;
                            res_qt += phred_symbol(c_err);
                        }
                    } else if (nt == cnt) {
                        res_read += nt;
                        res_qt += qual[seq_pos];
                    } else if (occ_ratio >= min_occ && err_ratio * err_p > c_err) {
                        res_read += cnt;
                        res_qt += phred_symbol(c_err);
                    } else {
                        res_read += nt;
                        res_qt += qual[seq_pos];
                    }
                }
                if ((size_t)seq_pos == qual.size() - 1) ++seq_pos;
            }
        }
        if (!res_read.empty()) corrected.push_back(Read{reads[i].header, res_read, "+", res_qt});
        else uncorrected.push_back(reads[i]);
    }
}

}  // namespace
