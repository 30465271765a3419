// Host-side MSA end clean-up of hot path B (correct.cpp:32-92), plain C++ so that a CPU test can run it against the
// reference's own fix_msa_ends (tests/test_msa_ends_cpu.py).
#pragma once
#include <algorithm>
#include <string>
#include <vector>

namespace {

struct Read {
    std::string header, seq, ann, quality;
};

// MSA end clean-up of correct.cpp:32-92, restated as "trim the front, turn the row round, trim the front again, turn it
// back".  A read whose first few bases were aligned far away from the rest of it (a block of fewer than 10 bases —
// blocks are separated by four gap columns — followed by at least 20 gap columns in all) loses that block: the columns
// are blanked in the MSA row and the bases dropped from the read and its qualities.
//
// trim_row_front returns false when the trimming consumed the row up to its last column; the reference then leaves the
// row in whatever orientation it has at that moment (its loop ends without the reversal), which matters for the second
// pass: such a row stays reversed.  That quirk is part of the output and is kept.
bool trim_row_front(std::string &row, std::string &seq, std::string &qual) {
    const size_t n = row.size();
    size_t at = 0;
    while (at < n) {
        while (at < n && row[at] == '-') ++at;
        size_t stop = at;
        int gap_run = 0, block_bases = 0;
        for (; gap_run < 4 && stop < n; ++stop) {
            if (row[stop] == '-') {
                ++gap_run;
            } else {
                ++block_bases;
                gap_run = 0;
            }
        }
        if (block_bases >= 10) return true;  // a real block: nothing more to trim on this side
        for (; stop < n && row[stop] == '-'; ++stop) ++gap_run;
        if (gap_run < 20) return true;       // a short block, but not isolated enough
        std::fill(row.begin() + at, row.begin() + stop, '-');
        seq.erase(0, (size_t)block_bases);
        qual.erase(0, (size_t)block_bases);
        at = stop;
    }
    return false;
}

void fix_msa_ends(std::vector<Read> &reads, std::vector<std::string> &aln) {
    for (size_t i = 0; i < aln.size(); ++i) {
        std::string &row = aln[i];
        Read &rd = reads[i];
        for (int side = 0; side < 2; ++side) {  // front, then (row turned round) back
            if (!trim_row_front(row, rd.seq, rd.quality)) break;
            std::reverse(row.begin(), row.end());
            std::reverse(rd.quality.begin(), rd.quality.end());
            std::reverse(rd.seq.begin(), rd.seq.end());
        }
    }
}

}  // namespace
