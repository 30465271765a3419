// The device-resident POA chain (rattle_b200/csrc/poa_devchain.cuh) on the CPU, SIMT-emulated (cuda_emu.h): k_poa_chain —
// one CTA works through a whole pack (graph update, DP, traceback per read; the warp-cooperative Graph::add_alignment
// runs from the traceback's output, with NO host graph in the loop) — then k_chain_msa_rows.  Checked against the
// reference's MSA rows (file: sequences, the reference's alignments, its MSA rows); any wrong alignment or graph update
// changes the MSA.
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include <cstdio>
#include <fstream>
#include <string>

#include "../../rattle_b200/csrc/poa_devchain.cuh"

using namespace rtl;

static uint8_t code(char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        case 'U': return 4;
    }
    return 255;
}

int main(int argc, char **argv) {
    long bad_msa = 0, alignments = 0, failed = 0;
    const int cap_scale = argc > 1 ? atoi(argv[1]) : 100;  // percent: small values exercise DC_FAIL_CAP
    const bool sort_in_smem = argc > 2 ? atoi(argv[2]) != 0 : true;  // 0: the global-memory sort of oversized graphs
    for (int fi = 3; fi < argc; ++fi) {
        std::ifstream f(argv[fi]);
        int n;
        f >> n;
        std::vector<std::string> seqs(n);
        for (auto &s : seqs) f >> s;
        std::vector<std::vector<std::pair<int, int>>> ref(n);
        for (int i = 0; i < n; ++i) {
            int len;
            f >> len;
            ref[i].resize(len);
            for (auto &p : ref[i]) f >> p.first >> p.second;
        }
        std::vector<std::string> rows(n);
        for (auto &r : rows) f >> r;
        int maxlen = 0, total = 0;
        for (auto &s : seqs) {
            maxlen = std::max(maxlen, (int)s.size());
            total += (int)s.size();
        }
        const int cap_n = std::max(16, std::min(total + 8, 4 * maxlen + 1024) * cap_scale / 100), cap_e = 3 * cap_n, cap_a = 4 * cap_n;
        const int max_nst = (maxlen + PS_STRIP - 1) / PS_STRIP;
        const int spill_cap = cap_n / 8 + 16;
        std::vector<int32_t> pool(dg_words(cap_n, cap_e, cap_a));
        std::vector<uint32_t> arena(ps_hf_words(cap_n, max_nst, spill_cap) + ps_code_words(cap_n, max_nst) + 64, 0xdeadbeefu);
        std::vector<uint32_t> rec(4 * (cap_n + 1));
        std::vector<int32_t> preds(cap_e + 8), spill_rows(cap_n + 4), aln(2 * (cap_n + maxlen + 8)), aln_len(n, 0), path(total + 8),
            qnode(maxlen + 8);
        std::vector<DCSeq> sq(n);
        std::vector<uint8_t> q;
        uint32_t prel = 0;
        for (int i = 0; i < n; ++i) {
            const int L = (int)seqs[i].size(), nst = (L + PS_STRIP - 1) / PS_STRIP;
            sq[i].q_off = (uint32_t)q.size();
            sq[i].path_rel = prel;
            sq[i].L = L;
            sq[i].pad = 0;
            prel += L;
            for (int x = 0; x < nst * PS_STRIP; ++x) q.push_back(x < L ? code(seqs[i][x]) : 255);
        }
        q.resize(q.size() + 16, 255);
        DCPack P{};
        P.gbase = 0;
        P.hf_off = 0;
        P.code_off = (ps_hf_words(cap_n, max_nst, spill_cap) + 3) & ~(size_t)3;
        P.n_seq = n;
        P.cap_n = cap_n; P.cap_e = cap_e; P.cap_a = cap_a; P.spill_cap = spill_cap;
        unsigned long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        // one CTA, as wide as the pack's longest read needs (strips balanced over passes beyond 8)
        const int n_pass = (max_nst + PS_MAXW - 1) / PS_MAXW, nw = (max_nst + n_pass - 1) / n_pass;
        const int K = fi % 2 ? 5 : 6;  // both ring depths
        const int32_t list0 = 0;
        unsigned counter = 0;
        const int smem_cap_n = sort_in_smem ? dc_sort_cap(72 * 1024, nw * 32) : 0;
        emu::launch(1, (unsigned)nw * 32, [&]() {
            k_poa_chain<5, -4, -8, -6, 256, 3>(&P, &list0, 1, sq.data(), pool.data(), q.data(), reinterpret_cast<uint4 *>(rec.data()),
                                       preds.data(), spill_rows.data(), aln.data(), path.data(), qnode.data(), arena.data(),
                                       stats, &counter, K, smem_cap_n);
        });
        alignments += (long)stats[1];
        if (P.status != DC_OK) {
            ++failed;
            continue;
        }
        std::vector<char> out((size_t)n * P.ncol + 16);
        uint64_t off = 0;
        emu::launch(1, 256, [&]() { k_chain_msa_rows(&P, sq.data(), &off, pool.data(), path.data(), out.data()); });
        for (int i = 0; i < n; ++i)
            if (std::string(out.data() + (size_t)i * P.ncol, P.ncol) != rows[i]) {
                ++bad_msa;
                break;
            }
        if (stats[1] != (unsigned long long)(n - 1)) ++bad_msa;
    }
    printf("chains %d, emulated alignments %ld, failed packs %ld, msa mismatches %ld\n", argc - 3, alignments, failed, bad_msa);
    return bad_msa ? 1 : 0;
}
