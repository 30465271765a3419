// The POA kernels' SOURCE on the CPU (SIMT emulation, cuda_emu.h): k_poa_graph_fold -> k_poa_strip ->
// k_poa_strip_traceback for every read of recorded POA runs, alignments compared with the reference's
// (file format of devgraph_check.cpp: sequences, then the reference's alignment of every read as node/pos pairs).
// The host graph (PoaGraph::add_alignment) consumes the emulated kernels' own alignments, as the engine does.
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include <cstdio>
#include <fstream>
#include <string>

#include "../../rattle_b200/csrc/poa_graph.hpp"
#include "../../rattle_b200/csrc/poa_devgraph.cuh"
#include "../../rattle_b200/csrc/poa_strip_kernel.cuh"

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
    long bad = 0, alignments = 0, cells = 0;
    for (int fi = 1; fi < argc; ++fi) {
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
        PoaGraph g;
        g.defer_sort = true;
        const int cap_n = 20000, cap_e = 60000, cap_a = 80000;
        std::vector<int32_t> pool(dg_words(cap_n, cap_e, cap_a));
        int sn = 0, se = 0, sa = 0;
        for (int i = 0; i < n; ++i) {
            const int L = (int)seqs[i].size();
            std::vector<std::pair<int32_t, int32_t>> aln;
            if (g.n_nodes() > 0 && L > 0) {
                // ---- delta + k_poa_graph_fold
                const int n_new = g.n_nodes(), e_new = (int)g.e_begin.size(), a_new = (int)g.a_node.size();
                std::vector<int32_t> delta(dg_delta_words(n_new - sn, e_new - se, a_new - sa) + 1, 0);
                uint8_t *let = (uint8_t *)delta.data();
                for (int v = sn; v < n_new; ++v) let[v - sn] = code(g.letter[v]);
                int32_t *ed = delta.data() + (n_new - sn + 3) / 4;
                for (int x = se; x < e_new; ++x) {
                    ed[2 * (x - se)] = g.e_begin[x];
                    ed[2 * (x - se) + 1] = g.e_end[x];
                }
                int32_t *al = ed + 2 * (e_new - se);
                for (int x = sa; x < a_new; ++x) {
                    al[2 * (x - sa)] = g.a_owner[x];
                    al[2 * (x - sa) + 1] = g.a_node[x];
                }
                const int nst = (L + PS_STRIP - 1) / PS_STRIP;
                const int n_pass = (nst + PS_MAXW - 1) / PS_MAXW;
                const int nw = (nst + n_pass - 1) / n_pass;
                const int K = (fi + i) % 2 ? 5 : 6;  // both ring depths
                DFoldJob F{};
                F.gbase = 0;
                F.cap_n = cap_n; F.cap_e = cap_e; F.cap_a = cap_a;
                F.n_old = sn; F.n_new = n_new; F.e_old = se; F.e_new = e_new; F.a_old = sa; F.a_new = a_new;
                F.K = K;
                std::vector<uint32_t> rec(4 * (n_new + 1));
                std::vector<int32_t> preds(e_new + 8), spill_rows(n_new + 2), counts(2);
                emu::launch(1, 32, [&]() {
                    k_poa_graph_fold(&F, 1, pool.data(), delta.data(), rec.data(), preds.data(), spill_rows.data(), counts.data());
                });
                sn = n_new; se = e_new; sa = a_new;
                // ---- k_poa_strip
                const int n_spill = counts[0];
                std::vector<uint8_t> q((size_t)nst * PS_STRIP + 16, 255);
                for (int x = 0; x < L; ++x) q[x] = code(seqs[i][x]);
                PoaSJob J{};
                std::vector<uint32_t> arena(ps_hf_words(n_new, nst, n_spill) + ps_code_words(n_new, nst) + 64, 0xdeadbeefu);
                J.hf_off = 0;
                J.code_off = (ps_hf_words(n_new, nst, n_spill) + 3) & ~(size_t)3;
                J.q_off = 0; J.row_off = 0; J.pred_base = 0; J.aln_off = 0; J.spill_off = 0;
                J.L = L; J.n = n_new; J.n_strips = nst; J.n_spill = n_spill;
                J.order_off = 4 * (uint64_t)cap_n;  // DGView::order
                int4 best{};
                unsigned counter = 0;
                emu::launch(1, (unsigned)nw * 32, [&]() {
                    k_poa_strip<5, -4, -8, -6>(&J, 1, q.data(), reinterpret_cast<const uint4 *>(rec.data()), preds.data(),
                                               arena.data(), &best, &counter, K);
                });
                // ---- k_poa_strip_traceback
                std::vector<int32_t> out(2 * (n_new + L + 8)), len(1);
                emu::launch(1, 32, [&]() {
                    k_poa_strip_traceback(&J, 1, reinterpret_cast<const uint4 *>(rec.data()), preds.data(), spill_rows.data(),
                                          arena.data(), &best, pool.data(), out.data(), len.data());
                });
                aln.resize(len[0]);
                for (int x = 0; x < len[0]; ++x) {  // reverse (sisd_alignment_engine.cpp:655); node ids come from the kernel
                    aln[x].first = out[2 * (len[0] - 1 - x)];
                    aln[x].second = out[2 * (len[0] - 1 - x) + 1];
                }
                ++alignments;
                cells += (long)L * n_new;
            }
            if (aln.size() != ref[i].size()) ++bad;
            else
                for (size_t x = 0; x < aln.size(); ++x)
                    if (aln[x].first != ref[i][x].first || aln[x].second != ref[i][x].second) {
                        ++bad;
                        break;
                    }
            g.add_alignment(aln, seqs[i].data(), L);
        }
    }
    printf("emulated alignments %ld (%ld cells), mismatches %ld\n", alignments, cells, bad);
    return bad ? 1 : 0;
}
