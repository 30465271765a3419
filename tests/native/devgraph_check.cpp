// CPU validation of poa_devgraph.cuh against PoaGraph (host sort + host staging logic) on replayed POA runs
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "../../rattle_b200/csrc/poa_graph.hpp"
#include "../../rattle_b200/csrc/poa_devgraph.cuh"
using namespace rtl;
static uint8_t code(char c) { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'U': return 4; } return 255; }
int main(int argc, char **argv) {
    long bad_order = 0, bad_rec = 0, folds = 0, bad_msa = 0;
    for (int fi = 1; fi < argc; ++fi) {
        std::ifstream f(argv[fi]); int n; f >> n; std::vector<std::string> seqs(n); for (auto &s : seqs) f >> s;
        std::vector<std::vector<std::pair<int, int>>> alns(n);
        for (int i = 0; i < n; ++i) { int len; f >> len; alns[i].resize(len); for (auto &p : alns[i]) f >> p.first >> p.second; }
        {   // deferred sort (graphs whose rank order lives on the GPU): msa() catches up and gives the same rows
            PoaGraph a, b;
            b.defer_sort = true;
            for (int i = 0; i < n; ++i) {
                a.add_alignment(alns[i], seqs[i].data(), (int)seqs[i].size());
                b.add_alignment(alns[i], seqs[i].data(), (int)seqs[i].size());
            }
            std::vector<std::string> ma, mb;
            a.msa(ma);
            b.msa(mb);
            if (ma != mb || b.rank_to_node != a.rank_to_node) ++bad_msa;
        }
        for (int K = 5; K <= 6; ++K) {
            PoaGraph g;
            const int cap_n = 40000, cap_e = 80000, cap_a = 160000;
            std::vector<int32_t> block(dg_words(cap_n, cap_e, cap_a));
            DGView dv = dg_view(block.data(), cap_n, cap_e, cap_a);
            int n_old = 0, e_old = 0, a_old = 0;
            for (int i = 0; i < n; ++i) {
                g.add_alignment(alns[i], seqs[i].data(), (int)seqs[i].size());
                const int n_new = g.n_nodes(), e_new = (int)g.e_begin.size(), a_new = (int)g.a_node.size();
                std::vector<int32_t> delta(dg_delta_words(n_new - n_old, e_new - e_old, a_new - a_old) + 1, 0);
                uint8_t *let = (uint8_t *)delta.data();
                for (int v = n_old; v < n_new; ++v) let[v - n_old] = code(g.letter[v]);
                int32_t *ed = delta.data() + (n_new - n_old + 3) / 4;
                for (int x = e_old; x < e_new; ++x) { ed[2 * (x - e_old)] = g.e_begin[x]; ed[2 * (x - e_old) + 1] = g.e_end[x]; }
                int32_t *al = ed + 2 * (e_new - e_old);
                for (int x = a_old; x < a_new; ++x) { al[2 * (x - a_old)] = g.a_owner[x]; al[2 * (x - a_old) + 1] = g.a_node[x]; }
                dg_init_nodes(dv, n_old, n_new, delta.data(), 0, 1);
                dg_append(dv, n_old, n_new, e_old, e_new, a_old, a_new, delta.data());
                dg_toposort(dv, n_new);
                ++folds;
                bool ok = true;
                for (int r = 0; r < n_new; ++r) if (dv.order[r] != g.rank_to_node[r]) ok = false;
                if (!ok) { ++bad_order; n_old = n_new; e_old = e_new; a_old = a_new; continue; }
                // records
                int32_t counters[2] = {0, 0};
                std::vector<uint32_t> rec(4 * (n_new + 1)); std::vector<int32_t> preds(e_new + n_new + 4), spill_rows(n_new + 2);
                dg_ranks(dv, n_new, 0, 1);
                dg_plan_spills(dv, n_new, K, counters, spill_rows.data(), 0, 1);
                dg_build_recs(dv, n_new, K, counters, rec.data(), preds.data(), spill_rows.data(), 0, 1);
                // host expectation: rows of predecessors, spilled set
                std::vector<char> spilled(n_new + 1, 0);
                for (int r = 1; r <= n_new; ++r) { int v = g.rank_to_node[r - 1]; for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x]) { int pr = g.node_to_rank[g.e_begin[x]] + 1; if (r - pr > K) spilled[pr] = 1; } }
                int nsp = 0; for (int r = 1; r <= n_new; ++r) nsp += spilled[r];
                if (nsp != counters[0]) ok = false;
                for (int r = 1; r <= n_new && ok; ++r) {
                    int v = g.rank_to_node[r - 1]; int np = g.n_in[v];
                    const uint32_t *rc = &rec[4 * r];
                    if ((rc[0] & 0xff) != code(g.letter[v])) ok = false;
                    if ((int)((rc[0] >> 8) & 0xff) != (np == 0 ? 1 : np)) ok = false;
                    const int myslot = rc[0] >> 16;
                    if ((myslot != 0) != (spilled[r] != 0)) ok = false;
                    if (myslot && spill_rows[myslot] != r) ok = false;
                    int k = 0;
                    auto word_at = [&](int idx) -> uint32_t { const int npp = (rc[0] >> 8) & 0xff; if (idx == 0) return rc[1]; if (idx == 1) return rc[2]; if (npp <= 3) return rc[3]; return (uint32_t)preds[rc[3] + idx]; };
                    if (np == 0) { uint32_t w = word_at(0); int row = (w & DG_FAR) ? spill_rows[w & 0xffff] : r - (int)w; if (row != 0) ok = false; }
                    for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x], ++k) {
                        int pr = g.node_to_rank[g.e_begin[x]] + 1; uint32_t w = word_at(k);
                        int row = (w & DG_FAR) ? spill_rows[w & 0xffff] : r - (int)w;
                        if (row != pr) ok = false;
                        if (((w & DG_FAR) != 0) != (r - pr > K)) ok = false;
                    }
                }
                if (!ok) ++bad_rec;
                n_old = n_new; e_old = e_new; a_old = a_new;
            }
        }
    }
    printf("folds %ld, order mismatches %ld, record mismatches %ld, deferred-sort msa mismatches %ld\n", folds, bad_order,
           bad_rec, bad_msa);
    return (bad_order || bad_rec || bad_msa) ? 1 : 0;
}
