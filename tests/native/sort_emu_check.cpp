// dc_sort_blocks (rattle_b200/csrc/poa_devchain.cuh: spoa's topological sort restated as label propagation + independent
// per-block DFS passes, run by a whole CTA) under the SIMT emulation, against PoaGraph::topological_sort (the literal
// restatement of graph.cpp:293-353) after EVERY add_alignment of recorded POA runs.  No DP here, so the graphs can be
// large: long branches (blocks that overflow a thread's stacks and are redone serially), in-degree > 2, aligned groups
// of more than two nodes.  argv[1] = threads of the CTA.
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include <cstdio>
#include <fstream>
#include <string>

#include "../../rattle_b200/csrc/poa_graph.hpp"
#include "../../rattle_b200/csrc/poa_devchain.cuh"

using namespace rtl;

static void sort_kernel(DGView g, int n, int cap) {
    DCSort s = dc_sort_view(emu::g_smem, cap, (int)blockDim.x);
    dc_sort_blocks(g, s, n);
}

int main(int argc, char **argv) {
    const unsigned nt = (unsigned)atoi(argv[1]);
    long sorts = 0, bad = 0, bad_lead = 0, max_n = 0;
    for (int fi = 2; fi < argc; ++fi) {
        std::ifstream f(argv[fi]);
        int n;
        f >> n;
        std::vector<std::string> seqs(n);
        for (auto &s : seqs) f >> s;
        std::vector<std::vector<std::pair<int, int>>> alns(n);
        for (int i = 0; i < n; ++i) {
            int len;
            f >> len;
            alns[i].resize(len);
            for (auto &p : alns[i]) f >> p.first >> p.second;
        }
        PoaGraph g;
        const int cap_n = 12000, cap_e = 3 * cap_n, cap_a = 4 * cap_n;
        if (dc_sort_smem(cap_n, (int)nt) > sizeof(emu::g_smem)) return 2;
        std::vector<int32_t> block(dg_words(cap_n, cap_e, cap_a));
        DGView dv = dg_view(block.data(), cap_n, cap_e, cap_a);
        for (int i = 0; i < n; ++i) {
            g.add_alignment(alns[i], seqs[i].data(), (int)seqs[i].size());
            const int nn = g.n_nodes();
            if (nn > cap_n) return 3;
            // device view of the graph: in-edge / aligned lists in creation order + the compact records
            for (int v = 0; v < nn; ++v) {
                dv.in_head[v] = g.in_head[v];
                dv.al_head[v] = g.al_head[v];
                uint32_t x = 0xffffffffu, y = 0xffffffffu;
                for (int e = g.in_head[v]; e >= 0; e = g.e_next_in[e]) x = dc_rec_push(x, (uint32_t)g.e_begin[e]);
                for (int a = g.al_head[v]; a >= 0; a = g.a_next[a]) y = dc_rec_push(y, (uint32_t)g.a_node[a]);
                dv.nrec[2 * v] = x;
                dv.nrec[2 * v + 1] = y;
            }
            for (size_t e = 0; e < g.e_begin.size(); ++e) {
                dv.e_begin[e] = g.e_begin[e];
                dv.e_next_in[e] = g.e_next_in[e];
            }
            for (size_t a = 0; a < g.a_node.size(); ++a) {
                dv.a_node[a] = g.a_node[a];
                dv.a_next[a] = g.a_next[a];
            }
            emu::launch(1, nt, [&]() { sort_kernel(dv, nn, cap_n); });
            ++sorts;
            max_n = std::max<long>(max_n, nn);
            bool ok = true;
            for (int r = 0; r < nn && ok; ++r) ok = dv.order[r] == g.rank_to_node[r];
            if (!ok) ++bad;
            // lead flags: a rank opens a column iff it is not inside the aligned group of the rank that opened the last one
            int r = 0;
            bool lead_ok = true;
            while (r < nn && lead_ok) {
                lead_ok = dv.lead[r] == 1;
                const int k = g.n_al[g.rank_to_node[r]];
                for (int j = 1; j <= k && lead_ok; ++j) lead_ok = dv.lead[r + j] == 0;
                r += k + 1;
            }
            if (!lead_ok) ++bad_lead;
        }
    }
    printf("sorts %ld (up to %ld nodes), order mismatches %ld, lead mismatches %ld\n", sorts, max_n, bad, bad_lead);
    return (bad || bad_lead) ? 1 : 0;
}
