// Device mirror of the partial-order graphs of hot path B: the rank order (spoa's topological sort, graph.cpp:293-353)
// and the DP kernel's row records are produced ON THE GPU, so that the host keeps only Graph::add_alignment
// (graph.cpp:154-271, ≈30 µs per alignment) of the ≈200 µs of per-alignment host work.
//
// The host graph (poa_graph.hpp) stays authoritative: nodes, edges and aligned-node links are append-only logs, and
// after every add_alignment the host sends the new log entries ("delta") of each graph.  One warp per graph
//   1. appends them to the mirror (in-edge lists and aligned lists in creation order = the reference's in_edges_ /
//      aligned_nodes_ids_ order),
//   2. lane 0 runs the reference's iterative DFS (same pushes, same LIFO order, same `check_aligned` rule),
//   3. all lanes build the strip kernel's row records: predecessor words (ring distance or spill slot), spill slots of
//      the rows some later row needs from more than K ranks back, overflow lists of rows with more than 3 predecessors.
// The traceback kernel then reports node ids (order[row-1]) instead of rows, which is what add_alignment consumes.
//
// The core routines are plain functions of raw arrays so that they also compile for the host: tests replay recorded
// alignments through PoaGraph and through these routines and compare rank orders and row records (no GPU needed).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define DG_HD __host__ __device__ __forceinline__
#else
#define DG_HD inline
#endif

namespace rtl {

constexpr uint32_t DG_FAR = 0x80000000u;  // = PS_FAR (poa_strip_kernel.cuh)

struct DGView {
    int32_t *in_head, *in_tail, *al_head, *al_tail, *order, *rank, *slot, *e_begin, *e_next_in, *a_node, *a_next, *stack,
        *pending;
    uint32_t *nrec;  // compact node records of the device-resident chains (poa_devchain.cuh), 2 words per node
    uint8_t *letter, *mark, *check, *lead, *spillf;  // lead[r] = 1: rank r opens an aligned group (an MSA column)
};

// int32 words of one graph's block for the given capacities
DG_HD size_t dg_words(int cap_n, int cap_e, int cap_a) {
    const size_t w = (size_t)cap_n * 6 + (size_t)(cap_n + 1) + (size_t)cap_e * 2 + (size_t)cap_a * 2 + (size_t)(2 * cap_n + 2) +
                     (size_t)(cap_e + cap_a + 2) + (size_t)cap_n * 2 + 1;
    const size_t bytes = (size_t)cap_n * 4 + (size_t)(cap_n + 1);
    return ((w + (bytes + 3) / 4) + 3) & ~(size_t)3;
}

DG_HD DGView dg_view(int32_t *b, int cap_n, int cap_e, int cap_a) {
    DGView v;
    v.in_head = b; b += cap_n;
    v.in_tail = b; b += cap_n;
    v.al_head = b; b += cap_n;
    v.al_tail = b; b += cap_n;
    v.order = b; b += cap_n;
    v.rank = b; b += cap_n;
    v.slot = b; b += cap_n + 1;
    v.e_begin = b; b += cap_e;
    v.e_next_in = b; b += cap_e;
    v.a_node = b; b += cap_a;
    v.a_next = b; b += cap_a;
    v.stack = b; b += 2 * cap_n + 2;
    v.pending = b; b += cap_e + cap_a + 2;
    if (reinterpret_cast<uintptr_t>(b) & 4) ++b;  // records are read as 8-byte words
    v.nrec = reinterpret_cast<uint32_t *>(b); b += 2 * cap_n;
    uint8_t *c = reinterpret_cast<uint8_t *>(b);
    v.letter = c; c += cap_n;
    v.mark = c; c += cap_n;
    v.check = c; c += cap_n;
    v.lead = c; c += cap_n;
    v.spillf = c;
    return v;
}

// Delta of one graph, int32 words: [new letters, 4 per word][new edges: begin,end ...][new aligned links: node,other ...]
DG_HD size_t dg_delta_words(int dn, int de, int da) { return (size_t)(dn + 3) / 4 + 2 * (size_t)de + 2 * (size_t)da; }

// step 1a (all lanes): new nodes
DG_HD void dg_init_nodes(DGView &g, int n_old, int n_new, const int32_t *delta, int lane, int nl) {
    const uint8_t *let = reinterpret_cast<const uint8_t *>(delta);
    for (int v = n_old + lane; v < n_new; v += nl) {
        g.in_head[v] = g.in_tail[v] = g.al_head[v] = g.al_tail[v] = -1;
        g.letter[v] = let[v - n_old];
    }
}
// step 1b (one lane): append the new edges and aligned links in creation order
DG_HD void dg_append(DGView &g, int n_old, int n_new, int e_old, int e_new, int a_old, int a_new, const int32_t *delta) {
    const int32_t *ed = delta + (n_new - n_old + 3) / 4;
    for (int id = e_old; id < e_new; ++id) {
        const int b = ed[2 * (id - e_old)], e = ed[2 * (id - e_old) + 1];
        g.e_begin[id] = b;
        g.e_next_in[id] = -1;
        if (g.in_tail[e] < 0) g.in_head[e] = id;
        else g.e_next_in[g.in_tail[e]] = id;
        g.in_tail[e] = id;
    }
    const int32_t *al = ed + 2 * (e_new - e_old);
    for (int id = a_old; id < a_new; ++id) {
        const int node = al[2 * (id - a_old)], other = al[2 * (id - a_old) + 1];
        g.a_node[id] = other;
        g.a_next[id] = -1;
        if (g.al_tail[node] < 0) g.al_head[node] = id;
        else g.a_next[g.al_tail[node]] = id;
        g.al_tail[node] = id;
    }
}

// step 2 (one lane): graph.cpp:293-353.  The nodes the reference pushes at a node's visit (unfinished in-edge sources in
// order, then — if the node still has its check flag — unfinished aligned nodes, which lose theirs) are written to
// `pending` once and consumed from the back (LIFO = the reference's stack order); a node whose range is exhausted is a
// node the reference finds `valid`: it is marked, and emitted with its aligned group if it still has its check flag.
DG_HD void dg_toposort(DGView &g, int n) {
    for (int i = 0; i < n; ++i) {
        g.mark[i] = 0;
        g.check[i] = 1;
    }
    int emitted = 0, sp = 0, pn = 0;
    for (int i = 0; i < n; ++i) {
        if (g.mark[i] != 0) continue;
        int v = i;
        while (true) {
            const int begin = pn;
            for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x]) {
                const int b = g.e_begin[x];
                if (g.mark[b] != 2) g.pending[pn++] = b;
            }
            if (g.check[v]) {
                for (int x = g.al_head[v]; x >= 0; x = g.a_next[x]) {
                    const int a = g.a_node[x];
                    if (g.mark[a] != 2) {
                        g.pending[pn++] = a;
                        g.check[a] = 0;
                    }
                }
            }
            g.mark[v] = 1;
            g.stack[sp++] = v;
            g.stack[sp++] = begin;
            bool descended = false;
            while (sp > 0) {
                const int fb = g.stack[sp - 1], fv = g.stack[sp - 2];
                bool found = false;
                while (pn > fb) {
                    const int c = g.pending[--pn];
                    if (g.mark[c] != 2) {
                        v = c;
                        found = true;
                        break;
                    }
                }
                if (found) {
                    descended = true;
                    break;
                }
                g.mark[fv] = 2;
                if (g.check[fv]) {
                    g.lead[emitted] = 1;
                    g.order[emitted++] = fv;
                    for (int x = g.al_head[fv]; x >= 0; x = g.a_next[x]) {
                        g.lead[emitted] = 0;
                        g.order[emitted++] = g.a_node[x];
                    }
                }
                sp -= 2;
            }
            if (!descended) break;
        }
    }
}

#ifdef __CUDA_ARCH__
#define DG_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define DG_ATOMIC_EXCH_U8(arr, i) dg_claim_u8((arr), (i))
// one-byte test-and-set through the containing 32-bit word
__device__ __forceinline__ int dg_claim_u8(uint8_t *arr, int i) {
    unsigned int *w = reinterpret_cast<unsigned int *>(reinterpret_cast<uintptr_t>(arr + i) & ~(uintptr_t)3);
    const unsigned int sh = (unsigned int)((reinterpret_cast<uintptr_t>(arr + i) & 3) * 8);
    const unsigned int old = atomicOr(w, 1u << sh);
    return (int)((old >> sh) & 0xffu);
}
#elif defined(CUDA_EMU)  // SIMT emulation on the host (tests/native/cuda_emu.h): lanes are real threads
#define DG_ATOMIC_ADD(p, v) __atomic_fetch_add((p), (v), __ATOMIC_SEQ_CST)
#define DG_ATOMIC_EXCH_U8(arr, i) ((int)__atomic_exchange_n((arr) + (i), (uint8_t)1, __ATOMIC_SEQ_CST))
#else
#define DG_ATOMIC_ADD(p, v) dg_host_add((p), (v))
#define DG_ATOMIC_EXCH_U8(arr, i) dg_host_claim((arr), (i))
inline int dg_host_add(int32_t *p, int v) {
    const int o = *p;
    *p += v;
    return o;
}
inline int dg_host_claim(uint8_t *arr, int i) {
    const int o = arr[i];
    arr[i] = 1;
    return o;
}
#endif

// step 3a (all lanes): rank of every node, spill flags cleared.  (caller syncs the lanes between 3a, 3b and 3c)
DG_HD void dg_ranks(DGView &g, int n, int lane, int nl) {
    for (int r = lane; r < n; r += nl) g.rank[g.order[r]] = r;
    for (int r = lane; r <= n; r += nl) {
        g.spillf[r] = 0;
        g.slot[r] = 0;
    }
}
// step 3b (all lanes): rows needed from more than K ranks back get a spill slot (ids in claim order: any numbering of
// the slots is valid); counters[0] = spill slots so far, spill_rows[slot] = row
DG_HD void dg_plan_spills(DGView &g, int n, int K, int32_t *counters, int32_t *spill_rows, int lane, int nl) {
    for (int r = 1 + lane; r <= n; r += nl) {
        const int v = g.order[r - 1];
        for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x]) {
            const int pr = g.rank[g.e_begin[x]] + 1;
            if (r - pr > K && DG_ATOMIC_EXCH_U8(g.spillf, pr) == 0) {
                const int s = DG_ATOMIC_ADD(&counters[0], 1) + 1;
                g.slot[pr] = s;
                spill_rows[s] = pr;
            }
        }
    }
}
// step 3c (all lanes): row records (layout: poa_strip_kernel.cuh); rows with more than 3 predecessors claim room in
// `preds` (counters[1] = words used).  rec[0] and spill_rows[0] describe the virtual start row.
// Returns the largest in-degree this lane met (the strip kernel's codes hold 5-bit predecessor indices).
DG_HD int dg_build_recs(DGView &g, int n, int K, int32_t *counters, uint32_t *rec /* 4 words per row */, int32_t *preds,
                        int32_t *spill_rows, int lane, int nl) {
    int max_np = 0;
    if (lane == 0) {
        rec[0] = rec[1] = rec[2] = rec[3] = 0u;
        spill_rows[0] = 0;
    }
    for (int r = 1 + lane; r <= n; r += nl) {
        const int v = g.order[r - 1];
        int np = 0;
        for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x]) ++np;
        if (np > max_np) max_np = np;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (np == 0) {  // no in-edge: the virtual start row (row 0 = spill slot 0 unless within the ring)
            w[1] = (r <= K) ? (uint32_t)r : DG_FAR;
            np = 1;
        } else {
            int base = 0;
            if (np > 3) {
                base = DG_ATOMIC_ADD(&counters[1], np);
                w[3] = (uint32_t)base;
            }
            int k = 0;
            for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x], ++k) {
                const int pr = g.rank[g.e_begin[x]] + 1;
                const uint32_t word = (r - pr <= K) ? (uint32_t)(r - pr) : (DG_FAR | (uint32_t)g.slot[pr]);
                if (np > 3) {
                    preds[base + k] = (int32_t)word;
                    if (k < 2) w[1 + k] = word;
                } else {
                    w[1 + k] = word;
                }
            }
        }
        w[0] = (uint32_t)g.letter[v] | ((uint32_t)np << 8) | ((uint32_t)g.slot[r] << 16);
        rec[4 * r + 0] = w[0];
        rec[4 * r + 1] = w[1];
        rec[4 * r + 2] = w[2];
        rec[4 * r + 3] = w[3];
    }
    return max_np;
}

// One graph of one step: where its mirror lives, what is new, where its row records go.
struct DFoldJob {
    uint64_t gbase;      // int32 words into the mirror pool
    uint32_t delta_off;  // int32 words into the delta buffer
    uint32_t rec_off;    // rows into rec (n_new + 1 records)
    uint32_t pred_base;  // into preds (room for e_new words)
    uint32_t spill_off;  // into spill_rows (room for n_new + 1 entries)
    int32_t cap_n, cap_e, cap_a;
    int32_t n_old, n_new, e_old, e_new, a_old, a_new;
    int32_t K, pad;
};

#if defined(__CUDACC__) || defined(CUDA_EMU)
// one warp per graph; counts[2*job] = spill slots, counts[2*job+1] = overflow predecessor words
__global__ void __launch_bounds__(128) k_poa_graph_fold(const DFoldJob *__restrict__ jobs, int n_jobs, int32_t *pool,
                                                        const int32_t *__restrict__ delta, uint32_t *rec, int32_t *preds,
                                                        int32_t *spill_rows, int32_t *counts) {
    const int lane = threadIdx.x & 31;
    const int jb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (jb >= n_jobs) return;
    const DFoldJob J = jobs[jb];
    DGView g = dg_view(pool + J.gbase, J.cap_n, J.cap_e, J.cap_a);
    const int32_t *d = delta + J.delta_off;
    int32_t *cnt = counts + 2 * jb;
    dg_init_nodes(g, J.n_old, J.n_new, d, lane, 32);
    if (lane == 0) cnt[0] = cnt[1] = 0;
    __syncwarp();
    if (lane == 0) {
        dg_append(g, J.n_old, J.n_new, J.e_old, J.e_new, J.a_old, J.a_new, d);
        dg_toposort(g, J.n_new);
    }
    __syncwarp();
    dg_ranks(g, J.n_new, lane, 32);
    __syncwarp();
    dg_plan_spills(g, J.n_new, J.K, cnt, spill_rows + J.spill_off, lane, 32);
    __syncwarp();
    dg_build_recs(g, J.n_new, J.K, cnt, rec + 4 * (size_t)J.rec_off, preds + J.pred_base, spill_rows + J.spill_off, lane, 32);
}
#endif

}  // namespace rtl
