// Device-resident POA chains of hot path B: the WHOLE per-pack loop of correct.cpp:399-402 / :430-433 —
//   align read s to the graph of reads 0..s-1, add the alignment to the graph, sort the graph —
// runs on the GPU inside ONE CTA per pack (k_poa_chain), without kernel boundaries or host round trips:
//
//   dc_step_cta         Graph::add_alignment of the previous read's alignment (graph.cpp:154-271, add_sequence :273-291,
//                       add_edge :99-115) on the device mirror of the graph (one warp), spoa's topological sort
//                       (graph.cpp:293-353; one thread, out of the CTA's shared memory), the DP's row records (all threads)
//   ps_align_job        the int16 DP of the next read (poa_strip_kernel.cuh), all warps
//   ps_traceback_warp   its traceback (one warp), which leaves the alignment where the next dc_step_cta reads it
//
// and after the last read the MSA columns; k_chain_msa_rows then writes the multiple sequence alignment
// (graph.cpp:371-426) that the host-side column vote consumes.  The host launches one kernel per CTA width and
// synchronises twice per round (MSA sizes, MSA rows): every size that the GPU discovers on the way (graph nodes, spilled
// rows, alignment lengths) stays on the GPU, and every buffer is a fixed per-pack slot sized from capacities the host
// knows in advance.  A pack that outgrows a capacity, meets in-degree > 32 or more spilled rows than its slot holds is
// flagged and re-run on the host-driven path (poa_engine.cu).
//
// Graph::add_alignment, restated so that a warp does it cooperatively.  The reference walks the alignment serially;
// its result only depends on per-position facts, because an alignment visits every graph node at most once and the
// nodes it visits lie on one path (so no two of them share an aligned group):
//   * query positions before the first / after the last aligned position become two new chains, prefix first, then
//     suffix (graph.cpp:194-200) — new node ids n0.. and n0+|prefix|..;
//   * an aligned position (node a, letter c) re-uses a if its letter is c, else the first node of a's aligned list with
//     letter c, else it becomes a new node cross-linked with a's whole group (graph.cpp:212-243); positions aligned to
//     no node (-1) become new nodes; new ids are handed out in walk order after the two chains;
//   * consecutive nodes of the read's path are joined by an edge unless it exists (graph.cpp:99-115, :251-265); a node
//     receives at most one new in-edge per alignment, appended at the end of its in-edge list, which is the order the
//     DP's predecessor priority depends on.
// Node ids, in-edge list order and aligned-list order are reproduced exactly; edge ids and aligned-entry ids (which
// nothing observes) are handed out per 32-position chunk.
#pragma once
#include "poa_devgraph.cuh"
#include "poa_strip_kernel.cuh"

namespace rtl {

enum DCStatus { DC_OK = 0, DC_FAIL_CAP = 1, DC_FAIL_DEGREE = 2, DC_FAIL_SPILL = 3 };

struct DCPack {            // one per pack; n/e/a/status/ncol are updated on the device
    uint64_t gbase;        // int32 words into the mirror pool (DGView block)
    uint64_t hf_off;       // u32 words into the arena: fixed slot for ps_hf_words(cap_n, max strips, spill_cap)
    uint64_t code_off;     // u32 words into the arena: fixed slot for ps_code_words(cap_n, max strips)
    uint32_t rec_off;      // rows into rec (cap_n + 1)
    uint32_t pred_base;    // into preds (cap_e + 4)
    uint32_t spill_off;    // into spill_rows (cap_n + 2)
    uint32_t aln_off;      // pairs into aln (cap_n + longest read + 8)
    uint32_t path_off;     // int32 words into the path pool: node id of every base of every read of the pack
    uint32_t qnode_off;    // int32 words into the scratch pool (longest read)
    uint32_t seq_base;     // first entry of the pack in the sequence table
    int32_t n_seq;
    int32_t cap_n, cap_e, cap_a, spill_cap;
    int32_t n, e, a;       // graph size
    int32_t status;        // DCStatus
    int32_t ncol;          // MSA columns (k_chain_msa_cols)
    int32_t pad;
};

struct DCSeq {             // one per read of a pack (host-written once)
    uint32_t q_off;        // bytes into the query buffer: strips*256 letter codes 0..4, padded with 255
    uint32_t path_rel;     // int32 words: this read's path inside the pack's path block
    int32_t L;
    int32_t pad;
};

#if defined(__CUDACC__) || defined(CUDA_EMU)

// Compact node record (DGView::nrec, 2 words per node), kept up to date by dc_add_alignment so that the topological sort
// can run out of shared memory: x = first two in-edge sources, y = first two aligned nodes, 16 bits each, DC_NONE = no
// entry, DC_MORE in the high half = the list is longer than two entries (walk the linked list in global memory).
constexpr uint32_t DC_NONE = 0xffffu, DC_MORE = 0xfffeu;
constexpr int DC_THREADS = 128;    // one CTA per pack
constexpr int DC_NDEF = 256;       // roots of blocks the parallel pass could not hold (more: every block is redone serially)
constexpr int DC_TP = 16, DC_TF = 8;  // per-thread pending entries / DFS frames of the parallel pass
constexpr int DC_MAXT = PS_MAXW * 32; // threads of the widest CTA

__host__ __device__ __forceinline__ uint32_t dc_rec_push(uint32_t w, uint32_t id) {
    if ((w & 0xffffu) == DC_NONE) return (w & 0xffff0000u) | id;
    if ((w >> 16) == DC_NONE) return (w & 0xffffu) | (id << 16);
    return (w & 0xffffu) | (DC_MORE << 16);
}
// shared-memory bytes of the sort for graphs of at most cap_n nodes in a CTA of nt threads: label + rank (16 bit each),
// mark/check byte, spill flag bits, and the DFS stacks — per-thread ones for the parallel pass; the serial pass of
// oversized blocks reuses the area (its capacities follow from the area's size; beyond them it works in DGView::pending /
// DGView::stack) — plus the list of deferred blocks
__host__ __device__ __forceinline__ size_t dc_stack_area(int nt) {
    const size_t a = (size_t)nt * (DC_TP * 2 + DC_TF * 4);
    return a < 4096 ? 4096 : a;
}
__host__ __device__ __forceinline__ int dc_serial_pcap(int nt) { return (int)(dc_stack_area(nt) / 5) & ~1; }   // 2-byte entries
__host__ __device__ __forceinline__ int dc_serial_fcap(int nt) { return (int)(dc_stack_area(nt) / 10) & ~1; }  // 6-byte frames
__host__ __device__ __forceinline__ size_t dc_stack_bytes(int nt) { return dc_stack_area(nt) + (size_t)DC_NDEF * 2; }
__host__ __device__ __forceinline__ size_t dc_sort_smem(int cap_n, int nt) {
    const size_t even = ((size_t)cap_n + 3) & ~(size_t)3;
    return even * 2 + even * 2 + even + ((size_t)(cap_n + 1 + 31) / 32 + 1) * 4 + dc_stack_bytes(nt) + 64;
}
// largest graph whose sort fits `bytes` of shared memory (node ids must also fit the 16-bit records)
__host__ __device__ __forceinline__ int dc_sort_cap(size_t bytes, int nt) {
    int lo = 0, hi = 0xfff0;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (dc_sort_smem(mid, nt) <= bytes) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ long long dc_clock() {
#ifdef CUDA_EMU
    return 0;
#else
    return clock64();
#endif
}
__device__ __forceinline__ int dc_warp_sum(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ int dc_warp_excl_scan(int v, int lane, int &total) {
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}
__device__ __forceinline__ void dc_append_aligned(DGView &g, int node, int other, int id) {
    g.a_node[id] = other;
    g.a_next[id] = -1;
    if (g.al_tail[node] < 0) g.al_head[node] = id;
    else g.a_next[g.al_tail[node]] = id;
    g.al_tail[node] = id;
    g.nrec[2 * node + 1] = dc_rec_push(g.nrec[2 * node + 1], (uint32_t)other);
}

// Graph::add_alignment by one warp.  aln: (node id or -1, query position or -1) pairs as the traceback kernel wrote
// them (end-to-start; the order does not matter here), cnt pairs (0 = empty alignment: the read becomes a new chain,
// graph.cpp:174-182).  q: letter codes of the read.  path[0..L) receives the read's node ids.  qnode: L words of scratch.
// Returns DC_OK or DC_FAIL_CAP (graph untouched).
__device__ __forceinline__ int dc_add_alignment(DGView &g, int &n, int &e, int &a, int cap_n, int cap_e, int cap_a,
                                                const uint8_t *q, int L, const int32_t *aln, int cnt, int32_t *qnode,
                                                int32_t *path, int lane) {
    const int n0 = n, e0 = e, a0 = a;
    // ---- which graph node is every query position aligned to (-2: outside the alignment)
    for (int i = lane; i < L; i += 32) qnode[i] = -2;
    __syncwarp();
    int fv = L, lv = -1;
    for (int x = lane; x < cnt; x += 32) {
        const int node = aln[2 * x], pos = aln[2 * x + 1];
        if (pos >= 0) {
            qnode[pos] = node;
            fv = min(fv, pos);
            lv = max(lv, pos);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        fv = min(fv, __shfl_xor_sync(0xffffffffu, fv, d));
        lv = max(lv, __shfl_xor_sync(0xffffffffu, lv, d));
    }
    if (lv < 0) {  // nothing aligned: the whole read is the "prefix" chain
        fv = L;
        lv = L - 1;
    }
    __syncwarp();
    const int n_chain = fv + (L - 1 - lv);  // prefix + suffix nodes, created before the aligned part
    // ---- decisions for the aligned positions, in walk order (reads the graph as it was)
    int n_add = 0, a_need = 0;
    for (int q0 = fv; q0 <= lv; q0 += 32) {
        const int qi = q0 + lane;
        bool is_new = false;
        int nid = -1, links = 0;
        if (qi <= lv) {
            const int anchor = qnode[qi];
            const uint8_t c = q[qi];
            if (anchor < 0) {
                is_new = true;
                qnode[qi] = -1;  // no aligned group to join
            } else if (g.letter[anchor] == c) {
                nid = anchor;
            } else {
                int len = 0;
                for (int x = g.al_head[anchor]; x >= 0; x = g.a_next[x]) {
                    ++len;
                    if (nid < 0 && g.letter[g.a_node[x]] == c) nid = g.a_node[x];
                }
                if (nid < 0) {
                    is_new = true;
                    links = 2 * (len + 1);
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, is_new);
        if (is_new) nid = n0 + n_chain + n_add + __popc(m & ((1u << lane) - 1u));
        if (qi <= lv) path[qi] = nid;
        n_add += __popc(m);
        a_need += links;
    }
    a_need = dc_warp_sum(a_need);
    if (n0 + n_chain + n_add > cap_n || e0 + L > cap_e || a0 + a_need > cap_a) return DC_FAIL_CAP;
    // ---- the two chains
    for (int i = lane; i < fv; i += 32) path[i] = n0 + i;
    for (int i = lv + 1 + lane; i < L; i += 32) path[i] = n0 + fv + (i - lv - 1);
    __syncwarp();
    // ---- new nodes
    for (int i = lane; i < L; i += 32) {
        const int v = path[i];
        if (v >= n0) {
            g.in_head[v] = g.in_tail[v] = g.al_head[v] = g.al_tail[v] = -1;
            g.letter[v] = q[i];
            g.nrec[2 * v] = g.nrec[2 * v + 1] = 0xffffffffu;
        }
    }
    __syncwarp();
    // ---- aligned-group links of the new mismatch nodes (graph.cpp:228-239): the new node copies the anchor's list in
    // order and appends itself to each of its members, then anchor and new node append each other
    int a_at = a0;
    for (int q0 = fv; q0 <= lv; q0 += 32) {
        const int qi = q0 + lane;
        int anchor = -1, nid = -1, links = 0;
        if (qi <= lv) {
            nid = path[qi];
            anchor = qnode[qi];
            if (nid >= n0 && anchor >= 0) {
                int len = 0;
                for (int x = g.al_head[anchor]; x >= 0; x = g.a_next[x]) ++len;
                links = 2 * (len + 1);
            }
        }
        int total;
        int id = a_at + dc_warp_excl_scan(links, lane, total);
        if (links) {
            for (int x = g.al_head[anchor]; x >= 0; x = g.a_next[x]) {
                const int aid = g.a_node[x];
                dc_append_aligned(g, nid, aid, id++);
                dc_append_aligned(g, aid, nid, id++);
            }
            dc_append_aligned(g, nid, anchor, id++);
            dc_append_aligned(g, anchor, nid, id++);
        }
        a_at += total;
    }
    __syncwarp();
    // ---- edges between consecutive nodes of the read's path, unless present (every target is met once)
    int e_at = e0;
    for (int q0 = 1; q0 < L; q0 += 32) {
        const int qi = q0 + lane;
        bool add = false;
        int b = -1, t = -1;
        uint32_t rx = 0xffffffffu;
        if (qi < L) {
            b = path[qi - 1];
            t = path[qi];
            add = true;
            if (t < n0) {
                rx = g.nrec[2 * t];
                if ((rx & 0xffffu) == (uint32_t)b || (rx >> 16) == (uint32_t)b) add = false;
                else if ((rx >> 16) == DC_MORE)
                    for (int x = g.in_head[t]; x >= 0; x = g.e_next_in[x])
                        if (g.e_begin[x] == b) {
                            add = false;
                            break;
                        }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, add);
        if (add) {
            const int id = e_at + __popc(m & ((1u << lane) - 1u));
            g.e_begin[id] = b;
            g.e_next_in[id] = -1;
            if (g.in_tail[t] < 0) g.in_head[t] = id;
            else g.e_next_in[g.in_tail[t]] = id;
            g.in_tail[t] = id;
            g.nrec[2 * t] = dc_rec_push(rx, (uint32_t)b);
        }
        e_at += __popc(m);
    }
    __syncwarp();
    n = n0 + n_chain + n_add;
    e = e_at;
    a = a_at;
    return DC_OK;
}

// Shared-memory working set of one graph's sort and row records
struct DCSort {
    uint16_t *label;   // [cap] block of every node (see dc_sort_blocks)
    uint16_t *rank;    // [cap] node -> rank; before the DFS pass: nodes per block, then first rank of every block
    uint8_t *mc;       // [cap] mark (bits 0-1) | check (bit 2)
    uint32_t *flag;    // spill flags of rows 0..n, one bit each
    unsigned char *stacks;  // DFS stacks: per thread [DC_TP pending | DC_TF frames], or the serial pass' big ones
    uint16_t *deferred;     // [DC_NDEF] roots of blocks the parallel pass could not hold
};
__device__ __forceinline__ DCSort dc_sort_view(unsigned char *base, int cap_n, int nt) {
    const size_t even = ((size_t)cap_n + 3) & ~(size_t)3;
    DCSort s;
    s.label = reinterpret_cast<uint16_t *>(base);
    base += even * 2;
    s.rank = reinterpret_cast<uint16_t *>(base);
    base += even * 2;
    s.flag = reinterpret_cast<uint32_t *>(base);
    base += ((size_t)(cap_n + 1 + 31) / 32 + 1) * 4;
    s.stacks = base;
    base += dc_stack_area(nt);
    s.deferred = reinterpret_cast<uint16_t *>(base);
    base += (size_t)DC_NDEF * 2;
    s.mc = base;
    return s;
}

// 16-bit shared-memory atomics through the containing 32-bit word
__device__ __forceinline__ bool dc_atomic_min16(uint16_t *arr, int idx, uint32_t val) {
    uint32_t *w = reinterpret_cast<uint32_t *>(arr) + (idx >> 1);
    const int sh = (idx & 1) * 16;
    uint32_t old = *reinterpret_cast<volatile uint32_t *>(w);
    while (((old >> sh) & 0xffffu) > val) {
        const uint32_t nw = (old & ~(0xffffu << sh)) | (val << sh);
        const uint32_t got = atomicCAS(w, old, nw);
        if (got == old) return true;
        old = got;
    }
    return false;
}
__device__ __forceinline__ void dc_atomic_inc16(uint16_t *arr, int idx) {
    atomicAdd(reinterpret_cast<uint32_t *>(arr) + (idx >> 1), 1u << ((idx & 1) * 16));
}

// predecessors of node v (in in_edges order) -> cb(index, source node); returns the in-degree
template <typename F>
__device__ __forceinline__ int dc_for_preds(const DGView &g, int v, F &&cb) {
    const uint32_t x = g.nrec[2 * v];
    if ((x >> 16) == DC_MORE) {
        int k = 0;
        for (int e = g.in_head[v]; e >= 0; e = g.e_next_in[e]) cb(k++, g.e_begin[e]);
        return k;
    }
    int k = 0;
    if ((x & 0xffffu) != DC_NONE) cb(k++, (int)(x & 0xffffu));
    if ((x >> 16) != DC_NONE) cb(k++, (int)(x >> 16));
    return k;
}
template <typename F>
__device__ __forceinline__ void dc_for_aligned(const DGView &g, int v, F &&cb) {
    const uint32_t y = g.nrec[2 * v + 1];
    if ((y >> 16) == DC_MORE) {
        for (int x = g.al_head[v]; x >= 0; x = g.a_next[x]) cb(g.a_node[x]);
        return;
    }
    if ((y & 0xffffu) != DC_NONE) cb((int)(y & 0xffffu));
    if ((y >> 16) != DC_NONE) cb((int)(y >> 16));
}

// spoa's topological sort (graph.cpp:293-353) of ONE block (see dc_sort_blocks) by one thread: the reference's
// iterative DFS from `root` — same pushes, same LIFO order, same check_aligned rule as dg_toposort (poa_devgraph.cuh),
// whose structure this follows — over the nodes whose label is `root`; every other node it meets belongs to an earlier
// block and counts as finished.  Emits from rank `emitted` on.  pend/sv/sb: the thread's stacks (capacities pcap /
// fcap; with `big`, entries beyond them go to DGView::pending / stack).  Returns false when a stack is full.
__device__ __forceinline__ bool dc_block_dfs(DGView &g, DCSort &s, int root, int emitted, uint16_t *pend, uint16_t *sv,
                                             uint16_t *sb16, uint32_t *sb32, int pcap, int fcap, bool big) {
    int sp = 0, pn = 0;
    bool ok = true;
    auto done = [&](int x) -> bool { return (int)s.label[x] != root || (s.mc[x] & 3) == 2; };
    auto push = [&](int c) {
        if (pn < pcap) pend[pn] = (uint16_t)c;
        else if (big) g.pending[pn] = c;
        else ok = false;
        ++pn;
    };
    auto emit = [&](int v, int lead) {
        g.lead[emitted] = (uint8_t)lead;
        g.order[emitted] = v;
        s.rank[v] = (uint16_t)emitted;
        ++emitted;
    };
    int v = root;
    while (true) {
        const int begin = pn;
        dc_for_preds(g, v, [&](int, int b) {
            if (!done(b)) push(b);
        });
        if (s.mc[v] & 4)
            dc_for_aligned(g, v, [&](int a) {
                if (!done(a)) {
                    push(a);
                    s.mc[a] &= (uint8_t)~4;
                }
            });
        s.mc[v] = (uint8_t)((s.mc[v] & 4) | 1);
        if (sp < fcap) {
            sv[sp] = (uint16_t)v;
            if (big) sb32[sp] = (uint32_t)begin;
            else sb16[sp] = (uint16_t)begin;
        } else if (big) {
            g.stack[2 * sp] = v;
            g.stack[2 * sp + 1] = begin;
        } else
            ok = false;
        ++sp;
        if (!ok) return false;
        bool descended = false;
        while (sp > 0) {
            const int fv = sp <= fcap ? (int)sv[sp - 1] : g.stack[2 * (sp - 1)];
            const int fb = sp <= fcap ? (big ? (int)sb32[sp - 1] : (int)sb16[sp - 1]) : g.stack[2 * (sp - 1) + 1];
            bool found = false;
            while (pn > fb) {
                --pn;
                const int c = pn < pcap ? (int)pend[pn] : g.pending[pn];
                if (!done(c)) {
                    v = c;
                    found = true;
                    break;
                }
            }
            if (found) {
                descended = true;
                break;
            }
            s.mc[fv] = (uint8_t)((s.mc[fv] & 4) | 2);
            if (s.mc[fv] & 4) {
                emit(fv, 1);
                dc_for_aligned(g, fv, [&](int a) { emit(a, 0); });
            }
            --sp;
        }
        if (!descended) break;
    }
    return true;
}

// spoa's topological sort, restated so that the whole CTA works on it.  The reference runs its DFS from the nodes in
// id order; the DFS from node i finishes exactly the nodes that are ancestors of i — through in-edges and aligned
// links — and are not ancestors of a smaller id.  So every node v belongs to the BLOCK
//     label(v) = smallest id among the nodes that have v in their ancestor closure (v itself included),
// the reference's order is "block after block in id order of their roots", and inside a block it is the DFS from the
// block's root, which only ever pushes nodes of that block.  Labels are a min-propagation against the edges
// (label(p) <- min over successors), block sizes a histogram, block offsets a prefix sum — all parallel; the blocks'
// DFS passes are independent and tiny (a backbone node, its aligned siblings, the bubble hanging off them), one thread
// each.  A block that overflows a thread's small stacks is redone afterwards by one thread with the big stacks.
// Writes order / lead (global) and rank (shared).  All threads call this; it ends with a block barrier.
__device__ __forceinline__ void dc_sort_blocks(DGView &g, DCSort &s, int n) {
    __shared__ int s_changed, s_ndef, s_carry[DC_MAXT / 32];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    for (int v = tid; v < n; v += nt) {
        s.label[v] = (uint16_t)v;
        s.mc[v] = 4;  // mark 0, check 1
    }
    for (int w = tid; w < (n + 1) / 2 + 1; w += nt) reinterpret_cast<uint32_t *>(s.rank)[w] = 0u;
    for (int w = tid; w < (n + 1 + 31) / 32 + 1; w += nt) s.flag[w] = 0u;
    if (tid == 0) s_ndef = 0;
    // ---- labels: every thread sweeps a contiguous range of ids downwards (a read's new nodes have increasing ids along
    // the read, so one sweep carries a label back along a whole run of them), until nothing changes
    const int chunk = (n + nt - 1) / nt;
    const int lo = min(n, tid * chunk), hi = min(n, lo + chunk);
    while (true) {
        __syncthreads();
        if (tid == 0) s_changed = 0;
        __syncthreads();
        bool changed = false;
        for (int v = hi - 1; v >= lo; --v) {
            const uint32_t lv = *reinterpret_cast<volatile uint16_t *>(&s.label[v]);
            dc_for_preds(g, v, [&](int, int b) {
                if (*reinterpret_cast<volatile uint16_t *>(&s.label[b]) > lv) changed |= dc_atomic_min16(s.label, b, lv);
            });
            dc_for_aligned(g, v, [&](int a) {
                if (*reinterpret_cast<volatile uint16_t *>(&s.label[a]) > lv) changed |= dc_atomic_min16(s.label, a, lv);
            });
        }
        if (changed) s_changed = 1;
        __syncthreads();
        if (!s_changed) break;
    }
    // ---- nodes per block, then the first rank of every block (exclusive prefix sum over the ids, in place)
    for (int v = tid; v < n; v += nt) dc_atomic_inc16(s.rank, (int)s.label[v]);
    __syncthreads();
    {
        int sum = 0;
        for (int v = lo; v < hi; ++v) sum += (int)s.rank[v];
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_carry[wid] = incl;
        __syncthreads();
        int base = incl - sum;
        for (int w = 0; w < wid; ++w) base += s_carry[w];
        for (int v = lo; v < hi; ++v) {
            const int c = (int)s.rank[v];
            s.rank[v] = (uint16_t)base;
            base += c;
        }
    }
    __syncthreads();
    // ---- the blocks' DFS passes (a root's rank entry is read before anything of its block is written; the entries of
    // the other nodes of a block are dead: they are not roots)
    {
        uint16_t *pend = reinterpret_cast<uint16_t *>(s.stacks) + (size_t)tid * (DC_TP + 2 * DC_TF);
        uint16_t *sv = pend + DC_TP, *sb = sv + DC_TF;
        for (int i = tid; i < n; i += nt) {
            if ((int)s.label[i] != i) continue;
            const int first = (int)s.rank[i];
            if (!dc_block_dfs(g, s, i, first, pend, sv, sb, nullptr, DC_TP, DC_TF, false)) {
                const int k = atomicAdd(&s_ndef, 1);
                if (k < DC_NDEF) s.deferred[k] = (uint16_t)i;
                s.rank[i] = (uint16_t)first;  // (the pass may have overwritten it)
            }
        }
    }
    __syncthreads();
    const int ndef = s_ndef;
    if (ndef) {
        // blocks that did not fit a thread's stacks: undo what their passes left, then one thread redoes them with the
        // big stacks.  More than DC_NDEF such blocks (never seen): every block is redone.
        const bool all = ndef > DC_NDEF;
        for (int v = tid; v < n; v += nt) {
            bool redo = all;
            if (!all)
                for (int k = 0; k < ndef && !redo; ++k) redo = s.label[v] == s.deferred[k];
            if (redo) s.mc[v] = 4;
        }
        __syncthreads();
        if (tid == 0) {
            const int spc = dc_serial_pcap(nt), ssc = dc_serial_fcap(nt);
            uint16_t *pend = reinterpret_cast<uint16_t *>(s.stacks);
            uint16_t *sv = pend + spc;
            uint32_t *sb = reinterpret_cast<uint32_t *>(sv + ssc);
            if (all) {
                // positions: recount from the labels (rank entries of redone non-roots are gone)
                int at = 0;
                for (int i = 0; i < n; ++i) {
                    if ((int)s.label[i] != i) continue;
                    int c = 0;
                    for (int v = i; v < n; ++v) c += (int)s.label[v] == i;
                    dc_block_dfs(g, s, i, at, pend, sv, nullptr, sb, spc, ssc, true);
                    at += c;
                }
            } else {
                for (int k = 0; k < ndef; ++k) {
                    const int i = (int)s.deferred[k];
                    dc_block_dfs(g, s, i, (int)s.rank[i], pend, sv, nullptr, sb, spc, ssc, true);
                }
            }
        }
        __syncthreads();
    }
}

// One chain step of one pack by the whole CTA: add read `step-1` (cnt pairs of its alignment lie in the pack's aln block,
// written by the traceback of the previous step; read 0 is a plain chain), sort, and — if the pack has a read `step` —
// build the DP kernel's row records for it with ring depth K.  Results in shared memory: r[0] = status, r[1] = nodes,
// r[2] = spilled rows; valid for every thread after the call (it ends with a block barrier).
// dyn = the CTA's dynamic shared memory, laid out for graphs of up to smem_cap_n nodes (dc_sort_smem); larger graphs
// sort through the global-memory routines of poa_devgraph.cuh.
__device__ __forceinline__ void dc_step_cta(DCPack &P, const DCSeq *__restrict__ sq, int step, int cnt, int K, int32_t *pool,
                                            const uint8_t *__restrict__ qcodes, uint32_t *rec, int32_t *preds,
                                            int32_t *spill_rows, const int32_t *aln, int32_t *path, int32_t *qnode,
                                            unsigned char *dyn, int smem_cap_n, int *r, long long *tph) {
    __shared__ int s_cnt[2], s_maxdeg;
    const int tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
    DGView g = dg_view(pool + P.gbase, P.cap_n, P.cap_e, P.cap_a);
    const long long tc0 = dc_clock();
    if (tid < 32) {
        int status = P.status;
        int n = P.n, e = P.e, a = P.a;
        if (status == DC_OK) {
            const DCSeq S = sq[step - 1];
            status = dc_add_alignment(g, n, e, a, P.cap_n, P.cap_e, P.cap_a, qcodes + S.q_off, S.L,
                                      aln + 2 * (size_t)P.aln_off, cnt, qnode + P.qnode_off, path + P.path_off + S.path_rel, lane);
        }
        if (lane == 0) {
            P.n = n;
            P.e = e;
            P.a = a;
            r[0] = status;
            r[1] = n;
            r[2] = 0;
            s_cnt[0] = s_cnt[1] = 0;
            s_maxdeg = 0;
        }
    }
    __syncthreads();
    const long long tc1 = dc_clock();
    tph[0] += tc1 - tc0;  // Graph::add_alignment
    int status = r[0];
    const int n = r[1];
    const bool in_smem = n <= smem_cap_n;
    const bool want_recs = step < P.n_seq;
    int n_spill = 0;
    if (status == DC_OK && in_smem) {
        DCSort s = dc_sort_view(dyn, smem_cap_n, nt);
        dc_sort_blocks(g, s, n);
        tph[1] += dc_clock() - tc1;  // sort
        if (want_recs) {
            int32_t *srows = spill_rows + P.spill_off;
            // rows needed from more than K ranks back get a spill slot (any numbering of the slots is valid)
            for (int rr = 1 + tid; rr <= n; rr += nt) {
                const int v = g.order[rr - 1];
                dc_for_preds(g, v, [&](int, int b) {
                    const int pr = (int)s.rank[b] + 1;
                    if (rr - pr > K) {
                        const uint32_t bit = 1u << (pr & 31);
                        if (!(atomicOr(&s.flag[pr >> 5], bit) & bit)) {
                            const int sl = atomicAdd(&s_cnt[0], 1) + 1;
                            g.slot[pr] = sl;
                            srows[sl] = pr;
                        }
                    }
                });
            }
            __syncthreads();
            // row records (layout: poa_strip_kernel.cuh; same content as dg_build_recs)
            uint32_t *rc = rec + 4 * (size_t)P.rec_off;
            int32_t *pw = preds + P.pred_base;
            if (tid == 0) {
                rc[0] = rc[1] = rc[2] = rc[3] = 0u;
                srows[0] = 0;
            }
            int maxdeg = 0;
            for (int rr = 1 + tid; rr <= n; rr += nt) {
                const int v = g.order[rr - 1];
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                auto word = [&](int b) -> uint32_t {
                    const int pr = (int)s.rank[b] + 1;
                    return (rr - pr <= K) ? (uint32_t)(rr - pr) : (DG_FAR | (uint32_t)g.slot[pr]);
                };
                int np = dc_for_preds(g, v, [&](int k, int b) {
                    if (k < 3) w[1 + k] = word(b);
                });
                maxdeg = max(maxdeg, np);
                if (np == 0) {  // no in-edge: the virtual start row (row 0 = spill slot 0 unless within the ring)
                    w[1] = (rr <= K) ? (uint32_t)rr : DG_FAR;
                    np = 1;
                } else if (np > 3) {
                    const int base = atomicAdd(&s_cnt[1], np);
                    w[3] = (uint32_t)base;
                    dc_for_preds(g, v, [&](int k, int b) { pw[base + k] = (int32_t)word(b); });
                }
                const uint32_t myslot = ((s.flag[rr >> 5] >> (rr & 31)) & 1u) ? (uint32_t)g.slot[rr] : 0u;
                w[0] = (uint32_t)g.letter[v] | ((uint32_t)np << 8) | (myslot << 16);
                reinterpret_cast<uint4 *>(rc)[rr] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (maxdeg > 32) atomicAdd(&s_maxdeg, 1);
            __syncthreads();
            n_spill = s_cnt[0];
            if (s_maxdeg) status = DC_FAIL_DEGREE;  // predecessor indices are 5-bit in the traceback codes
            else if (n_spill > P.spill_cap || n_spill >= 65535) status = DC_FAIL_SPILL;
        }
    } else if (status == DC_OK) {  // graph too large for shared memory
        if (tid < 32) {
            if (lane == 0) dg_toposort(g, n);
            __syncwarp();
            if (want_recs) {
                int32_t *cnt2 = spill_rows + P.spill_off + P.cap_n + 2;  // two counters behind the pack's spill_rows block
                if (lane == 0) cnt2[0] = cnt2[1] = 0;
                dg_ranks(g, n, lane, 32);
                __syncwarp();
                dg_plan_spills(g, n, K, cnt2, spill_rows + P.spill_off, lane, 32);
                __syncwarp();
                const int deg = dg_build_recs(g, n, K, cnt2, rec + 4 * (size_t)P.rec_off, preds + P.pred_base,
                                              spill_rows + P.spill_off, lane, 32);
                __syncwarp();
                int maxdeg = deg;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, d));
                if (lane == 0) {
                    s_cnt[0] = cnt2[0];
                    s_maxdeg = maxdeg > 32 ? 1 : 0;
                }
            }
        }
        __syncthreads();
        if (want_recs) {
            n_spill = s_cnt[0];
            if (s_maxdeg) status = DC_FAIL_DEGREE;
            else if (n_spill > P.spill_cap || n_spill >= 65535) status = DC_FAIL_SPILL;
        }
    }
    __syncthreads();  // every thread has read r[] and s_cnt before they change
    if (tid == 0) {
        P.status = status;
        r[0] = status;
        r[2] = n_spill;
    }
    __syncthreads();
}

// MSA columns (graph.cpp:371-388) by one warp: column of every node = index of its aligned group in rank order.  The
// sort left lead[r] = 1 where rank r opens a group.  col[] reuses the sort's stack array (DGView::stack).
__device__ __forceinline__ void dc_msa_cols_warp(DCPack &P, int32_t *pool) {
    const int lane = threadIdx.x & 31;
    DGView g = dg_view(pool + P.gbase, P.cap_n, P.cap_e, P.cap_a);
    int32_t *col = g.stack;
    const int n = P.n;
    int run = 0;
    for (int r0 = 0; r0 < n; r0 += 32) {
        const int r = r0 + lane;
        const bool lead = r < n && g.lead[r] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, lead);
        if (r < n) col[g.order[r]] = run + __popc(m & ((2u << lane) - 1u)) - 1;
        run += __popc(m);
    }
    if (lane == 0) P.ncol = run;
}

// The whole chain of a pack in ONE CTA (one block per cluster): for every read, graph update -> DP -> traceback, with
// the graph's sort in the CTA's shared memory between two DPs.  `list` = the packs of this launch (all of one CTA
// width), largest first; CTAs fetch packs from it until it is empty.  stats: [0] DP cells, [1] alignments, [2] bytes the
// DP writes by construction.
// MAXT / MINB: launch bounds (threads per CTA, CTAs per SM) — narrower CTAs are compiled for four CTAs per SM, which
// keeps more DP warps resident while other CTAs of the SM are in their one-warp phases (graph update, traceback).
template <int SM, int SN, int SG, int SE, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_poa_chain(DCPack *packs, const int32_t *__restrict__ list, int n_list, const DCSeq *__restrict__ seqs, int32_t *pool,
            const uint8_t *__restrict__ qcodes, uint4 *rec, int32_t *preds, int32_t *spill_rows, int32_t *aln, int32_t *path,
            int32_t *qnode, uint32_t *arena, unsigned long long *stats, unsigned int *counter, int K, int smem_cap_n) {
    PS_DYNAMIC_SHARED(uint4, s_dyn);
    __shared__ int s_item, s_r[3], s_aln_cnt;
    __shared__ int4 s_cell;
    const int tid = threadIdx.x;
    while (true) {
        if (tid == 0) s_item = (int)atomicAdd(counter, 1u);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= n_list) break;
        DCPack &P = packs[list[item]];
        const DCSeq *sq = seqs + P.seq_base;
        const int n_seq = P.n_seq;
        int cnt = 0;
        unsigned long long cells = 0, bytes = 0;
        int n_aln = 0;
        long long t_graph = 0, t_dp = 0, t_tb = 0, t0 = dc_clock();  // phase clocks of this CTA (stats[3..7])
        long long tph[2] = {0, 0};
        for (int step = 1; step <= n_seq; ++step) {
            dc_step_cta(P, sq, step, cnt, K, pool, qcodes, reinterpret_cast<uint32_t *>(rec), preds, spill_rows, aln, path,
                        qnode, reinterpret_cast<unsigned char *>(s_dyn), smem_cap_n, s_r, tph);
            {
                const long long t1 = dc_clock();
                t_graph += t1 - t0;
                t0 = t1;
            }
            if (s_r[0] != DC_OK || step == n_seq) break;
            const DCSeq S = sq[step];
            PoaSJob J;
            J.hf_off = P.hf_off;
            J.code_off = P.code_off;
            J.q_off = S.q_off;
            J.row_off = P.rec_off;
            J.pred_base = P.pred_base;
            J.aln_off = P.aln_off;
            J.spill_off = P.spill_off;
            J.L = S.L;
            J.n = s_r[1];
            J.n_strips = (S.L + PS_STRIP - 1) / PS_STRIP;
            J.n_spill = s_r[2];
            J.pad = 0;
            J.order_off = P.gbase + 4 * (uint64_t)P.cap_n;  // DGView::order
            ps_align_job<SM, SN, SG, SE>(J, qcodes, rec, preds, arena, K, s_dyn, &s_cell);
            {
                const long long t1 = dc_clock();
                t_dp += t1 - t0;
                t0 = t1;
            }
            if (tid < 32) {
                const int c = ps_traceback_warp(J, s_cell, rec, preds, spill_rows, arena, pool, aln);
                if (tid == 0) s_aln_cnt = c;
            }
            __syncthreads();
            {
                const long long t1 = dc_clock();
                t_tb += t1 - t0;
                t0 = t1;
            }
            cnt = s_aln_cnt;
            cells += (unsigned long long)J.L * (unsigned long long)J.n;
            bytes += (unsigned long long)ps_code_words(J.n, J.n_strips) * 4ull + (unsigned long long)(J.n_spill + 1) * J.n_strips * 1028ull;
            ++n_aln;
        }
        if (tid < 32 && s_r[0] == DC_OK) dc_msa_cols_warp(P, pool);
        if (tid == 0) {
            atomicAdd(&stats[0], cells);
            atomicAdd(&stats[1], (unsigned long long)n_aln);
            atomicAdd(&stats[2], bytes);
            atomicAdd(&stats[3], (unsigned long long)t_graph);
            atomicAdd(&stats[4], (unsigned long long)t_dp);
            atomicAdd(&stats[5], (unsigned long long)t_tb);
            atomicAdd(&stats[6], (unsigned long long)tph[0]);
            atomicAdd(&stats[7], (unsigned long long)tph[1]);
        }
        __syncthreads();
    }
}

// MSA rows (graph.cpp:390-426 without the consensus row): row s = '-' everywhere except letter(v) at col[v] for the
// nodes v of read s.  One CTA per pack; rows go to out + msa_off[pack], n_seq x ncol chars.
__global__ void __launch_bounds__(256) k_chain_msa_rows(const DCPack *__restrict__ packs, const DCSeq *__restrict__ seqs,
                                                        const uint64_t *__restrict__ msa_off, int32_t *pool,
                                                        const int32_t *__restrict__ path, char *out) {
    const DCPack &P = packs[blockIdx.x];
    if (P.status != DC_OK) return;
    DGView g = dg_view(pool + P.gbase, P.cap_n, P.cap_e, P.cap_a);
    const int32_t *col = g.stack;
    char *rows = out + msa_off[blockIdx.x];
    const size_t total = (size_t)P.n_seq * P.ncol;
    for (size_t i = threadIdx.x; i < total; i += blockDim.x) rows[i] = '-';
    __syncthreads();
    for (int s = 0; s < P.n_seq; ++s) {
        const DCSeq S = seqs[P.seq_base + s];
        const int32_t *p = path + P.path_off + S.path_rel;
        char *row = rows + (size_t)s * P.ncol;
        for (int i = threadIdx.x; i < S.L; i += blockDim.x) {
            const int v = p[i];
            row[col[v]] = "ACGTU"[g.letter[v]];
        }
    }
}

#endif  // __CUDACC__ || CUDA_EMU

}  // namespace rtl
