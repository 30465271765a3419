// Host-side partial-order graph of hot path B, flat-array restatement of the parts of spoa::Graph that
// correct.cpp uses (spoa/src/graph.cpp):
//   add_alignment       graph.cpp:154-271   (+ add_sequence :273-291, add_edge :99-115)
//   topological_sort    graph.cpp:293-353   (row order of the DP; aligned nodes are emitted as one group)
//   MSA                 graph.cpp:371-426
// Edge weights and sequence labels only feed spoa's heaviest-bundle consensus, which RATTLE never calls
// (it derives its consensus from the MSA columns), so they are not kept: each sequence remembers its node path
// instead, which is what Node::successor(label) walks (graph.cpp:30-41).
#pragma once
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

namespace rtl {

struct PoaGraph {
    // nodes
    std::vector<char> letter;
    std::vector<int32_t> in_head, in_tail, out_head, out_tail, al_head, al_tail;
    std::vector<int32_t> n_in, n_al;
    // edges (append order = creation order; per-node lists keep that order)
    std::vector<int32_t> e_begin, e_end, e_next_in, e_next_out;
    // aligned-node lists (a_owner[x] = the node whose list entry x belongs to: the append log the device mirror replays)
    std::vector<int32_t> a_node, a_next, a_owner;
    // sequences
    std::vector<std::vector<int32_t>> paths;
    // topological order
    std::vector<int32_t> rank_to_node, node_to_rank;
    // scratch for the sort
    std::vector<uint8_t> mark, check;
    std::vector<int32_t> stack;
    int max_in_degree = 0;
    // The rank order is kept on the GPU for graphs that have a device mirror (poa_devgraph.cuh): add_alignment then
    // skips the sort, and msa() sorts once at the end.
    bool defer_sort = false, sorted = true;

    int n_nodes() const { return (int)letter.size(); }

    void clear() {
        letter.clear(); in_head.clear(); in_tail.clear(); out_head.clear(); out_tail.clear(); al_head.clear();
        al_tail.clear(); n_in.clear(); n_al.clear(); e_begin.clear(); e_end.clear(); e_next_in.clear();
        e_next_out.clear(); a_node.clear(); a_next.clear(); a_owner.clear(); paths.clear(); rank_to_node.clear(); node_to_rank.clear();
        max_in_degree = 0;
        defer_sort = false;
        sorted = true;
    }

    int add_node(char c) {
        letter.push_back(c);
        in_head.push_back(-1); in_tail.push_back(-1); out_head.push_back(-1); out_tail.push_back(-1);
        al_head.push_back(-1); al_tail.push_back(-1); n_in.push_back(0); n_al.push_back(0);
        return (int)letter.size() - 1;
    }

    // graph.cpp:99-115 (existing edge: only labels/weights change, which we do not keep)
    void add_edge(int b, int e) {
        for (int x = out_head[b]; x >= 0; x = e_next_out[x])
            if (e_end[x] == e) return;
        const int id = (int)e_begin.size();
        e_begin.push_back(b); e_end.push_back(e); e_next_in.push_back(-1); e_next_out.push_back(-1);
        if (out_tail[b] < 0) out_head[b] = id; else e_next_out[out_tail[b]] = id;
        out_tail[b] = id;
        if (in_tail[e] < 0) in_head[e] = id; else e_next_in[in_tail[e]] = id;
        in_tail[e] = id;
        if (++n_in[e] > max_in_degree) max_in_degree = n_in[e];
    }

    void add_aligned(int node, int other) {
        const int id = (int)a_node.size();
        a_node.push_back(other); a_next.push_back(-1); a_owner.push_back(node);
        if (al_tail[node] < 0) al_head[node] = id; else a_next[al_tail[node]] = id;
        al_tail[node] = id;
        ++n_al[node];
    }

    // graph.cpp:273-291: linear chain for seq[begin,end); returns first node or -1
    int add_chain(const char *seq, int begin, int end, std::vector<int32_t> &path) {
        if (begin == end) return -1;
        const int first = add_node(seq[begin]);
        path.push_back(first);
        for (int i = begin + 1; i < end; ++i) {
            const int id = add_node(seq[i]);
            add_edge(id - 1, id);
            path.push_back(id);
        }
        return first;
    }

    // graph.cpp:154-271.  aln = (node id or -1, query pos or -1) pairs, start-to-end.
    void add_alignment(const std::vector<std::pair<int32_t, int32_t>> &aln, const char *seq, int len) {
        if (len == 0) return;
        std::vector<int32_t> path;
        path.reserve(len);
        if (aln.empty()) {
            add_chain(seq, 0, len, path);
            paths.push_back(std::move(path));
            after_change();
            return;
        }
        int first_valid = -1, last_valid = -1;
        for (const auto &it : aln)
            if (it.second != -1) {
                if (first_valid < 0) first_valid = it.second;
                last_valid = it.second;
            }
        const int before = n_nodes();
        add_chain(seq, 0, first_valid, path);
        int head = before == n_nodes() ? -1 : n_nodes() - 1;
        std::vector<int32_t> tail_path;
        const int tail = add_chain(seq, last_valid + 1, len, tail_path);  // created before the aligned part (:199-200)
        for (const auto &it : aln) {
            if (it.second == -1) continue;
            const char c = seq[it.second];
            int nid;
            if (it.first == -1) {
                nid = add_node(c);
            } else if (letter[it.first] == c) {
                nid = it.first;
            } else {
                int found = -1;
                for (int x = al_head[it.first]; x >= 0; x = a_next[x])
                    if (letter[a_node[x]] == c) {
                        found = a_node[x];
                        break;
                    }
                if (found < 0) {
                    nid = add_node(c);
                    // snapshot: the anchor's list must be walked as it was before the new node is appended to anyone
                    for (int x = al_head[it.first]; x >= 0; x = a_next[x]) {
                        const int aid = a_node[x];
                        add_aligned(nid, aid);
                        add_aligned(aid, nid);
                    }
                    add_aligned(nid, it.first);
                    add_aligned(it.first, nid);
                } else
                    nid = found;
            }
            if (head != -1) add_edge(head, nid);
            head = nid;
            path.push_back(nid);
        }
        if (tail != -1) add_edge(head, tail);
        path.insert(path.end(), tail_path.begin(), tail_path.end());
        paths.push_back(std::move(path));
        after_change();
    }

    void after_change() {
        if (defer_sort) sorted = false;
        else topological_sort();
    }

    // graph.cpp:293-353
    void topological_sort() {
        sorted = true;
        const int n = n_nodes();
        rank_to_node.clear();
        rank_to_node.reserve(n);
        mark.assign(n, 0);
        check.assign(n, 1);
        stack.clear();
        for (int i = 0; i < n; ++i) {
            if (mark[i] != 0) continue;
            stack.push_back(i);
            while (!stack.empty()) {
                const int v = stack.back();
                bool valid = true;
                if (mark[v] != 2) {
                    for (int x = in_head[v]; x >= 0; x = e_next_in[x]) {
                        const int b = e_begin[x];
                        if (mark[b] != 2) {
                            stack.push_back(b);
                            valid = false;
                        }
                    }
                    if (check[v]) {
                        for (int x = al_head[v]; x >= 0; x = a_next[x]) {
                            const int a = a_node[x];
                            if (mark[a] != 2) {
                                stack.push_back(a);
                                check[a] = 0;
                                valid = false;
                            }
                        }
                    }
                    if (valid) {
                        mark[v] = 2;
                        if (check[v]) {
                            rank_to_node.push_back(v);
                            for (int x = al_head[v]; x >= 0; x = a_next[x]) rank_to_node.push_back(a_node[x]);
                        }
                    } else
                        mark[v] = 1;
                }
                if (valid) stack.pop_back();
            }
        }
        node_to_rank.assign(n, 0);
        for (int r = 0; r < n; ++r) node_to_rank[rank_to_node[r]] = r;
    }

    // graph.cpp:371-426 (without the consensus row)
    void msa(std::vector<std::string> &dst) {
        if (!sorted) topological_sort();
        const int n = n_nodes();
        std::vector<int32_t> col(n, 0);
        int ncol = 0;
        for (int i = 0; i < n; ++i) {
            const int v = rank_to_node[i];
            col[v] = ncol;
            for (int j = 0; j < n_al[v]; ++j) col[rank_to_node[++i]] = ncol;
            ++ncol;
        }
        for (const auto &p : paths) {
            std::string row((size_t)ncol, '-');
            for (int v : p) row[col[v]] = letter[v];
            dst.push_back(std::move(row));
        }
    }
};

}  // namespace rtl
