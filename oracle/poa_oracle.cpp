// TEST INFRASTRUCTURE ONLY — see rattle_oracle.h.  CPU restatement ("oracle") of hot path B: spoa's local (kSW)
// affine-gap sequence-to-graph alignment, the partial-order graph update / topological order / MSA, and RATTLE's
// correct_reads on top of them.  Plain scalar code with full H/F/E matrices, written from the behaviour of
//   spoa/src/sisd_alignment_engine.cpp:94-200,437-657   (scalar engine; byte-identical to the AVX2 engine RATTLE
//                                                         links — SURVEY.md §6 — and checked against it in tests)
//   spoa/src/graph.cpp:99-115,154-353,371-426
//   correct.cpp:32-563, utils.cpp:6-24, fasta.cpp:458-464
// Pinned in tests/test_oracle_poa.py against oracle/_ref/libref_shim.so (the unmodified reference) and against the
// golden vectors under tests/golden/ generated from it.
#include <math.h>

#include <algorithm>
#include <climits>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "rattle_oracle.h"

namespace {

typedef std::vector<std::pair<int, int>> alignment_t;
int64_t g_cells = 0;

struct Node {
    char c;
    std::vector<int> in, out, aligned;  // in/out hold node ids in edge-creation order
};

struct Graph {
    std::vector<Node> nodes;
    std::vector<int> order;                // rank -> node
    std::vector<std::vector<int>> walks;   // node path of every sequence

    int new_node(char c) {
        nodes.push_back(Node{c, {}, {}, {}});
        return (int)nodes.size() - 1;
    }
    void link(int a, int b) {  // graph.cpp:99-115
        for (int e : nodes[a].out)
            if (e == b) return;
        nodes[a].out.push_back(b);
        nodes[b].in.push_back(a);
    }
    int chain(const std::string &s, int b, int e, std::vector<int> &walk) {  // graph.cpp:273-291
        if (b == e) return -1;
        int first = new_node(s[b]);
        walk.push_back(first);
        for (int i = b + 1; i < e; ++i) {
            int id = new_node(s[i]);
            link(id - 1, id);
            walk.push_back(id);
        }
        return first;
    }
    void add(const alignment_t &al, const std::string &s) {  // graph.cpp:154-271
        if (s.empty()) return;
        std::vector<int> walk;
        if (al.empty()) {
            chain(s, 0, (int)s.size(), walk);
            walks.push_back(walk);
            sort();
            return;
        }
        std::vector<int> valid;
        for (auto &p : al)
            if (p.second != -1) valid.push_back(p.second);
        size_t before = nodes.size();
        chain(s, 0, valid.front(), walk);
        int head = before == nodes.size() ? -1 : (int)nodes.size() - 1;
        std::vector<int> tail_walk;
        int tail = chain(s, valid.back() + 1, (int)s.size(), tail_walk);
        for (auto &p : al) {
            if (p.second == -1) continue;
            char c = s[p.second];
            int id = -1;
            if (p.first == -1) id = new_node(c);
            else if (nodes[p.first].c == c) id = p.first;
            else {
                for (int a : nodes[p.first].aligned)
                    if (nodes[a].c == c) {
                        id = a;
                        break;
                    }
                if (id == -1) {
                    id = new_node(c);
                    std::vector<int> grp = nodes[p.first].aligned;
                    for (int a : grp) {
                        nodes[id].aligned.push_back(a);
                        nodes[a].aligned.push_back(id);
                    }
                    nodes[id].aligned.push_back(p.first);
                    nodes[p.first].aligned.push_back(id);
                }
            }
            if (head != -1) link(head, id);
            head = id;
            walk.push_back(id);
        }
        if (tail != -1) link(head, tail);
        walk.insert(walk.end(), tail_walk.begin(), tail_walk.end());
        walks.push_back(walk);
        sort();
    }
    void sort() {  // graph.cpp:293-353
        order.clear();
        std::vector<char> mark(nodes.size(), 0), chk(nodes.size(), 1);
        std::vector<int> st;
        for (size_t i = 0; i < nodes.size(); ++i) {
            if (mark[i]) continue;
            st.push_back((int)i);
            while (!st.empty()) {
                int v = st.back();
                bool ok = true;
                if (mark[v] != 2) {
                    for (int b : nodes[v].in)
                        if (mark[b] != 2) {
                            st.push_back(b);
                            ok = false;
                        }
                    if (chk[v])
                        for (int a : nodes[v].aligned)
                            if (mark[a] != 2) {
                                st.push_back(a);
                                chk[a] = 0;
                                ok = false;
                            }
                    if (ok) {
                        mark[v] = 2;
                        if (chk[v]) {
                            order.push_back(v);
                            for (int a : nodes[v].aligned) order.push_back(a);
                        }
                    } else
                        mark[v] = 1;
                }
                if (ok) st.pop_back();
            }
        }
    }
    std::vector<std::string> msa() const {  // graph.cpp:371-426
        std::vector<int> col(nodes.size(), 0);
        int nc = 0;
        for (size_t i = 0; i < order.size(); ++i) {
            int v = order[i];
            col[v] = nc;
            for (size_t j = 0; j < nodes[v].aligned.size(); ++j) col[order[++i]] = nc;
            ++nc;
        }
        std::vector<std::string> rows;
        for (auto &w : walks) {
            std::string r(nc, '-');
            for (int v : w) r[col[v]] = nodes[v].c;
            rows.push_back(r);
        }
        return rows;
    }
};

const int NEG = INT_MIN / 2;

// sisd_alignment_engine.cpp:437-657 for (kSW, affine)
alignment_t align(const std::string &q, const Graph &G, int m, int n, int g, int e) {
    alignment_t out;
    if (G.nodes.empty() || q.empty()) return out;
    const int W = (int)q.size() + 1, Hh = (int)G.nodes.size() + 1;
    g_cells += (int64_t)q.size() * (int64_t)G.nodes.size();
    std::vector<int> H((size_t)W * Hh, 0), F((size_t)W * Hh, 0), E((size_t)W * Hh, 0), rank(G.nodes.size());
    for (size_t r = 0; r < G.order.size(); ++r) rank[G.order[r]] = (int)r;
    for (int j = 1; j < W; ++j) {
        F[j] = NEG;
        E[j] = g + (j - 1) * e;
    }
    for (int i = 1; i < Hh; ++i) E[(size_t)i * W] = NEG;
    int best = 0, bi = -1, bj = -1;
    auto preds = [&](int node, std::vector<int> &rows) {
        rows.clear();
        for (int b : G.nodes[node].in) rows.push_back(rank[b] + 1);
        if (rows.empty()) rows.push_back(0);
    };
    std::vector<int> pr;
    for (int r = 0; r < (int)G.order.size(); ++r) {
        const int node = G.order[r], i = r + 1;
        preds(node, pr);
        int *Hr = &H[(size_t)i * W], *Fr = &F[(size_t)i * W], *Er = &E[(size_t)i * W];
        for (int j = 1; j < W; ++j) {
            int f = NEG, h = NEG;
            const int s = G.nodes[node].c == q[j - 1] ? m : n;
            for (int p : pr) {
                f = std::max(f, std::max(H[(size_t)p * W + j] + g, F[(size_t)p * W + j] + e));
                h = std::max(h, H[(size_t)p * W + j - 1] + s);
            }
            Fr[j] = f;
            Hr[j] = h;
        }
        for (int j = 1; j < W; ++j) {
            Er[j] = std::max(Hr[j - 1] + g, Er[j - 1] + e);
            Hr[j] = std::max(std::max(Hr[j], std::max(Fr[j], Er[j])), 0);
            if (best < Hr[j]) {
                best = Hr[j];
                bi = i;
                bj = j;
            }
        }
    }
    if (bi == -1) return out;
    int i = bi, j = bj;
    while (H[(size_t)i * W + j] != 0) {
        const int h = H[(size_t)i * W + j];
        int pi = 0, pj = 0;
        bool found = false, left = false, up = false;
        const int node = G.order[i - 1];
        preds(node, pr);
        if (i != 0 && j != 0) {
            const int s = G.nodes[node].c == q[j - 1] ? m : n;
            for (int p : pr)
                if (h == H[(size_t)p * W + j - 1] + s) {
                    pi = p;
                    pj = j - 1;
                    found = true;
                    break;
                }
        }
        if (!found && i != 0)
            for (int p : pr)
                if ((up = h == F[(size_t)p * W + j] + e) || h == H[(size_t)p * W + j] + g) {
                    pi = p;
                    pj = j;
                    found = true;
                    break;
                }
        if (!found && j != 0)
            if ((left = h == E[(size_t)i * W + j - 1] + e) || h == H[(size_t)i * W + j - 1] + g) {
                pi = i;
                pj = j - 1;
                found = true;
            }
        out.emplace_back(i == pi ? -1 : node, j == pj ? -1 : j - 1);
        i = pi;
        j = pj;
        if (left) {
            while (true) {
                out.emplace_back(-1, j - 1);
                --j;
                if (E[(size_t)i * W + j] + e != E[(size_t)i * W + j + 1]) break;
            }
        } else if (up) {
            while (true) {
                bool stop = false;
                int nx = 0;
                for (int b : G.nodes[G.order[i - 1]].in) {
                    const int p = rank[b] + 1;
                    if ((stop = F[(size_t)i * W + j] == H[(size_t)p * W + j] + g) ||
                        F[(size_t)i * W + j] == F[(size_t)p * W + j] + e) {
                        nx = p;
                        break;
                    }
                }
                out.emplace_back(G.order[i - 1], -1);
                i = nx;
                if (stop || i == 0) break;
            }
        }
    }
    std::reverse(out.begin(), out.end());
    return out;
}

struct Rd {
    std::string header, seq, ann, qual;
};

char phred_symbol(double p) { return (char)(-10 * log10(p) + 33); }  // utils.cpp:6-8
double phred_err(char c) {                                            // utils.cpp:10-13
    double q = c - 33;
    return pow(10.0, -q / 10.0);
}
std::string revcomp(const std::string &s) {  // utils.cpp:15-24
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[s.size() - 1 - i];
        r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';  // T and U -> A
    }
    return r;
}

void trim_ends(std::vector<Rd> &reads, std::vector<std::string> &aln) {  // correct.cpp:32-92
    for (size_t i = 0; i < aln.size(); ++i) {
        for (int pass = 0; pass < 2; ++pass) {
            std::string &r = aln[i];
            size_t pos = 0;
            bool flipped = false;
            while (pos < r.size()) {
                while (pos < r.size() && r[pos] == '-') ++pos;
                size_t end = pos;
                int gaps = 0, sz = 0;
                while (gaps < 4 && end < r.size()) {
                    if (r[end] == '-') ++gaps;
                    else {
                        ++sz;
                        gaps = 0;
                    }
                    ++end;
                }
                if (sz < 10) {
                    while (end < r.size() && r[end] == '-') {
                        ++end;
                        ++gaps;
                    }
                    if (gaps >= 20) {
                        for (size_t j = pos; j < end; ++j) r[j] = '-';
                        reads[i].qual.erase(0, sz);
                        reads[i].seq.erase(0, sz);
                        pos = end;
                        continue;
                    }
                }
                std::reverse(r.begin(), r.end());
                std::reverse(reads[i].qual.begin(), reads[i].qual.end());
                std::reverse(reads[i].seq.begin(), reads[i].seq.end());
                flipped = true;
                break;
            }
            if (!flipped) break;  // row exhausted without a reversal: the reference stops here too
        }
    }
}

// symbols in the iteration order of the reference's unordered_map<char,pos_info_t>: U - G T C A
const char ORDER[6] = {'U', '-', 'G', 'T', 'C', 'A'};
int slot(char c) {
    for (int i = 0; i < 6; ++i)
        if (ORDER[i] == c) return i;
    return -1;
}
struct Cols {
    std::vector<int> occ, tot;
    std::vector<double> err;
    std::string cons;
};
Cols vote(const std::vector<Rd> &reads, const std::vector<std::string> &aln) {  // correct.cpp:94-193
    Cols C;
    if (reads.empty() || aln.empty()) return C;
    size_t nc = aln[0].size();
    C.occ.assign(nc * 6, 0);
    C.tot.assign(nc * 6, 0);
    C.err.assign(nc * 6, 0.0);
    for (size_t i = 0; i < reads.size(); ++i) {
        int sp = -1;
        for (size_t k = 0; k < aln[i].size(); ++k) {
            char nt = aln[i][k];
            double ep = 0.0;
            if (nt != '-') {
                ++sp;
                ep = phred_err(reads[i].qual[sp]);
            }
            if (sp >= 0 && sp < (int)reads[i].qual.size()) {
                int s = slot(nt);
                if (s < 0) continue;
                C.occ[k * 6 + s]++;
                C.err[k * 6 + s] += ep;
                if (sp == (int)reads[i].qual.size() - 1) ++sp;
            }
        }
    }
    C.cons.assign(nc, '-');
    for (size_t k = 0; k < nc; ++k) {
        int all = 0;
        for (int s = 0; s < 6; ++s) all += C.occ[k * 6 + s];
        int mo = 0;
        char mc = 0;
        for (int s = 0; s < 6; ++s) {
            if (C.occ[k * 6 + s] > 0) {
                C.tot[k * 6 + s] += all;
                C.err[k * 6 + s] /= double(C.occ[k * 6 + s]);
            }
            if (C.occ[k * 6 + s] > mo) {
                mo = C.occ[k * 6 + s];
                mc = ORDER[s];
            }
        }
        C.cons[k] = mc ? mc : '-';
    }
    return C;
}
std::string degap(const std::string &s) {
    std::string r;
    for (char c : s)
        if (c != '-') r += c;
    return r;
}

void fix_pack(const std::vector<Rd> &reads, const std::vector<std::string> &aln, double min_occ, double gap_occ,
              std::vector<Rd> &good, std::vector<Rd> &bad) {  // correct.cpp:196-309
    Cols C = vote(reads, aln);
    for (size_t i = 0; i < reads.size(); ++i) {
        int sp = -1;
        std::string rs, rq;
        for (size_t k = 0; k < aln[i].size(); ++k) {
            char nt = aln[i][k];
            double ep = 0.0;
            if (nt != '-') {
                ++sp;
                ep = phred_err(reads[i].qual[sp]);
            }
            if (!(sp >= 0 && sp < (int)reads[i].qual.size())) continue;
            char cn = C.cons[k];
            int ci = slot(cn);
            double ratio = double(C.occ[k * 6 + ci]) / double(C.tot[k * 6 + ci]);
            double cerr = C.err[k * 6 + ci];
            if (cn == '-') {
                if (nt != '-' && !(ratio >= gap_occ)) {
                    rs += nt;
                    rq += reads[i].qual[sp];
                }
            } else if (nt == '-') {
                if (ratio >= gap_occ) {
                    rs += cn;
                    rq += phred_symbol(cerr);
                }
            } else if (nt == cn || !(ratio >= min_occ && 30.0 * ep > cerr)) {
                rs += nt;
                rq += reads[i].qual[sp];
            } else {
                rs += cn;
                rq += phred_symbol(cerr);
            }
            if (sp == (int)reads[i].qual.size() - 1) ++sp;
        }
        if (!rs.empty()) good.push_back(Rd{reads[i].header, rs, "+", rq});
        else bad.push_back(reads[i]);
    }
}

std::vector<std::string> poa_rows(const std::vector<Rd> &rs, std::vector<alignment_t> *alns = nullptr) {
    Graph G;
    for (auto &r : rs) {
        alignment_t a = align(r.seq, G, 5, -4, -8, -6);
        if (alns) alns->push_back(a);
        G.add(a, r.seq);
    }
    return G.msa();
}

void dump(const std::vector<Rd> &rs, std::string &o) {
    for (auto &r : rs) o += r.header + "\n" + r.seq + "\n" + r.ann + "\n" + r.qual + "\n";
}

}  // namespace

extern "C" {

int orc_poa_msa(const char *bases, const uint64_t *offsets, uint32_t n, int m, int nn, int g, int e, char *msa_out,
                int64_t cap, int *msa_cols, int64_t *aln_off, int32_t *aln_pairs, int64_t aln_cap) {
    g_cells = 0;
    Graph G;
    int64_t ao = 0;
    for (uint32_t i = 0; i < n; ++i) {
        std::string s(bases + offsets[i], bases + offsets[i + 1]);
        alignment_t a = align(s, G, m, nn, g, e);
        if (aln_off) {
            aln_off[i] = ao;
            for (auto &p : a) {
                if (ao + 1 <= aln_cap / 2) {
                    aln_pairs[2 * ao] = p.first;
                    aln_pairs[2 * ao + 1] = p.second;
                }
                ++ao;
            }
        }
        G.add(a, s);
    }
    if (aln_off) aln_off[n] = ao;
    std::vector<std::string> rows = G.msa();
    *msa_cols = rows.empty() ? 0 : (int)rows[0].size();
    if ((int64_t)rows.size() * (*msa_cols) > cap) return -1;
    for (size_t i = 0; i < rows.size(); ++i) memcpy(msa_out + i * (size_t)(*msa_cols), rows[i].data(), *msa_cols);
    return (int)rows.size();
}

int orc_correct_reads(const char *bases, const char *quals, const uint64_t *offsets, uint32_t n_reads,
                      const int32_t *main_id, const uint8_t *main_rev, const int32_t *main_gene, const int64_t *cl_off,
                      const int32_t *mem_id, const uint8_t *mem_rev, const int32_t *mem_gene, int n_clusters,
                      double min_occ, double gap_occ, double err_ratio, int split, int min_reads, char *corrected,
                      int64_t *corrected_len, char *uncorrected, int64_t *uncorrected_len, char *consensi,
                      int64_t *consensi_len) {
    (void)main_id; (void)main_rev; (void)mem_gene; (void)err_ratio;
    g_cells = 0;
    std::vector<Rd> reads(n_reads);
    for (uint32_t i = 0; i < n_reads; ++i) {
        reads[i].header = "@r" + std::to_string(i);
        reads[i].seq.assign(bases + offsets[i], bases + offsets[i + 1]);
        reads[i].ann = "+";
        reads[i].qual.assign(quals + offsets[i], quals + offsets[i + 1]);
    }
    const bool gene_mode = main_gene[0] == -1;
    std::vector<Rd> out_bad, out_good, out_cons;
    struct PackT { int cid; std::vector<Rd> rs; };
    std::vector<PackT> queue;
    for (int cid = 0; cid < n_clusters; ++cid) {  // correct.cpp:328-370
        const size_t n = (size_t)(cl_off[cid + 1] - cl_off[cid]);
        const int files = (int)((n - 1) / split + 1);
        const int gid = main_gene[cid];
        for (int f = 0; f < files; ++f) {
            std::vector<Rd> rs;
            for (size_t j = f; j < n; j += files) {
                const int id = mem_id[cl_off[cid] + j];
                if (mem_rev[cl_off[cid] + j]) {
                    reads[id].seq = revcomp(reads[id].seq);
                    std::reverse(reads[id].qual.begin(), reads[id].qual.end());
                }
                reads[id].header += gid == -1 ? ",gene_cluster_" + std::to_string(cid)
                                              : ",gene_cluster_" + std::to_string(gid) + ",transcript_cluster_" + std::to_string(cid);
                rs.push_back(reads[id]);
            }
            if ((int)rs.size() > min_reads) queue.push_back(PackT{cid, rs});
            else out_bad.insert(out_bad.end(), rs.begin(), rs.end());
        }
    }
    std::vector<std::vector<Rd>> per_cluster(n_clusters);
    for (auto &pk : queue) {  // correct.cpp:379-471, one worker
        std::vector<Rd> rs = pk.rs;
        std::vector<std::string> aln = poa_rows(rs);
        trim_ends(rs, aln);
        std::vector<Rd> good, bad;
        fix_pack(rs, aln, min_occ, gap_occ, good, bad);
        out_good.insert(out_good.end(), good.begin(), good.end());
        out_bad.insert(out_bad.end(), bad.begin(), bad.end());
        std::stable_sort(good.begin(), good.end(), [](const Rd &a, const Rd &b) { return a.seq.size() > b.seq.size(); });
        aln = poa_rows(good);
        trim_ends(good, aln);
        std::string cons = degap(vote(good, aln).cons);
        std::string gid;
        for (auto &r : rs) gid = std::to_string(std::stoi(r.header.substr(r.header.find("gene_cluster") + 13)));
        per_cluster[pk.cid].push_back(Rd{gid + "," + std::to_string(rs.size()) + ",", cons, "+", std::string(cons.size(), 'K')});
    }
    for (int cid = 0; cid < n_clusters; ++cid) {  // correct.cpp:488-556
        auto &it = per_cluster[cid];
        int total = 0, gid = 0;
        for (auto &r : it) {
            std::stringstream ss(r.header);
            std::string a, b;
            getline(ss, a, ',');
            getline(ss, b, ',');
            gid = std::stoi(a);
            total += std::stoi(b);
        }
        std::string head = gene_mode ? "@gene_cluster_" + std::to_string(cid)
                                     : "@transcript_cluster_" + std::to_string(cid) + " gene_cluster_" + std::to_string(gid);
        head += " reads=" + std::to_string(total) + " labels=";
        if (it.size() > 1) {
            std::vector<std::string> aln = poa_rows(it);
            trim_ends(it, aln);
            std::string cons = degap(vote(it, aln).cons);
            out_cons.push_back(Rd{head, cons, "+", std::string(cons.size(), 'K')});
        } else if (it.size() == 1)
            out_cons.push_back(Rd{head, it[0].seq, "+", it[0].qual});
    }
    std::string a, b, c;
    dump(out_good, a);
    dump(out_bad, b);
    dump(out_cons, c);
    int rc = 0;
    auto put = [&rc](const std::string &s, char *buf, int64_t *len) {
        if ((int64_t)s.size() > *len) rc = -1;
        else memcpy(buf, s.data(), s.size());
        *len = (int64_t)s.size();
    };
    put(a, corrected, corrected_len);
    put(b, uncorrected, uncorrected_len);
    put(c, consensi, consensi_len);
    return rc;
}

int64_t orc_poa_cells(void) { return g_cells; }

}  // extern "C"
