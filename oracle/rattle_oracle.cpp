// TEST INFRASTRUCTURE ONLY — see rattle_oracle.h.  CPU restatement of hot path A (greedy k-mer/bitvector
// clustering).  Written from the behaviour of the reference, flat-array style; every function cites the
// reference lines it follows.  Checked against the reference itself in tests/test_oracle_cluster.py.
#include "rattle_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// kmer.hpp:25-31 (A=0,C=1,T=U=2,G=3); anything else -> -1 (the reference dereferences end(), i.e. UB)
inline int base_code(char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'T': case 'U': return 2;
        case 'G': return 3;
        default: return -1;
    }
}
// utils.hpp:7-13: complement in code space.  A<->T/U, C<->G  =>  0<->2, 1<->3
inline int comp_code(int c) { return c ^ 2; }

struct ReadKmers {
    int len = 0, n = 0;  // n = len-k list entries
    std::vector<uint32_t> fh, rh;
    std::vector<int32_t> fp, rp;
    uint64_t bvf[ORC_BV_WORDS], bvr[ORC_BV_WORDS];
    int pcf = 0;  // popcount of forward bitvector
};

// kmer.cpp:6-42: (hash,pos) for pos in [0,len-k) and 6-mer bits for pos in [0,len-6), both strands
void hashes_of(const std::vector<int> &codes, int k, std::vector<uint32_t> &h, std::vector<int32_t> &p, uint64_t *bv) {
    const int len = (int)codes.size();
    const int n = len - k;
    const uint32_t kmask = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    std::vector<std::pair<uint32_t, int32_t>> v(n);
    uint32_t roll = 0, roll6 = 0;
    for (int i = 0; i < len; ++i) {
        roll = ((roll << 2) | (uint32_t)codes[i]) & kmask;
        roll6 = ((roll6 << 2) | (uint32_t)codes[i]) & 0xFFFu;
        int pk = i - k + 1;  // start position of the k-mer ending at i
        if (pk >= 0 && pk < n) v[pk] = std::make_pair(roll, pk);
        int p6 = i - 5;
        if (p6 >= 0 && p6 < len - 6) bv[roll6 >> 6] |= (uint64_t)1 << (roll6 & 63);
    }
    std::sort(v.begin(), v.end());  // kmer.cpp:39-40: lexicographic (hash,pos)
    h.resize(n);
    p.resize(n);
    for (int i = 0; i < n; ++i) {
        h[i] = v[i].first;
        p[i] = v[i].second;
    }
}

int extract(const char *seq, int len, int k, bool both, ReadKmers &out) {
    if (len <= k || len <= 6 || k < 1 || k > 16) return -1;
    std::vector<int> codes(len);
    for (int i = 0; i < len; ++i) {
        codes[i] = base_code(seq[i]);
        if (codes[i] < 0) return -2;
    }
    out.len = len;
    out.n = len - k;
    memset(out.bvf, 0, sizeof(out.bvf));
    memset(out.bvr, 0, sizeof(out.bvr));
    hashes_of(codes, k, out.fh, out.fp, out.bvf);
    out.pcf = 0;
    for (int w = 0; w < ORC_BV_WORDS; ++w) out.pcf += __builtin_popcountll(out.bvf[w]);
    if (both) {  // utils.cpp:15-24 reverse complement, then the same extraction (kmer.cpp:7,23-24,31-33)
        std::vector<int> rc(len);
        for (int i = 0; i < len; ++i) rc[i] = comp_code(codes[len - 1 - i]);
        hashes_of(rc, k, out.rh, out.rp, out.bvr);
    } else {
        out.rh.clear();
        out.rp.clear();
    }
    return out.n;
}

typedef std::pair<int32_t, int32_t> match_t;

// kmer.cpp:45-67: every cross pair of equal hashes, then lexicographic sort
void common_kmers(const uint32_t *h1, const int32_t *p1, int n1, const uint32_t *h2, const int32_t *p2, int n2,
                  std::vector<match_t> &out) {
    out.clear();
    int a = 0, b = 0;
    while (a < n1 && b < n2) {
        if (h1[a] < h2[b]) ++a;
        else if (h2[b] < h1[a]) ++b;
        else {
            int ae = a, be = b;
            while (ae < n1 && h1[ae] == h1[a]) ++ae;
            while (be < n2 && h2[be] == h2[b]) ++be;
            for (int x = a; x < ae; ++x)
                for (int y = b; y < be; ++y) out.push_back(match_t(p1[x], p2[y]));
            a = ae;
            b = be;
        }
    }
    std::sort(out.begin(), out.end());
}

// similarity.cpp:4-97
int similarity(const std::vector<match_t> &c, int k, std::vector<int32_t> &dist) {
    dist.clear();
    const size_t n = c.size();
    if (n == 0) return 0;
    std::vector<int> prev(n), tail(n + 1, 0);
    int l = 0;
    for (size_t i = 0; i < n; ++i) {
        // similarity.cpp:11-23: number of tails whose .second is strictly smaller (tails are strictly increasing)
        int lo = 1, hi = l;
        while (lo <= hi) {
            int mid = (lo + hi + 1) / 2;
            if (c[tail[mid]].second < c[i].second) lo = mid + 1;
            else hi = mid - 1;
        }
        prev[i] = tail[lo - 1];
        tail[lo] = (int)i;
        if (lo > l) l = lo;
    }
    std::vector<match_t> s(l);
    int at = tail[l];
    for (int i = l - 1; i >= 0; --i) {  // similarity.cpp:40-44
        s[i] = c[at];
        at = prev[at];
    }
    int bases = k;  // similarity.cpp:79-83
    match_t last = s[0];
    for (int i = 1; i < l; ++i) {
        int df = s[i].first - last.first, ds = s[i].second - last.second;
        if ((df < k && ds < k) || (df >= k && ds >= k)) {  // similarity.cpp:54-59
            bases += k;
            int ex = k - (s[i].second - s[i - 1].second);  // overlap w.r.t. previous LIS element (:62)
            if (ex > 0) bases -= ex;
            dist.push_back(ds - df);  // :69-71
            last = s[i];
        }
    }
    return bases;
}

// utils.cpp:26-55
double variance(const std::vector<int32_t> &s) {
    if (s.empty()) return 0;
    double sum = 0.0;
    for (int v : s) sum += double(v);
    double m = sum / double(s.size());
    double ss = 0.0, comp = 0.0;
    for (int v : s) {
        double d = v - m;
        ss += d * d;
        comp += d;
    }
    return (ss - comp * comp / double(s.size())) / double(s.size() - 1);
}

struct Params {
    int k;
    double t_s, t_v;
    bool is_rna;
};

struct Stats {
    int64_t bv_tests = 0, full = 0, accepted = 0, rounds = 0;
};
Stats g_stats;

// cluster.cpp:12-65.  Returns -1 / 0 (fwd) / 1 (rev).
int pair_match(const ReadKmers &a, const ReadKmers &b, const Params &P, double thr, int64_t *n_full) {
    int cf = 0, cr = 0;
    for (int w = 0; w < ORC_BV_WORDS; ++w) {
        cf += __builtin_popcountll(a.bvf[w] & b.bvf[w]);
        cr += __builtin_popcountll(a.bvf[w] & b.bvr[w]);
    }
    double mmax = (double)std::max(a.pcf, b.pcf);
    double mn = (double)std::min(a.len, b.len);
    std::vector<match_t> common;
    std::vector<int32_t> dist;
    if (thr == 0 || (double)cf / mmax >= thr) {
        ++*n_full;
        common_kmers(a.fh.data(), a.fp.data(), a.n, b.fh.data(), b.fp.data(), b.n, common);
        int bases = similarity(common, P.k, dist);
        if (double(bases) / mn >= P.t_s && variance(dist) < P.t_v) return 0;
    }
    if (P.is_rna) return -1;
    if ((double)cr / mmax >= thr) {
        ++*n_full;
        common_kmers(a.fh.data(), a.fp.data(), a.n, b.rh.data(), b.rp.data(), b.n, common);
        int bases = similarity(common, P.k, dist);
        if (double(bases) / mn >= P.t_s && variance(dist) < P.t_v) return 1;
    }
    return -1;
}

struct Member {
    int32_t id;
    uint8_t rev;
};

// cluster.cpp:67-91 (sorts members in place)
Member pick_main(std::vector<Member> &m, const std::vector<ReadKmers> &R, double pct) {
    Member old = m[0];
    std::stable_sort(m.begin(), m.end(), [](const Member &a, const Member &b) { return a.id > b.id; });
    std::stable_sort(m.begin(), m.end(), [&R](const Member &a, const Member &b) { return R[a.id].len > R[b.id].len; });
    size_t at = (size_t)(int)(m.size() * pct);
    Member pick = m[at];
    while (pick.rev != old.rev && at < m.size() - 1) pick = m[++at];
    if (at == m.size() - 1) return old;
    return pick;
}

struct Cluster {
    Member main;
    std::vector<Member> mem;
};

// One greedy pass (cluster.cpp:124-166 with items=reads, :174-245 with items=cluster representatives):
// item i (ascending) not yet taken becomes a seed and takes every later untaken j with match(rep[i],rep[j]).
// hits[i] = list of (j,rev) taken by seed i, ascending j.
void greedy_pass(const std::vector<int32_t> &rep, const std::vector<ReadKmers> &R, const Params &P, double thr,
                 int n_threads, std::vector<std::vector<Member>> &hits, std::vector<char> &is_seed) {
    const int M = (int)rep.size();
    std::vector<char> taken(M, 0);
    hits.assign(M, std::vector<Member>());
    is_seed.assign(M, 0);
    std::vector<int> res(M);
    for (int i = 0; i < M; ++i) {
        if (taken[i]) continue;
        taken[i] = 1;
        is_seed[i] = 1;
        std::vector<std::thread> th;
        std::vector<int64_t> nb(n_threads, 0), nf(n_threads, 0);
        for (int t = 0; t < n_threads; ++t)
            th.emplace_back([&, t]() {
                for (int j = i + 1 + t; j < M; j += n_threads) {
                    if (taken[j]) { res[j] = -1; continue; }
                    ++nb[t];
                    res[j] = pair_match(R[rep[i]], R[rep[j]], P, thr, &nf[t]);
                }
            });
        for (auto &x : th) x.join();
        for (int t = 0; t < n_threads; ++t) { g_stats.bv_tests += nb[t]; g_stats.full += nf[t]; }
        for (int j = i + 1; j < M; ++j)
            if (res[j] >= 0 && !taken[j]) {
                taken[j] = 1;
                hits[i].push_back(Member{(int32_t)j, (uint8_t)res[j]});
                ++g_stats.accepted;
            }
    }
    ++g_stats.rounds;
}

void put_varint(std::vector<uint8_t> &o, uint64_t v) {  // hps uint_serializer.h:16-32
    while (v >= 0x80) {
        o.push_back((uint8_t)(v | 0x80));
        v >>= 7;
    }
    o.push_back((uint8_t)v);
}
void put_zigzag(std::vector<uint8_t> &o, int32_t n) {  // hps int_serializer.h:18-23
    put_varint(o, (uint32_t)((n << 1) ^ (n >> 31)));
}

}  // namespace

extern "C" {

int orc_extract_kmers(const char *seq, int len, int k, int both, uint32_t *fwd_hash, int32_t *fwd_pos,
                      uint32_t *rev_hash, int32_t *rev_pos, uint64_t *bv_fwd, uint64_t *bv_rev) {
    ReadKmers r;
    int n = extract(seq, len, k, both != 0, r);
    if (n < 0) return n;
    memcpy(fwd_hash, r.fh.data(), n * 4);
    memcpy(fwd_pos, r.fp.data(), n * 4);
    if (both) {
        memcpy(rev_hash, r.rh.data(), n * 4);
        memcpy(rev_pos, r.rp.data(), n * 4);
    }
    memcpy(bv_fwd, r.bvf, sizeof(r.bvf));
    memcpy(bv_rev, r.bvr, sizeof(r.bvr));
    return n;
}

int64_t orc_common_kmers(const uint32_t *h1, const int32_t *p1, int n1, const uint32_t *h2, const int32_t *p2,
                         int n2, int32_t *out_first, int32_t *out_second, int64_t cap) {
    std::vector<match_t> c;
    common_kmers(h1, p1, n1, h2, p2, n2, c);
    for (size_t i = 0; i < c.size() && (int64_t)i < cap; ++i) {
        out_first[i] = c[i].first;
        out_second[i] = c[i].second;
    }
    return (int64_t)c.size();
}

int orc_similarity(const int32_t *first, const int32_t *second, int64_t n, int k, int *bases, int32_t *distances,
                   int dist_cap) {
    std::vector<match_t> c(n);
    for (int64_t i = 0; i < n; ++i) c[i] = match_t(first[i], second[i]);
    std::vector<int32_t> d;
    *bases = similarity(c, k, d);
    for (size_t i = 0; i < d.size() && (int)i < dist_cap; ++i) distances[i] = d[i];
    return (int)d.size();
}

double orc_var(const int32_t *d, int n) { return variance(std::vector<int32_t>(d, d + n)); }

int orc_pair_match(const char *s1, int l1, const char *s2, int l2, int k, double t_s, double t_v, double bv_threshold,
                   int is_rna) {
    ReadKmers a, b;
    if (extract(s1, l1, k, !is_rna, a) < 0 || extract(s2, l2, k, !is_rna, b) < 0) return -2;
    Params P{k, t_s, t_v, is_rna != 0};
    int64_t nf = 0;
    return pair_match(a, b, P, bv_threshold, &nf);
}

int orc_cluster_reads(const char *bases, const uint64_t *offsets, uint32_t n_reads, int k, double t_s, double t_v,
                      double bv_thr, double bv_min, double bv_falloff, double repr_pct, int is_rna, int n_threads,
                      int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id, uint8_t *mem_rev) {
    g_stats = Stats();
    if (n_threads < 1) n_threads = 1;
    const int N = (int)n_reads;
    std::vector<ReadKmers> R(N);
    {
        std::vector<std::thread> th;
        std::vector<int> bad(n_threads, 0);
        for (int t = 0; t < n_threads; ++t)
            th.emplace_back([&, t]() {
                for (int i = t; i < N; i += n_threads)
                    if (extract(bases + offsets[i], (int)(offsets[i + 1] - offsets[i]), k, !is_rna, R[i]) < 0) bad[t] = 1;
            });
        for (auto &x : th) x.join();
        for (int b : bad)
            if (b) return -1;
    }
    Params P{k, t_s, t_v, is_rna != 0};
    std::vector<std::vector<Member>> hits;
    std::vector<char> is_seed;

    // initial pass over reads (cluster.cpp:124-166)
    std::vector<int32_t> rep(N);
    for (int i = 0; i < N; ++i) rep[i] = i;
    greedy_pass(rep, R, P, bv_thr, n_threads, hits, is_seed);
    std::vector<Cluster> cl;
    for (int i = 0; i < N; ++i) {
        if (!is_seed[i]) continue;
        Cluster c;
        c.mem.push_back(Member{i, 0});
        for (auto &h : hits[i]) c.mem.push_back(h);
        c.main = pick_main(c.mem, R, repr_pct);
        cl.push_back(c);
    }

    // merge rounds over a falling threshold schedule, then one unfiltered round (cluster.cpp:171-256)
    double thr = bv_thr - bv_falloff;
    bool last = false;
    while (thr >= bv_min || last) {
        const int M = (int)cl.size();
        rep.resize(M);
        for (int i = 0; i < M; ++i) rep[i] = cl[i].main.id;
        greedy_pass(rep, R, P, thr, n_threads, hits, is_seed);
        std::vector<Cluster> next;
        for (int i = 0; i < M; ++i) {
            if (!is_seed[i]) continue;
            Cluster c;
            c.mem = cl[i].mem;
            for (auto &h : hits[i])
                for (Member s : cl[h.id].mem) {
                    if (h.rev) s.rev = !s.rev;  // cluster.cpp:232-234
                    c.mem.push_back(s);
                }
            c.main = pick_main(c.mem, R, repr_pct);
            next.push_back(c);
        }
        cl.swap(next);
        if (last) break;
        thr -= bv_falloff;
        if (thr < bv_min && !last) {
            last = true;
            thr = 0.0;
        }
    }

    int64_t o = 0;
    for (size_t c = 0; c < cl.size(); ++c) {
        main_id[c] = cl[c].main.id;
        main_rev[c] = cl[c].main.rev;
        cl_off[c] = o;
        for (auto &m : cl[c].mem) {
            mem_id[o] = m.id;
            mem_rev[o] = m.rev;
            ++o;
        }
    }
    cl_off[cl.size()] = o;
    return (int)cl.size();
}

void orc_cluster_stats(int64_t out[4]) {
    out[0] = g_stats.bv_tests;
    out[1] = g_stats.full;
    out[2] = g_stats.accepted;
    out[3] = g_stats.rounds;
}

int64_t orc_hps_encode(int n_clusters, const int32_t *main_id, const uint8_t *main_rev, const int32_t *main_gene,
                       const int64_t *cl_off, const int32_t *mem_id, const uint8_t *mem_rev, const int32_t *mem_gene,
                       uint8_t *out, int64_t cap) {
    std::vector<uint8_t> o;
    put_varint(o, (uint64_t)n_clusters);
    for (int c = 0; c < n_clusters; ++c) {
        put_zigzag(o, main_id[c]);
        put_varint(o, main_rev[c] ? 1 : 0);
        put_zigzag(o, main_gene ? main_gene[c] : -1);
        put_varint(o, (uint64_t)(cl_off[c + 1] - cl_off[c]));
        for (int64_t i = cl_off[c]; i < cl_off[c + 1]; ++i) {
            put_zigzag(o, mem_id[i]);
            put_varint(o, mem_rev[i] ? 1 : 0);
            put_zigzag(o, mem_gene ? mem_gene[i] : -1);
        }
    }
    if ((int64_t)o.size() > cap) return -(int64_t)o.size();
    memcpy(out, o.data(), o.size());
    return (int64_t)o.size();
}

}  // extern "C"
