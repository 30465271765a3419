// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product path.
//
// Thin extern "C" wrapper around the UNMODIFIED reference sources as they lie under
// /root/reference (compiled in place by oracle/Makefile into oracle/_ref/libref_shim.so).
// It exists so that tests can call the reference's own functions at function granularity
// (the reference binary only gives file-level outputs) and pin oracle/rattle_oracle.cpp and
// the CUDA kernels against them.  Nothing here restates an algorithm: every entry point
// forwards to a reference symbol:
//   extract_kmers_from_read   /root/reference/kmer.cpp:6
//   get_common_kmers          /root/reference/kmer.cpp:45
//   calc_similarity           /root/reference/similarity.cpp:4
//   var                       /root/reference/utils.cpp:36
//   cluster_together          /root/reference/cluster.cpp:12   (external linkage, not in a header)
//   cluster_reads             /root/reference/cluster.cpp:93
//   spoa align/add_alignment/generate_multiple_sequence_alignment
//                             /root/reference/spoa/include/spoa/*.hpp
//   fix_msa_ends / generate_consensus_vector / correct_read_pack / correct_reads
//                             /root/reference/correct.cpp:32,94,196,311
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "cluster.hpp"
#include "correct.hpp"
#include "kmer.hpp"
#include "similarity.hpp"
#include "utils.hpp"
#include "spoa/spoa.hpp"

cseq_t cluster_together(const read_set_t &reads, const std::vector<std::vector<kmer_t>> &kmers,
                        const std::vector<std::vector<kmer_t>> &rev_kmers,
                        const std::vector<kmer_bv_t> &bv_kmers, const std::vector<kmer_bv_t> &rev_bv_kmers,
                        int i, int j, int kmer_size, double t_s, double t_v, double bv_threshold,
                        bool use_hc, bool is_rna);

static void bv_to_words(const kmer_bv_t &bv, uint64_t *out) {
    for (int w = 0; w < BV_SIZE / 64; ++w) {
        uint64_t x = 0;
        for (int b = 0; b < 64; ++b)
            if (bv[w * 64 + b]) x |= (uint64_t)1 << b;
        out[w] = x;
    }
}

static read_set_t make_reads(const char *bases, const uint64_t *offsets, uint32_t n) {
    read_set_t reads(n);
    for (uint32_t i = 0; i < n; ++i) {
        reads[i].header = "@r" + std::to_string(i);
        reads[i].seq.assign(bases + offsets[i], bases + offsets[i + 1]);
        reads[i].ann = std::to_string(i);
    }
    return reads;
}

extern "C" {

// lists are written as (hash, pos) pairs; each buffer must hold len-k entries.
int ref_extract_kmers(const char *seq, int len, int k, int both_strands,
                      uint32_t *fwd_hash, int32_t *fwd_pos, uint32_t *rev_hash, int32_t *rev_pos,
                      uint64_t *bv_fwd, uint64_t *bv_rev) {
    read_kmers_t r = extract_kmers_from_read(std::string(seq, len), k, both_strands != 0);
    for (size_t i = 0; i < r.list_forward.size(); ++i) {
        fwd_hash[i] = r.list_forward[i].first;
        fwd_pos[i] = r.list_forward[i].second;
    }
    if (both_strands)
        for (size_t i = 0; i < r.list_reverse.size(); ++i) {
            rev_hash[i] = r.list_reverse[i].first;
            rev_pos[i] = r.list_reverse[i].second;
        }
    bv_to_words(r.bv_forward, bv_fwd);
    bv_to_words(r.bv_reverse, bv_rev);
    return (int)r.list_forward.size();
}

// returns number of common pairs; writes at most cap of them
int64_t ref_common_kmers(const uint32_t *h1, const int32_t *p1, int n1, const uint32_t *h2,
                         const int32_t *p2, int n2, int32_t *out_first, int32_t *out_second, int64_t cap) {
    std::vector<kmer_t> a(n1), b(n2);
    for (int i = 0; i < n1; ++i) a[i] = kmer_t(h1[i], p1[i]);
    for (int i = 0; i < n2; ++i) b[i] = kmer_t(h2[i], p2[i]);
    auto c = get_common_kmers(a, b);
    for (size_t i = 0; i < c.size() && (int64_t)i < cap; ++i) {
        out_first[i] = c[i].first;
        out_second[i] = c[i].second;
    }
    return (int64_t)c.size();
}

// returns number of distances; bases via out param
int ref_similarity(const int32_t *first, const int32_t *second, int64_t n, int k, int *bases,
                   int32_t *distances, int dist_cap) {
    std::vector<kmer_match_t> c(n);
    for (int64_t i = 0; i < n; ++i) c[i] = kmer_match_t(first[i], second[i]);
    auto sim = calc_similarity(c, k);
    *bases = sim.bases;
    for (size_t i = 0; i < sim.distances.size() && (int)i < dist_cap; ++i) distances[i] = sim.distances[i];
    return (int)sim.distances.size();
}

double ref_var(const int32_t *d, int n) { return var(std::vector<int>(d, d + n)); }

// pair test on two reads: returns -1 (no), 0 (forward), 1 (reverse)
int ref_pair_match(const char *s1, int l1, const char *s2, int l2, int k, double t_s, double t_v,
                   double bv_threshold, int is_rna) {
    read_set_t reads(2);
    reads[0].seq.assign(s1, l1);
    reads[1].seq.assign(s2, l2);
    std::vector<std::vector<kmer_t>> kmers(2), rev_kmers(2);
    std::vector<kmer_bv_t> bv(2), rbv(2);
    for (int i = 0; i < 2; ++i) {
        read_kmers_t r = extract_kmers_from_read(reads[i].seq, k, !is_rna);
        kmers[i] = r.list_forward;
        rev_kmers[i] = r.list_reverse;
        bv[i] = r.bv_forward;
        rbv[i] = r.bv_reverse;
    }
    cseq_t c = cluster_together(reads, kmers, rev_kmers, bv, rbv, 0, 1, k, t_s, t_v, bv_threshold, false, is_rna != 0);
    if (c.seq_id == -1) return -1;
    return c.rev ? 1 : 0;
}

// Full clustering through the reference's cluster_reads.  Flat outputs:
//  main_id[c], main_rev[c], cl_off[c..c+1], mem_id[], mem_rev[]; returns n_clusters.
int ref_cluster_reads(const char *bases, const uint64_t *offsets, uint32_t n, int k, double t_s, double t_v,
                      double bv_thr, double bv_min, double bv_falloff, double repr_pct, int is_rna,
                      int n_threads, int32_t *main_id, uint8_t *main_rev, int64_t *cl_off, int32_t *mem_id,
                      uint8_t *mem_rev) {
    read_set_t reads = make_reads(bases, offsets, n);
    cluster_set_t cs = cluster_reads(reads, k, t_s, t_v, bv_thr, bv_min, bv_falloff, 0, false, repr_pct,
                                     is_rna != 0, false, n_threads);
    int64_t o = 0;
    for (size_t c = 0; c < cs.size(); ++c) {
        main_id[c] = cs[c].main_seq.seq_id;
        main_rev[c] = cs[c].main_seq.rev;
        cl_off[c] = o;
        for (auto &s : cs[c].seqs) {
            mem_id[o] = s.seq_id;
            mem_rev[o] = s.rev;
            ++o;
        }
    }
    cl_off[cs.size()] = o;
    return (int)cs.size();
}

// ---------------------------------------------------------------------------------------------
// spoa: POA of a list of sequences in order with (kSW, m, n, g, e) exactly as correct.cpp:395-405
// Emits the MSA rows (concatenated, each of length *msa_cols) into msa_out (cap bytes).
// If aln_out != NULL also dumps every alignment as (node,pos) pairs: aln_off[n+1], aln_pairs.
int ref_poa_msa(const char *bases, const uint64_t *offsets, uint32_t n, int m, int nn, int g, int e,
                char *msa_out, int64_t cap, int *msa_cols, int64_t *aln_off, int32_t *aln_pairs,
                int64_t aln_cap) {
    auto engine = spoa::createAlignmentEngine(spoa::AlignmentType::kSW, (int8_t)m, (int8_t)nn, (int8_t)g, (int8_t)e);
    auto graph = spoa::createGraph();
    int64_t ao = 0;
    for (uint32_t i = 0; i < n; ++i) {
        std::string s(bases + offsets[i], bases + offsets[i + 1]);
        auto aln = engine->align(s, graph);
        if (aln_off) {
            aln_off[i] = ao;
            for (auto &p : aln) {
                if (ao + 1 <= aln_cap / 2) {
                    aln_pairs[2 * ao] = p.first;
                    aln_pairs[2 * ao + 1] = p.second;
                }
                ++ao;
            }
        }
        graph->add_alignment(aln, s);
    }
    if (aln_off) aln_off[n] = ao;
    std::vector<std::string> msa;
    graph->generate_multiple_sequence_alignment(msa);
    *msa_cols = msa.empty() ? 0 : (int)msa[0].size();
    int64_t need = (int64_t)msa.size() * (*msa_cols);
    if (need > cap) return -1;
    for (size_t i = 0; i < msa.size(); ++i) memcpy(msa_out + i * (*msa_cols), msa[i].data(), *msa_cols);
    return (int)msa.size();
}

// correct_reads through the reference (correct.cpp:311).  Inputs: all reads (seq+qual+header), clusters flat.
// Outputs are serialised as FASTQ text into three caller buffers; returns 0 or -1 when a buffer is too small.
int ref_correct_reads(const char *bases, const char *quals, const uint64_t *offsets, uint32_t n_reads,
                      const int32_t *main_id, const uint8_t *main_rev, const int32_t *main_gene,
                      const int64_t *cl_off, const int32_t *mem_id, const uint8_t *mem_rev,
                      const int32_t *mem_gene, int n_clusters, double min_occ, double gap_occ, double err_ratio,
                      int split, int min_reads, int n_threads, char *corrected, int64_t *corrected_len,
                      char *uncorrected, int64_t *uncorrected_len, char *consensi, int64_t *consensi_len) {
    read_set_t reads(n_reads);
    for (uint32_t i = 0; i < n_reads; ++i) {
        reads[i].header = "@r" + std::to_string(i);
        reads[i].seq.assign(bases + offsets[i], bases + offsets[i + 1]);
        reads[i].ann = "+";
        reads[i].quality.assign(quals + offsets[i], quals + offsets[i + 1]);
    }
    cluster_set_t cs(n_clusters);
    for (int c = 0; c < n_clusters; ++c) {
        cs[c].main_seq = cseq_t{main_id[c], main_rev[c] != 0, main_gene[c]};
        for (int64_t o = cl_off[c]; o < cl_off[c + 1]; ++o)
            cs[c].seqs.push_back(cseq_t{mem_id[o], mem_rev[o] != 0, mem_gene[o]});
    }
    auto res = correct_reads(cs, reads, min_occ, gap_occ, err_ratio, split, min_reads, n_threads, false,
                             std::vector<std::string>());
    auto dump = [](const read_set_t &rs, char *buf, int64_t *len) -> int {
        std::string s;
        for (auto &r : rs) {
            s += r.header; s += '\n'; s += r.seq; s += '\n'; s += r.ann; s += '\n'; s += r.quality; s += '\n';
        }
        if ((int64_t)s.size() > *len) { *len = (int64_t)s.size(); return -1; }
        memcpy(buf, s.data(), s.size());
        *len = (int64_t)s.size();
        return 0;
    };
    int rc = 0;
    rc |= dump(res.corrected, corrected, corrected_len);
    rc |= dump(res.uncorrected, uncorrected, uncorrected_len);
    rc |= dump(res.consensi, consensi, consensi_len);
    return rc;
}

// fix_msa_ends (correct.cpp:32) on caller buffers: rows = n x ncol chars (edited in place); read i's bases / qualities
// at off[i]..off[i+1] (edited in place, new_len[i] = what is left of them)
int ref_fix_msa_ends(char *rows, int n, int ncol, char *seqs, char *quals, const int64_t *off, int32_t *new_len) {
    read_set_t reads(n);
    msa_t aln(n);
    for (int i = 0; i < n; ++i) {
        reads[i].seq.assign(seqs + off[i], seqs + off[i + 1]);
        reads[i].quality.assign(quals + off[i], quals + off[i + 1]);
        aln[i].assign(rows + (size_t)i * ncol, ncol);
    }
    fix_msa_ends(reads, aln);
    for (int i = 0; i < n; ++i) {
        memcpy(rows + (size_t)i * ncol, aln[i].data(), ncol);
        new_len[i] = (int32_t)reads[i].seq.size();
        memcpy(seqs + off[i], reads[i].seq.data(), reads[i].seq.size());
        memcpy(quals + off[i], reads[i].quality.data(), reads[i].quality.size());
    }
    return 0;
}

}  // extern "C"
