// Drop-in replacements for RATTLE's two hot entry points, with the reference's own signatures, on top of the
// C ABI of librattle_b200 (include/rattle_b200.h):
//
//   cluster_set_t cluster_reads(...)            /root/reference/cluster.hpp:44   (defined in cluster.cpp:93-259)
//   correction_results_t correct_reads(...)     /root/reference/correct.hpp:44   (defined in correct.cpp:311-563)
//   std::vector<std::string> splitString(...)   /root/reference/correct.hpp:47   (correct.cpp:20-30; main.cpp uses it)
//
// This file is what INTEGRATION.md asks a RATTLE maintainer to add: it is compiled against the reference's HEADERS
// (-I$(REF), nothing is copied) and linked with the reference's unmodified main.cpp / fasta.cpp / utils.cpp INSTEAD OF
// cluster.cpp, kmer.cpp, similarity.cpp, correct.cpp and spoa (integration/Makefile).  The result is the `rattle`
// CLI with `cluster`, `correct` and `polish` running on the GPU: same arguments, same clusters.out / *.fq bytes.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "cluster.hpp"
#include "correct.hpp"
#include "utils.hpp"

#include "rattle_b200.h"

namespace {

rtl_ctx *context() {
    static rtl_ctx *ctx = nullptr;
    if (!ctx) {
        const char *dev = getenv("RATTLE_B200_DEVICE");
        if (rtl_init(dev ? atoi(dev) : 0, &ctx) != RTL_OK) {
            fprintf(stderr, "rattle_b200: %s\n", rtl_last_error(nullptr));
            exit(EXIT_FAILURE);  // no CPU fallback
        }
    }
    return ctx;
}

void check(int rc, const char *what) {
    if (rc < 0) {
        fprintf(stderr, "rattle_b200: %s failed: %s\n", what, rtl_last_error(context()));
        exit(EXIT_FAILURE);
    }
}

struct FlatReads {
    std::string bases, quals, headers;
    std::vector<uint64_t> off, hoff;
};

FlatReads flatten(const read_set_t &reads, bool with_quals) {
    FlatReads f;
    size_t total = 0, htotal = 0;
    for (const auto &r : reads) {
        total += r.seq.size();
        htotal += r.header.size();
    }
    f.bases.reserve(total);
    if (with_quals) {
        f.quals.reserve(total);
        f.headers.reserve(htotal);
    }
    f.off.push_back(0);
    f.hoff.push_back(0);
    for (const auto &r : reads) {
        f.bases += r.seq;
        f.off.push_back(f.bases.size());
        if (with_quals) {
            // FASTA input has no qualities; the reference would index an empty string (undefined): refuse instead
            if (r.quality.size() != r.seq.size()) {
                fprintf(stderr, "rattle_b200: read without a quality string of its length (%s)\n", r.header.c_str());
                exit(EXIT_FAILURE);
            }
            f.quals += r.quality;
            f.headers += r.header;
            f.hoff.push_back(f.headers.size());
        }
    }
    return f;
}

read_set_t parse_fastq(const std::string &text) {
    read_set_t out;
    size_t at = 0;
    auto line = [&](std::string &dst) {
        const size_t e = text.find('\n', at);
        dst.assign(text, at, e - at);
        at = e + 1;
    };
    while (at < text.size()) {
        read_t r;
        line(r.header);
        line(r.seq);
        line(r.ann);
        line(r.quality);
        out.push_back(std::move(r));
    }
    return out;
}

}  // namespace

// correct.cpp:20-30
std::vector<std::string> splitString(std::string str, char delimiter) {
    std::vector<std::string> out;
    std::stringstream ss(str);
    std::string tok;
    while (getline(ss, tok, delimiter)) out.push_back(tok);
    return out;
}

// cluster.hpp:44.  min_reads_cluster and use_hc are ignored by the reference as well (cluster.cpp:93-259 never reads
// the former; main.cpp always passes use_hc=false); n_threads and verbose only affect the CPU build.
cluster_set_t cluster_reads(const read_set_t &reads, int kmer_size, double t_s, double t_v, double bv_threshold,
                            double min_bv_threshold, double bv_falloff, int /*min_reads_cluster*/, bool /*use_hc*/,
                            double repr_percentile, bool is_rna, bool /*verbose*/, int /*n_threads*/) {
    cluster_set_t result;
    const uint32_t n = (uint32_t)reads.size();
    if (n == 0) return result;
    const FlatReads f = flatten(reads, false);
    std::vector<int32_t> main_id(n), mem_id(n);
    std::vector<uint8_t> main_rev(n), mem_rev(n);
    std::vector<int64_t> cl_off((size_t)n + 1);
    int32_t n_clusters = 0;
    check(rtl_cluster_reads(context(), f.bases.data(), f.off.data(), n, kmer_size, t_s, t_v, bv_threshold,
                            min_bv_threshold, bv_falloff, repr_percentile, is_rna ? 1 : 0, main_id.data(), main_rev.data(),
                            cl_off.data(), mem_id.data(), mem_rev.data(), &n_clusters),
          "rtl_cluster_reads");
    result.resize(n_clusters);
    for (int c = 0; c < n_clusters; ++c) {
        cluster_t &cl = result[c];
        cl.main_seq = cseq_t{main_id[c], main_rev[c] != 0};
        cl.seqs.reserve((size_t)(cl_off[c + 1] - cl_off[c]));
        for (int64_t i = cl_off[c]; i < cl_off[c + 1]; ++i) cl.seqs.push_back(cseq_t{mem_id[i], mem_rev[i] != 0});
    }
    return result;
}

// correct.hpp:44.  err_ratio is ignored by the reference too (correct.cpp:409 passes the literal 30.0).
correction_results_t correct_reads(const cluster_set_t &clusters, read_set_t &reads, double min_occ, double gap_occ,
                                   double err_ratio, int split, int min_reads, int /*n_threads*/, bool /*verbose*/,
                                   std::vector<std::string> labels) {
    const FlatReads f = flatten(reads, true);
    const int nc = (int)clusters.size();
    std::vector<int32_t> main_id(nc), main_gene(nc);
    std::vector<uint8_t> main_rev(nc);
    std::vector<int64_t> cl_off((size_t)nc + 1, 0);
    size_t total = 0;
    for (const auto &c : clusters) total += c.seqs.size();
    std::vector<int32_t> mem_id(total), mem_gene(total);
    std::vector<uint8_t> mem_rev(total);
    size_t at = 0;
    for (int c = 0; c < nc; ++c) {
        main_id[c] = clusters[c].main_seq.seq_id;
        main_rev[c] = clusters[c].main_seq.rev;
        main_gene[c] = clusters[c].main_seq.gene_id;
        for (const auto &s : clusters[c].seqs) {
            mem_id[at] = s.seq_id;
            mem_rev[at] = s.rev;
            mem_gene[at] = s.gene_id;
            ++at;
        }
        cl_off[c + 1] = (int64_t)at;
    }
    std::vector<const char *> lab;
    for (const auto &l : labels) lab.push_back(l.c_str());
    check(rtl_set_labels(context(), lab.data(), (int)lab.size()), "rtl_set_labels");
    // generous first guess (corrected reads are about as long as the raw ones), exact sizes on RTL_ERR_CAPACITY
    std::unique_ptr<char[]> out[3];  // uninitialised: untouched pages cost nothing
    const int64_t guess = 3 * (int64_t)f.bases.size() + (int64_t)f.headers.size() + 64 * (int64_t)(reads.size() + nc) + 1024;
    int64_t len[3] = {guess, guess, guess};
    for (int attempt = 0; attempt < 2; ++attempt) {
        for (int i = 0; i < 3; ++i) out[i].reset(new char[(size_t)len[i] + 1]);
        const int rc = rtl_correct_reads(context(), f.bases.data(), f.quals.data(), f.off.data(), (uint32_t)reads.size(),
                                         f.headers.data(), f.hoff.data(), main_id.data(), main_rev.data(), main_gene.data(),
                                         cl_off.data(), mem_id.data(), mem_rev.data(), mem_gene.data(), nc, min_occ, gap_occ,
                                         err_ratio, split, min_reads, out[0].get(), &len[0], out[1].get(), &len[1],
                                         out[2].get(), &len[2]);
        if (rc == RTL_OK) break;
        if (rc != RTL_ERR_CAPACITY || attempt == 1) check(rc, "rtl_correct_reads");
    }
    // the reference edits the caller's reads in place (correct.cpp:338-350): reverse members are reverse-complemented,
    // every member's header gets the cluster suffix; polish reads those headers afterwards (main.cpp:685)
    for (int cid = 0; cid < nc; ++cid) {
        const int n_files = (int)((clusters[cid].seqs.size() - 1) / split + 1);
        const int gid = clusters[cid].main_seq.gene_id;
        for (int nf = 0; nf < n_files; ++nf)
            for (size_t j = nf; j < clusters[cid].seqs.size(); j += n_files) {
                const cseq_t &ts = clusters[cid].seqs[j];
                read_t &r = reads[ts.seq_id];
                if (ts.rev) {
                    r.seq = reverse_complement(r.seq);
                    std::reverse(r.quality.begin(), r.quality.end());
                }
                if (gid == -1) r.header = r.header + ",gene_cluster_" + std::to_string(cid);
                else r.header = r.header + ",gene_cluster_" + std::to_string(gid) + ",transcript_cluster_" + std::to_string(cid);
            }
    }
    return correction_results_t{parse_fastq(std::string(out[0].get(), (size_t)len[0])),
                                parse_fastq(std::string(out[1].get(), (size_t)len[1])),
                                parse_fastq(std::string(out[2].get(), (size_t)len[2]))};
}
