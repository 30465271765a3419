// correct_reads (correct.cpp:311-563) on top of the batched GPU POA driver.  The POA (91 % of the reference's time)
// runs on the device for all packs at once; the MSA post-processing (fix_msa_ends, column vote, read correction —
// <2 % of the reference's time, double arithmetic whose summation order and libm calls must match) runs on host
// threads, one pack per thread, and follows the reference's single-threaded pack order.
#include <math.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "common.cuh"
#include "poa_engine.hpp"
#include "msa_ends.hpp"
#include "vote_host.hpp"

namespace {

// utils.cpp:15-24 with utils.hpp:7-13
std::string reverse_complement(const std::string &s) {
    std::string r = s;
    const int len = (int)s.size();
    for (int i = 0; i < len; ++i) {
        char c = s[len - 1 - i], o;
        switch (c) {
            case 'A': o = 'T'; break;
            case 'C': o = 'G'; break;
            case 'T': o = 'A'; break;
            case 'G': o = 'C'; break;
            case 'U': o = 'A'; break;
            default: throw InputError("base outside ACGTU in a reverse-strand cluster member (utils.cpp:20)");
        }
        r[i] = o;
    }
    return r;
}

std::vector<std::string> split_string(const std::string &str, char delim) {  // correct.cpp:20-30
    std::vector<std::string> out;
    std::stringstream ss(str);
    std::string tok;
    while (getline(ss, tok, delim)) out.push_back(tok);
    return out;
}

void append_fastq(std::string &dst, const std::vector<Read> &rs) {  // fasta.cpp:436-445
    for (const auto &r : rs) {
        dst += r.header;
        dst += '\n';
        dst += r.seq;
        dst += '\n';
        dst += r.ann;
        dst += '\n';
        dst += r.quality;
        dst += '\n';
    }
}

struct Pack {
    int cid;
    int64_t first;  // offset of the cluster in the member arrays
    int nf, n_files;
    size_t n;       // cluster size
    bool small = false;
    std::vector<Read> creads;
    std::vector<Read> corrected, uncorrected, sorted_corrected;
    PoaTask t1, t2;
    std::string consensus;
    std::string fq_corrected, fq_uncorrected;  // FASTQ text of this pack's outputs
};

void set_task(PoaTask &t, const std::vector<Read> &rs) {
    t.seq.clear();
    t.len.clear();
    for (const auto &r : rs) {
        t.seq.push_back(r.seq.data());
        t.len.push_back((int)r.seq.size());
    }
}

}  // namespace

int correct_reads_impl(rtl_ctx *ctx, const char *bases, const char *quals, const uint64_t *offsets, uint32_t n_reads,
                       const char *headers, const uint64_t *header_off, const int32_t *main_id, const uint8_t *main_rev,
                       const int32_t *main_gene, const int64_t *cl_off, const int32_t *mem_id, const uint8_t *mem_rev,
                       const int32_t *mem_gene, int n_clusters, double min_occ, double gap_occ, double err_ratio,
                       int split, int min_reads, char *corrected_out, int64_t *corrected_len, char *uncorrected_out,
                       int64_t *uncorrected_len, char *consensi_out, int64_t *consensi_len) {
    (void)main_id;
    (void)main_rev;
    (void)mem_gene;
    (void)err_ratio;  // the reference ignores it too: correct.cpp:409 passes the literal 30.0
    const double t0 = now_ms();
    const bool trace = getenv("RTL_TRACE") != nullptr;
    double t_last = t0;
    auto lap = [&](const char *what) {
        if (!trace) return;
        const double t = now_ms();
        fprintf(stderr, "[rtl] correct %-28s %9.1f ms\n", what, t - t_last);
        t_last = t;
    };
    ctx->stats = rtl_stats{};
    if (n_clusters < 1) throw InputError("empty cluster set (correct.cpp:322 reads clusters[0])");
    if (split < 1) throw InputError("split must be >= 1");
    if (!cl_off || !mem_id || !mem_rev || !offsets || !bases || !quals) throw InputError("null input buffers");
    if (cl_off[0] != 0) throw InputError("cl_off must start at 0");
    for (int c = 0; c < n_clusters; ++c)
        if (cl_off[c + 1] < cl_off[c]) throw InputError("cl_off is not monotonic");
    for (uint32_t i = 0; i < n_reads; ++i)
        if (offsets[i + 1] < offsets[i]) throw InputError("read offsets are not monotonic");
    const int nthreads = host_threads();
    const bool gene_mode = (main_gene ? main_gene[0] : -1) == -1;

    // ---- pack construction (correct.cpp:328-370).  Every read belongs to one pack, so packs are built in parallel
    // straight from the flat input; a read listed twice would be reverse-complemented / re-labelled twice by the
    // reference (it edits the shared read in place), which the serial fallback below reproduces.
    std::vector<Pack> packs;
    for (int cid = 0; cid < n_clusters; ++cid) {
        const size_t n = (size_t)(cl_off[cid + 1] - cl_off[cid]);
        if (n == 0) throw InputError("cluster without members");
        const int n_files = (int)((n - 1) / split + 1);
        for (int nf = 0; nf < n_files; ++nf) {
            packs.emplace_back();
            Pack &p = packs.back();
            p.cid = cid;
            p.first = cl_off[cid];
            p.nf = nf;
            p.n_files = n_files;
            p.n = n;
        }
    }
    bool duplicates = false;
    {
        std::vector<uint8_t> seen(n_reads, 0);
        const int64_t total = cl_off[n_clusters];
        for (int64_t i = 0; i < total; ++i) {
            const int id = mem_id[i];
            if (id < 0 || (uint32_t)id >= n_reads) throw InputError("cluster member index out of range");
            if (seen[id]) duplicates = true;
            seen[id] = 1;
        }
    }
    // sharded correction (rtl_set_cluster_ids): cluster c of this call is cluster out_cid(c) of the whole cluster set,
    // and that id is what the headers carry (correct.cpp:344-349,540-549 write the index in clusters.out)
    if (!ctx->cluster_ids.empty() && (int)ctx->cluster_ids.size() != n_clusters)
        throw InputError("rtl_set_cluster_ids: id count differs from n_clusters");
    auto out_cid = [&](int cid) { return ctx->cluster_ids.empty() ? cid : (int)ctx->cluster_ids[cid]; };
    auto suffix = [&](int cid) {
        const int gid = main_gene ? main_gene[cid] : -1;
        return gid == -1 ? ",gene_cluster_" + std::to_string(out_cid(cid))
                         : ",gene_cluster_" + std::to_string(gid) + ",transcript_cluster_" + std::to_string(out_cid(cid));
    };
    if (!duplicates) {
        parallel_for(nthreads, packs.size(), [&](size_t pi) {
            Pack &p = packs[pi];
            const std::string suf = suffix(p.cid);
            for (size_t j = p.nf; j < p.n; j += p.n_files) {
                const int id = mem_id[p.first + j];
                Read r;
                if (headers) r.header.assign(headers + header_off[id], headers + header_off[id + 1]);
                else r.header = "@r" + std::to_string(id);
                r.header += suf;
                r.seq.assign(bases + offsets[id], bases + offsets[id + 1]);
                r.ann = "+";
                r.quality.assign(quals + offsets[id], quals + offsets[id + 1]);
                if (mem_rev[p.first + j]) {
                    r.seq = reverse_complement(r.seq);
                    std::reverse(r.quality.begin(), r.quality.end());
                }
                p.creads.push_back(std::move(r));
            }
        });
    } else {
        std::vector<Read> reads(n_reads);
        for (uint32_t i = 0; i < n_reads; ++i) {
            if (headers) reads[i].header.assign(headers + header_off[i], headers + header_off[i + 1]);
            else reads[i].header = "@r" + std::to_string(i);
            reads[i].seq.assign(bases + offsets[i], bases + offsets[i + 1]);
            reads[i].ann = "+";
            reads[i].quality.assign(quals + offsets[i], quals + offsets[i + 1]);
        }
        for (auto &p : packs) {
            const std::string suf = suffix(p.cid);
            for (size_t j = p.nf; j < p.n; j += p.n_files) {
                Read &r = reads[mem_id[p.first + j]];
                if (mem_rev[p.first + j]) {
                    r.seq = reverse_complement(r.seq);
                    std::reverse(r.quality.begin(), r.quality.end());
                }
                r.header = r.header + suf;
                p.creads.push_back(r);
            }
        }
    }
    // packs that are too small are not corrected (correct.cpp:360-366); their reads go to `uncorrected` first
    std::string out_u, out_s;
    {
        std::vector<Pack> keep;
        keep.reserve(packs.size());
        for (auto &p : packs) {
            if ((int)p.creads.size() > min_reads) keep.push_back(std::move(p));
            else append_fastq(out_u, p.creads);
        }
        packs.swap(keep);
    }
    lap("pack construction");
    // ---- per-pack pipeline (correct.cpp:395-445): POA round 1 on the raw reads, MSA -> fix_msa_ends -> column vote
    // -> corrected reads, POA round 2 on the corrected reads (length-sorted), MSA -> consensus.  Packs are
    // independent, so they are split into units that run the whole pipeline concurrently (poa_engine.cu): while some
    // units are in their host phases the kernels of the others keep the GPU busy.
    {
        CK(cudaStreamSynchronize(ctx->stream));
        const double tp0 = now_ms();
        const int U = poa_unit_count(ctx, packs.size());
        const int nt = nthreads;  // the shared worker pool arbitrates between the units
        std::vector<std::vector<Pack *>> unit(U);
        for (size_t i = 0; i < packs.size(); ++i) unit[i % U].push_back(&packs[i]);
        run_units(U, [&](int u) {
            std::vector<Pack *> mine = unit[u];
            // ---- the whole pipeline on the GPU (poa_vote.cuh): MSAs stay on the device, the host gets the corrected reads
            // and the consensus.  Packs the device path does not take (or flags) fall through to the host pipeline below.
            if (ctx->poa_device_vote != 0 && ctx->poa_device_chain != 0) {
                const double td0 = now_ms();
                std::vector<VotePack> vps(mine.size());
                std::vector<VotePack *> ptrs(mine.size());
                for (size_t i = 0; i < mine.size(); ++i) {
                    VotePack &v = vps[i];
                    for (const auto &r : mine[i]->creads) {
                        v.seq.push_back(r.seq.data());
                        v.qual.push_back(r.quality.data());
                        v.len.push_back((int)r.seq.size());
                        if (r.quality.size() != r.seq.size()) v.seq.clear();  // (never: FASTQ) -> host path
                    }
                    ptrs[i] = &v;
                }
                poa_correct_unit(ctx, u, ptrs, min_occ, gap_occ, nt);
                const double td1 = now_ms();
                parallel_for(nt, mine.size(), [&](size_t i) {
                    VotePack &v = vps[i];
                    if (!v.done) return;
                    Pack &p = *mine[i];
                    for (size_t r = 0; r < p.creads.size(); ++r) {
                        Read &rd = p.creads[r];
                        if (v.tf[r] || v.tb[r]) {  // fix_msa_ends edits the reads as well (correct.cpp:64-65)
                            const size_t keep = rd.seq.size() - (size_t)v.tf[r] - (size_t)v.tb[r];
                            rd.seq = rd.seq.substr((size_t)v.tf[r], keep);
                            rd.quality = rd.quality.substr((size_t)v.tf[r], keep);
                        }
                        if (!v.cseq[r].empty()) p.corrected.push_back(Read{rd.header, std::move(v.cseq[r]), "+", std::move(v.cqual[r])});
                        else p.uncorrected.push_back(rd);
                    }
                    append_fastq(p.fq_corrected, p.corrected);
                    append_fastq(p.fq_uncorrected, p.uncorrected);
                    std::vector<Read>().swap(p.corrected);
                    std::vector<Read>().swap(p.uncorrected);
                    p.consensus = std::move(v.consensus);
                });
                std::vector<Pack *> rest;
                for (size_t i = 0; i < mine.size(); ++i)
                    if (!vps[i].done) rest.push_back(mine[i]);
                if (trace)
                    fprintf(stderr, "[rtl] correct unit %d: device pipeline %.1f ms, FASTQ text of its packs %.1f ms, %zu of %zu packs "
                            "left for the host pipeline\n", u, td1 - td0, now_ms() - td1, rest.size(), mine.size());
                mine.swap(rest);
                if (mine.empty()) return;
            }
            std::vector<PoaTask *> tasks;
            for (Pack *p : mine) {
                set_task(p->t1, p->creads);
                tasks.push_back(&p->t1);
            }
            poa_chain(ctx, u, tasks, 5, -4, -8, -6, false, nt);
            const double tv0 = now_ms();
            parallel_for(nt, mine.size(), [&](size_t i) {
                Pack &p = *mine[i];
                std::vector<std::string> msa;
                p.t1.take_msa(msa);
                fix_msa_ends(p.creads, msa);
                correct_pack(p.creads, msa, min_occ, gap_occ, 30.0, p.corrected, p.uncorrected);
                append_fastq(p.fq_corrected, p.corrected);
                append_fastq(p.fq_uncorrected, p.uncorrected);
                p.sorted_corrected = p.corrected;
                std::stable_sort(p.sorted_corrected.begin(), p.sorted_corrected.end(),
                                 [](const Read &a, const Read &b) { return a.seq.size() > b.seq.size(); });  // fasta.cpp:458-464
                std::vector<Read>().swap(p.corrected);
                std::vector<Read>().swap(p.uncorrected);
            });
            const double tv1 = now_ms();
            tasks.clear();
            for (Pack *p : mine) {
                set_task(p->t2, p->sorted_corrected);
                tasks.push_back(&p->t2);
            }
            poa_chain(ctx, u, tasks, 5, -4, -8, -6, false, nt);
            const double tv2 = now_ms();
            parallel_for(nt, mine.size(), [&](size_t i) {
                Pack &p = *mine[i];
                std::vector<std::string> msa;
                p.t2.take_msa(msa);
                fix_msa_ends(p.sorted_corrected, msa);
                ColStats cs;
                consensus_vector(p.sorted_corrected, msa, cs);
                p.consensus = strip_gaps(cs.consensus);
            });
            if (trace)
                fprintf(stderr, "[rtl] correct unit %d: host column vote + read correction %.1f ms, consensus vote %.1f ms (%d threads)\n",
                        u, tv1 - tv0, now_ms() - tv2, nt);
        });
        ctx->stats.poa_wall_ms += now_ms() - tp0;
        poa_account_busy(ctx);
    }
    lap("POA rounds 1+2 with correction");
    // queue order = the reference's -t 1 order (correct.cpp:413-424); the per-pack FASTQ texts are copied into the
    // caller's buffers at the end, in parallel
    lap("output assembly");
    // ---- pack consensus headers (correct.cpp:447-470), literally, file labels included (rtl_set_labels)
    const std::vector<std::string> &labels = ctx->labels;
    std::vector<std::vector<Read>> consensi(n_clusters);
    for (auto &p : packs) {
        std::string gid;
        std::vector<std::string> labelset;
        for (const auto &r : p.creads) {
            if (!labels.empty()) {
                int index = (int)r.header.find_first_of(",");
                int i = (int)r.header.substr(index + 1).find_first_of(",");
                labelset.push_back(r.header.substr(index + 1, i));
            }
            const size_t index = r.header.find("gene_cluster");
            gid = std::to_string(std::stoi(r.header.substr(index + 13)));
        }
        std::string label_result;
        for (const auto &label : labels)
            label_result = label_result + " " + label + ":" + std::to_string(std::count(labelset.begin(), labelset.end(), label));
        consensi[p.cid].push_back(Read{gid + "," + std::to_string(p.creads.size()) + "," + label_result, p.consensus, "+",
                                       std::string(p.consensus.size(), 'K')});
    }

    std::vector<Read> consensus_set;
    // ---- clusters with several packs: third POA over the pack consensi (correct.cpp:518-538)
    std::vector<PoaTask> t3(n_clusters);
    std::vector<PoaTask *> tasks;
    for (int cid = 0; cid < n_clusters; ++cid)
        if (consensi[cid].size() > 1) {
            set_task(t3[cid], consensi[cid]);
            tasks.push_back(&t3[cid]);
        }
    if (!tasks.empty()) poa_run(ctx, tasks, 5, -4, -8, -6, false);
    for (int cid = 0; cid < n_clusters; ++cid) {
        auto &it = consensi[cid];
        int total_reads = 0, gid = 0;
        std::vector<int> label_counts(labels.size());
        for (const auto &r : it) {
            auto num = split_string(r.header, ',');
            gid = std::stoi(num[0]);
            total_reads += std::stoi(num[1]);
            int i = 0;
            for (const auto &label : labels) {  // correct.cpp:498-509
                if (r.header.find(label) != std::string::npos) {
                    int index = (int)r.header.find(label);
                    const std::string sub = r.header.substr(index + 1);
                    index = (int)sub.find_first_of(":");
                    try {
                        label_counts[i] += std::stoi(sub.substr(index + 1));
                    } catch (const std::exception &) {
                        throw InputError("file label that cannot be counted in a consensus header (correct.cpp:506)");
                    }
                }
                ++i;
            }
        }
        std::string labels_result;
        for (size_t i = 0; i < labels.size(); ++i) labels_result += labels[i] + ":" + std::to_string(label_counts[i]) + ",";
        const std::string head =
            (gene_mode ? "@gene_cluster_" + std::to_string(out_cid(cid)) + " reads=" + std::to_string(total_reads) + " labels="
                       : "@transcript_cluster_" + std::to_string(out_cid(cid)) + " gene_cluster_" + std::to_string(gid) + " reads=" +
                             std::to_string(total_reads) + " labels=") +
            labels_result;
        if (it.size() > 1) {
            std::vector<std::string> msa;
            t3[cid].take_msa(msa);
            fix_msa_ends(it, msa);
            ColStats cs;
            consensus_vector(it, msa, cs);
            const std::string consensus = strip_gaps(cs.consensus);
            consensus_set.push_back(Read{head, consensus, "+", std::string(consensus.size(), 'K')});
        } else if (it.size() == 1) {
            consensus_set.push_back(Read{head, it[0].seq, "+", it[0].quality});
        }
    }

    lap("third POA + headers");
    append_fastq(out_s, consensus_set);
    int rc = RTL_OK;
    // corrected = packs in order; uncorrected = reads of the too-small packs (out_u), then the packs in order
    std::vector<size_t> off_c(packs.size() + 1, 0), off_u(packs.size() + 1, out_u.size());
    for (size_t i = 0; i < packs.size(); ++i) {
        off_c[i + 1] = off_c[i] + packs[i].fq_corrected.size();
        off_u[i + 1] = off_u[i] + packs[i].fq_uncorrected.size();
    }
    const size_t need[3] = {off_c[packs.size()], off_u[packs.size()], out_s.size()};
    char *const bufs[3] = {corrected_out, uncorrected_out, consensi_out};
    int64_t *const lens[3] = {corrected_len, uncorrected_len, consensi_len};
    for (int i = 0; i < 3; ++i)
        if ((int64_t)need[i] > *lens[i] || !bufs[i]) rc = RTL_ERR_CAPACITY;
    if (rc == RTL_OK) {
        memcpy(uncorrected_out, out_u.data(), out_u.size());
        memcpy(consensi_out, out_s.data(), out_s.size());
        parallel_for(nthreads, packs.size(), [&](size_t i) {
            memcpy(corrected_out + off_c[i], packs[i].fq_corrected.data(), packs[i].fq_corrected.size());
            memcpy(uncorrected_out + off_u[i], packs[i].fq_uncorrected.data(), packs[i].fq_uncorrected.size());
        });
    }
    for (int i = 0; i < 3; ++i) *lens[i] = (int64_t)need[i];
    lap("output");
    ctx->stats.total_ms = now_ms() - t0;
    ctx->stats.d2h_bytes += (int64_t)(need[0] + need[1] + need[2]);
    if (rc == RTL_ERR_CAPACITY) ctx->err = "output buffer too small (needed sizes returned in *_len)";
    return rc;
}
