// Batched POA driver of hot path B: many independent partial-order alignments (one per read pack) advance in
// lock-step — step s aligns the s-th sequence of every pack to that pack's graph on the GPU (k_poa_align, one CTA
// per alignment), then the host threads fold the alignments into the graphs (PoaGraph::add_alignment, which also
// re-sorts the graph) and emit the next step's graphs in rank-order CSR form.  This is the loop of
// correct.cpp:399-402 / :430-433 / :525-528 turned inside out so that thousands of clusters share one launch.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "common.cuh"
#include "poa_engine.hpp"
#include "poa_kernels.cuh"

using namespace rtl;

static double g_t_finish_wait = 0, g_t_fold = 0, g_t_stage = 0;  // RTL_TRACE phase timers

struct JobRef {
    PoaTask *task;
    int seq_index;
    bool wide;
    size_t hf_cells, codes;  // memory need in elements
    int L, n;
};

// One pipeline slot: its own staging buffers, stream and half of the device arena.  While the kernel of one slot
// runs, the host folds the alignments of the other slot into their graphs and stages that slot's next step.
struct PoaSlot {
    DevBuf<uint8_t> d_q;
    DevBuf<uint32_t> d_row_info, d_row_poff;
    DevBuf<int32_t> d_preds, d_aln, d_aln_len;
    DevBuf<PoaJob> d_jobs;
    DevBuf<unsigned int> d_counter;
    PinBuf<uint8_t> h_q;
    PinBuf<uint32_t> h_row_info, h_row_poff;
    PinBuf<int32_t> h_preds, h_aln, h_aln_len;
    PinBuf<PoaJob> h_jobs;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned char *hf = nullptr, *code = nullptr;  // this slot's arena halves
    size_t hf_bytes = 0, code_bytes = 0;
    std::vector<JobRef> jobs;  // in flight
    std::vector<size_t> aln_off;
    bool pending = false;
    bool keep_alns = false;
};

struct PoaState {
    DevBuf<unsigned char> hf_arena, code_arena;
    PoaSlot slot[2];
    int occ[2] = {0, 0};
    int n_threads = 0;
};

void poa_state_free(rtl_ctx *ctx) {
    if (ctx->poa) {
        for (auto &sl : ctx->poa->slot) {
            if (sl.ev0) cudaEventDestroy(sl.ev0);
            if (sl.ev1) cudaEventDestroy(sl.ev1);
            if (sl.stream) cudaStreamDestroy(sl.stream);
        }
        delete ctx->poa;
    }
    ctx->poa = nullptr;
}

void parallel_for(int n_threads, size_t n, const std::function<void(size_t)> &fn) {
    if (n == 0) return;
    if (n_threads <= 1 || n == 1) {
        for (size_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<size_t> next(0);
    std::vector<std::thread> th;
    const int nt = (int)std::min<size_t>((size_t)n_threads, n);
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&]() {
            while (true) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                fn(i);
            }
        });
    for (auto &x : th) x.join();
}

int host_threads() {
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    return (int)std::min(hc, 64u);
}

static PoaState &pstate(rtl_ctx *ctx) {
    if (!ctx->poa) {
        ctx->poa = new PoaState();
        PoaState &P = *ctx->poa;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&P.occ[0], k_poa_align<false>, POA_T, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&P.occ[1], k_poa_align<true>, POA_T, 0));
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        size_t budget = ctx->poa_arena_mb > 0 ? ((size_t)ctx->poa_arena_mb << 20) : std::min<size_t>(free_b * 2 / 5, 64ull << 30);
        budget = std::max<size_t>(budget, 64ull << 20);
        const size_t hf_half = (budget / 3) & ~(size_t)255, code_half = (budget / 6) & ~(size_t)255;
        P.hf_arena.need(2 * hf_half);
        P.code_arena.need(2 * code_half);
        for (int i = 0; i < 2; ++i) {
            PoaSlot &sl = P.slot[i];
            CK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
            CK(cudaEventCreate(&sl.ev0));
            CK(cudaEventCreate(&sl.ev1));
            sl.hf = P.hf_arena.p + i * hf_half;
            sl.code = P.code_arena.p + i * code_half;
            sl.hf_bytes = hf_half;
            sl.code_bytes = code_half;
        }
        P.n_threads = host_threads();
    }
    return *ctx->poa;
}

// rank-order CSR of the graph + query into the staging buffers
static void stage_job(const JobRef &jr, PoaJob &J, uint8_t *q, uint32_t *row_info, uint32_t *row_poff, int32_t *preds) {
    const PoaGraph &g = jr.task->g;
    const int L = jr.L, n = jr.n;
    const int Lp = poa_lp(L);
    memcpy(q, jr.task->seq[jr.seq_index], L);
    memset(q + L, 0, Lp + 4 - L);
    row_info[0] = 0;
    row_poff[0] = 0;
    uint32_t at = 0;
    for (int r = 1; r <= n; ++r) {
        const int v = g.rank_to_node[r - 1];
        const int np = g.n_in[v];
        row_poff[r] = at;
        if (np == 0) {
            preds[at++] = 0;
            row_info[r] = (uint32_t)(uint8_t)g.letter[v] | (1u << 8);
        } else {
            for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x]) preds[at++] = g.node_to_rank[g.e_begin[x]] + 1;
            row_info[r] = (uint32_t)(uint8_t)g.letter[v] | ((uint32_t)np << 8);
        }
    }
    J.L = L;
    J.n = n;
}

// stage + H2D + launch + D2H of one group (all narrow or all wide, fits the slot's arena); returns immediately
static void submit(rtl_ctx *ctx, PoaState &P, PoaSlot &S, std::vector<JobRef> &&group, int sm, int sn, int sg, int se,
                   bool keep_alns) {
    const double ts0 = now_ms();
    S.jobs = std::move(group);
    S.keep_alns = keep_alns;
    std::vector<JobRef> &jobs = S.jobs;
    cudaStream_t st = S.stream;
    const bool wide = jobs[0].wide;
    const size_t nj = jobs.size();
    std::sort(jobs.begin(), jobs.end(), [](const JobRef &a, const JobRef &b) {
        return (int64_t)a.L * a.n > (int64_t)b.L * b.n;
    });
    std::vector<size_t> q_off(nj + 1, 0), row_off(nj + 1, 0), pred_off(nj + 1, 0), hf_off(nj + 1, 0), code_off(nj + 1, 0);
    S.aln_off.assign(nj + 1, 0);
    for (size_t i = 0; i < nj; ++i) {
        const JobRef &jr = jobs[i];
        q_off[i + 1] = q_off[i] + poa_lp(jr.L) + 4;
        row_off[i + 1] = row_off[i] + jr.n + 1;
        pred_off[i + 1] = pred_off[i] + jr.task->g.e_begin.size() + (size_t)jr.n;  // upper bound (sources count 1 each)
        S.aln_off[i + 1] = S.aln_off[i] + jr.n + jr.L + 8;
        hf_off[i + 1] = hf_off[i] + jr.hf_cells;
        code_off[i + 1] = code_off[i] + jr.codes;
    }
    if (q_off[nj] >= (1ull << 32) || row_off[nj] >= (1ull << 32) || pred_off[nj] >= (1ull << 32) ||
        S.aln_off[nj] >= (1ull << 31))
        throw CapacityError("POA batch too large for 32-bit staging offsets");
    uint8_t *hq = S.h_q.need_geo(q_off[nj]);
    uint32_t *hri = S.h_row_info.need_geo(row_off[nj]);
    uint32_t *hrp = S.h_row_poff.need_geo(row_off[nj]);
    int32_t *hpr = S.h_preds.need_geo(pred_off[nj]);
    PoaJob *hj = S.h_jobs.need_geo(nj);
    const std::vector<size_t> &aln_off = S.aln_off;
    parallel_for(P.n_threads, nj, [&](size_t i) {
        PoaJob &J = hj[i];
        J.hf_off = hf_off[i];
        J.code_off = code_off[i];
        J.q_off = (uint32_t)q_off[i];
        J.row_off = (uint32_t)row_off[i];
        J.pred_base = (uint32_t)pred_off[i];
        J.aln_off = (uint32_t)aln_off[i];
        stage_job(jobs[i], J, hq + q_off[i], hri + row_off[i], hrp + row_off[i], hpr + pred_off[i]);
    });
    CK(cudaMemcpyAsync(S.d_q.need_geo(q_off[nj]), hq, q_off[nj], cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_row_info.need_geo(row_off[nj]), hri, row_off[nj] * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_row_poff.need_geo(row_off[nj]), hrp, row_off[nj] * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_preds.need_geo(pred_off[nj]), hpr, pred_off[nj] * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_jobs.need_geo(nj), hj, nj * sizeof(PoaJob), cudaMemcpyHostToDevice, st));
    ctx->stats.h2d_bytes += (int64_t)(q_off[nj] + row_off[nj] * 8 + pred_off[nj] * 4 + nj * sizeof(PoaJob));
    S.d_aln.need_geo(aln_off[nj] * 2);
    S.d_aln_len.need_geo(nj);
    CK(cudaMemsetAsync(S.d_counter.need_geo(1), 0, 4, st));
    const int occ = std::max(1, P.occ[wide ? 1 : 0]);
    const int grid = (int)std::min<size_t>(nj, (size_t)ctx->n_sm * occ);
    CK(cudaEventRecord(S.ev0, st));
    if (!wide)
        k_poa_align<false><<<grid, POA_T, 0, st>>>(S.d_jobs.p, (int)nj, S.d_q.p, S.d_row_info.p, S.d_row_poff.p, S.d_preds.p,
                                                   (short2 *)S.hf, (uint16_t *)S.code, S.d_aln.p, S.d_aln_len.p, sm, sn,
                                                   sg, se, S.d_counter.p);
    else
        k_poa_align<true><<<grid, POA_T, 0, st>>>(S.d_jobs.p, (int)nj, S.d_q.p, S.d_row_info.p, S.d_row_poff.p, S.d_preds.p,
                                                  (int2 *)S.hf, (uint32_t *)S.code, S.d_aln.p, S.d_aln_len.p, sm, sn, sg,
                                                  se, S.d_counter.p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(S.ev1, st));
    int32_t *haln = S.h_aln.need_geo(aln_off[nj] * 2);
    int32_t *hlen = S.h_aln_len.need_geo(nj);
    CK(cudaMemcpyAsync(hlen, S.d_aln_len.p, nj * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(haln, S.d_aln.p, aln_off[nj] * 8, cudaMemcpyDeviceToHost, st));
    S.pending = true;
    ctx->stats.poa_launches++;
    ctx->stats.kernel_launches++;
    ctx->stats.poa_alignments += (int64_t)nj;
    ctx->stats.d2h_bytes += (int64_t)(aln_off[nj] * 8 + nj * 4);
    for (size_t i = 0; i < nj; ++i) ctx->stats.poa_cells += (int64_t)jobs[i].L * jobs[i].n;
    g_t_stage += now_ms() - ts0;
}

// wait for the slot's launch and fold its alignments into the graphs (host threads)
static void finish(rtl_ctx *ctx, PoaState &P, PoaSlot &S) {
    if (!S.pending) return;
    const double tw0 = now_ms();
    CK(cudaStreamSynchronize(S.stream));
    const double tw1 = now_ms();
    g_t_finish_wait += tw1 - tw0;
    S.pending = false;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, S.ev0, S.ev1));
    ctx->stats.poa_ms += ms;
    const int32_t *haln = S.h_aln.p;
    const int32_t *hlen = S.h_aln_len.p;
    const bool keep = S.keep_alns;
    parallel_for(P.n_threads, S.jobs.size(), [&](size_t i) {
        const JobRef &jr = S.jobs[i];
        PoaGraph &g = jr.task->g;
        const int len = hlen[i];
        const int32_t *src = haln + 2 * S.aln_off[i];
        std::vector<std::pair<int32_t, int32_t>> aln((size_t)len);
        for (int x = 0; x < len; ++x) {  // reverse (sisd_alignment_engine.cpp:655) and map rows to node ids
            const int row = src[2 * (len - 1 - x)], pos = src[2 * (len - 1 - x) + 1];
            aln[x].first = row < 0 ? -1 : g.rank_to_node[row - 1];
            aln[x].second = pos;
        }
        g.add_alignment(aln, jr.task->seq[jr.seq_index], jr.L);
        if (keep) jr.task->alns[jr.seq_index] = std::move(aln);
    });
    g_t_fold += now_ms() - tw1;
    S.jobs.clear();
}

// All tasks advance in lock-step.  The tasks are split into two units that alternate between the two slots:
// unit u's step s runs on the device while unit 1-u's step s (or s-1) is folded and re-staged on the host.
void poa_run(rtl_ctx *ctx, std::vector<PoaTask *> &tasks, int sm, int sn, int sg, int se, bool keep_alns) {
    PoaState &P = pstate(ctx);
    g_t_finish_wait = g_t_fold = g_t_stage = 0;
    CK(cudaStreamSynchronize(ctx->stream));  // inputs produced on the ctx stream are complete
    size_t max_steps = 0;
    for (auto *t : tasks) {
        max_steps = std::max(max_steps, t->seq.size());
        t->g.clear();
        if (keep_alns) t->alns.assign(t->seq.size(), {});
    }
    const int maxabs = std::max(std::max(std::abs(sm), std::abs(sn)), std::max(std::abs(sg), std::abs(se)));
    // two units of similar total work (tasks arrive in cluster order; alternate)
    std::vector<PoaTask *> unit[2];
    for (size_t i = 0; i < tasks.size(); ++i) unit[tasks.size() > 1 ? (i & 1) : 0].push_back(tasks[i]);
    for (size_t step = 0; step < max_steps; ++step) {
        for (int u = 0; u < 2; ++u) {
            PoaSlot &S = P.slot[u];
            finish(ctx, P, S);  // step-1 of this unit
            std::vector<JobRef> narrow, wide;
            std::vector<PoaTask *> direct;
            for (auto *t : unit[u]) {
                if (step >= t->seq.size()) continue;
                const int L = t->len[step];
                // simd_alignment_engine.cpp:652-654: empty graph or empty sequence -> empty alignment
                if (t->g.n_nodes() == 0 || L == 0) {
                    direct.push_back(t);
                    continue;
                }
                JobRef jr;
                jr.task = t;
                jr.seq_index = (int)step;
                jr.L = L;
                jr.n = t->g.n_nodes();
                jr.wide = t->g.max_in_degree > 32 || (int64_t)maxabs * (L + 16) >= 32000;
                jr.hf_cells = (size_t)(jr.n + 1) * poa_ws(L);
                jr.codes = (size_t)jr.n * poa_lp(L);
                (jr.wide ? wide : narrow).push_back(jr);
            }
            parallel_for(P.n_threads, direct.size(), [&](size_t i) {
                PoaTask *t = direct[i];
                t->g.add_alignment({}, t->seq[step], t->len[step]);
            });
            // groups that fit this slot's arena; every group but the last is completed synchronously
            std::vector<std::vector<JobRef>> groups;
            for (int w = 0; w < 2; ++w) {
                std::vector<JobRef> &all = w ? wide : narrow;
                const size_t cell_b = w ? 8 : 4, code_b = w ? 4 : 2;
                size_t i = 0;
                while (i < all.size()) {
                    std::vector<JobRef> group;
                    size_t hf = 0, cd = 0;
                    while (i < all.size()) {
                        const size_t nh = hf + all[i].hf_cells * cell_b, nc = cd + all[i].codes * code_b;
                        if (nh > S.hf_bytes || nc > S.code_bytes) {
                            if (group.empty())
                                throw CapacityError("one POA alignment does not fit the device arena: raise option poa_arena_mb");
                            break;
                        }
                        hf = nh;
                        cd = nc;
                        group.push_back(all[i++]);
                    }
                    groups.push_back(std::move(group));
                }
            }
            for (size_t gi = 0; gi < groups.size(); ++gi) {
                submit(ctx, P, S, std::move(groups[gi]), sm, sn, sg, se, keep_alns);
                if (gi + 1 < groups.size()) finish(ctx, P, S);
            }
        }
    }
    finish(ctx, P, P.slot[0]);
    finish(ctx, P, P.slot[1]);
    if (getenv("RTL_TRACE"))
        fprintf(stderr, "[rtl] poa_run: host waited for GPU %.1f ms, fold %.1f ms, stage+submit %.1f ms, host threads %d\n",
                g_t_finish_wait, g_t_fold, g_t_stage, P.n_threads);
}

int poa_msa(rtl_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n, int m, int nn, int g, int e,
            char *msa_out, int64_t cap, int *msa_cols, int64_t *aln_off, int32_t *aln_pairs, int64_t aln_cap) {
    const double t0 = now_ms();
    PoaTask task;
    for (uint32_t i = 0; i < n; ++i) {
        task.seq.push_back(bases + offsets[i]);
        task.len.push_back((int)(offsets[i + 1] - offsets[i]));
    }
    std::vector<PoaTask *> tasks{&task};
    ctx->stats = rtl_stats{};
    poa_run(ctx, tasks, m, nn, g, e, aln_off != nullptr);
    std::vector<std::string> msa;
    task.g.msa(msa);
    *msa_cols = msa.empty() ? 0 : (int)msa[0].size();
    if ((int64_t)msa.size() * (*msa_cols) > cap) throw CapacityError("msa_out too small");
    for (size_t i = 0; i < msa.size(); ++i) memcpy(msa_out + i * (size_t)(*msa_cols), msa[i].data(), *msa_cols);
    if (aln_off) {
        int64_t ao = 0;
        for (uint32_t i = 0; i < n; ++i) {
            aln_off[i] = ao;
            for (auto &p : task.alns[i]) {
                if (ao + 1 <= aln_cap) {
                    aln_pairs[2 * ao] = p.first;
                    aln_pairs[2 * ao + 1] = p.second;
                }
                ++ao;
            }
        }
        aln_off[n] = ao;
        if (ao > aln_cap) throw CapacityError("aln_pairs too small");
    }
    ctx->stats.total_ms = now_ms() - t0;
    return (int)msa.size();
}
