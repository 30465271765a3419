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
#include "poa_strip_kernel.cuh"

using namespace rtl;

static double g_t_finish_wait = 0, g_t_fold = 0, g_t_stage = 0;  // RTL_TRACE phase timers

enum JobKind { JK_STRIP = 0, JK_NARROW = 1, JK_WIDE = 2 };

struct JobRef {
    PoaTask *task;
    int seq_index;
    int kind;                    // JobKind
    int nst;                     // JK_STRIP: strips of 256 query columns
    size_t hf_bytes, code_bytes; // device arena need
    int L, n;
};

// One pipeline slot: its own staging buffers, stream and half of the device arena.  While the kernel of one slot
// runs, the host folds the alignments of the other slot into their graphs and stages that slot's next step.
struct PoaSlot {
    DevBuf<uint8_t> d_q;
    DevBuf<uint32_t> d_row_info, d_row_poff;
    DevBuf<int32_t> d_preds, d_aln, d_aln_len;
    DevBuf<PoaJob> d_jobs;
    DevBuf<PoaSJob> d_sjobs;
    DevBuf<uint4> d_rec;
    DevBuf<unsigned int> d_counter;
    PinBuf<uint8_t> h_q;
    PinBuf<uint32_t> h_row_info, h_row_poff;
    PinBuf<int32_t> h_preds, h_aln, h_aln_len;
    PinBuf<PoaJob> h_jobs;
    PinBuf<PoaSJob> h_sjobs;
    PinBuf<uint4> h_rec;
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

// rank-order CSR of the graph + query into the staging buffers (int32 kernel)
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

// letter -> code of the strip kernel's score profile (255 = not representable: the int32 kernel takes the job)
static const uint8_t *letter_codes() {
    static uint8_t tab[256];
    static bool init = false;
    if (!init) {
        memset(tab, 255, sizeof(tab));
        tab[(unsigned char)'A'] = 0;
        tab[(unsigned char)'C'] = 1;
        tab[(unsigned char)'G'] = 2;
        tab[(unsigned char)'T'] = 3;
        tab[(unsigned char)'U'] = 4;
        init = true;
    }
    return tab;
}

// strip kernel: letter codes of the query (padded with 255 to whole strips), one 16-byte record per row
// {letter | n_pred << 8, pred0, pred1, pred2 or offset into preds}, and the predecessor CSR for rows with > 3
static void stage_strip_job(const JobRef &jr, PoaSJob &J, uint8_t *q, uint4 *rec, int32_t *preds) {
    const PoaGraph &g = jr.task->g;
    const uint8_t *tab = letter_codes();
    const int L = jr.L, n = jr.n;
    const char *src = jr.task->seq[jr.seq_index];
    for (int i = 0; i < L; ++i) q[i] = tab[(unsigned char)src[i]];
    memset(q + L, 255, (size_t)jr.nst * PS_STRIP - L);
    rec[0] = make_uint4(0, 0, 0, 0);
    uint32_t at = 0;
    for (int r = 1; r <= n; ++r) {
        const int v = g.rank_to_node[r - 1];
        const int np = g.n_in[v];
        uint4 rc;
        rc.x = (uint32_t)tab[(unsigned char)g.letter[v]] | ((uint32_t)(np == 0 ? 1 : np) << 8);
        rc.y = rc.z = rc.w = 0;
        if (np <= 3) {
            uint32_t *dst = &rc.y;
            int k = 0;
            for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x]) dst[k++] = (uint32_t)(g.node_to_rank[g.e_begin[x]] + 1);
        } else {
            rc.w = at;
            int k = 0;
            for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x], ++k) {
                const int pr = g.node_to_rank[g.e_begin[x]] + 1;
                preds[at++] = pr;
                if (k == 0) rc.y = (uint32_t)pr;
                if (k == 1) rc.z = (uint32_t)pr;
            }
        }
        rec[r] = rc;
    }
    J.L = L;
    J.n = n;
    J.n_strips = jr.nst;
    J.pad = 0;
}

// stage + H2D + launch + D2H of one group (fits the slot's arena); returns immediately.  The jobs are sorted into
// segments of equal kernel configuration (strip kernel by warps per CTA, int32 kernel narrow / wide), each with its
// own job list and work counter; the segments' launches follow each other on the slot's stream.
static void submit(rtl_ctx *ctx, PoaState &P, PoaSlot &S, std::vector<JobRef> &&group, int sm, int sn, int sg, int se,
                   bool keep_alns) {
    const double ts0 = now_ms();
    S.jobs = std::move(group);
    S.keep_alns = keep_alns;
    std::vector<JobRef> &jobs = S.jobs;
    cudaStream_t st = S.stream;
    const size_t nj = jobs.size();
    auto seg_key = [](const JobRef &a) { return a.kind == JK_STRIP ? std::min(a.nst, PS_MAXW) : 100 + a.kind; };
    std::sort(jobs.begin(), jobs.end(), [&](const JobRef &a, const JobRef &b) {
        const int ka = seg_key(a), kb = seg_key(b);
        if (ka != kb) return ka > kb;  // widest CTAs first
        return (int64_t)a.L * a.n > (int64_t)b.L * b.n;
    });
    struct Seg {
        size_t begin, end;
        int key;
    };
    std::vector<Seg> segs;
    for (size_t i = 0; i < nj; ++i) {
        const int k = seg_key(jobs[i]);
        if (segs.empty() || segs.back().key != k) segs.push_back(Seg{i, i, k});
        segs.back().end = i + 1;
    }
    // offsets: strip jobs and int32-kernel jobs use separate staging arrays, one shared arena and output
    std::vector<size_t> q_off(nj + 1, 0), row_off(nj, 0), pred_off(nj, 0), hf_off(nj + 1, 0), code_off(nj + 1, 0);
    S.aln_off.assign(nj + 1, 0);
    size_t rows_old = 0, rows_strip = 0, preds_total = 0;
    for (size_t i = 0; i < nj; ++i) {
        const JobRef &jr = jobs[i];
        const bool strip = jr.kind == JK_STRIP;
        q_off[i + 1] = q_off[i] + (strip ? (size_t)jr.nst * PS_STRIP : (size_t)((poa_lp(jr.L) + 4 + 15) & ~15));
        row_off[i] = strip ? rows_strip : rows_old;
        (strip ? rows_strip : rows_old) += (size_t)jr.n + 1;
        pred_off[i] = preds_total;
        preds_total += jr.task->g.e_begin.size() + (size_t)jr.n;  // upper bound (sources count 1 each)
        S.aln_off[i + 1] = S.aln_off[i] + jr.n + jr.L + 8;
        hf_off[i + 1] = hf_off[i] + ((jr.hf_bytes + 255) & ~(size_t)255);
        code_off[i + 1] = code_off[i] + ((jr.code_bytes + 255) & ~(size_t)255);
    }
    if (q_off[nj] >= (1ull << 32) || rows_old >= (1ull << 32) || rows_strip >= (1ull << 32) ||
        preds_total >= (1ull << 32) || S.aln_off[nj] >= (1ull << 31))
        throw CapacityError("POA batch too large for 32-bit staging offsets");
    uint8_t *hq = S.h_q.need_geo(q_off[nj] + 16);
    uint32_t *hri = S.h_row_info.need_geo(rows_old + 1);
    uint32_t *hrp = S.h_row_poff.need_geo(rows_old + 1);
    uint4 *hrec = S.h_rec.need_geo(rows_strip + 1);
    int32_t *hpr = S.h_preds.need_geo(preds_total + 1);
    PoaJob *hj = S.h_jobs.need_geo(nj);
    PoaSJob *hsj = S.h_sjobs.need_geo(nj);
    const std::vector<size_t> &aln_off = S.aln_off;
    parallel_for(P.n_threads, nj, [&](size_t i) {
        const JobRef &jr = jobs[i];
        if (jr.kind == JK_STRIP) {
            PoaSJob &J = hsj[i];
            J.hf_off = hf_off[i] / 4;
            J.code_off = code_off[i] / 4;
            J.q_off = (uint32_t)q_off[i];
            J.row_off = (uint32_t)row_off[i];
            J.pred_base = (uint32_t)pred_off[i];
            J.aln_off = (uint32_t)aln_off[i];
            stage_strip_job(jr, J, hq + q_off[i], hrec + row_off[i], hpr + pred_off[i]);
        } else {
            PoaJob &J = hj[i];
            J.hf_off = hf_off[i] / (jr.kind == JK_WIDE ? 8 : 4);
            J.code_off = code_off[i] / (jr.kind == JK_WIDE ? 4 : 2);
            J.q_off = (uint32_t)q_off[i];
            J.row_off = (uint32_t)row_off[i];
            J.pred_base = (uint32_t)pred_off[i];
            J.aln_off = (uint32_t)aln_off[i];
            stage_job(jr, J, hq + q_off[i], hri + row_off[i], hrp + row_off[i], hpr + pred_off[i]);
        }
    });
    CK(cudaMemcpyAsync(S.d_q.need_geo(q_off[nj] + 16), hq, q_off[nj], cudaMemcpyHostToDevice, st));
    S.d_row_info.need_geo(rows_old + 1);
    S.d_row_poff.need_geo(rows_old + 1);
    if (rows_old) {
        CK(cudaMemcpyAsync(S.d_row_info.p, hri, rows_old * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.d_row_poff.p, hrp, rows_old * 4, cudaMemcpyHostToDevice, st));
    }
    S.d_rec.need_geo(rows_strip + 1);
    if (rows_strip) CK(cudaMemcpyAsync(S.d_rec.p, hrec, rows_strip * sizeof(uint4), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_preds.need_geo(preds_total + 1), hpr, preds_total * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_jobs.need_geo(nj), hj, nj * sizeof(PoaJob), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_sjobs.need_geo(nj), hsj, nj * sizeof(PoaSJob), cudaMemcpyHostToDevice, st));
    ctx->stats.h2d_bytes += (int64_t)(q_off[nj] + rows_old * 8 + rows_strip * 16 + preds_total * 4 +
                                      nj * (sizeof(PoaJob) + sizeof(PoaSJob)));
    S.d_aln.need_geo(aln_off[nj] * 2);
    S.d_aln_len.need_geo(nj);
    CK(cudaMemsetAsync(S.d_counter.need_geo(segs.size()), 0, 4 * segs.size(), st));
    CK(cudaEventRecord(S.ev0, st));
    for (size_t si = 0; si < segs.size(); ++si) {
        const Seg &sg_ = segs[si];
        const int cnt = (int)(sg_.end - sg_.begin);
        unsigned int *counter = S.d_counter.p + si;
        if (sg_.key < 100) {
            const int nw = sg_.key;
            const size_t smem = (size_t)nw * PS_NLET * 32 * sizeof(uint4);
            int occ = 1;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_poa_strip, nw * 32, smem));
            const int grid = (int)std::min<size_t>((size_t)cnt, (size_t)ctx->n_sm * std::max(1, occ));
            k_poa_strip<<<grid, nw * 32, smem, st>>>(S.d_sjobs.p + sg_.begin, cnt, S.d_q.p, S.d_rec.p, S.d_preds.p,
                                                    (uint32_t *)S.hf, (uint32_t *)S.code, S.d_aln.p,
                                                    S.d_aln_len.p + sg_.begin, sm, sn, sg, se, counter);
        } else {
            const bool wide = sg_.key == 100 + JK_WIDE;
            const int occ = std::max(1, P.occ[wide ? 1 : 0]);
            const int grid = (int)std::min<size_t>((size_t)cnt, (size_t)ctx->n_sm * occ);
            if (!wide)
                k_poa_align<false><<<grid, POA_T, 0, st>>>(S.d_jobs.p + sg_.begin, cnt, S.d_q.p, S.d_row_info.p,
                                                           S.d_row_poff.p, S.d_preds.p, (short2 *)S.hf, (uint16_t *)S.code,
                                                           S.d_aln.p, S.d_aln_len.p + sg_.begin, sm, sn, sg, se, counter);
            else
                k_poa_align<true><<<grid, POA_T, 0, st>>>(S.d_jobs.p + sg_.begin, cnt, S.d_q.p, S.d_row_info.p,
                                                          S.d_row_poff.p, S.d_preds.p, (int2 *)S.hf, (uint32_t *)S.code,
                                                          S.d_aln.p, S.d_aln_len.p + sg_.begin, sm, sn, sg, se, counter);
        }
        CK(cudaGetLastError());
        ctx->stats.poa_launches++;
        ctx->stats.kernel_launches++;
    }
    CK(cudaEventRecord(S.ev1, st));
    int32_t *haln = S.h_aln.need_geo(aln_off[nj] * 2);
    int32_t *hlen = S.h_aln_len.need_geo(nj);
    CK(cudaMemcpyAsync(hlen, S.d_aln_len.p, nj * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(haln, S.d_aln.p, aln_off[nj] * 8, cudaMemcpyDeviceToHost, st));
    S.pending = true;
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
    // the int16 strip kernel assumes m > 0 > n,g,e of small magnitude (poa_strip_kernel.cuh); option poa_kernel=1
    // forces the int32 kernel
    const bool strip_scores = ctx->poa_kernel != 1 && sm > 0 && sn < 0 && sg < 0 && se < 0 && maxabs < 100;
    {
        const uint8_t *tab = letter_codes();
        parallel_for(P.n_threads, tasks.size(), [&](size_t i) {
            PoaTask *t = tasks[i];
            bool ok = true;
            for (size_t s = 0; s < t->seq.size() && ok; ++s)
                for (int x = 0; x < t->len[s]; ++x)
                    if (tab[(unsigned char)t->seq[s][x]] == 255) {
                        ok = false;
                        break;
                    }
            t->acgtu = ok;
        });
    }
    // two units of similar total work (tasks arrive in cluster order; alternate)
    std::vector<PoaTask *> unit[2];
    for (size_t i = 0; i < tasks.size(); ++i) unit[tasks.size() > 1 ? (i & 1) : 0].push_back(tasks[i]);
    for (size_t step = 0; step < max_steps; ++step) {
        for (int u = 0; u < 2; ++u) {
            PoaSlot &S = P.slot[u];
            finish(ctx, P, S);  // step-1 of this unit
            std::vector<JobRef> all;
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
                const bool wide = t->g.max_in_degree > 32 || (int64_t)maxabs * (L + 16) >= 32000;
                jr.kind = wide ? JK_WIDE : ((strip_scores && t->acgtu) ? JK_STRIP : JK_NARROW);
                jr.nst = (L + PS_STRIP - 1) / PS_STRIP;
                if (jr.kind == JK_STRIP) {
                    jr.hf_bytes = ps_hf_words(jr.n, jr.nst) * 4;
                    jr.code_bytes = ps_code_words(jr.n, jr.nst) * 4;
                } else {
                    jr.hf_bytes = (size_t)(jr.n + 1) * poa_ws(L) * (wide ? 8 : 4);
                    jr.code_bytes = (size_t)jr.n * poa_lp(L) * (wide ? 4 : 2);
                }
                all.push_back(jr);
            }
            parallel_for(P.n_threads, direct.size(), [&](size_t i) {
                PoaTask *t = direct[i];
                t->g.add_alignment({}, t->seq[step], t->len[step]);
            });
            // groups that fit this slot's arena; every group but the last is completed synchronously
            std::vector<std::vector<JobRef>> groups;
            {
                size_t i = 0;
                while (i < all.size()) {
                    std::vector<JobRef> group;
                    size_t hf = 0, cd = 0;
                    while (i < all.size()) {
                        const size_t nh = hf + ((all[i].hf_bytes + 255) & ~(size_t)255);
                        const size_t nc = cd + ((all[i].code_bytes + 255) & ~(size_t)255);
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
