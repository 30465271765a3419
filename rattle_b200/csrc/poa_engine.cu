// Batched POA driver of hot path B.  The partial-order alignments of many read packs are independent chains
// (read s of a pack is aligned to the graph built from reads 0..s-1: correct.cpp:399-402 / :430-433 / :525-528).
// The packs are split into UNITS; each unit is driven by its own host thread on its own CUDA stream and slice of
// the device arena: step s of a unit aligns the s-th sequence of every pack of the unit in one launch group
// (one CTA per alignment), then the thread folds the alignments into the graphs (PoaGraph::add_alignment, which
// also re-sorts the graph) and stages the next step in rank-order CSR form.  Units do not wait for each other, so
// the kernels of some units keep the GPU busy while other units are in their host phase; callers with more host
// work per pack (correct_engine.cu) run that work inside the unit's thread as well.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

#include "common.cuh"
#include "poa_engine.hpp"
#include "poa_kernels.cuh"
#include "poa_strip_kernel.cuh"
#include "poa_devgraph.cuh"
#include "poa_devchain.cuh"
#include "poa_vote.cuh"

using namespace rtl;


enum JobKind { JK_STRIP = 0, JK_NARROW = 1, JK_WIDE = 2 };

struct JobRef {
    PoaTask *task;
    int seq_index;
    int kind;                    // JobKind
    int nst;                     // JK_STRIP: strips of 256 query columns
    int n_spill = 0;             // JK_STRIP: rows that are also written to HBM (needed more than PS_K rows later)
    size_t hf_bytes, code_bytes; // device arena need
    int L, n;
    // graphs with a device mirror (poa_devgraph.cuh): row records / predecessor words / spill rows are produced on the
    // device by k_poa_graph_fold at these (absolute) offsets of the slot's d_rec / d_preds / d_spill
    bool mirror = false;
    uint32_t rec_off = 0, pred_base = 0, spill_off = 0;
};

// One unit's slot: its own staging buffers, stream and slice of the device arena.
struct PoaSlot {
    DevBuf<uint8_t> d_q;
    DevBuf<uint32_t> d_row_info, d_row_poff;
    DevBuf<int32_t> d_preds, d_aln, d_aln_len;
    DevBuf<PoaJob> d_jobs;
    DevBuf<PoaSJob> d_sjobs;
    DevBuf<uint4> d_rec;
    DevBuf<unsigned int> d_counter;
    DevBuf<int4> d_best;
    DevBuf<int32_t> d_spill;
    PinBuf<int32_t> h_spill;
    // device mirrors of this unit's graphs and the per-step fold staging
    DevBuf<int32_t> d_pool, d_delta, d_counts;
    PinBuf<int32_t> h_delta, h_counts;
    DevBuf<DFoldJob> d_fjobs;
    PinBuf<DFoldJob> h_fjobs;
    size_t base_rows = 0, base_preds = 0, base_spill = 0;  // first host-staged entry of d_rec / d_preds / d_spill this step
    PinBuf<uint8_t> h_q;
    PinBuf<uint32_t> h_row_info, h_row_poff;
    PinBuf<int32_t> h_preds, h_aln, h_aln_len;
    PinBuf<PoaJob> h_jobs;
    PinBuf<PoaSJob> h_sjobs;
    PinBuf<uint4> h_rec;
    cudaStream_t stream = nullptr;
    cudaStream_t hi = nullptr;  // high priority: the small kernels between two chain launches (MSA rows, vote) must not queue
                                // behind the other units' long-running chain CTAs
    static constexpr int N_SUB = 4, N_SEG_EV = 16;
    cudaStream_t sub[N_SUB] = {nullptr, nullptr, nullptr, nullptr};  // segments of one group run side by side
    cudaEvent_t ev_seg[N_SEG_EV] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_h2d = nullptr, ev_dp = nullptr;
    cudaEvent_t ev_wait = nullptr;  // blocking-sync event: a unit's driver thread SLEEPS while its kernels run (a dozen
                                    // spinning cudaStreamSynchronize per rank would eat the cores of an 8-rank node)
    double t_submit0 = 0, t_submit1 = 0;  // host clock around the last submit (timeline trace)
    int epoch = 0;                        // chains run on this slot so far (timeline trace)
    unsigned char *arena = nullptr;  // this unit's slice of the device arena (score rows + traceback codes)
    size_t arena_bytes = 0;
    std::vector<JobRef> jobs;  // in flight
    std::vector<size_t> aln_off;
    bool pending = false;
    bool keep_alns = false;
    int n_threads = 1;                                  // host threads of the unit driving this slot
    double t_wait = 0, t_fold = 0, t_stage = 0;         // RTL_TRACE phase timers of the running chain
    rtl_stats st{};                                     // counters of the running chain (merged under stats_mu)
    // device-resident chains (poa_devchain.cuh): per-pack descriptors, sequence table, fixed per-pack slots
    DevBuf<DCPack> c_packs;
    PinBuf<DCPack> ch_packs;
    DevBuf<DCSeq> c_seqs;
    PinBuf<DCSeq> ch_seqs;
    DevBuf<uint8_t> c_q;
    PinBuf<uint8_t> ch_q;
    DevBuf<int32_t> c_pool, c_preds, c_spill, c_aln, c_path, c_qnode, c_list;
    DevBuf<uint4> c_rec;
    DevBuf<unsigned int> c_counter;
    DevBuf<unsigned long long> c_stats;
    PinBuf<unsigned long long> ch_stats;
    DevBuf<uint64_t> c_msa_off;
    PinBuf<uint64_t> ch_msa_off;
    DevBuf<char> c_msa;
    PinBuf<char> ch_msa;
    // MSA post-processing on the device (poa_vote.cuh)
    DevBuf<uint8_t> c_qual;            // the reads' qualities, laid out like c_q
    PinBuf<uint8_t> ch_qual;
    DevBuf<DVPack> v_packs;
    PinBuf<DVPack> vh_packs;
    DevBuf<DVRead> v_reads;
    PinBuf<DVRead> vh_reads;
    DevBuf<DVRow> v_rows;
    PinBuf<DVRow> vh_rows;
    DevBuf<DVCol> v_cols;
    DevBuf<char> v_qm, v_out_seq, v_out_qual, v_cons;
    PinBuf<char> vh_out_seq, vh_out_qual, vh_cons;
    DevBuf<int32_t> v_len;
    PinBuf<int32_t> vh_len;
    DevBuf<double> v_tab;
    DevBuf<unsigned char> v_symtab;
    bool v_tab_ready = false;
    DevBuf<int4> v_flagged;            // (pack, column, cerr as two words)
    PinBuf<int4> vh_flagged;
    DevBuf<unsigned int> v_nflag;
    PinBuf<unsigned int> vh_nflag;
    DevBuf<DVStage> v_stage;
    PinBuf<DVStage> vh_stage;
};

constexpr int POA_MAX_UNITS = 16;

struct PoaState {
    DevBuf<unsigned char> arena;
    PoaSlot slot_store[POA_MAX_UNITS];
    int n_units = 0;
    int occ[2] = {0, 0};
    int n_threads = 0;
    std::mutex stats_mu;
    cudaEvent_t ev_ref = nullptr;  // device-time origin of the kernel intervals below
    double t_ref = 0;
    FILE *trace = nullptr;         // RTL_TRACE_FILE: one JSON line per launch group (tools/timeline.py)
    std::vector<std::pair<float, float>> intervals;  // [first kernel start, last kernel end] of every launch group
};

void poa_state_free(rtl_ctx *ctx) {
    if (ctx->poa) {
        for (auto &sl : ctx->poa->slot_store) {
            if (sl.ev0) cudaEventDestroy(sl.ev0);
            if (sl.ev1) cudaEventDestroy(sl.ev1);
            if (sl.stream) cudaStreamDestroy(sl.stream);
            if (sl.hi) cudaStreamDestroy(sl.hi);
            for (auto &x : sl.sub)
                if (x) cudaStreamDestroy(x);
            for (auto &x : sl.ev_seg)
                if (x) cudaEventDestroy(x);
            if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
            if (sl.ev_dp) cudaEventDestroy(sl.ev_dp);
            if (sl.ev_wait) cudaEventDestroy(sl.ev_wait);
        }
        delete ctx->poa;
    }
    ctx->poa = nullptr;
}

int host_threads();

// Shared worker pool: every parallel_for of every unit thread feeds the same workers, so that the cores follow
// the units that are in their host phase (a unit waiting for its kernel uses none).  A job is an index range with
// an atomic cursor; the caller works on its own job too and returns when all indices are done.
namespace {
struct PoolJob {
    const std::function<void(size_t)> *fn;
    size_t n;
    int max_workers;
    std::atomic<size_t> next{0}, done{0};
    std::atomic<int> workers{0};
    std::exception_ptr err;
    std::mutex err_mu;
    void run_some() {
        while (true) {
            const size_t i = next.fetch_add(1);
            if (i >= n) break;
            try {
                (*fn)(i);
            } catch (...) {
                std::lock_guard<std::mutex> lk(err_mu);
                if (!err) err = std::current_exception();
            }
            done.fetch_add(1);
        }
    }
};

class WorkerPool {
  public:
    static WorkerPool &get() {
        static WorkerPool *p = new WorkerPool();  // lives until process exit (workers are detached)
        return *p;
    }
    void run(PoolJob &job) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_.push_back(&job);
        }
        cv_.notify_all();
        job.run_some();
        // workers that took the job are still inside run_some(): sleep until the last one reports (no spinning — the
        // cores are needed by the other units' host phases and, with several ranks per node, by the other ranks)
        std::unique_lock<std::mutex> lk(mu_);
        jobs_.erase(std::find(jobs_.begin(), jobs_.end(), &job));
        done_cv_.wait(lk, [&]() { return job.done.load() >= job.n && job.workers.load() == 0; });
    }

  private:
    WorkerPool() {
        const int n = std::max(1, host_threads() - 1);
        for (int i = 0; i < n; ++i) std::thread([this]() { loop(); }).detach();
    }
    void loop() {
        std::unique_lock<std::mutex> lk(mu_);
        while (true) {
            PoolJob *job = nullptr;
            for (PoolJob *j : jobs_)
                if (j->next.load() < j->n && j->workers.load() < j->max_workers) {
                    job = j;
                    break;
                }
            if (!job) {
                cv_.wait(lk);  // woken by run() when a job is posted (posted under mu_: no lost wake-up)
                continue;
            }
            job->workers.fetch_add(1);
            lk.unlock();
            job->run_some();
            lk.lock();
            job->workers.fetch_sub(1);  // under mu_, so that the owner's predicate check cannot miss it
            done_cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<PoolJob *> jobs_;
};
}  // namespace

void parallel_for(int n_threads, size_t n, const std::function<void(size_t)> &fn) {
    if (n == 0) return;
    if (n_threads <= 1 || n == 1) {
        for (size_t i = 0; i < n; ++i) fn(i);
        return;
    }
    PoolJob job;
    job.fn = &fn;
    job.n = n;
    job.max_workers = n_threads - 1;  // plus the caller
    WorkerPool::get().run(job);
    if (job.err) std::rethrow_exception(job.err);
}

// Host threads of this process: all cores, or this rank's share of them when several ranks run on one node
// (torchrun exports LOCAL_WORLD_SIZE; RATTLE_B200_THREADS overrides) — oversubscribed cores only add contention.
int host_threads() {
    if (const char *e = getenv("RATTLE_B200_THREADS")) {
        const int v = atoi(e);
        if (v > 0) return std::min(v, 256);
    }
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
        const int w = atoi(e);
        if (w > 1) hc = std::max(4u, hc / (unsigned)w);
    }
    return (int)std::min(hc, 64u);
}

// wait for everything enqueued on the slot's stream without spinning
static void slot_wait(PoaSlot &S, cudaStream_t st = nullptr) {
    CK(cudaEventRecord(S.ev_wait, st ? st : S.stream));
    CK(cudaEventSynchronize(S.ev_wait));
}

static PoaState &pstate(rtl_ctx *ctx) {
    if (!ctx->poa) {
        // one context at a time: contexts of one process may share a device (the drop-in's device slots), and each sizes
        // its arena from the memory that is free at that moment
        static std::mutex create_mu;
        std::lock_guard<std::mutex> create_lock(create_mu);
        ctx->poa = new PoaState();
        PoaState &P = *ctx->poa;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&P.occ[0], k_poa_align<false>, POA_T, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&P.occ[1], k_poa_align<true>, POA_T, 0));
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        size_t budget = ctx->poa_arena_mb > 0 ? ((size_t)ctx->poa_arena_mb << 20) : std::min<size_t>(free_b / 2, 96ull << 30);
        budget = std::max<size_t>(budget, 64ull << 20);
        const int n_units = std::max(1, std::min(ctx->poa_units > 0 ? ctx->poa_units : 12, POA_MAX_UNITS));
        size_t part = (budget / n_units) & ~(size_t)255;
        while (true) {  // (another process may have taken memory since cudaMemGetInfo: settle for less)
            try {
                P.arena.need(n_units * part);
                break;
            } catch (const CudaError &) {
                cudaGetLastError();
                if (ctx->poa_arena_mb > 0 || part * n_units <= (256ull << 20)) throw;
                part = (part / 2) & ~(size_t)255;
            }
        }
        P.n_units = n_units;
        for (int i = 0; i < n_units; ++i) {
            PoaSlot &sl = P.slot_store[i];
            CK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
            {
                int lo = 0, hi = 0;
                CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                CK(cudaStreamCreateWithPriority(&sl.hi, cudaStreamNonBlocking, hi));
            }
            CK(cudaEventCreate(&sl.ev0));
            CK(cudaEventCreate(&sl.ev1));
            CK(cudaEventCreate(&sl.ev_h2d));
            CK(cudaEventCreate(&sl.ev_dp));
            CK(cudaEventCreateWithFlags(&sl.ev_wait, cudaEventBlockingSync | cudaEventDisableTiming));
            for (auto &x : sl.sub) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
            for (auto &x : sl.ev_seg) CK(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
            sl.arena = P.arena.p + i * part;
            sl.arena_bytes = part;
        }
        CK(cudaFuncSetAttribute(k_poa_strip<5, -4, -8, -6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)ps_smem_bytes(PS_MAXW, PS_K)));
        CK(cudaFuncSetAttribute(k_poa_chain<5, -4, -8, -6, 256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_poa_chain<5, -4, -8, -6, 224, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_poa_chain<5, -4, -8, -6, 192, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_poa_chain<5, -4, -8, -6, 192, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_poa_chain<5, -4, -8, -6, 256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_poa_chain<5, -4, -8, -6, 128, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaFuncSetAttribute(k_poa_chain<5, -4, -8, -6, 96, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        P.n_threads = host_threads();
        if (const char *tf = getenv("RTL_TRACE_FILE")) P.trace = fopen(tf, "w");
        CK(cudaEventCreate(&P.ev_ref));
        CK(cudaEventRecord(P.ev_ref, ctx->stream));
        CK(cudaEventSynchronize(P.ev_ref));
        P.t_ref = now_ms();
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
    struct Table {
        uint8_t v[256];
        Table() {
            memset(v, 255, sizeof(v));
            v[(unsigned char)'A'] = 0;
            v[(unsigned char)'C'] = 1;
            v[(unsigned char)'G'] = 2;
            v[(unsigned char)'T'] = 3;
            v[(unsigned char)'U'] = 4;
        }
    };
    static const Table t;  // thread-safe initialisation: unit threads call this concurrently
    return t.v;
}

// strip kernel: warps per CTA = strips per pass, passes balanced (9 strips -> 2 passes of 5 and 4, not 8 and 1)
static int strip_warps(int nst) {
    const int n_pass = (nst + PS_MAXW - 1) / PS_MAXW;
    return (nst + n_pass - 1) / n_pass;
}
// Device-resident chains, narrow CTAs: a pack's CTA spends a third of its time in phases that occupy one warp (graph
// update, traceback) while its other warps wait, so CTAs of at most four warps — reads of more than four strips run in
// balanced passes, 6 strips as 3 + 3 — leave fewer warps idle and twice as many packs share an SM.
static int chain_warps(int nst) {
    static const bool narrow = getenv("RATTLE_B200_NARROW") != nullptr && atoi(getenv("RATTLE_B200_NARROW")) != 0;
    if (!narrow) return strip_warps(nst);
    const int n_pass = (nst + 3) / 4;
    return (nst + n_pass - 1) / n_pass;
}
// rows of the shared-memory ring for a CTA of nw warps: 6 where the register file limits the CTAs per SM anyway,
// 5 for 6-warp CTAs, where one row less lets a fourth CTA fit into the SM's shared memory
static int strip_ring_rows(int nw) {
    static const int force = getenv("RATTLE_B200_RING") ? atoi(getenv("RATTLE_B200_RING")) : 0;  // experiments
    if (force >= 2 && force <= PS_K) return force;
    return nw == 6 ? 5 : PS_K;
}

// strip kernel, step 1 (before the arena is divided): which rows must be spilled to HBM?  Row p is read from the
// shared-memory ring by rows up to PS_K ranks later; if a successor is further away the row gets a spill slot
// (slot 0 is the virtual start row).  Result: task->spill_slot[r] for r = 0..n, returns the number of spilled rows.
static int plan_spills(PoaTask *t, int K) {
    const PoaGraph &g = t->g;
    const int n = g.n_nodes();
    std::vector<int32_t> &slot = t->spill_slot;
    slot.assign((size_t)n + 1, 0);
    for (int r = 1; r <= n; ++r) {
        const int v = g.rank_to_node[r - 1];
        if (g.n_in[v] == 0) continue;  // predecessor = virtual start row = spill slot 0 unless near
        for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x]) {
            const int pr = g.node_to_rank[g.e_begin[x]] + 1;
            if (r - pr > K) slot[pr] = 1;
        }
    }
    int ns = 0;
    for (int r = 1; r <= n; ++r)
        if (slot[r]) slot[r] = ++ns;
    return ns;
}

// strip kernel, graphs with a device mirror: only the query is staged by the host
static void stage_strip_query(const JobRef &jr, PoaSJob &J, uint8_t *q) {
    const uint8_t *tab = letter_codes();
    const char *src = jr.task->seq[jr.seq_index];
    for (int i = 0; i < jr.L; ++i) q[i] = tab[(unsigned char)src[i]];
    memset(q + jr.L, 255, (size_t)jr.nst * PS_STRIP - jr.L);
    J.L = jr.L;
    J.n = jr.n;
    J.n_strips = jr.nst;
    J.n_spill = jr.n_spill;
    J.pad = 0;
}

// strip kernel, step 2: letter codes of the query (padded with 255 to whole strips), one 16-byte record per row
// (layout: poa_strip_kernel.cuh), the predecessor words of rows with more than 3 predecessors, and the row of
// every spill slot (for the traceback)
static void stage_strip_job(const JobRef &jr, PoaSJob &J, uint8_t *q, uint4 *rec, int32_t *preds, int32_t *spill_rows) {
    const PoaGraph &g = jr.task->g;
    const std::vector<int32_t> &slot = jr.task->spill_slot;
    const uint8_t *tab = letter_codes();
    const int L = jr.L, n = jr.n;
    const int K = strip_ring_rows(strip_warps(jr.nst));
    const char *src = jr.task->seq[jr.seq_index];
    for (int i = 0; i < L; ++i) q[i] = tab[(unsigned char)src[i]];
    memset(q + L, 255, (size_t)jr.nst * PS_STRIP - L);
    rec[0] = make_uint4(0, 0, 0, 0);
    spill_rows[0] = 0;
    uint32_t at = 0;
    for (int r = 1; r <= n; ++r) {
        const int v = g.rank_to_node[r - 1];
        const int np = g.n_in[v];
        if (slot[r]) spill_rows[slot[r]] = r;
        uint4 rc;
        rc.x = (uint32_t)tab[(unsigned char)g.letter[v]] | ((uint32_t)(np == 0 ? 1 : np) << 8) | ((uint32_t)slot[r] << 16);
        rc.y = rc.z = rc.w = 0;
        auto word = [&](int pr) -> uint32_t { return (r - pr <= K) ? (uint32_t)(r - pr) : (PS_FAR | (uint32_t)slot[pr]); };
        if (np == 0) {
            rc.y = word(0);
        } else if (np <= 3) {
            uint32_t *dst = &rc.y;
            int k = 0;
            for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x]) dst[k++] = word(g.node_to_rank[g.e_begin[x]] + 1);
        } else {
            rc.w = at;
            int k = 0;
            for (int x = g.in_head[v]; x >= 0; x = g.e_next_in[x], ++k) {
                const uint32_t w = word(g.node_to_rank[g.e_begin[x]] + 1);
                preds[at++] = (int32_t)w;
                if (k == 0) rc.y = w;
                if (k == 1) rc.z = w;
            }
        }
        rec[r] = rc;
    }
    J.L = L;
    J.n = n;
    J.n_strips = jr.nst;
    J.n_spill = jr.n_spill;
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
    auto seg_key = [](const JobRef &a) { return a.kind != JK_STRIP ? 100 + a.kind : strip_warps(a.nst); };
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
    std::vector<size_t> q_off(nj + 1, 0), row_off(nj, 0), pred_off(nj, 0), hf_off(nj, 0), code_off(nj, 0), spill_off(nj + 1, 0);
    size_t arena_at = 0;
    S.aln_off.assign(nj + 1, 0);
    size_t rows_old = 0, rows_strip = 0, preds_total = 0;
    for (size_t i = 0; i < nj; ++i) {
        const JobRef &jr = jobs[i];
        const bool strip = jr.kind == JK_STRIP;
        q_off[i + 1] = q_off[i] + (strip ? (size_t)jr.nst * PS_STRIP : (size_t)((poa_lp(jr.L) + 4 + 15) & ~15));
        if (!jr.mirror) {  // host-staged graph description
            row_off[i] = strip ? rows_strip : rows_old;
            (strip ? rows_strip : rows_old) += (size_t)jr.n + 1;
            pred_off[i] = preds_total;
            preds_total += jr.task->g.e_begin.size() + (size_t)jr.n;  // upper bound (sources count 1 each)
        }
        S.aln_off[i + 1] = S.aln_off[i] + jr.n + jr.L + 8;
        hf_off[i] = arena_at;
        arena_at += (jr.hf_bytes + 255) & ~(size_t)255;
        code_off[i] = arena_at;
        arena_at += (jr.code_bytes + 255) & ~(size_t)255;
        spill_off[i + 1] = spill_off[i] + ((strip && !jr.mirror) ? (size_t)jr.n_spill + 1 : 0);
    }
    if (arena_at > S.arena_bytes) throw StateError("POA group exceeds the unit's arena");
    if (q_off[nj] >= (1ull << 32) || rows_old >= (1ull << 32) || rows_strip >= (1ull << 32) ||
        preds_total >= (1ull << 32) || S.aln_off[nj] >= (1ull << 31))
        throw CapacityError("POA batch too large for 32-bit staging offsets");
    uint8_t *hq = S.h_q.need_geo(q_off[nj] + 16);
    uint32_t *hri = S.h_row_info.need_geo(rows_old + 1);
    uint32_t *hrp = S.h_row_poff.need_geo(rows_old + 1);
    uint4 *hrec = S.h_rec.need_geo(rows_strip + 1);
    int32_t *hsp = S.h_spill.need_geo(spill_off[nj] + 1);
    int32_t *hpr = S.h_preds.need_geo(preds_total + 1);
    PoaJob *hj = S.h_jobs.need_geo(nj);
    PoaSJob *hsj = S.h_sjobs.need_geo(nj);
    const std::vector<size_t> &aln_off = S.aln_off;
    parallel_for(S.n_threads, nj, [&](size_t i) {
        const JobRef &jr = jobs[i];
        if (jr.kind == JK_STRIP) {
            PoaSJob &J = hsj[i];
            J.hf_off = hf_off[i] / 4;
            J.code_off = code_off[i] / 4;
            J.q_off = (uint32_t)q_off[i];
            J.aln_off = (uint32_t)aln_off[i];
            if (jr.mirror) {  // records are already on the device (k_poa_graph_fold); only the query is staged
                J.row_off = jr.rec_off;
                J.pred_base = jr.pred_base;
                J.spill_off = jr.spill_off;
                J.order_off = jr.task->gbase + 4 * (uint64_t)jr.task->cap_n;  // DGView::order
                stage_strip_query(jr, J, hq + q_off[i]);
            } else {
                J.row_off = (uint32_t)(S.base_rows + row_off[i]);
                J.pred_base = (uint32_t)(S.base_preds + pred_off[i]);
                J.spill_off = (uint32_t)(S.base_spill + spill_off[i]);
                J.order_off = ~0ull;
                stage_strip_job(jr, J, hq + q_off[i], hrec + row_off[i], hpr + pred_off[i], hsp + spill_off[i]);
            }
        } else {
            PoaJob &J = hj[i];
            J.hf_off = hf_off[i] / (jr.kind == JK_WIDE ? 8 : 4);
            J.code_off = code_off[i] / (jr.kind == JK_WIDE ? 4 : 2);
            J.q_off = (uint32_t)q_off[i];
            J.row_off = (uint32_t)row_off[i];
            J.pred_base = (uint32_t)(S.base_preds + pred_off[i]);
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
    // (when the step has mirrored graphs, run_fold sized these buffers for the whole step: no reallocation here)
    S.d_rec.need_geo(S.base_rows + rows_strip + 1);
    if (rows_strip) CK(cudaMemcpyAsync(S.d_rec.p + S.base_rows, hrec, rows_strip * sizeof(uint4), cudaMemcpyHostToDevice, st));
    S.d_spill.need_geo(S.base_spill + spill_off[nj] + 1);
    if (spill_off[nj]) CK(cudaMemcpyAsync(S.d_spill.p + S.base_spill, hsp, spill_off[nj] * 4, cudaMemcpyHostToDevice, st));
    S.d_preds.need_geo(S.base_preds + preds_total + 1);
    if (preds_total) CK(cudaMemcpyAsync(S.d_preds.p + S.base_preds, hpr, preds_total * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_jobs.need_geo(nj), hj, nj * sizeof(PoaJob), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_sjobs.need_geo(nj), hsj, nj * sizeof(PoaSJob), cudaMemcpyHostToDevice, st));
    S.st.h2d_bytes += (int64_t)(q_off[nj] + rows_old * 8 + rows_strip * 16 + preds_total * 4 + spill_off[nj] * 4 +
                                      nj * (sizeof(PoaJob) + sizeof(PoaSJob)));
    S.d_aln.need_geo(aln_off[nj] * 2);
    S.d_aln_len.need_geo(nj);
    S.d_best.need_geo(nj);
    CK(cudaMemsetAsync(S.d_counter.need_geo(segs.size()), 0, 4 * segs.size(), st));
    CK(cudaEventRecord(S.ev0, st));
    S.t_submit0 = ts0;
    // the segments are independent: fork them over the slot's sub-streams so that the tail of one segment's
    // kernel overlaps the others, and join before the D2H copy
    const bool fork = segs.size() > 1 && segs.size() <= (size_t)PoaSlot::N_SEG_EV;
    if (fork) CK(cudaEventRecord(S.ev_h2d, st));
    for (size_t si = 0; si < segs.size(); ++si) {
        const Seg &sg_ = segs[si];
        cudaStream_t st = fork ? S.sub[si % PoaSlot::N_SUB] : S.stream;
        if (fork) CK(cudaStreamWaitEvent(st, S.ev_h2d, 0));
        const int cnt = (int)(sg_.end - sg_.begin);
        unsigned int *counter = S.d_counter.p + si;
        if (sg_.key < 100) {
            const int nw = sg_.key;
            const int K = strip_ring_rows(nw);
            const size_t smem = ps_smem_bytes(nw, K);
            int occ = 1;
            auto kern = k_poa_strip<5, -4, -8, -6>;  // correct.cpp:395 createAlignmentEngine(kSW, 5, -4, -8, -6)
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nw * 32, smem));
            const int grid = (int)std::min<size_t>((size_t)cnt, (size_t)ctx->n_sm * std::max(1, occ));
            kern<<<grid, nw * 32, smem, st>>>(S.d_sjobs.p + sg_.begin, cnt, S.d_q.p, S.d_rec.p, S.d_preds.p,
                                              (uint32_t *)S.arena, S.d_best.p + sg_.begin, counter, K);
            CK(cudaGetLastError());
            k_poa_strip_traceback<<<(cnt + 3) / 4, 128, 0, st>>>(S.d_sjobs.p + sg_.begin, cnt, S.d_rec.p, S.d_preds.p,
                                                                 S.d_spill.p, (const uint32_t *)S.arena, S.d_best.p + sg_.begin,
                                                                 S.d_pool.p, S.d_aln.p, S.d_aln_len.p + sg_.begin);
            S.st.kernel_launches++;
        } else {
            const bool wide = sg_.key == 100 + JK_WIDE;
            const int occ = std::max(1, P.occ[wide ? 1 : 0]);
            const int grid = (int)std::min<size_t>((size_t)cnt, (size_t)ctx->n_sm * occ);
            if (!wide)
                k_poa_align<false><<<grid, POA_T, 0, st>>>(S.d_jobs.p + sg_.begin, cnt, S.d_q.p, S.d_row_info.p,
                                                           S.d_row_poff.p, S.d_preds.p, (short2 *)S.arena, (uint16_t *)S.arena,
                                                           S.d_aln.p, S.d_aln_len.p + sg_.begin, sm, sn, sg, se, counter);
            else
                k_poa_align<true><<<grid, POA_T, 0, st>>>(S.d_jobs.p + sg_.begin, cnt, S.d_q.p, S.d_row_info.p,
                                                          S.d_row_poff.p, S.d_preds.p, (int2 *)S.arena, (uint32_t *)S.arena,
                                                          S.d_aln.p, S.d_aln_len.p + sg_.begin, sm, sn, sg, se, counter);
        }
        CK(cudaGetLastError());
        S.st.poa_launches++;
        S.st.kernel_launches++;
        if (fork) {
            CK(cudaEventRecord(S.ev_seg[si], st));
            CK(cudaStreamWaitEvent(S.stream, S.ev_seg[si], 0));
        }
    }
    CK(cudaEventRecord(S.ev1, st));
    S.t_submit1 = now_ms();
    int32_t *haln = S.h_aln.need_geo(aln_off[nj] * 2);
    int32_t *hlen = S.h_aln_len.need_geo(nj);
    CK(cudaMemcpyAsync(hlen, S.d_aln_len.p, nj * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(haln, S.d_aln.p, aln_off[nj] * 8, cudaMemcpyDeviceToHost, st));
    S.pending = true;
    S.st.poa_alignments += (int64_t)nj;
    S.st.d2h_bytes += (int64_t)(aln_off[nj] * 8 + nj * 4);
    for (size_t i = 0; i < nj; ++i) {
        S.st.poa_cells += (int64_t)jobs[i].L * jobs[i].n;
        if (jobs[i].kind == JK_STRIP)  // codes + spilled rows (H, F, halo) + pass hand-over words
            S.st.poa_dram_bytes += (int64_t)jobs[i].code_bytes + (int64_t)(jobs[i].n_spill + 1) * jobs[i].nst * 1028 +
                                   (jobs[i].nst > PS_MAXW ? 4 * (int64_t)jobs[i].n : 0);
        else
            S.st.poa_dram_bytes += (int64_t)jobs[i].code_bytes + (int64_t)jobs[i].hf_bytes;
    }
    S.t_stage += now_ms() - ts0;
}

// wait for the slot's launch and fold its alignments into the graphs (host threads)
static void finish(rtl_ctx *ctx, PoaState &P, PoaSlot &S) {
    if (!S.pending) return;
    const double tw0 = now_ms();
    slot_wait(S);
    const double tw1 = now_ms();
    S.t_wait += tw1 - tw0;
    S.pending = false;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, S.ev0, S.ev1));
    S.st.poa_ms += ms;
    const int32_t *haln = S.h_aln.p;
    const int32_t *hlen = S.h_aln_len.p;
    const bool keep = S.keep_alns;
    parallel_for(S.n_threads, S.jobs.size(), [&](size_t i) {
        const JobRef &jr = S.jobs[i];
        PoaGraph &g = jr.task->g;
        const int len = hlen[i];
        const int32_t *src = haln + 2 * S.aln_off[i];
        std::vector<std::pair<int32_t, int32_t>> aln((size_t)len);
        for (int x = 0; x < len; ++x) {  // reverse (sisd_alignment_engine.cpp:655) and map rows to node ids
            const int row = src[2 * (len - 1 - x)], pos = src[2 * (len - 1 - x) + 1];
            aln[x].first = row < 0 ? -1 : (jr.mirror ? row : g.rank_to_node[row - 1]);  // mirrored graphs: node ids
            aln[x].second = pos;
        }
        g.add_alignment(aln, jr.task->seq[jr.seq_index], jr.L);
        if (keep) jr.task->alns[jr.seq_index] = std::move(aln);
    });
    S.t_fold += now_ms() - tw1;
    {
        float k0 = 0, k1 = 0;
        cudaEventElapsedTime(&k0, P.ev_ref, S.ev0);
        cudaEventElapsedTime(&k1, P.ev_ref, S.ev1);
        std::lock_guard<std::mutex> lk(P.stats_mu);
        P.intervals.emplace_back(k0, k1);
        if (P.trace) {
        fprintf(P.trace, "{\"unit\": %d, \"epoch\": %d, \"jobs\": %zu, \"stage0\": %.3f, \"stage1\": %.3f, \"k0\": %.3f, \"k1\": %.3f, "
                "\"sync\": %.3f, \"fold\": %.3f}\n", (int)(&S - P.slot_store), S.epoch, S.jobs.size(), S.t_submit0 - P.t_ref,
                S.t_submit1 - P.t_ref, k0, k1, tw1 - P.t_ref, now_ms() - P.t_ref);
            fflush(P.trace);
        }
    }
    S.jobs.clear();
}

// A graph leaves the device-mirror path (capacity exceeded, in-degree > 32, a read the strip kernel cannot take):
// from now on the host sorts it itself.
static void demote_mirror(PoaTask *t) {
    t->mirror = false;
    t->g.defer_sort = false;
    if (!t->g.sorted) t->g.topological_sort();
}

// Step prologue for the graphs with a device mirror (`mj` = their indices in `all`): send the new entries of the
// graphs' append logs, let the GPU fold them in, sort the graphs and build the row records (k_poa_graph_fold), and
// read back how many rows each graph spills (the arena is divided with that).  `all` holds every job of the step: the
// slot's d_rec / d_preds / d_spill are sized for all of them here, the mirrored graphs first.
static void run_fold(rtl_ctx *ctx, PoaSlot &S, std::vector<JobRef> &all, const std::vector<size_t> &mj) {
    (void)ctx;
    const double ts0 = now_ms();
    cudaStream_t st = S.stream;
    const size_t nm = mj.size();
    size_t rows_all = 0, preds_all = 0;
    for (const auto &jr : all) {
        rows_all += (size_t)jr.n + 1;
        preds_all += jr.task->g.e_begin.size() + (size_t)jr.n;
    }
    std::vector<size_t> delta_off(nm + 1, 0);
    size_t rows = 0, preds = 0;
    DFoldJob *hf = S.h_fjobs.need_geo(nm);
    for (size_t k = 0; k < nm; ++k) {
        JobRef &jr = all[mj[k]];
        PoaTask *t = jr.task;
        const PoaGraph &g = t->g;
        const int n_new = g.n_nodes(), e_new = (int)g.e_begin.size(), a_new = (int)g.a_node.size();
        delta_off[k + 1] = delta_off[k] + dg_delta_words(n_new - t->sync_n, e_new - t->sync_e, a_new - t->sync_a);
        jr.rec_off = (uint32_t)rows;
        jr.pred_base = (uint32_t)preds;
        jr.spill_off = (uint32_t)rows;  // one spill_rows entry per row is always enough
        rows += (size_t)n_new + 1;
        preds += (size_t)e_new + 4;
        DFoldJob &F = hf[k];
        F.gbase = t->gbase;
        F.delta_off = (uint32_t)delta_off[k];
        F.rec_off = jr.rec_off;
        F.pred_base = jr.pred_base;
        F.spill_off = jr.spill_off;
        F.cap_n = t->cap_n;
        F.cap_e = t->cap_e;
        F.cap_a = t->cap_a;
        F.n_old = t->sync_n;
        F.n_new = n_new;
        F.e_old = t->sync_e;
        F.e_new = e_new;
        F.a_old = t->sync_a;
        F.a_new = a_new;
        F.K = strip_ring_rows(strip_warps(jr.nst));
        F.pad = 0;
    }
    if (delta_off[nm] >= (1ull << 32) || rows_all >= (1ull << 31) || preds_all + 4 * nm >= (1ull << 31))
        throw CapacityError("POA step too large for 32-bit staging offsets");
    S.base_rows = rows;
    S.base_preds = preds;
    S.base_spill = rows;
    S.d_rec.need_geo(rows_all + 1);
    S.d_preds.need_geo(preds_all + 4 * nm + 1);
    S.d_spill.need_geo(rows_all + 1);
    int32_t *hd = S.h_delta.need_geo(delta_off[nm] + 1);
    const uint8_t *tab = letter_codes();
    parallel_for(S.n_threads, nm, [&](size_t k) {
        const JobRef &jr = all[mj[k]];
        PoaTask *t = jr.task;
        const PoaGraph &g = t->g;
        const DFoldJob &F = hf[k];
        int32_t *d = hd + delta_off[k];
        const int dn = F.n_new - F.n_old;
        if (dn) d[(dn - 1) / 4] = 0;
        uint8_t *let = reinterpret_cast<uint8_t *>(d);
        for (int v = F.n_old; v < F.n_new; ++v) let[v - F.n_old] = tab[(unsigned char)g.letter[v]];
        int32_t *ed = d + (dn + 3) / 4;
        for (int x = F.e_old; x < F.e_new; ++x) {
            ed[2 * (x - F.e_old)] = g.e_begin[x];
            ed[2 * (x - F.e_old) + 1] = g.e_end[x];
        }
        int32_t *al = ed + 2 * (F.e_new - F.e_old);
        for (int x = F.a_old; x < F.a_new; ++x) {
            al[2 * (x - F.a_old)] = g.a_owner[x];
            al[2 * (x - F.a_old) + 1] = g.a_node[x];
        }
        t->sync_n = F.n_new;
        t->sync_e = F.e_new;
        t->sync_a = F.a_new;
    });
    CK(cudaMemcpyAsync(S.d_delta.need_geo(delta_off[nm] + 1), hd, delta_off[nm] * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.d_fjobs.need_geo(nm), hf, nm * sizeof(DFoldJob), cudaMemcpyHostToDevice, st));
    S.d_counts.need_geo(2 * nm);
    k_poa_graph_fold<<<(unsigned)((nm + 3) / 4), 128, 0, st>>>(S.d_fjobs.p, (int)nm, S.d_pool.p, S.d_delta.p,
                                                           reinterpret_cast<uint32_t *>(S.d_rec.p), S.d_preds.p,
                                                           S.d_spill.p, S.d_counts.p);
    CK(cudaGetLastError());
    int32_t *hc = S.h_counts.need_geo(2 * nm);
    CK(cudaMemcpyAsync(hc, S.d_counts.p, 2 * nm * 4, cudaMemcpyDeviceToHost, st));
    const double ts1 = now_ms();
    slot_wait(S);
    S.t_wait += now_ms() - ts1;
    S.t_stage += ts1 - ts0;
    for (size_t k = 0; k < nm; ++k) all[mj[k]].n_spill = hc[2 * k];
    S.st.kernel_launches++;
    S.st.h2d_bytes += (int64_t)(delta_off[nm] * 4 + nm * sizeof(DFoldJob));
    S.st.d2h_bytes += (int64_t)(2 * nm * 4);
}

// ------------------------------------------------------------------------------------------------ device-resident chains
// Whole per-pack chains on the GPU (poa_devchain.cuh): the host lays out fixed per-pack slots from what it knows in
// advance (read lengths -> capacities, strips, CTA width), uploads the packs' reads once and launches k_poa_chain once
// per CTA width — one CTA works through a pack's whole chain.  It synchronises twice at the end (MSA sizes, MSA rows).
// Packs the device flags (capacity, in-degree > 32, spill slots) are returned in `failed` for the host-driven path.

struct ChainPlanSeg {
    size_t begin, end;  // range of the launch list
    int key;            // warps per CTA
};

struct ChainSizes {
    int cap_n, cap_e, cap_a, spill_cap, max_nst, maxlen;
    long long total;
    size_t hf_words, code_words, arena_bytes;
};

static ChainSizes chain_sizes(const rtl_ctx *ctx, const PoaTask *t) {
    ChainSizes z{};
    for (int l : t->len) {
        z.maxlen = std::max(z.maxlen, l);
        z.total += l;
    }
    // graph nodes: at most one per base; in practice the longest read plus what every further read adds (its share of new
    // bases: errors and isoform differences): 2.9-3.5 k nodes for 50 x 1.5 kb reads at 7 % error, ~7.7 k for 100 at 10 %.
    // The slot allows (2 + reads/25) x the longest read (50 reads: 4 x, as measured above with 2 x headroom)
    const double per_len = std::max(3.0, std::min(12.0, 2.0 + (double)t->len.size() / 25.0));
    const long long cn = std::max<long long>(16, std::min<long long>(z.total + 8, (long long)(per_len * z.maxlen) + 1024) * ctx->poa_mirror_pct / 100);
    z.cap_n = (int)cn;
    z.cap_e = (int)(3 * cn);
    z.cap_a = (int)(4 * cn);
    // rows that some later row needs from beyond the ring: 1 % of the rows of a 50-read pack, but packs of 100+ noisy reads
    // are bushier (an eighth of the rows was not enough for 45 of 50 such packs): the share grows with the read count
    const size_t nr = t->len.size();
    const double spill_share = nr <= 50 ? 0.125 : std::min(0.5, 0.125 + (double)(nr - 50) / 200.0);
    z.spill_cap = (int)std::min<long long>(65000, (long long)(cn * spill_share) + 16);
    z.max_nst = (z.maxlen + PS_STRIP - 1) / PS_STRIP;
    z.hf_words = (ps_hf_words(z.cap_n, z.max_nst, z.spill_cap) + 63) & ~(size_t)63;
    z.code_words = (ps_code_words(z.cap_n, z.max_nst) + 63) & ~(size_t)63;
    z.arena_bytes = (z.hf_words + z.code_words) * 4;
    return z;
}

// Where a chain launch takes its reads from: host strings (PoaTask::seq), optionally with their qualities (uploaded next
// to the letter codes for the device-side vote), or corrected reads that already lie on the device (round 2 after a
// device-side vote: `stage` lists them in table order, k_vote_stage_queries turns them into letter codes).
struct ChainInput {
    const std::vector<std::vector<const char *>> *quals = nullptr;  // per pack, per read
    const std::vector<DVStage> *stage = nullptr;                     // q_off is filled in here
    const char *stage_src = nullptr;
};
struct ChainRun {
    double ts0 = 0, ts1 = 0, ts2 = 0;
};

// plan + upload + launch + wait: afterwards S.ch_packs holds every pack's status / graph size / MSA columns and the
// graphs, paths and MSA column ids lie in the slot's device buffers
static ChainRun chain_run(rtl_ctx *ctx, PoaState &P, PoaSlot &S, const std::vector<PoaTask *> &batch,
                          const std::vector<ChainSizes> &sizes, const ChainInput &in) {
    const double ts0 = now_ms();
    cudaStream_t st = S.stream;
    const size_t np = batch.size();
    const uint8_t *tab = letter_codes();
    // ---- per-pack slots
    size_t n_seq_total = 0, q_bytes = 0, pool_words = 0, rec_rows = 0, pred_words = 0, spill_words = 0, aln_pairs = 0,
           path_words = 0, qnode_words = 0, arena_words = 0;
    DCPack *hp = S.ch_packs.need_geo(np);
    for (size_t i = 0; i < np; ++i) {
        const PoaTask *t = batch[i];
        const ChainSizes &z = sizes[i];
        DCPack &K = hp[i];
        memset(&K, 0, sizeof(K));
        K.gbase = pool_words;
        K.hf_off = arena_words;
        K.code_off = arena_words + z.hf_words;
        K.rec_off = (uint32_t)rec_rows;
        K.pred_base = (uint32_t)pred_words;
        K.spill_off = (uint32_t)spill_words;
        K.aln_off = (uint32_t)aln_pairs;
        K.path_off = (uint32_t)path_words;
        K.qnode_off = (uint32_t)qnode_words;
        K.seq_base = (uint32_t)n_seq_total;
        K.n_seq = (int32_t)t->len.size();
        K.cap_n = z.cap_n;
        K.cap_e = z.cap_e;
        K.cap_a = z.cap_a;
        K.spill_cap = z.spill_cap;
        pool_words += dg_words(z.cap_n, z.cap_e, z.cap_a);
        arena_words += z.hf_words + z.code_words;
        rec_rows += (size_t)z.cap_n + 1;
        pred_words += (size_t)z.cap_e + 4;
        spill_words += (size_t)z.cap_n + 4;
        aln_pairs += (size_t)z.cap_n + z.maxlen + 8;
        path_words += (size_t)z.total;
        qnode_words += (size_t)z.maxlen + 8;
        n_seq_total += t->len.size();
        for (int l : t->len) q_bytes += (size_t)((l + PS_STRIP - 1) / PS_STRIP) * PS_STRIP;
    }
    if (arena_words * 4 > S.arena_bytes) throw StateError("device-chain batch exceeds the unit's arena");
    if (rec_rows >= (1ull << 31) || pred_words >= (1ull << 31) || aln_pairs >= (1ull << 30) || path_words >= (1ull << 31) ||
        q_bytes >= (1ull << 32))
        throw CapacityError("device-chain batch too large for 32-bit offsets");
    // ---- sequence table and query codes
    DCSeq *hs = S.ch_seqs.need_geo(n_seq_total);
    uint8_t *hq = S.ch_q.need_geo(q_bytes + 16);
    {
        std::vector<size_t> qo(np + 1, 0);
        for (size_t i = 0; i < np; ++i) {
            size_t b = 0;
            for (int l : batch[i]->len) b += (size_t)((l + PS_STRIP - 1) / PS_STRIP) * PS_STRIP;
            qo[i + 1] = qo[i] + b;
        }
        uint8_t *hql = in.quals ? S.ch_qual.need_geo(q_bytes + 16) : nullptr;
        parallel_for(S.n_threads, np, [&](size_t i) {
            const PoaTask *t = batch[i];
            size_t at = qo[i];
            uint32_t prel = 0;
            for (size_t s = 0; s < t->len.size(); ++s) {
                const int L = t->len[s], nst = (L + PS_STRIP - 1) / PS_STRIP;
                DCSeq &Q = hs[hp[i].seq_base + s];
                Q.q_off = (uint32_t)at;
                Q.path_rel = prel;
                Q.L = L;
                Q.pad = 0;
                if (!in.stage) {
                    const char *src = t->seq[s];
                    for (int x = 0; x < L; ++x) hq[at + x] = tab[(unsigned char)src[x]];
                    memset(hq + at + L, 255, (size_t)nst * PS_STRIP - L);
                    if (hql) memcpy(hql + at, (*in.quals)[i][s], (size_t)L);
                }
                at += (size_t)nst * PS_STRIP;
                prel += (uint32_t)L;
            }
        });
    }
    // ---- launch plan: one k_poa_chain launch per CTA width (warps = strips of the pack's longest read, balanced over
    // passes beyond 8 strips); inside a launch the packs with the most DP work are fetched first
    struct Ent {
        int key;
        double work;
        int32_t pack;
    };
    std::vector<Ent> ents(np);
    for (size_t i = 0; i < np; ++i) {
        double w = 0, sum = 0;
        for (int l : batch[i]->len) {
            w += (double)l * sum;  // read x (nodes so far, at most the bases so far)
            sum += l;
        }
        ents[i] = Ent{chain_warps(sizes[i].max_nst), w, (int32_t)i};
    }
    std::sort(ents.begin(), ents.end(), [](const Ent &a, const Ent &b) {
        if (a.key != b.key) return a.key > b.key;
        if (a.work != b.work) return a.work > b.work;
        return a.pack < b.pack;
    });
    std::vector<ChainPlanSeg> segs;
    std::vector<int32_t> list(np);
    for (size_t i = 0; i < np; ++i) {
        list[i] = ents[i].pack;
        if (segs.empty() || segs.back().key != ents[i].key) segs.push_back(ChainPlanSeg{i, i, ents[i].key});
        segs.back().end = i + 1;
    }
    const size_t n_segs = segs.size();
    // ---- device buffers + upload
    S.c_pool.need_geo(pool_words);
    S.c_rec.need_geo(rec_rows + 1);
    S.c_preds.need_geo(pred_words + 1);
    S.c_spill.need_geo(spill_words + 1);
    S.c_aln.need_geo(2 * aln_pairs + 2);
    S.c_path.need_geo(path_words + 1);
    S.c_qnode.need_geo(qnode_words + 1);
    S.c_counter.need_geo(n_segs + 1);
    S.c_stats.need_geo(8);
    CK(cudaMemcpyAsync(S.c_packs.need_geo(np), hp, np * sizeof(DCPack), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.c_seqs.need_geo(n_seq_total), hs, n_seq_total * sizeof(DCSeq), cudaMemcpyHostToDevice, st));
    S.c_q.need_geo(q_bytes + 16);
    if (!in.stage) {
        CK(cudaMemcpyAsync(S.c_q.p, hq, q_bytes, cudaMemcpyHostToDevice, st));
        if (in.quals) {
            CK(cudaMemcpyAsync(S.c_qual.need_geo(q_bytes + 16), S.ch_qual.p, q_bytes, cudaMemcpyHostToDevice, st));
            S.st.h2d_bytes += (int64_t)q_bytes;
        }
    } else {  // the reads are on the device already: one CTA per read writes its letter codes
        const size_t nr = in.stage->size();
        if (nr != n_seq_total) throw StateError("chain_run: staging table does not match the packs");
        DVStage *hst2 = S.vh_stage.need_geo(nr + 1);
        for (size_t r = 0; r < nr; ++r) {
            hst2[r] = (*in.stage)[r];
            hst2[r].q_off = hs[r].q_off;
        }
        CK(cudaMemcpyAsync(S.v_stage.need_geo(nr + 1), hst2, nr * sizeof(DVStage), cudaMemcpyHostToDevice, st));
        if (nr) k_vote_stage_queries<<<(unsigned)nr, 256, 0, st>>>(S.v_stage.p, in.stage_src, S.c_q.p, PS_STRIP);
        CK(cudaGetLastError());
        S.st.kernel_launches++;
        S.st.h2d_bytes += (int64_t)(nr * sizeof(DVStage)) - (int64_t)q_bytes;
    }
    CK(cudaMemcpyAsync(S.c_list.need_geo(np), list.data(), np * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(S.c_counter.p, 0, (n_segs + 1) * sizeof(unsigned int), st));
    CK(cudaMemsetAsync(S.c_stats.p, 0, 8 * sizeof(unsigned long long), st));
    S.st.h2d_bytes += (int64_t)(np * sizeof(DCPack) + n_seq_total * sizeof(DCSeq) + q_bytes + np * 4);
    // ---- the chains: one launch per CTA width, side by side on the unit's sub-streams; no host synchronisation until
    // the MSA sizes are needed
    CK(cudaEventRecord(S.ev0, st));
    const bool fork = n_segs > 1 && n_segs <= (size_t)PoaSlot::N_SEG_EV;
    if (fork) CK(cudaEventRecord(S.ev_h2d, st));
    for (size_t si = 0; si < n_segs; ++si) {
        const ChainPlanSeg &sg_ = segs[si];
        cudaStream_t ss = fork ? S.sub[si % PoaSlot::N_SUB] : st;
        if (fork) CK(cudaStreamWaitEvent(ss, S.ev_h2d, 0));
        const int cnt = (int)(sg_.end - sg_.begin);
        const int nw = sg_.key;
        // CTAs per SM: 8-warp CTAs are register-limited to three; narrower ones run four per SM (kernel variant compiled
        // for that), with a ring of 5 rows where that is what fits 54 KB.  Dynamic shared memory = the DP's profiles and
        // row ring, or — between two DPs — the working set of the graph's sort.
        const bool four = nw <= 7 && getenv("RATTLE_B200_CTAS3") == nullptr;
        // experiments: one more CTA per SM (64 registers per thread, a shallower ring) for CTAs of up to 6 / of 8 warps
        const bool five = nw <= 6 && getenv("RATTLE_B200_CTAS5") != nullptr;
        // 8-warp CTAs: four per SM at 64 registers (8 bytes of spill) beat three at 80: +8 % on 2 kb reads (config 4)
        const bool four8 = nw == 8 && getenv("RATTLE_B200_CTAS3") == nullptr;
        const bool narrow = getenv("RATTLE_B200_NARROW") != nullptr && atoi(getenv("RATTLE_B200_NARROW")) != 0 && nw <= 4;
        // narrow CTAs: 8 per SM for up to 3 warps (26 KB each), 6 per SM for 4 warps (35 KB)
        const size_t budget = narrow ? (nw <= 3 ? (size_t)26 * 1024 : (size_t)35 * 1024)
                                     : (five ? (size_t)43 * 1024 : (four8 ? (size_t)54 * 1024 : (four ? (size_t)54 * 1024 : (size_t)72 * 1024)));
        int K = strip_ring_rows(nw);
        while (K > 3 && ps_smem_bytes(nw, K) > budget) --K;
        const size_t smem = std::max(ps_smem_bytes(nw, K), budget);
        const int smem_cap_n = dc_sort_cap(smem, nw * 32);
        auto kern = narrow ? (nw <= 3 ? k_poa_chain<5, -4, -8, -6, 96, 8> : k_poa_chain<5, -4, -8, -6, 128, 6>) : nw == 8 ? (four8 ? k_poa_chain<5, -4, -8, -6, 256, 4> : k_poa_chain<5, -4, -8, -6, 256, 3>)
                            : (!four ? k_poa_chain<5, -4, -8, -6, 256, 3>
                                     : (nw == 7 ? k_poa_chain<5, -4, -8, -6, 224, 4>
                                                : (five ? k_poa_chain<5, -4, -8, -6, 192, 5> : k_poa_chain<5, -4, -8, -6, 192, 4>)));
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nw * 32, smem));
        const int grid = (int)std::min<size_t>((size_t)cnt, (size_t)ctx->n_sm * std::max(1, occ));
        kern<<<grid, nw * 32, smem, ss>>>(S.c_packs.p, S.c_list.p + sg_.begin, cnt, S.c_seqs.p, S.c_pool.p, S.c_q.p, S.c_rec.p,
                                          S.c_preds.p, S.c_spill.p, S.c_aln.p, S.c_path.p, S.c_qnode.p, (uint32_t *)S.arena,
                                          S.c_stats.p, S.c_counter.p + si, K, smem_cap_n);
        if (getenv("RTL_TRACE") && S.epoch <= 2 && (&S - P.slot_store) == 0)
            fprintf(stderr, "[rtl] k_poa_chain: %d packs, %d warps per CTA, ring %d, %zu B smem, %d CTAs per SM\n", cnt, nw, K, smem, occ);
        CK(cudaGetLastError());
        S.st.poa_launches++;
        S.st.kernel_launches++;
        if (fork) {
            CK(cudaEventRecord(S.ev_seg[si], ss));
            CK(cudaStreamWaitEvent(st, S.ev_seg[si], 0));
        }
    }
    CK(cudaEventRecord(S.ev1, st));
    CK(cudaMemcpyAsync(hp, S.c_packs.p, np * sizeof(DCPack), cudaMemcpyDeviceToHost, st));
    unsigned long long *hst = S.ch_stats.need(8);
    CK(cudaMemcpyAsync(hst, S.c_stats.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    const double ts1 = now_ms();
    S.t_stage += ts1 - ts0;
    slot_wait(S);
    const double ts2 = now_ms();
    S.t_wait += ts2 - ts1;
    {
        float ms = 0, k0 = 0, k1 = 0;
        CK(cudaEventElapsedTime(&ms, S.ev0, S.ev1));
        cudaEventElapsedTime(&k0, P.ev_ref, S.ev0);
        cudaEventElapsedTime(&k1, P.ev_ref, S.ev1);
        S.st.poa_ms += ms;
        S.st.poa_cells += (int64_t)hst[0];
        S.st.poa_alignments += (int64_t)hst[1];
        S.st.poa_dram_bytes += (int64_t)hst[2];
        if (getenv("RTL_TRACE")) {
            const double tot = (double)(hst[3] + hst[4] + hst[5]);
            fprintf(stderr, "[rtl] device chains unit %d: %zu packs, CTA clocks: graph update %.1f %% (add_alignment %.1f, sort %.1f), "
                    "DP %.1f %%, traceback %.1f %% (%.0f ms of CTA time at 1.9 GHz)\n", (int)(&S - P.slot_store), np,
                    100.0 * hst[3] / std::max(1.0, tot), 100.0 * hst[6] / std::max(1.0, tot), 100.0 * hst[7] / std::max(1.0, tot),
                    100.0 * hst[4] / std::max(1.0, tot), 100.0 * hst[5] / std::max(1.0, tot), tot / 1.9e6);
        }
        S.st.d2h_bytes += (int64_t)(np * sizeof(DCPack) + 32);
        std::lock_guard<std::mutex> lk(P.stats_mu);
        P.intervals.emplace_back(k0, k1);
        if (P.trace) {
            fprintf(P.trace, "{\"unit\": %d, \"epoch\": %d, \"jobs\": %zu, \"stage0\": %.3f, \"stage1\": %.3f, \"k0\": %.3f, \"k1\": %.3f, "
                    "\"sync\": %.3f, \"fold\": %.3f, \"device_chain\": 1}\n", (int)(&S - P.slot_store), S.epoch, np, ts0 - P.t_ref,
                    ts1 - P.t_ref, k0, k1, ts2 - P.t_ref, now_ms() - P.t_ref);
            fflush(P.trace);
        }
    }
    ChainRun R;
    R.ts0 = ts0;
    R.ts1 = ts1;
    R.ts2 = ts2;
    return R;
}

// MSA rows (graph.cpp:390-426) of the packs of the last chain_run that made it, compact in S.c_msa: pack i's n_seq x ncol
// chars start at offset hmo[i] (S.ch_msa_off); returns the total size.  Enqueued on the slot's stream.
static size_t chain_msa_rows(PoaSlot &S, size_t np, cudaStream_t st) {
    const DCPack *hp = S.ch_packs.p;
    uint64_t *hmo = S.ch_msa_off.need_geo(np + 1);
    size_t msa_bytes = 0;
    for (size_t i = 0; i < np; ++i) {
        hmo[i] = msa_bytes;
        if (hp[i].status == DC_OK) msa_bytes += (size_t)hp[i].n_seq * (size_t)hp[i].ncol;
    }
    hmo[np] = msa_bytes;
    CK(cudaMemcpyAsync(S.c_msa_off.need_geo(np + 1), hmo, np * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    S.c_msa.need_geo(msa_bytes + 16);
    k_chain_msa_rows<<<(unsigned)np, 256, 0, st>>>(S.c_packs.p, S.c_seqs.p, S.c_msa_off.p, S.c_pool.p, S.c_path.p, S.c_msa.p);
    CK(cudaGetLastError());
    S.st.kernel_launches++;
    S.st.h2d_bytes += (int64_t)(np * sizeof(uint64_t));
    return msa_bytes;
}

// Whole chains on the GPU, MSA rows back to the host (PoaTask::msa_rows).
static void dev_chain_batch(rtl_ctx *ctx, PoaState &P, PoaSlot &S, const std::vector<PoaTask *> &batch,
                            const std::vector<ChainSizes> &sizes, std::vector<PoaTask *> &failed) {
    const size_t np = batch.size();
    cudaStream_t st = S.stream;
    const ChainRun R = chain_run(ctx, P, S, batch, sizes, ChainInput());
    const double ts2 = R.ts2;
    const DCPack *hp = S.ch_packs.p;
    const size_t msa_bytes = chain_msa_rows(S, np, st);
    const uint64_t *hmo = S.ch_msa_off.p;
    // ---- MSA rows of the packs that made it, compact
    char *hm = S.ch_msa.need_geo(msa_bytes + 16);
    if (msa_bytes) CK(cudaMemcpyAsync(hm, S.c_msa.p, msa_bytes, cudaMemcpyDeviceToHost, st));
    slot_wait(S);
    S.t_wait += now_ms() - ts2;
    const double tf0 = now_ms();
    parallel_for(S.n_threads, np, [&](size_t i) {
        PoaTask *t = batch[i];
        if (hp[i].status != DC_OK) return;
        const int rows = hp[i].n_seq, ncol = hp[i].ncol;
        t->msa_rows.resize(rows);
        for (int r = 0; r < rows; ++r) t->msa_rows[r].assign(hm + hmo[i] + (size_t)r * ncol, (size_t)ncol);
        t->have_msa = true;
    });
    for (size_t i = 0; i < np; ++i)
        if (hp[i].status != DC_OK) failed.push_back(batch[i]);
    S.t_fold += now_ms() - tf0;
    S.st.d2h_bytes += (int64_t)msa_bytes;
}

static void poa_chain_host(rtl_ctx *ctx, PoaState &P, PoaSlot &S, int unit, std::vector<PoaTask *> &tasks, int sm, int sn,
                           int sg, int se, bool keep_alns);

// One unit's chain on slot `unit`, synchronously (the calling thread is the unit's driver).  Packs that qualify run
// wholly on the GPU (device-resident chains), in batches that fit the unit's arena slice; the others — and the packs
// the device flags — take the host-driven lock-step path.  n_threads = host threads this chain may use.
void poa_chain(rtl_ctx *ctx, int unit, std::vector<PoaTask *> &tasks, int sm, int sn, int sg, int se, bool keep_alns,
               int n_threads) {
    PoaState &P = pstate(ctx);
    if (unit < 0 || unit >= P.n_units) throw StateError("poa_chain: no such unit");
    CK(cudaSetDevice(ctx->device));
    PoaSlot &S = P.slot_store[unit];
    S.n_threads = std::max(1, n_threads);
    S.t_wait = S.t_fold = S.t_stage = 0;
    S.st = rtl_stats{};
    S.epoch++;
    const double t_begin = now_ms();
    const uint8_t *tab = letter_codes();
    const int maxabs = std::max(std::max(std::abs(sm), std::abs(sn)), std::max(std::abs(sg), std::abs(se)));
    parallel_for(S.n_threads, tasks.size(), [&](size_t i) {
        PoaTask *t = tasks[i];
        t->g.clear();
        t->have_msa = false;
        t->msa_rows.clear();
        if (keep_alns) t->alns.assign(t->seq.size(), {});
        bool ok = true;
        for (size_t s = 0; s < t->seq.size() && ok; ++s)
            for (int x = 0; x < t->len[s]; ++x)
                if (tab[(unsigned char)t->seq[s][x]] == 255) {
                    ok = false;
                    break;
                }
        t->acgtu = ok;
    });
    std::vector<PoaTask *> rest;
    size_t n_dev = 0, n_failed = 0;
    const bool dev_ok = ctx->poa_device_chain != 0 && ctx->poa_gpu_sort != 0 && ctx->poa_kernel != 1 && !keep_alns && sm == 5 &&
                        sn == -4 && sg == -8 && se == -6;
    if (dev_ok) {
        std::vector<PoaTask *> batch;
        std::vector<ChainSizes> sizes;
        size_t used = 0;
        auto flush = [&]() {
            if (batch.empty()) return;
            n_dev += batch.size();
            const size_t before = rest.size();
            dev_chain_batch(ctx, P, S, batch, sizes, rest);
            n_failed += rest.size() - before;
            batch.clear();
            sizes.clear();
            used = 0;
        };
        for (PoaTask *t : tasks) {
            bool ok = t->acgtu && !t->seq.empty();
            for (int l : t->len)
                if (l < 1 || (int64_t)maxabs * (l + 16) >= 32000) ok = false;
            if (!ok) {
                rest.push_back(t);
                continue;
            }
            const ChainSizes z = chain_sizes(ctx, t);
            if (z.arena_bytes > S.arena_bytes) {
                rest.push_back(t);
                continue;
            }
            if (used + z.arena_bytes > S.arena_bytes) flush();
            batch.push_back(t);
            sizes.push_back(z);
            used += z.arena_bytes;
        }
        flush();
    } else {
        rest = tasks;
    }
    if (!rest.empty()) poa_chain_host(ctx, P, S, unit, rest, sm, sn, sg, se, keep_alns);
    {
        std::lock_guard<std::mutex> lk(P.stats_mu);
        rtl_stats &d = ctx->stats;
        d.h2d_bytes += S.st.h2d_bytes;
        d.d2h_bytes += S.st.d2h_bytes;
        d.poa_launches += S.st.poa_launches;
        d.kernel_launches += S.st.kernel_launches;
        d.poa_alignments += S.st.poa_alignments;
        d.poa_cells += S.st.poa_cells;
        d.poa_dram_bytes += S.st.poa_dram_bytes;
        d.poa_ms += S.st.poa_ms;
        if (getenv("RTL_TRACE"))
            fprintf(stderr, "[rtl] poa_chain unit %d: %zu tasks (%zu on the device chain, %zu of them sent back, %zu host-driven), "
                    "%.1f ms (waited for GPU %.1f, host fold/MSA %.1f, stage+submit %.1f), %d host threads\n", unit, tasks.size(),
                    n_dev, n_failed, rest.size(), now_ms() - t_begin, S.t_wait, S.t_fold, S.t_stage, S.n_threads);
    }
}

// ------------------------------------------------------------------------------------------------ device-side vote
// correct.cpp:395-445 for one batch of packs without the MSAs leaving the GPU: chains of round 1 -> MSA rows -> fix_msa_ends
// + per-cell qualities (k_vote_rows) -> column statistics (k_vote_cols) -> the few quality symbols that need the host's
// log10 -> corrected reads (k_vote_apply) -> [host: lengths, order of round 2] -> chains of round 2 on the corrected reads,
// staged device to device -> MSA rows -> fix_msa_ends + vote -> consensus.  D2H traffic: corrected reads and consensi.

static void vote_tables(PoaSlot &S, double *tab, unsigned char *symtab) {
    for (int i = 0; i < 256; ++i) {
        const double q = (char)i - 33;  // utils.cpp:11-13 (phred_err of a char)
        tab[i] = pow(10.0, -q / 10.0);
        symtab[i] = (unsigned char)(char)(-10 * log10(tab[i]) + 33);  // utils.cpp:6-9 (phred_symbol)
    }
    (void)S;
}

struct VoteRound1 {           // what round 2 needs from round 1, per pack of the batch
    std::vector<int> order;   // rows of the corrected reads, longest first (stable: fasta.cpp:458-464)
    uint64_t out_off = 0;     // the pack's block in v_out_seq
    int ncol = 0;
    bool ok = false;
};

static void vote_batch(rtl_ctx *ctx, PoaState &P, PoaSlot &S, const std::vector<VotePack *> &packs, std::vector<PoaTask> &tasks,
                       const std::vector<ChainSizes> &sizes, double min_occ, double gap_occ) {
    // chain_run waits for its chains, and everything below waits for its own work before the next chain_run: the
    // high-priority stream needs no events against the unit's stream
    cudaStream_t st = S.hi;
    const size_t np = packs.size();
    if (!S.v_tab_ready) {
        double tab[256];
        unsigned char symtab[256];
        vote_tables(S, tab, symtab);
        CK(cudaMemcpyAsync(S.v_tab.need(256), tab, sizeof(tab), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.v_symtab.need(256), symtab, sizeof(symtab), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));  // the tables are on this stack frame
        S.v_tab_ready = true;
    }
    // ---- round 1: chains
    std::vector<PoaTask *> batch(np);
    std::vector<std::vector<const char *>> quals(np);
    for (size_t i = 0; i < np; ++i) {
        batch[i] = &tasks[i];
        quals[i] = packs[i]->qual;
    }
    ChainInput in1;
    in1.quals = &quals;
    chain_run(ctx, P, S, batch, sizes, in1);
    const size_t msa_bytes = chain_msa_rows(S, np, st);
    // ---- round 1: vote
    const double tv0 = now_ms();
    std::vector<VoteRound1> r1(np);
    int r1_chain[4] = {0, 0, 0, 0}, r1_vote[3] = {0, 0, 0};  // (RTL_TRACE: why packs leave the device pipeline)
    size_t n_rows = 0, n_cols = 0;
    int max_ncol = 1;
    {
        const DCPack *hp = S.ch_packs.p;
        const DCSeq *hs = S.ch_seqs.p;
        const uint64_t *hmo = S.ch_msa_off.p;
        for (size_t i = 0; i < np; ++i) n_rows = std::max<size_t>(n_rows, (size_t)hp[i].seq_base + (size_t)hp[i].n_seq);
        DVPack *vp = S.vh_packs.need_geo(np);
        DVRead *vr = S.vh_reads.need_geo(n_rows + 1);
        for (size_t i = 0; i < np; ++i) {
            DVPack &V = vp[i];
            memset(&V, 0, sizeof(V));
            const bool ok = hp[i].status == DC_OK;
            r1_chain[std::min(3, std::max(0, (int)hp[i].status))]++;
            V.msa_off = hmo[i];
            V.out_off = hmo[i];
            V.col_off = n_cols;
            V.row_base = hp[i].seq_base;
            V.seq_base = hp[i].seq_base;
            V.n_seq = hp[i].n_seq;
            V.ncol = ok ? hp[i].ncol : 0;
            V.status = ok ? DV_OK : DV_DEGENERATE;
            if (ok) {
                n_cols += (size_t)hp[i].ncol;
                max_ncol = std::max(max_ncol, hp[i].ncol);
            }
            for (int s2 = 0; s2 < hp[i].n_seq; ++s2) {
                DVRead &R = vr[hp[i].seq_base + s2];
                R.qual_off = hs[hp[i].seq_base + s2].q_off;
                R.L = hs[hp[i].seq_base + s2].L;
                R.pad = 0;
            }
            r1[i].out_off = hmo[i];
            r1[i].ncol = V.ncol;
        }
        CK(cudaMemcpyAsync(S.v_packs.need_geo(np), vp, np * sizeof(DVPack), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.v_reads.need_geo(n_rows + 1), vr, n_rows * sizeof(DVRead), cudaMemcpyHostToDevice, st));
        S.st.h2d_bytes += (int64_t)(np * sizeof(DVPack) + n_rows * sizeof(DVRead));
    }
    S.v_rows.need_geo(n_rows + 1);
    S.v_cols.need_geo(n_cols + 1);
    S.v_qm.need_geo(msa_bytes + 16);
    S.v_out_seq.need_geo(msa_bytes + 16);
    S.v_out_qual.need_geo(msa_bytes + 16);
    S.v_len.need_geo(n_rows + 1);
    S.v_flagged.need_geo(n_cols + 1);
    CK(cudaMemsetAsync(S.v_nflag.need(1), 0, sizeof(unsigned int), st));
    const unsigned int flag_cap = (unsigned int)std::min<size_t>(n_cols, 0x7fffffffu);
    k_vote_rows<<<(unsigned)np, 64, 0, st>>>(S.v_packs.p, S.v_reads.p, S.c_msa.p, S.v_qm.p, reinterpret_cast<const char *>(S.c_qual.p),
                                             S.v_rows.p, 1);
    CK(cudaGetLastError());
    k_vote_cols<<<dim3((unsigned)((max_ncol + 127) / 128), (unsigned)np), 128, 0, st>>>(
        S.v_packs.p, S.c_msa.p, S.v_qm.p, S.v_rows.p, S.v_tab.p, S.v_symtab.p, S.v_cols.p, 1, S.v_flagged.p, S.v_nflag.p, flag_cap);
    CK(cudaGetLastError());
    S.st.kernel_launches += 2;
    unsigned int *hnf = S.vh_nflag.need(1);
    CK(cudaMemcpyAsync(hnf, S.v_nflag.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    slot_wait(S, st);
    if (*hnf > flag_cap) throw StateError("vote: more flagged columns than columns");
    if (*hnf) {  // the host's libm decides these quality symbols (poa_vote.cuh)
        const unsigned int nf = *hnf;
        int4 *hf = S.vh_flagged.need_geo(nf);
        CK(cudaMemcpyAsync(hf, S.v_flagged.p, (size_t)nf * sizeof(int4), cudaMemcpyDeviceToHost, st));
        slot_wait(S, st);
        for (unsigned int x = 0; x < nf; ++x) {
            const unsigned long long bits = (unsigned long long)(unsigned int)hf[x].z | ((unsigned long long)(unsigned int)hf[x].w << 32);
            double cerr;
            memcpy(&cerr, &bits, 8);
            hf[x].z = (int)(char)(-10 * log10(cerr) + 33);  // utils.cpp:6-9
            hf[x].w = 0;
        }
        CK(cudaMemcpyAsync(S.v_flagged.p, hf, (size_t)nf * sizeof(int4), cudaMemcpyHostToDevice, st));
        k_vote_patch<<<(nf + 255) / 256, 256, 0, st>>>(S.v_flagged.p, nf, S.v_cols.p);
        CK(cudaGetLastError());
        S.st.kernel_launches++;
        S.st.d2h_bytes += (int64_t)nf * 16;
        S.st.h2d_bytes += (int64_t)nf * 16;
    }
    k_vote_apply<<<(unsigned)np, 64, 0, st>>>(S.v_packs.p, S.c_msa.p, S.v_qm.p, S.v_rows.p, S.v_cols.p, S.v_tab.p, min_occ, gap_occ,
                                              S.v_out_seq.p, S.v_out_qual.p, S.v_len.p);
    CK(cudaGetLastError());
    S.st.kernel_launches++;
    DVPack *vp = S.vh_packs.p;
    DVRow *hrows = S.vh_rows.need_geo(n_rows + 1);
    int32_t *hlen = S.vh_len.need_geo(n_rows + 1);
    char *hos = S.vh_out_seq.need_geo(msa_bytes + 16);
    char *hoq = S.vh_out_qual.need_geo(msa_bytes + 16);
    CK(cudaMemcpyAsync(vp, S.v_packs.p, np * sizeof(DVPack), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hrows, S.v_rows.p, n_rows * sizeof(DVRow), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hlen, S.v_len.p, n_rows * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (msa_bytes) {
        CK(cudaMemcpyAsync(hos, S.v_out_seq.p, msa_bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hoq, S.v_out_qual.p, msa_bytes, cudaMemcpyDeviceToHost, st));
    }
    slot_wait(S, st);
    S.st.d2h_bytes += (int64_t)(np * sizeof(DVPack) + n_rows * (sizeof(DVRow) + 4) + 2 * msa_bytes);
    const double tv1 = now_ms();
    // ---- host: corrected reads of every pack, order of round 2
    parallel_for(S.n_threads, np, [&](size_t i) {
        VotePack &K = *packs[i];
        const DVPack &V = vp[i];
        r1[i].ok = V.status == DV_OK;
        if (!r1[i].ok) return;
        const int n = V.n_seq;
        K.tf.resize(n);
        K.tb.resize(n);
        K.cseq.assign(n, std::string());
        K.cqual.assign(n, std::string());
        for (int r = 0; r < n; ++r) {
            K.tf[r] = hrows[V.row_base + r].tf;
            K.tb[r] = hrows[V.row_base + r].tb;
            const int len = hlen[V.row_base + r];
            if (len > 0) {
                K.cseq[r].assign(hos + V.out_off + (size_t)r * V.ncol, (size_t)len);
                K.cqual[r].assign(hoq + V.out_off + (size_t)r * V.ncol, (size_t)len);
                r1[i].order.push_back(r);
            }
        }
        std::stable_sort(r1[i].order.begin(), r1[i].order.end(),
                         [&](int a, int b) { return K.cseq[a].size() > K.cseq[b].size(); });
    });
    for (size_t i = 0; i < np; ++i)
        if (S.ch_packs.p[i].status == DC_OK) r1_vote[std::min(2, std::max(0, (int)vp[i].status))]++;
    const double tv2 = now_ms();
    // ---- round 2: chains on the corrected reads (device to device), in sub-batches that fit the arena
    std::vector<size_t> todo;
    for (size_t i = 0; i < np; ++i) {
        if (!r1[i].ok) continue;
        if (r1[i].order.empty()) {  // nothing was corrected: empty consensus (correct.cpp:427-445 on an empty read set)
            packs[i]->consensus.clear();
            packs[i]->done = true;
            continue;
        }
        todo.push_back(i);
    }
    std::vector<PoaTask> t2(todo.size());
    std::vector<ChainSizes> z2(todo.size());
    const int maxabs = 8;
    for (size_t x = 0; x < todo.size(); ++x) {
        const size_t i = todo[x];
        for (int r : r1[i].order) t2[x].len.push_back((int)packs[i]->cseq[r].size());
        z2[x] = chain_sizes(ctx, &t2[x]);
        for (int l : t2[x].len)
            if ((int64_t)maxabs * (l + 16) >= 32000) z2[x].arena_bytes = ~(size_t)0;  // leaves int16: host path
    }
    size_t x0 = 0;
    while (x0 < todo.size()) {
        std::vector<PoaTask *> b2;
        std::vector<ChainSizes> s2;
        std::vector<DVStage> stage;
        std::vector<size_t> who;
        size_t used = 0, x1 = x0;
        for (; x1 < todo.size(); ++x1) {
            if (z2[x1].arena_bytes > S.arena_bytes) {  // cannot run here: the caller redoes the pack on the host path
                if (b2.empty()) {
                    ++x1;
                    break;
                }
                break;
            }
            if (used + z2[x1].arena_bytes > S.arena_bytes) break;
            used += z2[x1].arena_bytes;
            const size_t i = todo[x1];
            b2.push_back(&t2[x1]);
            s2.push_back(z2[x1]);
            who.push_back(i);
            for (int r : r1[i].order) {
                DVStage e;
                e.src_off = r1[i].out_off + (uint64_t)r * (uint64_t)r1[i].ncol;
                e.q_off = 0;
                e.L = (int32_t)packs[i]->cseq[r].size();
                stage.push_back(e);
            }
        }
        x0 = x1;
        if (b2.empty()) continue;
        ChainInput in2;
        in2.stage = &stage;
        in2.stage_src = S.v_out_seq.p;
        chain_run(ctx, P, S, b2, s2, in2);
        const size_t nb = b2.size();
        const size_t msa2 = chain_msa_rows(S, nb, st);
        const DCPack *hp = S.ch_packs.p;
        const uint64_t *hmo = S.ch_msa_off.p;
        DVPack *v2 = S.vh_packs.need_geo(nb);
        size_t rows2 = 0, cols2 = 0;
        int max2 = 1;
        for (size_t k = 0; k < nb; ++k) {
            DVPack &V = v2[k];
            memset(&V, 0, sizeof(V));
            const bool ok = hp[k].status == DC_OK;
            V.msa_off = hmo[k];
            V.out_off = cols2;  // the consensus goes to v_cons + out_off
            V.col_off = cols2;
            V.row_base = hp[k].seq_base;
            V.seq_base = hp[k].seq_base;
            V.n_seq = hp[k].n_seq;
            V.ncol = ok ? hp[k].ncol : 0;
            V.status = ok ? DV_OK : DV_DEGENERATE;
            if (ok) {
                cols2 += (size_t)hp[k].ncol;
                max2 = std::max(max2, hp[k].ncol);
            }
            rows2 = std::max<size_t>(rows2, (size_t)hp[k].seq_base + (size_t)hp[k].n_seq);
        }
        CK(cudaMemcpyAsync(S.v_packs.need_geo(nb), v2, nb * sizeof(DVPack), cudaMemcpyHostToDevice, st));
        S.v_rows.need_geo(rows2 + 1);
        S.v_cols.need_geo(cols2 + 1);
        S.v_cons.need_geo(cols2 + 16);
        S.v_flagged.need_geo(1);
        CK(cudaMemsetAsync(S.v_nflag.need(1), 0, sizeof(unsigned int), st));
        k_vote_rows<<<(unsigned)nb, 64, 0, st>>>(S.v_packs.p, S.v_reads.p, S.c_msa.p, S.v_qm.p, nullptr, S.v_rows.p, 0);
        CK(cudaGetLastError());
        k_vote_cols<<<dim3((unsigned)((max2 + 127) / 128), (unsigned)nb), 128, 0, st>>>(
            S.v_packs.p, S.c_msa.p, S.v_qm.p, S.v_rows.p, S.v_tab.p, S.v_symtab.p, S.v_cols.p, 0, S.v_flagged.p, S.v_nflag.p, 0u);
        CK(cudaGetLastError());
        k_vote_consensus<<<(unsigned)nb, 32, 0, st>>>(S.v_packs.p, S.v_cols.p, S.v_cons.p);
        CK(cudaGetLastError());
        S.st.kernel_launches += 3;
        char *hc = S.vh_cons.need_geo(cols2 + 16);
        CK(cudaMemcpyAsync(v2, S.v_packs.p, nb * sizeof(DVPack), cudaMemcpyDeviceToHost, st));
        if (cols2) CK(cudaMemcpyAsync(hc, S.v_cons.p, cols2, cudaMemcpyDeviceToHost, st));
        slot_wait(S, st);
        S.st.h2d_bytes += (int64_t)(nb * sizeof(DVPack));
        S.st.d2h_bytes += (int64_t)(nb * sizeof(DVPack) + cols2);
        (void)msa2;
        for (size_t k = 0; k < nb; ++k) {
            if (v2[k].status != DV_OK) continue;  // stays !done: host path
            VotePack &K = *packs[who[k]];
            K.consensus.assign(hc + v2[k].out_off, (size_t)v2[k].cons_len);
            K.done = true;
        }
    }
    if (getenv("RTL_TRACE")) {
        size_t n_done = 0;
        for (auto *k : packs) n_done += k->done;
        fprintf(stderr, "[rtl] device vote unit %d: %zu packs (%zu done; round-1 chain status ok/cap/degree/spill %d/%d/%d/%d, round-1 vote "
                "ok/degenerate/letter %d/%d/%d), round-1 vote kernels + D2H %.1f ms, host strings %.1f ms, round 2 %.1f ms\n",
                (int)(&S - P.slot_store), np, n_done, r1_chain[0], r1_chain[1], r1_chain[2], r1_chain[3], r1_vote[0], r1_vote[1], r1_vote[2],
                tv1 - tv0, tv2 - tv1, now_ms() - tv2);
    }
}

void poa_correct_unit(rtl_ctx *ctx, int unit, std::vector<VotePack *> &packs, double min_occ, double gap_occ, int n_threads) {
    PoaState &P = pstate(ctx);
    if (unit < 0 || unit >= P.n_units) throw StateError("poa_correct_unit: no such unit");
    CK(cudaSetDevice(ctx->device));
    PoaSlot &S = P.slot_store[unit];
    S.n_threads = std::max(1, n_threads);
    S.t_wait = S.t_fold = S.t_stage = 0;
    S.st = rtl_stats{};
    S.epoch++;
    const double t_begin = now_ms();
    const uint8_t *tab = letter_codes();
    std::vector<PoaTask> tasks(packs.size());
    std::vector<uint8_t> elig(packs.size(), 0);
    const bool dev_ok = ctx->poa_device_chain != 0 && ctx->poa_gpu_sort != 0 && ctx->poa_kernel != 1;
    parallel_for(S.n_threads, packs.size(), [&](size_t i) {
        VotePack &K = *packs[i];
        K.done = false;
        PoaTask &t = tasks[i];
        t.seq = K.seq;
        t.len = K.len;
        bool ok = dev_ok && !K.seq.empty() && K.seq.size() == K.qual.size();
        for (size_t s2 = 0; s2 < K.seq.size() && ok; ++s2) {
            const int l = K.len[s2];
            if (l < 1 || (int64_t)8 * (l + 16) >= 32000) ok = false;
            for (int x = 0; x < l && ok; ++x)
                if (tab[(unsigned char)K.seq[s2][x]] == 255) ok = false;
        }
        elig[i] = ok;
    });
    std::vector<VotePack *> bp;
    std::vector<PoaTask> bt;
    std::vector<ChainSizes> bz;
    size_t used = 0;
    size_t n_dev = 0;
    auto flush = [&]() {
        if (bp.empty()) return;
        n_dev += bp.size();
        vote_batch(ctx, P, S, bp, bt, bz, min_occ, gap_occ);
        bp.clear();
        bt.clear();
        bz.clear();
        used = 0;
    };
    for (size_t i = 0; i < packs.size(); ++i) {
        if (!elig[i]) continue;
        const ChainSizes z = chain_sizes(ctx, &tasks[i]);
        if (z.arena_bytes > S.arena_bytes) continue;
        if (used + z.arena_bytes > S.arena_bytes) flush();
        bp.push_back(packs[i]);
        bt.push_back(tasks[i]);
        bz.push_back(z);
        used += z.arena_bytes;
    }
    flush();
    {
        std::lock_guard<std::mutex> lk(P.stats_mu);
        rtl_stats &d = ctx->stats;
        d.h2d_bytes += S.st.h2d_bytes;
        d.d2h_bytes += S.st.d2h_bytes;
        d.poa_launches += S.st.poa_launches;
        d.kernel_launches += S.st.kernel_launches;
        d.poa_alignments += S.st.poa_alignments;
        d.poa_cells += S.st.poa_cells;
        d.poa_dram_bytes += S.st.poa_dram_bytes;
        d.poa_ms += S.st.poa_ms;
        size_t n_done = 0;
        for (auto *k : packs) n_done += k->done;
        if (getenv("RTL_TRACE"))
            fprintf(stderr, "[rtl] poa_correct_unit %d: %zu packs (%zu on the device pipeline, %zu finished there), %.1f ms "
                    "(waited for GPU %.1f, stage+submit %.1f), %d host threads\n", unit, packs.size(), n_dev, n_done,
                    now_ms() - t_begin, S.t_wait, S.t_stage, S.n_threads);
    }
}

// Host-driven lock-step chain of one unit (the first-generation path, and the fallback of the device-resident chains):
// every step synchronises with the host, which runs Graph::add_alignment and stages the next step.
static void poa_chain_host(rtl_ctx *ctx, PoaState &P, PoaSlot &S, int unit, std::vector<PoaTask *> &tasks, int sm, int sn,
                           int sg, int se, bool keep_alns) {
    (void)unit;
    size_t max_steps = 0;
    for (auto *t : tasks) max_steps = std::max(max_steps, t->seq.size());
    const int maxabs = std::max(std::max(std::abs(sm), std::abs(sn)), std::max(std::abs(sg), std::abs(se)));
    // Device mirrors (poa_devgraph.cuh) for the graphs whose every read the strip kernel can take: the GPU then sorts
    // the graph and builds the row records, the host only runs add_alignment.  Capacities are generous multiples of the
    // longest read; a graph that outgrows them (or meets in-degree > 32) is demoted to the host path.
    {
        const bool want = ctx->poa_gpu_sort != 0 && ctx->poa_kernel != 1 && sm == 5 && sn == -4 && sg == -8 && se == -6;
        size_t words = 0;
        const size_t budget_words = (size_t)6 << 28;  // 6 GB of mirrors per unit at most
        for (auto *t : tasks) {
            t->mirror = false;
            t->sync_n = t->sync_e = t->sync_a = 0;
            if (!want || !t->acgtu) continue;
            int maxlen = 0;
            long long total = 0;
            for (int l : t->len) {
                maxlen = std::max(maxlen, l);
                total += l;
            }
            if (maxlen == 0 || (int64_t)maxabs * (maxlen + 16) >= 32000) continue;
            // (option poa_mirror_pct shrinks the capacity: tests use it to exercise the demotion to the host path)
            const long long cn = std::max<long long>(16, std::min<long long>(total + 8, 6ll * maxlen + 2048) * ctx->poa_mirror_pct / 100);
            const size_t w = dg_words((int)cn, (int)(3 * cn), (int)(4 * cn));
            if (words + w > budget_words) continue;
            t->mirror = true;
            t->cap_n = (int)cn;
            t->cap_e = (int)(3 * cn);
            t->cap_a = (int)(4 * cn);
            t->gbase = words;
            t->g.defer_sort = true;
            words += w;
        }
        if (words) S.d_pool.need(words);
    }
    // the int16 strip kernel is compiled for RATTLE's scores (correct.cpp:395: 5,-4,-8,-6); other scores take the
    // int32 kernel, and option poa_kernel=1 forces it
    const bool strip_scores = ctx->poa_kernel != 1 && sm == 5 && sn == -4 && sg == -8 && se == -6;
    for (size_t step = 0; step < max_steps; ++step) {
        std::vector<JobRef> all;
        std::vector<PoaTask *> direct;
        for (auto *t : tasks) {
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
                jr.hf_bytes = 0;  // after plan_spills
                jr.code_bytes = ps_code_words(jr.n, jr.nst) * 4;
            } else {
                jr.hf_bytes = (size_t)(jr.n + 1) * poa_ws(L) * (wide ? 8 : 4);
                jr.code_bytes = (size_t)jr.n * poa_lp(L) * (wide ? 4 : 2);
            }
            all.push_back(jr);
        }
        parallel_for(S.n_threads, direct.size(), [&](size_t i) {
            PoaTask *t = direct[i];
            t->g.add_alignment({}, t->seq[step], t->len[step]);
        });
        // graphs with a device mirror: fold + sort + row records on the GPU; the others are sorted by the host
        S.base_rows = S.base_preds = S.base_spill = 0;
        std::vector<size_t> mj;
        for (size_t k = 0; k < all.size(); ++k) {
            JobRef &jr = all[k];
            PoaTask *t = jr.task;
            if (!t->mirror) continue;
            const PoaGraph &g = t->g;
            if (jr.kind != JK_STRIP || g.n_nodes() > t->cap_n || (int)g.e_begin.size() > t->cap_e ||
                (int)g.a_node.size() > t->cap_a) {
                demote_mirror(t);
                continue;
            }
            jr.mirror = true;
            mj.push_back(k);
        }
        if (!mj.empty()) run_fold(ctx, S, all, mj);
        parallel_for(S.n_threads, all.size(), [&](size_t k) {
            JobRef &jr = all[k];
            if (jr.kind != JK_STRIP) return;
            if (jr.mirror) {
                jr.hf_bytes = ps_hf_words(jr.n, jr.nst, jr.n_spill) * 4;
                return;
            }
            jr.n_spill = plan_spills(jr.task, strip_ring_rows(strip_warps(jr.nst)));
            if (jr.n_spill >= 65535) {  // spill slots are 16-bit in the row records
                jr.kind = JK_NARROW;
                jr.hf_bytes = (size_t)(jr.n + 1) * poa_ws(jr.L) * 4;
                jr.code_bytes = (size_t)jr.n * poa_lp(jr.L) * 2;
            } else {
                jr.hf_bytes = ps_hf_words(jr.n, jr.nst, jr.n_spill) * 4;
            }
        });
        // groups that fit this unit's arena slice, one after the other
        size_t i = 0;
        while (i < all.size()) {
            std::vector<JobRef> group;
            size_t used = 0;
            while (i < all.size()) {
                const size_t need = used + ((all[i].hf_bytes + 255) & ~(size_t)255) + ((all[i].code_bytes + 255) & ~(size_t)255);
                if (need > S.arena_bytes) {
                    if (group.empty())
                        throw CapacityError("one POA alignment does not fit the device arena: raise option poa_arena_mb");
                    break;
                }
                used = need;
                group.push_back(all[i++]);
            }
            submit(ctx, P, S, std::move(group), sm, sn, sg, se, keep_alns);
            finish(ctx, P, S);
        }
    }
}

// device time during which at least one POA launch group was running since the last call (union of the groups'
// [start, end] intervals, CUDA events); added to stats.poa_busy_ms
void poa_account_busy(rtl_ctx *ctx) {
    PoaState &P = pstate(ctx);
    std::lock_guard<std::mutex> lk(P.stats_mu);
    std::sort(P.intervals.begin(), P.intervals.end());
    double busy = 0;
    float cur_a = 0, cur_b = -1;
    for (auto &iv : P.intervals) {
        if (cur_b < cur_a || iv.first > cur_b) {
            if (cur_b >= cur_a) busy += cur_b - cur_a;
            cur_a = iv.first;
            cur_b = iv.second;
        } else if (iv.second > cur_b) {
            cur_b = iv.second;
        }
    }
    if (cur_b >= cur_a) busy += cur_b - cur_a;
    P.intervals.clear();
    ctx->stats.poa_busy_ms += busy;
}

int poa_unit_count(rtl_ctx *ctx, size_t n_tasks) {
    PoaState &P = pstate(ctx);
    // a unit should keep a fair share of the GPU's CTA slots busy by itself
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)P.n_units, n_tasks / 48));
}

void run_units(int n_units, const std::function<void(int)> &fn) {
    if (n_units <= 1) {
        fn(0);
        return;
    }
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> err(n_units);
    for (int u = 0; u < n_units; ++u)
        th.emplace_back([&, u]() {
            try {
                fn(u);
            } catch (...) {
                err[u] = std::current_exception();
            }
        });
    for (auto &x : th) x.join();
    for (auto &e : err)
        if (e) std::rethrow_exception(e);
}

// All tasks, split round-robin into units that run concurrently.
void poa_run(rtl_ctx *ctx, std::vector<PoaTask *> &tasks, int sm, int sn, int sg, int se, bool keep_alns) {
    PoaState &P = pstate(ctx);
    CK(cudaStreamSynchronize(ctx->stream));  // inputs produced on the ctx stream are complete
    const double t0 = now_ms();
    const int U = poa_unit_count(ctx, tasks.size());
    std::vector<std::vector<PoaTask *>> unit(U);
    for (size_t i = 0; i < tasks.size(); ++i) unit[i % U].push_back(tasks[i]);
    const int nt = P.n_threads;  // the shared pool arbitrates between the units
    run_units(U, [&](int u) { poa_chain(ctx, u, unit[u], sm, sn, sg, se, keep_alns, nt); });
    ctx->stats.poa_wall_ms += now_ms() - t0;
    poa_account_busy(ctx);
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
    task.take_msa(msa);
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
