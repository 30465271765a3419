// Interface between the POA batch driver (poa_engine.cu) and its callers (rtl_poa_msa, correct_engine.cu).
#pragma once
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "poa_graph.hpp"

struct rtl_ctx;

// one partial-order alignment: an ordered list of sequences (views; the caller keeps them alive)
struct PoaTask {
    std::vector<const char *> seq;
    std::vector<int> len;
    rtl::PoaGraph g;
    bool acgtu = false;  // every letter is one of A,C,G,T,U (set by poa_chain)
    std::vector<int32_t> spill_slot;  // scratch of the strip kernel's staging (poa_engine.cu:plan_spills)
    // device mirror of the graph (poa_devgraph.cuh; set up by poa_chain): block offset and capacities in the unit's pool,
    // and how much of the graph's append logs the mirror has seen
    bool mirror = false;
    uint64_t gbase = 0;
    int cap_n = 0, cap_e = 0, cap_a = 0;
    int sync_n = 0, sync_e = 0, sync_a = 0;
    std::vector<std::vector<std::pair<int32_t, int32_t>>> alns;  // filled when keep_alns
    // packs whose whole chain ran on the GPU (poa_devchain.cuh) come back as MSA rows; `g` then stays empty
    std::vector<std::string> msa_rows;
    bool have_msa = false;
    // the multiple sequence alignment of the task (graph.cpp:390-426 without the consensus row); releases the graph
    void take_msa(std::vector<std::string> &dst) {
        if (have_msa) dst = std::move(msa_rows);
        else g.msa(dst);
        msa_rows.clear();
        have_msa = false;
        g.clear();
    }
};

// Runs all tasks on the GPU (split into concurrently running units); afterwards task->g holds the final graph
// (call g.msa()).
void poa_run(rtl_ctx *ctx, std::vector<PoaTask *> &tasks, int m, int n, int g, int e, bool keep_alns);
// The pieces of poa_run for callers that keep more per-pack host work inside the unit threads:
// poa_unit_count = units to split n_tasks into, run_units = one thread per unit (exceptions re-thrown after join),
// poa_chain = the synchronous chain of one unit's tasks on that unit's stream / arena slice.
int poa_unit_count(rtl_ctx *ctx, size_t n_tasks);
void poa_account_busy(rtl_ctx *ctx);  // folds the launch intervals since the last call into stats.poa_busy_ms
void run_units(int n_units, const std::function<void(int)> &fn);
void poa_chain(rtl_ctx *ctx, int unit, std::vector<PoaTask *> &tasks, int m, int n, int g, int e, bool keep_alns,
               int n_threads);
void parallel_for(int n_threads, size_t n, const std::function<void(size_t)> &fn);

// One pack of correct_reads whose whole pipeline (correct.cpp:395-445: POA round 1, fix_msa_ends, column vote, read
// correction, POA round 2 on the corrected reads, fix_msa_ends, vote -> consensus) runs on the GPU (poa_vote.cuh): the MSAs
// never travel to the host.
struct VotePack {
    std::vector<const char *> seq, qual;  // the pack's reads in POA order (views; the caller keeps them alive)
    std::vector<int> len;
    bool done = false;                    // false after the call: the caller runs the host pipeline for this pack
    std::vector<int32_t> tf, tb;          // bases fix_msa_ends (round 1) removed at the front / the back of every read
    std::vector<std::string> cseq, cqual; // corrected read per input read; empty = the read stays uncorrected
    std::string consensus;
};
// one unit's packs through that pipeline, synchronously, on the unit's stream / arena slice (like poa_chain)
void poa_correct_unit(rtl_ctx *ctx, int unit, std::vector<VotePack *> &packs, double min_occ, double gap_occ, int n_threads);
int host_threads();
