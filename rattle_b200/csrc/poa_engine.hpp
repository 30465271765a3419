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
    bool acgtu = false;  // every letter is one of A,C,G,T,U (set by poa_run)
    std::vector<std::vector<std::pair<int32_t, int32_t>>> alns;  // filled when keep_alns
};

// Runs all tasks in lock-step on the GPU; afterwards task->g holds the final graph (call g.msa()).
void poa_run(rtl_ctx *ctx, std::vector<PoaTask *> &tasks, int m, int n, int g, int e, bool keep_alns);
void parallel_for(int n_threads, size_t n, const std::function<void(size_t)> &fn);
int host_threads();
