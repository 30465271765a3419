// Shared host-side plumbing of librattle_b200: context, device buffers, error handling, timing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rattle_b200.h"

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};
struct InputError : std::runtime_error {
    explicit InputError(const std::string &m) : std::runtime_error(m) {}
};
struct CapacityError : std::runtime_error {
    explicit CapacityError(const std::string &m) : std::runtime_error(m) {}
};
struct StateError : std::runtime_error {
    explicit StateError(const std::string &m) : std::runtime_error(m) {}
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            throw CudaError(std::string(#call) + " -> " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                            std::to_string(__LINE__) + ")");                                             \
    } while (0)

// grow-only device buffer
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    T *need(size_t n) {
        if (n > cap) {
            release();
            size_t want = n + n / 8 + 64;
            CK(cudaMalloc((void **)&p, want * sizeof(T)));
            cap = want;
        }
        return p;
    }
    // for buffers that grow step by step inside a pipeline: cudaFree synchronises the device, so grow geometrically
    T *need_geo(size_t n) {
        if (n > cap) {
            release();
            size_t want = 2 * n + 64;
            CK(cudaMalloc((void **)&p, want * sizeof(T)));
            cap = want;
        }
        return p;
    }
};

// pinned host staging buffer (grow-only)
template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    ~PinBuf() {
        if (p) cudaFreeHost(p);
    }
    T *need(size_t n) {
        if (n > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            CK(cudaMallocHost((void **)&p, (n + 64) * sizeof(T)));
            cap = n + 64;
        }
        return p;
    }
    T *need_geo(size_t n) {
        if (n > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            CK(cudaMallocHost((void **)&p, (2 * n + 64) * sizeof(T)));
            cap = 2 * n + 64;
        }
        return p;
    }
};

struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    void init() {
        if (!a) {
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
        }
    }
    ~EventTimer() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

struct ClusterState;  // cluster_engine.cu
struct PoaState;      // poa_engine.cu

struct rtl_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;      // stream in use (own_stream unless rtl_set_stream)
    cudaStream_t own_stream = nullptr;
    std::string err;
    rtl_stats stats{};
    // options
    int wave = 512;
    int64_t task_cap = 32ll << 20;
    int64_t scratch_mb = 1024;
    int bv_kernel = 0;         // 2 = scans with at most 16 seeds use the bulk-copy ring kernel k_bv_stream (default: k_bv_scan)
    int poa_batch = 0;
    int poa_units = 0;         // 0 = 12 concurrently running units (set before the first POA call)
    int poa_gpu_sort = 1;      // 1 = graphs are sorted and their row records built on the GPU (poa_devgraph.cuh), 0 = on the host
    int poa_mirror_pct = 100;  // capacity of the device graph mirrors, percent of the default (tests)
    int poa_kernel = 0;        // 0 = int16 strip kernel where eligible, 1 = int32 kernel only
    int poa_device_chain = 1;  // 1 = whole per-pack chains on the GPU (poa_devchain.cuh), 0 = host-driven lock-steps
    int poa_device_vote = 1;   // 1 = fix_msa_ends / column vote / read correction on the GPU (poa_vote.cuh), 0 = on the host
    int64_t poa_arena_mb = 0;  // 0 = 40 % of free device memory, at most 64 GB
    std::vector<int32_t> cluster_ids;  // global ids of the clusters of the next rtl_correct_reads calls (rtl_set_cluster_ids)
    std::vector<std::string> labels;  // file labels of `rattle correct -l` (rtl_set_labels; correct.cpp:447-470,488-512)
    // sharding
    int rank = 0, world = 1;
    rtl_allreduce_min_fn allreduce = nullptr;
    void *allreduce_user = nullptr;
    rtl_broadcast_fn broadcast = nullptr;  // sharded extraction (rtl_set_broadcast)
    void *broadcast_user = nullptr;
    ClusterState *cl = nullptr;
    PoaState *poa = nullptr;
};

inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
