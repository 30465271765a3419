// Minimal SIMT emulation for CPU tests: runs a CUDA kernel's SOURCE on the host, one OS thread per CUDA thread.
// Covers exactly what rattle_b200/csrc/poa_strip_kernel.cuh and poa_devgraph.cuh use: block barrier, warp collectives
// (shuffles, ballot), the DPX packed-int16 intrinsics, volatile shared-memory mailboxes (real threads, so the
// kernel's producer/consumer spins between warps work as on the device), atomics.  One CTA runs at a time.
// TEST INFRASTRUCTURE ONLY — nothing under rattle_b200/ includes this.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define CUDA_EMU 1

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct int4 { int32_t x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
using std::max;
using std::min;

namespace emu {
struct Barrier {  // reusable barrier
    std::mutex m;
    std::condition_variable cv;
    int count = 0, waiting = 0, gen = 0;
    void init(int n) { count = n; waiting = 0; gen = 0; }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const int g = gen;
        if (++waiting == count) {
            waiting = 0;
            ++gen;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};
struct Warp {
    Barrier bar;
    uint64_t slot[32];
};
struct Block {
    Barrier bar;
    std::vector<Warp> warps;
};
extern Block *g_block;
extern thread_local int t_tid;
}  // namespace emu

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

inline void __syncthreads() { emu::g_block->bar.wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::g_block->warps[emu::t_tid >> 5].bar.wait(); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }

// warp collectives: every lane publishes, barrier, reads, barrier
template <typename T>
inline T emu_exchange(T v, int src_lane) {
    emu::Warp &w = emu::g_block->warps[emu::t_tid >> 5];
    const int lane = emu::t_tid & 31;
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.slot[lane] = raw;
    w.bar.wait();
    uint64_t got = w.slot[src_lane & 31];
    w.bar.wait();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, int d) {
    const int lane = emu::t_tid & 31;
    return emu_exchange(v, lane >= d ? lane - d : lane);
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (emu::t_tid & 31) ^ m); }
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
inline unsigned __ballot_sync(unsigned, bool p) {
    emu::Warp &w = emu::g_block->warps[emu::t_tid >> 5];
    const int lane = emu::t_tid & 31;
    w.slot[lane] = p ? 1 : 0;
    w.bar.wait();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned)(w.slot[i] & 1) << i;
    w.bar.wait();
    return m;
}
inline int __ffs(unsigned x) { return x ? __builtin_ctz(x) + 1 : 0; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }

// packed int16 (DPX) intrinsics and friends
inline int16_t emu_lo(uint32_t v) { return (int16_t)(v & 0xffffu); }
inline int16_t emu_hi(uint32_t v) { return (int16_t)(v >> 16); }
inline uint32_t emu_pk(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
inline uint32_t __vadd2(uint32_t a, uint32_t b) { return emu_pk(emu_lo(a) + emu_lo(b), emu_hi(a) + emu_hi(b)); }  // wraps
inline uint32_t __vmaxs2(uint32_t a, uint32_t b) { return emu_pk(max<int>(emu_lo(a), emu_lo(b)), max<int>(emu_hi(a), emu_hi(b))); }
inline uint32_t __vminu2(uint32_t a, uint32_t b) {
    return emu_pk((int)min<uint32_t>(a & 0xffffu, b & 0xffffu), (int)min<uint32_t>(a >> 16, b >> 16));
}
inline uint32_t __vimax_s16x2_relu(uint32_t a, uint32_t b) {
    return emu_pk(max<int>(max<int>(emu_lo(a), emu_lo(b)), 0), max<int>(max<int>(emu_hi(a), emu_hi(b)), 0));
}
inline uint32_t __viaddmax_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxs2(__vadd2(a, b), c); }
inline uint32_t __viaddmin_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t s = __vadd2(a, b);
    return emu_pk(max<int>(min<int>(emu_lo(s), emu_lo(c)), 0), max<int>(min<int>(emu_hi(s), emu_hi(c)), 0));
}
inline uint32_t __vimax3_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxs2(__vmaxs2(a, b), c); }
inline int __viaddmax_s32(int a, int b, int c) { return max(a + b, c); }
inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t both = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((both >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicCAS(unsigned *p, unsigned expected, unsigned desired) {
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;  // the value found
}
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// shared memory: the dynamic part is one host buffer (emu::g_smem), `__shared__` statics are function statics (one CTA
// runs at a time); "shared-space addresses" are plain pointers (poa_strip_kernel.cuh: ps_saddr)
namespace emu {
extern unsigned char g_smem[256 * 1024];
}
#define __cvta_generic_to_shared(p) (reinterpret_cast<uintptr_t>(p))

namespace emu {
// run `kernel` as <<<grid, block>>> (CTAs one after the other, threads concurrently)
inline void launch(unsigned grid, unsigned block, const std::function<void()> &kernel) {
    for (unsigned b = 0; b < grid; ++b) {
        Block blk;
        blk.bar.init((int)block);
        blk.warps = std::vector<Warp>((block + 31) / 32);
        for (unsigned w = 0; w < blk.warps.size(); ++w) blk.warps[w].bar.init((int)std::min(32u, block - 32 * w));
        g_block = &blk;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t]() {
                t_tid = (int)t;
                threadIdx.x = t;
                blockIdx.x = b;
                blockDim.x = block;
                gridDim.x = grid;
                kernel();
            });
        for (auto &x : th) x.join();
        g_block = nullptr;
    }
}
}  // namespace emu

#ifdef CUDA_EMU_IMPLEMENTATION
namespace emu {
Block *g_block = nullptr;
thread_local int t_tid = 0;
alignas(16) unsigned char g_smem[256 * 1024];
}  // namespace emu
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#endif
