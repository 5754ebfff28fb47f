// Thread-per-CUDA-thread shim for the host emulation of kernels whose threads cooperate (warp shuffles, block barriers) —
// TEST TOOL ONLY.  Every CUDA thread of a block is a host thread; a shuffle is an exchange through a slot array between two
// waits of the (warp, group) barrier, __syncthreads a wait of the block barrier.
#pragma once
#include <math.h>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
using std::isfinite;

#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__
#define __constant__ static const
#define __host__
#define CUDART_INF INFINITY
#define CUDART_NAN NAN
struct emu_dim3 { int x, y, z; };
static thread_local emu_dim3 threadIdx;
static emu_dim3 blockIdx, blockDim, gridDim;
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }

struct Barrier {
    std::mutex m; std::condition_variable cv; int count = 0; unsigned gen = 0;
    void wait(int n) {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
// a warp's lanes cooperate in groups of `width` (= lanes per trajectory, 32 / 16 / 8): one barrier per (warp, group)
static Barrier g_block_bar, g_group_bar[4][4];
static unsigned long long g_slot[4][32];
static int g_block_threads = 128;
static inline void __syncthreads() { g_block_bar.wait(g_block_threads); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    const int width = __builtin_popcount(mask), gi = __builtin_ctz(mask) / width;
    g_group_bar[threadIdx.x >> 5][gi].wait(width);
}
template <class T> static inline T emu_shfl(T v, int src_lane, int width) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, gi = l / width;
    unsigned long long raw = 0; memcpy(&raw, &v, sizeof(T));
    g_slot[w][l] = raw;
    g_group_bar[w][gi].wait(width);
    raw = g_slot[w][gi * width + (src_lane & (width - 1))];
    g_group_bar[w][gi].wait(width);
    T r; memcpy(&r, &raw, sizeof(T)); return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) { return emu_shfl(v, src, width); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o, int width = 32) { return emu_shfl(v, ((threadIdx.x & 31) & (width - 1)) ^ o, width); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
