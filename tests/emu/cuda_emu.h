// Minimal CUDA execution-model shim for the CPU (TEST INFRASTRUCTURE ONLY).
//
// tests/emu/ compiles the very kernel sources of the product (pyro_b200/csrc/pyrodp_device.cuh,
// sweep_fused.cuh) with g++ and runs them on the host: one coroutine per CUDA thread of a block (a
// hand-rolled x86-64 context switch, all on one OS thread, so the run is deterministic), blocks one after
// the other, __syncthreads / warp votes / shuffles as real rendezvous between those coroutines.  It lets
// the CPU test-suite check the kernels' ARITHMETIC and control flow (cell walks, lane splits, argmin
// ties, padding) bit for bit against the reference fixtures without a GPU.  It is not a fallback: nothing
// under pyro_b200/ knows about it, it is orders of magnitude slower than the reference itself, and it says
// nothing about the kernels' performance or about CUDA-specific behaviour (memory model, launch limits).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static   // one block runs at a time: block-shared == process-global ("extern __shared__" is rewritten by the build)

struct emu_uint3 { unsigned x, y, z; };
typedef emu_uint3 dim3_emu;
extern emu_uint3 threadIdx, blockIdx;   // of the coroutine that is running (set by the scheduler on every switch)
extern emu_uint3 blockDim, gridDim;

struct __attribute__((aligned(16))) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }

// ---- block / warp rendezvous -------------------------------------------------------------------------
struct EmuGroup {              // a set of coroutines that meet at barriers: the block, or one warp
    int size = 0, arrived = 0;
    unsigned long long generation = 0;
    unsigned long long slot[32];
};
struct EmuBlock {
    EmuGroup all;
    std::vector<EmuGroup> warps;
};
extern EmuBlock* emu_block;
void emu_barrier(EmuGroup& g);   // returns when every member of g has called it
static inline int emu_lane() { return (int)((threadIdx.x + threadIdx.y * blockDim.x) & 31); }
static inline EmuGroup& emu_warp() { return emu_block->warps[(threadIdx.x + threadIdx.y * blockDim.x) >> 5]; }

static inline void __syncthreads() { emu_barrier(emu_block->all); }
static inline void __syncwarp() { emu_barrier(emu_warp()); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <typename T>
static inline T emu_exchange(T v, int src_lane) {   // every lane of the warp calls this (full mask)
    static_assert(sizeof(T) <= 8, "shuffle payload");
    EmuGroup& w = emu_warp();
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.slot[emu_lane()] = bits;
    emu_barrier(w);
    const unsigned long long got = w.slot[src_lane & 31];
    emu_barrier(w);
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = emu_lane(), base = lane & ~(width - 1);
    return emu_exchange(v, base + (src & (width - 1)));
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int lane = emu_lane(), base = lane & ~(width - 1), src = lane + (int)delta;
    return emu_exchange(v, src < base + width ? src : lane);   // out of the segment: own value, as the hardware does
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
    const int lane = emu_lane(), base = lane & ~(width - 1), src = lane ^ mask;
    return emu_exchange(v, (src >= base && src < base + width) ? src : lane);
}
static inline int __any_sync(unsigned, int pred) {
    EmuGroup& w = emu_warp();
    w.slot[emu_lane()] = pred ? 1ull : 0ull;
    emu_barrier(w);
    unsigned long long any = 0;
    for (int i = 0; i < 32; ++i) any |= w.slot[i];
    emu_barrier(w);
    return any != 0;
}

static inline int emu_reduce_int(int v, bool want_max) {
    EmuGroup& w = emu_warp();
    w.slot[emu_lane()] = (unsigned long long)(long long)v;
    emu_barrier(w);
    int r = (int)(long long)w.slot[0];
    for (int i = 1; i < 32; ++i) {
        const int x = (int)(long long)w.slot[i];
        r = want_max ? (x > r ? x : r) : (x < r ? x : r);
    }
    emu_barrier(w);
    return r;
}
static inline int __reduce_min_sync(unsigned, int v) { return emu_reduce_int(v, false); }
static inline int __reduce_max_sync(unsigned, int v) { return emu_reduce_int(v, true); }

// ---- memory / conversion intrinsics ----------------------------------------------------------------------
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v) {
    unsigned long long old = __atomic_load_n(a, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned int atomicAdd(unsigned int* a, unsigned int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }

// saturating double -> int conversions (cvt.rpi / cvt.rmi .s32.f64: NaN -> 0)
static inline int emu_sat_int(double v) {
    if (!(v == v)) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (int)(-2147483647 - 1);
    return (int)v;
}
static inline int __double2int_ru(double x) { return emu_sat_int(std::ceil(x)); }
static inline int __double2int_rd(double x) { return emu_sat_int(std::floor(x)); }
using std::fabs;

// CUDA's overloaded min / max
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
using std::fma;
using std::fmax;
using std::fmin;
using std::sqrt;

// ---- launch: blocks one after the other, the threads of a block as coroutines -----------------------------------
void emu_launch(emu_uint3 grid, emu_uint3 block, const std::function<void()>& kernel_body);
