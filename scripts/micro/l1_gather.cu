// L1 data-pipe micro-benchmark (B200, sm_100a): what a warp-wide 8-byte gather costs for the address patterns of the 4-D
// sweep kernels (all data L1-resident, so this is the throughput of the L1 pipe itself, not of L2/HBM).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/l1_gather scripts/micro/l1_gather.cu
// Patterns (lane -> element index inside a per-block window of doubles):
//   0  32 consecutive doubles, 128-byte aligned                      (2 lines: the floor of a 256-byte warp access)
//   1  32 consecutive doubles, misaligned by 8 bytes                 (3 lines: the usual case, k3 is arbitrary)
//   2  lanes 0-19 and 20-31 in two different rows, misaligned        (cfg3 / cfg4: the (c0,c1) plane changes every ~20 lanes)
//   3  groups of 5 lanes in 7 different rows, misaligned             (cfg5: dt*step3/step1 = 0.2, the plane changes every 5 lanes)
//   4  as 0, from shared memory (LDS.64)
// Prints warp-level loads per clock per SM (launch duration by CUDA events, SM clock measured in-kernel) and clocks per load.
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 4096
#define ROW 512          // doubles per row of the window
#define ROWS 8
#define UNROLL 8
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int PATTERN>
__global__ void __launch_bounds__(256) k(const double* __restrict__ buf, double* out, long long* cycles, long long* nanos) {
    __shared__ double sm[ROW * 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < ROW * 2; i += blockDim.x) sm[i] = buf[i];
    __syncthreads();
    int idx;
    if (PATTERN == 0 || PATTERN == 4) idx = lane;
    else if (PATTERN == 1) idx = lane + 1;
    else if (PATTERN == 2) idx = (lane < 20 ? 0 : ROW) + lane + 1;
    else idx = (lane / 5) * ROW + lane + 1;
    const double* base = buf + (size_t)blockIdx.x * 0 + warp * 16;   // every block reads the same 32 KB window: L1 hits
    double acc[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc[u] = 0.0;
    unsigned long long g0 = gtime();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int o = idx + ((it + u) & 7) * 16;     // stays inside the row, keeps alignment class (16 doubles = 128 bytes)
            double v;   // volatile asm: the compiler may neither hoist nor merge the loads (the window repeats every 8 iterations)
            if (PATTERN == 4) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(sm + o);
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa));
            } else {
                asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(base + o));
            }
            acc[u] += v;
        }
    }
    long long t1 = clock64();
    unsigned long long g1 = gtime();
    double s = 0;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) s += acc[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; nanos[blockIdx.x] = (long long)(g1 - g0); }
}

template <int PATTERN>
static void run(const char* name, const double* buf) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks_per_sm = 4, threads = 256, blocks = sms * blocks_per_sm;
    double* out; long long* cyc; long long* ns;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(double));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    cudaMalloc(&ns, blocks * sizeof(long long));
    k<PATTERN><<<blocks, threads>>>(buf, out, cyc, ns);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<PATTERN><<<blocks, threads>>>(buf, out, cyc, ns);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long hc[4096], hn[4096];
    cudaMemcpy(hc, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(hn, ns, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double c = 0, n = 0; for (int i = 0; i < blocks; ++i) { c += hc[i]; n += hn[i]; }
    const double ticks_per_ns = c / n;
    const double launch_clocks = (double)ms * 1e6 * ticks_per_ns;
    const double loads_per_sm = (double)blocks_per_sm * (threads / 32) * ITERS * UNROLL;
    printf("%-58s : %.3f warp-loads/clk/SM = %.2f clocks per warp-load | launch %.3f ms at %.3f GHz\n", name, loads_per_sm / launch_clocks,
           launch_clocks / loads_per_sm, ms, ticks_per_ns);
    cudaFree(out); cudaFree(cyc); cudaFree(ns);
}

int main() {
    double* buf;
    cudaMalloc(&buf, (size_t)ROW * ROWS * sizeof(double) + 4096);
    cudaMemset(buf, 0, (size_t)ROW * ROWS * sizeof(double) + 4096);
    run<0>("warm-up", buf);
    run<0>("LDG.64 32 consecutive doubles, aligned (2 lines)", buf);
    run<1>("LDG.64 32 consecutive doubles, misaligned (3 lines)", buf);
    run<2>("LDG.64 two rows 20+12 lanes, misaligned (cfg3/cfg4)", buf);
    run<3>("LDG.64 seven rows of 5 lanes, misaligned (cfg5)", buf);
    run<4>("LDS.64 32 consecutive doubles (shared memory)", buf);
    return 0;
}
