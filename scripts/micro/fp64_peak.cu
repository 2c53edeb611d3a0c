// FP64 pipe micro-benchmark (B200, sm_100a): what the sweep kernels can at best issue.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/fp64_peak scripts/micro/fp64_peak.cu
// Prints warp-level FP64 instructions per clock per SM (launch duration by CUDA events, SM clock measured in-kernel) for: independent DFMA chains, DADD chains, DMUL chains,
// DSETP, a DADD/DMUL/DFMA mix with the sweep's proportions, and the same mix with one integer/select
// instruction interleaved per FP64 instruction.  Time by CUDA events; clock from cudaDevAttrClockRate is
// not trusted: SM cycles are read with clock64() inside the kernel.
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 32768
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k(double* out, double seed, long long* cycles, long long* nanos) {
    double a[ILP], b = seed, c = seed * 0.5;
    int sel = threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = seed + i + threadIdx.x;
    unsigned long long g0 = gtime();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) a[i] = fma(a[i], b, c);
            if (MODE == 1) a[i] = a[i] + b;
            if (MODE == 2) a[i] = a[i] * b;
            if (MODE == 3) { if (a[i] < b) sel += i; a[i] = a[i] + c; }        // DSETP + DADD
            if (MODE == 4) {                                                     // mix: mul, add, fma
                if ((i % 3) == 0) a[i] = a[i] * b; else if ((i % 3) == 1) a[i] = a[i] + c; else a[i] = fma(a[i], b, c);
            }
            if (MODE == 5) {                                                     // mix + one ALU op per FP64 op
                if ((i % 3) == 0) a[i] = a[i] * b; else if ((i % 3) == 1) a[i] = a[i] + c; else a[i] = fma(a[i], b, c);
                sel = (sel ^ (sel >> 3)) + i;
            }
        }
    }
    long long t1 = clock64();
    unsigned long long g1 = gtime();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + sel;
    if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; nanos[blockIdx.x] = (long long)(g1 - g0); }
}

template <int MODE, int ILP>
static void run(const char* name, int blocks_per_sm, int threads, double fp64_per_iter) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * blocks_per_sm;
    double* out; long long* cyc; long long* ns;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(double));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    cudaMalloc(&ns, blocks * sizeof(long long));
    k<MODE, ILP><<<blocks, threads>>>(out, 1.0000001, cyc, ns);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, ILP><<<blocks, threads>>>(out, 1.0000001, cyc, ns);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks];
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    cudaMemcpy(h, ns, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg_ns = 0; for (int i = 0; i < blocks; ++i) avg_ns += h[i]; avg_ns /= blocks;
    const double warp_inst_per_sm = (double)blocks_per_sm * (threads / 32) * ITERS * fp64_per_iter;
    // ONE throughput column: all FP64 warp instructions of the launch over the launch's duration (CUDA events) in SM
    // clocks (clock64 ticks per globaltimer ns, measured inside the kernel).  The per-block clock64 span is printed last
    // for reference only: blocks need not be co-resident for the whole launch, so it is not a throughput.
    const double ticks_per_ns = avg / avg_ns;
    const double launch_clocks = (double)ms * 1e6 * ticks_per_ns;
    const double rate = warp_inst_per_sm / launch_clocks;
    printf("%-30s ILP=%d warps/SM=%3d : %.3f FP64 warp-inst/clk/SM (%.1f lanes/clk/SM) | launch %.3f ms = %.0f SM clocks at %.3f GHz | per-block span %.0f clocks\n",
           name, ILP, blocks_per_sm * threads / 32, rate, 32.0 * rate, ms, launch_clocks, ticks_per_ns, avg);
    cudaFree(out); cudaFree(cyc); cudaFree(ns); delete[] h;
}

int main() {
    run<0, 8>("warm-up", 4, 256, 8);
    run<0, 8>("warm-up", 4, 256, 8);
    run<0, 8>("DFMA independent chains", 4, 256, 8);
    run<0, 8>("DFMA independent chains", 1, 128, 8);
    run<0, 1>("DFMA single chain (latency)", 1, 32, 1);
    run<1, 1>("DADD single chain (latency)", 1, 32, 1);
    run<1, 8>("DADD independent chains", 4, 256, 8);
    run<2, 8>("DMUL independent chains", 4, 256, 8);
    run<3, 8>("DSETP+DADD", 4, 256, 16);
    run<4, 9>("mix mul/add/fma", 4, 256, 9);
    run<5, 9>("mix + 1 ALU per FP64", 4, 256, 9);
    run<0, 2>("DFMA ILP2, 8 warps/SMSP", 4, 256, 2);
    run<0, 1>("DFMA ILP1, 8 warps/SMSP", 4, 256, 1);
    run<0, 1>("DFMA ILP1, 16 warps/SMSP", 8, 256, 1);
    return 0;
}
