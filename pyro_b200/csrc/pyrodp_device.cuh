// Device-side pieces of the Bellman sweep, shared by all kernels in pyrodp.cu.
//
// Arithmetic contract: every value that reaches J or the valid/out-of-bounds
// classification is produced by the same IEEE-754 double operations, in the same association,
// as the reference's NumPy/SciPy/OpenBLAS path.  The file is compiled with -fmad=false so the
// compiler never contracts a*b+c; fma() appears only where the reference's BLAS kernel fuses
// (np.dot of 2-vectors / 2x2 matrices, measured in oracle/probe notes, DESIGN.md "Arithmetic").
#pragma once
#include <stdint.h>

#define PDP_MAXN 4

struct DevProblem {
    int n, m, dof, A;
    int system_id, cost_id, ontarget_check, alpha_is_one;
    int all_act_ok, pad0;           // every action passes isavalidinput (the usual case)
    int dims[PDP_MAXN];
    long long stride[PDP_MAXN];     // node-id stride of each axis (C order)
    const double* level[PDP_MAXN];  // device copies of np.linspace levels
    const double* rinv[PDP_MAXN];   // correctly rounded 1/(level[i+1]-level[i]), dims-1 entries
    double lb[PDP_MAXN], ub[PDP_MAXN], inv_step[PDP_MAXN];
    double dt, alpha, INF, EPS;
    double Q[16], S[16], xbar[PDP_MAXN], par[8];
    const double* tab[4];
    const double* bu;          // [A*dof]
    const double* gu;          // [A]
    const unsigned char* act_ok;  // [A]
    const double* u_flat;      // [A*m] input_from_action_id
    long long node_begin, node_end, N;
    long long plane_begin;          // first (i0,i1) pair of this launch = first plane * dims[1] (4-D kernels)
    long long slab_node_begin;      // first node of the handle's slab (origin of slab-local tables, LUT mode)
    // 4-D fused kernels: blocks per (i0,i1) plane; structure of the action table for the range kernel (mech2_plan.h)
    int chunks, A0, A1, tile_rows;   // tile_rows: range kernel, rows of the (i2,i3) plane per block (1 = 128 consecutive nodes)
    double uv_first, uv_inv_step;
};

// ---- np.dot conventions of the reference's BLAS (OpenBLAS 0.3.30 SkylakeX kernels) -----------
// 2-vector / 2x2: r = fma(a0, b0, a1*b1) for matvec rows; ddot is a forward fma chain.
__device__ __forceinline__ double mv2(double m0, double m1, double v0, double v1) { return fma(m0, v0, m1 * v1); }
__device__ __forceinline__ double dot2(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }
__device__ __forceinline__ double dot4(const double* a, const double* b) {
    return fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])));
}
// 4x4 matvec row: products rounded separately, (p0+p2)+(p1+p3)
__device__ __forceinline__ double mv4(const double* row, const double* v) {
    return (row[0] * v[0] + row[2] * v[2]) + (row[1] * v[1] + row[3] * v[3]);
}

// Quadratic form dx^T W dx and ||dx|| exactly as costfunction.py:146-149,189-197 evaluates them:
// np.dot(dx.T, np.dot(W, dx)), np.linalg.norm(dx) = sqrt(dot(dx,dx)).
template <int N>
__device__ __forceinline__ double quad_form(const double* W, const double* dx) {
    if (N == 2) {
        double w0 = mv2(W[0], W[1], dx[0], dx[1]);
        double w1 = mv2(W[2], W[3], dx[0], dx[1]);
        return dot2(dx[0], dx[1], w0, w1);
    } else {
        double w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = mv4(W + 4 * i, dx);
        return dot4(dx, w);
    }
}
template <int N>
__device__ __forceinline__ double norm2(const double* dx) {
    if (N == 2) return sqrt(dot2(dx[0], dx[1], dx[0], dx[1]));
    return sqrt(dot4(dx, dx));
}

// Interval search of scipy's find_indices (scipy/interpolate/_rgi_cython: find_interval_ascending):
// largest i with level[i] <= x, clipped to [0, n-2]; x == level[n-1] lands in the last cell.
// Caller guarantees lb <= x <= ub.  The arithmetic guess is only a starting point; the level
// table decides, so the result is identical to the binary search.
__device__ __forceinline__ int find_cell(const double* __restrict__ lev, int nlev, double x, double lb, double inv_step) {
    int i = (int)((x - lb) * inv_step);
    i = min(max(i, 0), nlev - 2);
    while (i > 0 && x < lev[i]) --i;
    while (i < nlev - 2 && x >= lev[i + 1]) ++i;
    return i;
}

// Correctly rounded (x - lo) / (hi - lo) as find_indices computes the normalised distance.
// r is the correctly rounded reciprocal of den = hi - lo (host table).  q0 = a*r is within 1 ulp;
// one exact residual + one fma gives the correctly rounded quotient (Markstein 1990, Thm 8.3/8.4;
// checked against IEEE division in tests/test_kernels_gpu.py::test_exact_div).  3 FP64 issues
// instead of the ~25 of the division subroutine.
__device__ __forceinline__ double exact_div(double a, double den, double r) {
    double q0 = a * r;
    double e = fma(-den, q0, a);
    return fma(e, r, q0);
}

// ---- min/argmin with lowest-index tie break across a group of lanes (np.argmin semantics) ---
__device__ __forceinline__ void lane_group_argmin(double& q, int& a, int group) {
    for (int off = group >> 1; off > 0; off >>= 1) {
        double oq = __shfl_down_sync(0xffffffffu, q, off, group);
        int oa = __shfl_down_sync(0xffffffffu, a, off, group);
        if (oq < q || (oq == q && oa < a)) { q = oq; a = oa; }
    }
}

// ---- fused convergence statistics (finalize_backward_step, dynamicprogramming.py:247-250) ----
struct Stats3 {
    double jmax, dmax, dmin;
};
__device__ __forceinline__ Stats3 stats_identity() {
    Stats3 s;
    s.jmax = -__longlong_as_double(0x7ff0000000000000LL);
    s.dmax = s.jmax;
    s.dmin = -s.jmax;
    return s;
}
__device__ __forceinline__ void stats_merge(Stats3& a, const Stats3& b) {
    a.jmax = fmax(a.jmax, b.jmax);
    a.dmax = fmax(a.dmax, b.dmax);
    a.dmin = fmin(a.dmin, b.dmin);
}
__device__ __forceinline__ Stats3 warp_stats(Stats3 s) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Stats3 o;
        o.jmax = __shfl_xor_sync(0xffffffffu, s.jmax, off);
        o.dmax = __shfl_xor_sync(0xffffffffu, s.dmax, off);
        o.dmin = __shfl_xor_sync(0xffffffffu, s.dmin, off);
        stats_merge(s, o);
    }
    return s;
}

// Order-preserving map double -> uint64 so that max/min run as integer atomics (RED.MAX.U64).
__device__ __forceinline__ unsigned long long stats_key(double x) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
__device__ __forceinline__ double stats_unkey(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)u);
}

#define STATS_SLOTS 64
// Block reduce -> three RED.MAX.U64 into one of STATS_SLOTS slot triples {key(jmax), key(dmax),
// key(-dmin)} (spread by block id to keep the L2 atomic unit off a single address) -> ticket;
// the last block folds the slots, decodes into out[3] = {max J, max dJ, min dJ}
// (finalize_backward_step, dynamicprogramming.py:247-250) and re-arms slots and ticket.
// `slots` (3*STATS_SLOTS u64) and `counter` must be zero at launch.
__device__ __forceinline__ void block_stats_finish(Stats3 s, unsigned long long* __restrict__ slots, unsigned int* counter,
                                                   double* __restrict__ out) {
    __shared__ Stats3 sh[32];
    __shared__ bool is_last;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nthreads = blockDim.x * blockDim.y;
    const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
    const unsigned int bid = blockIdx.x + blockIdx.y * gridDim.x;
    const unsigned int nblocks = gridDim.x * gridDim.y;
    s = warp_stats(s);
    if (lane == 0) sh[warp] = s;
    __syncthreads();
    if (warp == 0) {
        Stats3 t = lane < nwarps ? sh[lane] : stats_identity();
        t = warp_stats(t);
        if (lane == 0) {
            unsigned long long* slot = slots + 3 * (bid % STATS_SLOTS);
            atomicMax(slot + 0, stats_key(t.jmax));
            atomicMax(slot + 1, stats_key(t.dmax));
            atomicMax(slot + 2, stats_key(-t.dmin));
            __threadfence();
            const unsigned int ticket = atomicAdd(counter, 1u);
            is_last = (ticket == nblocks - 1);
        }
    }
    __syncthreads();
    if (is_last && warp == 0) {
        __threadfence();
        unsigned long long k0 = 0, k1 = 0, k2 = 0;
        for (int b = lane; b < STATS_SLOTS; b += 32) {
            k0 = max(k0, __ldcg(slots + 3 * b + 0));
            k1 = max(k1, __ldcg(slots + 3 * b + 1));
            k2 = max(k2, __ldcg(slots + 3 * b + 2));
            slots[3 * b + 0] = 0; slots[3 * b + 1] = 0; slots[3 * b + 2] = 0;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            k0 = max(k0, __shfl_xor_sync(0xffffffffu, k0, off));
            k1 = max(k1, __shfl_xor_sync(0xffffffffu, k1, off));
            k2 = max(k2, __shfl_xor_sync(0xffffffffu, k2, off));
        }
        if (lane == 0) {
            out[0] = stats_unkey(k0);
            out[1] = stats_unkey(k1);
            out[2] = -stats_unkey(k2);
            *counter = 0u;
        }
    }
}
