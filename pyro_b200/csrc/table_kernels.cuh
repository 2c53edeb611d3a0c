// Table-mode and auxiliary kernels (sm_100a): the generic LUT sweep, the policy-evaluation sweep, the terminal
// cost, and the two post-processing kernels.  Included by pyrodp.cu (and, as test infrastructure, by the CPU
// emulation in tests/emu/).
#pragma once
#include "pyrodp_device.cuh"

#ifndef SWEEP_THREADS
#define SWEEP_THREADS 128
#endif

// ---- LUT mode: generic n in {2,3,4}, tables in HBM (dynamicprogramming.py:557-570) ---------------
// A group of G lanes owns one node and strides over its actions, so the x_next / G rows are read
// with contiguous, vectorisable accesses; the min/argmin over actions is a warp-shuffle reduction
// with lowest-index tie break (np.argmin).
template <int N>
__device__ __forceinline__ double rgi_linear(const DevProblem& P, const double* __restrict__ Jn, const double* x, bool& oob) {
    int c[N];
    double y[N];
    oob = false;
#pragma unroll
    for (int d = 0; d < N; ++d) oob = oob || (x[d] < P.lb[d]) || (x[d] > P.ub[d]);
    if (oob) return 0.0;  // fill_value (_rgi.py:476-477)
#pragma unroll
    for (int d = 0; d < N; ++d) {
        // scipy's find_interval_ascending: arithmetic guess, then the level table decides
        const double* __restrict__ lev = P.level[d];
        const int nlev = P.dims[d];
        int k = min(max((int)((x[d] - P.lb[d]) * P.inv_step[d]), 0), nlev - 2);
        double lo = __ldg(lev + k), hi = __ldg(lev + k + 1);
        while (x[d] < lo && k > 0) { --k; hi = lo; lo = __ldg(lev + k); }
        while (x[d] >= hi && k < nlev - 2) { ++k; lo = hi; hi = __ldg(lev + k + 1); }
        c[d] = k;
        y[d] = exact_div(x[d] - lo, hi - lo, __ldg(P.rinv[d] + k));   // correctly rounded quotient, 3 FP64 issues
    }
    if (N == 2) {
        const double* p = Jn + (long long)c[0] * P.dims[1] + c[1];
        const double v00 = __ldg(p), v01 = __ldg(p + 1), v10 = __ldg(p + P.dims[1]), v11 = __ldg(p + P.dims[1] + 1);
        double r = v00 * (1.0 - y[0]) * (1.0 - y[1]);
        r = r + v01 * (1.0 - y[0]) * y[1];
        r = r + v10 * y[0] * (1.0 - y[1]);
        r = r + v11 * y[0] * y[1];
        return r;
    }
    long long off = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) off += (long long)c[d] * P.stride[d];
    double value = 0.0;
#pragma unroll
    for (int corner = 0; corner < (1 << N); ++corner) {
        double w = 1.0;
        long long o = off;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            const int bit = (corner >> (N - 1 - d)) & 1;  // axis 0 slowest (itertools.product order)
            w = w * (bit ? y[d] : (1.0 - y[d]));
            o += bit ? P.stride[d] : 0;
        }
        value = value + __ldg(Jn + o) * w;
    }
    return value;
}

template <int N, int G>
__global__ void __launch_bounds__(SWEEP_THREADS)
sweep_lut_kernel(const __grid_constant__ DevProblem P, const double* __restrict__ Jn, double* __restrict__ Jo,
                 long long* __restrict__ pi, const double* __restrict__ xnext, const double* __restrict__ Gtab,
                 unsigned long long* __restrict__ partials, unsigned int* counter, double* __restrict__ stats) {
    // Persistent grid: one resident wave of blocks strides over the nodes, so the statistics epilogue (three
    // atomics and one ticket per BLOCK) stays negligible however many nodes there are — with one block per 128
    // nodes the ticket counter alone serialised a 16M-node policy-evaluation sweep (r01B).
    const int lane_in_group = threadIdx.x % G;
    const int A = P.A;
    const long long total = P.node_end - P.node_begin;
    const long long gstride = (long long)gridDim.x * blockDim.x / G;
    const long long iters = (total + gstride - 1) / gstride;   // the same trip count for every thread: shuffles inside
    Stats3 st = stats_identity();
    for (long long it = 0; it < iters; ++it) {
        const long long node = P.node_begin + it * gstride + ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
        const long long slot = node - P.slab_node_begin;  // row of the (slab-local) x_next / G tables
        const bool active = node < P.node_end;
        double best = __longlong_as_double(0x7ff0000000000000LL);
        int besta = 0x7fffffff;
        if (active) {
            const double* __restrict__ xrow = xnext + slot * (long long)A * N;
            const double* __restrict__ grow = Gtab + slot * (long long)A;
            for (int a = lane_in_group; a < A; a += G) {
                double x[N];
#pragma unroll
                for (int d = 0; d < N; ++d) x[d] = __ldcs(xrow + (long long)a * N + d);
                bool oob;
                const double Jx = rgi_linear<N>(P, Jn, x, oob);
                const double Qa = __ldcs(grow + a) + P.alpha * Jx;
                if (Qa < best) { best = Qa; besta = a; }
            }
        }
        lane_group_argmin(best, besta, G);
        if (active && lane_in_group == 0) {
            if (besta == 0x7fffffff) besta = 0;  // no Q below +inf: np.argmin of a constant row is 0
            Jo[node] = best;
            pi[node] = besta;
            const double d = best - Jn[node];
            Stats3 mine;
            mine.jmax = best; mine.dmax = d; mine.dmin = d;
            stats_merge(st, mine);
        }
    }
    block_stats_finish(st, partials, counter, stats);
}

// ---- policy evaluation: LUT mode with ONE table column per node (dynamicprogramming.py:743-752) --------
// J = G + alpha * RGI(J_next)(x_next_table): no min, 8n + 32 bytes of HBM traffic per node and a few dozen
// instructions — the streaming member of the family.  A persistent grid strides over the nodes; each thread
// takes U nodes per trip (a grid-strided tile, so every access of a warp is contiguous) and issues all their
// table loads before the first interpolation.  Measured (r01E, 4001^2 nodes): U = 1 at 32 registers and full
// occupancy wins — 0.162 ms, 4.74 TB/s of algorithmic traffic = 72 % of the measured HBM copy peak — over
// U = 2 (0.181), 4 (0.213), 8 (0.293).
#ifndef POLICY_U2
#define POLICY_U2 1   // nodes per thread and trip, n = 2: occupancy beats per-thread memory parallelism (profiles/r01E_policy_variants.txt)
#endif
#ifndef POLICY_U4
#define POLICY_U4 1   // the same, n = 3 and 4
#endif
template <int N, int U>
__global__ void __launch_bounds__(256)
sweep_policy_kernel(const __grid_constant__ DevProblem P, const double* __restrict__ Jn, double* __restrict__ Jo,
                    long long* __restrict__ pi, const double* __restrict__ xnext, const double* __restrict__ Gtab,
                    unsigned long long* __restrict__ partials, unsigned int* counter, double* __restrict__ stats) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tstride = (long long)gridDim.x * blockDim.x;
    Stats3 st = stats_identity();
    for (long long first = P.node_begin; first < P.node_end; first += tstride * U) {
        double x[U][N], g[U], jn[U];
        bool act[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long node = first + u * tstride + tid;
            act[u] = node < P.node_end;
            if (act[u]) {
                const long long slot = node - P.slab_node_begin;   // row of the (slab-local) tables
                const double2* __restrict__ xr = (const double2*)(xnext + slot * N);   // N even: 16-byte aligned rows
                if (N % 2 == 0) {
#pragma unroll
                    for (int d = 0; d < N; d += 2) {
                        const double2 v = __ldcs(xr + d / 2);
                        x[u][d] = v.x; x[u][d + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int d = 0; d < N; ++d) x[u][d] = __ldcs(xnext + slot * N + d);
                }
                g[u] = __ldcs(Gtab + slot);
                jn[u] = __ldg(Jn + node);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (act[u]) {
                const long long node = first + u * tstride + tid;
                bool oob;
                const double Jx = rgi_linear<N>(P, Jn, x[u], oob);
                const double Q = g[u] + P.alpha * Jx;
                Jo[node] = Q;
                pi[node] = 0;
                const double d = Q - jn[u];
                Stats3 mine;
                mine.jmax = Q; mine.dmax = d; mine.dmin = d;
                stats_merge(st, mine);
            }
        }
    }
    block_stats_finish(st, partials, counter, stats);
}

// ---- terminal cost (dynamicprogramming.py:159-171) -------------------------------------------------
template <int N>
__global__ void terminal_cost_kernel(const __grid_constant__ DevProblem P, double* __restrict__ J, long long* __restrict__ pi,
                                     long long first, long long last) {
    // J and pi are virtual bases indexed by global node id; [first,last) = allocated planes (slab + halo)
    const long long node = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= last) return;
    double dx[N];
    long long r = node;
#pragma unroll
    for (int d = N - 1; d >= 0; --d) {
        const int i = (int)(r % P.dims[d]);
        r /= P.dims[d];
        dx[d] = __ldg(P.level[d] + i) - P.xbar[d];
    }
    double h = 0.0;
    if (P.cost_id == PDP_COST_QUADRATIC) {
        h = quad_form<N>(P.S, dx);
        if (P.ontarget_check && norm2<N>(dx) < P.EPS) h = 0.0;
    }
    if (P.cost_id == PDP_COST_REACH) h = (norm2<N>(dx) < P.EPS) ? 0.0 : P.INF;   // Reachability.h / norm_test (costfunction.py:442-465)
    J[node] = h;
    if (node >= P.node_begin && node < P.node_end) pi[node] = 0;
}

// ---- after the sweep: pi -> u_k table (discretizer.py:616-633), infeasible-set cleaning (:322-334) ----
__global__ void input_from_policy_kernel(const long long* __restrict__ pi, const double* __restrict__ u_flat, int m, int k,
                                         double* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ldg(u_flat + pi[i] * m + k);
}
__global__ void clean_infeasible_kernel(double* __restrict__ J, long long* __restrict__ pi, double thr, double INF,
                                        long long def_action, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && J[i] > thr) {
        J[i] = INF;
        pi[i] = def_action;
    }
}

// ---- the step BEFORE the sweep: dense look-up tables on the device ------------------------------------------------
// GridDynamicSystem.compute_xnext_table (discretizer.py:342-376): x_next_table[s,a,:] = f(x_s,u_a)*dt + x_s and
// x_next_isok[s,a] = isavalidstate(x_next); DynamicProgrammingWithLookUpTable.compute_cost_lookuptable
// (dynamicprogramming.py:517-553): G[s,a] = g(x_s,u_a)*dt where the action and the arrival state are allowed, else INF.
// For the four fused systems, with the arithmetic of the sweep kernels (so the tables carry the reference's bits):
// one thread per (node, action) pair of the node range [node0, node0 + count).  Any output pointer may be null.
template <int N>
__global__ void build_tables_kernel(const __grid_constant__ DevProblem P, long long node0, long long count,
                                    double* __restrict__ x_next, unsigned char* __restrict__ x_ok, double* __restrict__ G) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int A = P.A;
    if (idx >= count * A) return;
    const long long s = idx / A;
    const int a = (int)(idx - s * A);
    int ix[N];
    double x[N], dx[N];
    long long r = node0 + s;
#pragma unroll
    for (int d = N - 1; d >= 0; --d) {
        ix[d] = (int)(r % P.dims[d]);
        r /= P.dims[d];
        x[d] = __ldg(P.level[d] + ix[d]);
        dx[d] = x[d] - P.xbar[d];
    }
    double f[N];
    if (N == 2) {        // SinglePendulum: ddq = inv(H) (B u - C dq - g - d), C = 0 (mechanical.py:222-234, pendulum.py:80-150)
        const double t = ((__ldg(P.bu + a) - 0.0 * x[1]) - __ldg(P.tab[0] + ix[0])) - P.par[1] * x[1];
        f[0] = x[1];
        f[1] = P.par[0] * t;
    } else {
        const double dq0 = x[2], dq1 = x[3];
        const double* __restrict__ Hi = P.tab[0] + 4 * ix[1];
        double cd0, cd1, g0, g1, d0, d1;
        if (P.system_id == PDP_SYS_TWOLINK) {
            const double h = __ldg(P.tab[1] + ix[1]);
            const double C00 = (-h) * dq1, C10 = h * dq0, C01 = (-h) * (dq0 + dq1);
            cd0 = mv2(C00, C01, dq0, dq1);
            cd1 = mv2(C10, 0.0, dq0, dq1);
            const double* __restrict__ Gq = P.tab[2] + 2 * ((long long)ix[0] * P.dims[1] + ix[1]);
            g0 = __ldg(Gq); g1 = __ldg(Gq + 1);
            d0 = mv2(P.par[0], 0.0, dq0, dq1);
            d1 = mv2(0.0, P.par[1], dq0, dq1);
        } else {
            const double C01 = __ldg(P.tab[1] + ix[1]) * dq1;
            cd0 = mv2(0.0, C01, dq0, dq1);
            cd1 = mv2(0.0, 0.0, dq0, dq1);
            g0 = 0.0; g1 = __ldg(P.tab[2] + ix[1]);
            d0 = 0.0; d1 = 0.0;
        }
        const double r0 = ((__ldg(P.bu + 2 * a) - cd0) - g0) - d0;
        const double r1 = ((__ldg(P.bu + 2 * a + 1) - cd1) - g1) - d1;
        f[0] = dq0; f[1] = dq1;
        f[N - 2] = mv2(__ldg(Hi), __ldg(Hi + 1), r0, r1);
        f[N - 1] = mv2(__ldg(Hi + 2), __ldg(Hi + 3), r0, r1);
    }
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const double xn = f[d] * P.dt + x[d];          // two roundings (discretizer.py:363)
        if (x_next) x_next[idx * N + d] = xn;
        if (xn < P.lb[d] || xn > P.ub[d]) ok = false;  // strict box test (system.py:198-205)
    }
    if (x_ok) x_ok[idx] = ok ? 1 : 0;
    if (G) {
        double g = 1.0;
        if (P.cost_id == PDP_COST_QUADRATIC) g = quad_form<N>(P.Q, dx) + __ldg(P.gu + a);
        if (P.cost_id == PDP_COST_REACH) g = 0.0;
        if (P.ontarget_check && norm2<N>(dx) < P.EPS) g = 0.0;
        G[idx] = (ok && P.act_ok[a]) ? g * P.dt : P.INF;
    }
}

