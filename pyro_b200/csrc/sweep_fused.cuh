// Fused on-the-fly Bellman sweep kernels (the north-star kernels), sm_100a.
//
// Replaces, per launch, one full sweep of
//   pyro/planning/dynamicprogramming.py:175-261 (initialize / compute / finalize_backward_step)
// with the dynamics evaluated on the fly as the base class does (:195-236) instead of the
// (N,A,n) x_next_table of the LUT variant (:557-570, discretizer.py:342-376).
//
// Mapping.  A group of G lanes owns one node; lanes of a group stride over the actions and the
// min/argmin over actions ends in a warp-shuffle reduction with lowest-index tie break
// (np.argmin, :236).  G = 1 for large grids (a thread scans all actions of its node, no
// reduction needed), G > 1 when the grid alone cannot fill 148 SMs.  Consecutive groups walk the
// last (contiguous) grid axis, so the J_next corner gathers of a warp fall into 2-3 cache lines
// and the J / pi stores are contiguous.
//
// Inner loop economy (ncu, profiles/): the sweep is bound by issue slots, the FP64 pipe and L1
// wavefronts, not by HBM.  Per action the loop therefore
//   * reads one packed record {B.u, du'R du} from shared memory (broadcast, one wavefront),
//   * keeps the current interpolation cell of every action-dependent axis in registers
//     {index, lo, hi, 1/(hi-lo)} and only walks it when x_next leaves the cell — the level table
//     decides, so the result equals scipy's binary search (find_interval_ascending),
//   * forms the normalised distance with the 3-instruction correctly rounded quotient exact_div,
//   * gathers the 2^n corners through the read-only path with immediate offsets.
#pragma once
#include "pyrodp_device.cuh"

#define SWEEP_THREADS 128

// ---- register-cached interpolation cell of one axis --------------------------------------------
struct CellCache {
    int c;
    double lo, hi, rinv;
};

// (re)position the cell so that lev[c] <= x < lev[c+1] (last cell closed on the right).
// Caller guarantees lev[0] <= x <= lev[nlev-1].  Identical to find_cell / scipy's search: the
// level table decides.  Fast paths: same cell (2 compares), neighbouring cell (one table read);
// otherwise an arithmetic guess followed by a table-checked walk of at most a step or two.
__device__ __forceinline__ void cell_seek(CellCache& cc, const double* __restrict__ s_lev,
                                          const double* __restrict__ s_rinv, int nlev, double x, double lb,
                                          double inv_step) {
    if (x >= cc.hi || x < cc.lo) {
        int c = min(max(cc.c + (x >= cc.hi ? 1 : -1), 0), nlev - 2);
        double lo = s_lev[c], hi = s_lev[c + 1];
        if (x >= hi || x < lo) {
            c = min(max((int)((x - lb) * inv_step), 0), nlev - 2);
            lo = s_lev[c]; hi = s_lev[c + 1];
            while (x < lo && c > 0) { --c; hi = lo; lo = s_lev[c]; }
            while (x >= hi && c < nlev - 2) { ++c; lo = hi; hi = s_lev[c + 1]; }
        }
        cc.c = c; cc.lo = lo; cc.hi = hi;
        cc.rinv = s_rinv[c];
    }
}

__device__ __forceinline__ void cell_init(CellCache& cc, const double* __restrict__ s_lev,
                                          const double* __restrict__ s_rinv, int nlev, int guess) {
    cc.c = min(max(guess, 0), nlev - 2);
    cc.lo = s_lev[cc.c];
    cc.hi = s_lev[cc.c + 1];
    cc.rinv = s_rinv[cc.c];
}

// Make a pointer opaque to the optimiser, so that ptr[int_index] compiles to one IMAD.WIDE from a
// register-resident base instead of re-deriving base + 64-bit element offset per access (SASS
// showed four integer instructions per gather address without it).
template <typename T>
__device__ __forceinline__ const T* opaque(const T* p) {
    asm("" : "+l"(p));
    return p;
}

// Pin a loop constant in a register.  ptxas otherwise re-reads kernel parameters from the
// constant bank inside the action loop (five LDC per eval in the first SASS); OR-ing in a
// run-time zero (blockIdx.z of a grid whose z extent is 1) makes the value a computed one.
__device__ __forceinline__ double pinned(double x) {
    return __longlong_as_double(__double_as_longlong(x) | (long long)blockIdx.z);
}

// Shared-memory record reads of the pendulum action loop through a 32-bit shared-window address that lives in a
// register: ptxas otherwise re-derives "window base + dynamic offset + 16*a" (eight integer / uniform-datapath
// instructions) on every pass.  The CPU emulation (tests/emu) has no address spaces: there the address is the pointer.
#ifdef __CUDACC__
typedef unsigned smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr_of(const void* p) {
    smem_addr_t a = (smem_addr_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
template <int BYTES>
__device__ __forceinline__ double2 lds_double2(smem_addr_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(BYTES) : "memory");
    return v;
}
#else
typedef uintptr_t smem_addr_t;
static inline smem_addr_t smem_addr_of(const void* p) { return (smem_addr_t)p; }
template <int BYTES>
static inline double2 lds_double2(smem_addr_t a) { return *(const double2*)(a + BYTES); }
#endif

// ---- shared-memory staging of the small tables -------------------------------------------------
__device__ __forceinline__ void stage(double* dst, const double* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
}

// =================================================================================================
// n = 2, 1-dof mechanical system (pyro/dynamic/pendulum.py:16 SinglePendulum)
// grid: blockIdx.x = row i0 of the launch's plane range, blockIdx.y = chunk of SWEEP_THREADS/G nodes
// of that row.
//
// FP64-pipe economy (the binding resource).  All nodes of a block share q, so
// the action-only part of the dynamics is tabulated once per block in shared memory:
//   NODAMP (d1 == 0):  t[a] = fl(fl(Hinv * fl(B.u_a - g(q))) * dt)   ->  x_next[1] = t[a] + dq, ONE add
//   otherwise:         t[a] = fl(B.u_a - g(q))                       ->  ((t - d) * Hinv) * dt + dq
// and, because x_next[1] moves by a fraction of a velocity cell per action, the interpolation cell
// {lo, hi, hi-lo, 1/(hi-lo)} and the four action-independent corner products
//   v00*(1-y0), v01*(1-y0), v10*y0, v11*y0      (first factor pair of evaluate_linear_2d's terms)
// stay in registers until x_next[1] leaves the cell: the common action costs 19 FP64 issues (17.5 in the
// MONO loop), one LDS and no global load; leaving the cell re-runs the table-checked search and four
// gathers.  What binds the kernel is the FP64 pipe (68 % of its peak rate, ncu r01K / r02m) — not the issue port, as
// round 1's count model had it (2*FP64 + other instructions = 57.5 "issue cycles" against 57.7 measured was a
// coincidence: see the A/B under MONO below).
// =================================================================================================
// MONO = the host verified that x_next[1] cannot decrease along the action list (B.u ascending,
// inv(H) > 0, dt > 0, every action allowed — the linspace input grid of a pendulum): then lo <= x
// holds by induction, one compare (x < hi) on the later action of a pair covers both, the cell only
// ever moves up, and a lane whose x_next passed the upper bound is finished.  MONO = 1: pairs of actions under one vote,
// the cell change inlined at both actions (round 1, the shipped loop).  MONO = 2 (PYRODP_PEND_LOOP=2): the loop nest
// further down — 11 % fewer non-FP64 instructions (ncu r02p: 16.5 instead of 18.7 per warp-eval, issue slots 67 -> 63 %)
// and the SAME sweep time (0.311 ms at cfg 2 for both, one process, profiles/r02p_pendulum_loops_ab.jsonl): the kernel
// does not sit on the issue port, as round 1's count model (above) had it, but on the FP64 pipe, which the four
// sub-partitions of an SM share (sm__pipe_shared_cycles_active = sm__pipe_fp64_cycles_active = 68 %; stall reasons:
// fixed-latency wait 3.2, math-pipe throttle 1.7 warps per issued instruction).  Kept as a measured alternative.
#ifndef PEND_MIN_BLOCKS
#define PEND_MIN_BLOCKS 8
#endif
template <int G, bool ALPHA1, bool NODAMP, int MONO>
__global__ void __launch_bounds__(SWEEP_THREADS, (MONO == 2) ? PEND_MIN_BLOCKS : 0)
sweep_pendulum_kernel(const __grid_constant__ DevProblem P, const double* __restrict__ Jn, double* __restrict__ Jo,
                      long long* __restrict__ pi, unsigned long long* __restrict__ partials, unsigned int* counter,
                      double* __restrict__ stats) {
    extern __shared__ __align__(16) double smem[];
    const int N0 = P.dims[0], N1 = P.dims[1], A = P.A;
    // [N1] MONO < 2: {lev[k], 1/(lev[k+1]-lev[k])}; MONO == 2: {1/(lev[k]-lev[k-1]), lev[k]}, so that entry k+1 holds
    // {1/step, upper level} of cell k.  Either way one LDS.128 per cell change.
    double2* s_cell = (double2*)smem;
    double2* s_act = s_cell + N1;       // [A_pad (+ 2G, MONO == 2)] {t[a] (NaN when isavalidinput fails), du'R du}; the
                                        // padding repeats the last action: an equal Q never beats an earlier index
    const int A_pad = ((A + 2 * G - 1) / (2 * G)) * (2 * G);
    const int A_stage = (MONO == 2) ? A_pad + 2 * G : A_pad;   // the loop nest may run one pair past A_pad

    const int i0 = (int)(P.plane_begin + (long long)blockIdx.x);  // row = axis-0 plane
    const double q = __ldg(P.level[0] + i0);
    const double grav = __ldg(P.tab[0] + i0);    // g(q) (pendulum.py:126-137)
    for (int i = threadIdx.x; i < N1; i += blockDim.x)
        s_cell[i] = (MONO == 2) ? make_double2(i > 0 ? __ldg(P.rinv[1] + i - 1) : 0.0, __ldg(P.level[1] + i))
                                : make_double2(__ldg(P.level[1] + i), __ldg(P.rinv[1] + i));
    for (int i = threadIdx.x; i < A_stage; i += blockDim.x) {
        const int a = min(i, A - 1);
        // ddq = inv(H) . (B u - C dq - g - d), C = 0 (mechanical.py:222-234)
        double t = __ldg(P.bu + a) - grav;
        if (NODAMP) t = (P.par[0] * t) * P.dt;
        s_act[i] = make_double2(t, __ldg(P.gu + a));
    }
    __syncthreads();

    const int i1 = (int)(((unsigned)blockIdx.y * SWEEP_THREADS + threadIdx.x) / G);
    const int g = threadIdx.x % G;
    const bool active = i1 < N1;
    const long long node = (long long)i0 * N1 + min(i1, N1 - 1);
    const double PINF = __longlong_as_double(0x7ff0000000000000LL);
    const double dt = pinned(P.dt);
    const double dq_node = (MONO == 2) ? s_cell[min(i1, N1 - 1)].y : s_cell[min(i1, N1 - 1)].x;

    // position row of x_next: f[0]*dt + x[0] = dq*dt + q (two roundings, discretizer.py:363)
    const double xn0 = dq_node * dt + q;
    // live = this lane evaluates actions.  Dead lanes (past the end of the row, or every action
    // leaves the box through the position row) still run the loop — the warp votes in it — on an
    // all-covering dummy cell with zero corner products, and their result is discarded.
    const bool live = active && !(xn0 < P.lb[0] || xn0 > P.ub[0]);
    const int c0 = live ? find_cell(P.level[0], N0, xn0, P.lb[0], P.inv_step[0]) : 0;
    const double lo0 = __ldg(P.level[0] + c0), hi0 = __ldg(P.level[0] + c0 + 1);
    const double y0 = (xn0 - lo0) / (hi0 - lo0);
    const double omy0 = 1.0 - y0;
    const double* __restrict__ row0 = opaque(Jn + (long long)c0 * N1);
    const double* __restrict__ row1 = opaque(row0 + N1);

    const double dq = live ? dq_node : 0.0;
    const double Hinv = pinned(P.par[0]);
    const double damp = P.par[1] * dq;   // d(q,dq) (pendulum.py:141-150)

    // state-only stage cost (costfunction.py:186-197)
    const double dx[2] = {q - P.xbar[0], dq_node - P.xbar[1]};
    double gx = 1.0;
    if (P.cost_id == PDP_COST_QUADRATIC) gx = quad_form<2>(P.Q, dx);
    if (P.cost_id == PDP_COST_REACH) gx = 0.0;   // Reachability.g is 0 on every node inside the box (costfunction.py:468-481)
    const bool ontarget = P.ontarget_check && (norm2<2>(dx) < P.EPS);
    // g = 0 inside the target zone (costfunction.py:193-197): 0*dt == (gx+gu)*0 == +0
    const double dt_cost = ontarget ? 0.0 : dt;
    const double INF = pinned(P.INF), alpha = P.alpha;

    // cached cell of axis 1 and its corner products; live lanes start with an empty cell
    // (every x fails lo <= x < hi), dead lanes with (-inf, +inf)
    int c = -1;
    double lo = live ? PINF : -PINF, hi = -lo, den = 1.0, rinv = 1.0;
    double p00 = 0.0, p01 = 0.0, p10 = 0.0, p11 = 0.0;
    double best = PINF;
    int besta = 0x7fffffff;
    if constexpr (MONO == 1) {
        // Two actions per iteration under one compare and one warp vote.  evalq = evaluate_linear_2d in
        // its value-first association on the cached products, Q = g*dt + alpha*J
        // (dynamicprogramming.py:223): 15 FP64 issues.
        double gxv = gx;
        auto evalq = [&](double x, double gu) {
            const double y1 = exact_div(x - lo, den, rinv);
            const double omy1 = 1.0 - y1;
            double Jx = p00 * omy1;
            Jx = Jx + p01 * y1;
            Jx = Jx + p10 * omy1;
            Jx = Jx + p11 * y1;
            return (gxv + gu) * dt_cost + (ALPHA1 ? Jx : alpha * Jx);
        };
        // x reached the top of the cached cell (or there is none yet): isavalidstate (system.py:198-205),
        // then walk the level table upwards — the interval scipy's search returns — and gather the corners
        auto slow = [&](double x, double gu) {
            bool oob = false;
            if (!(x < hi)) {
                const double lb1 = P.lb[1], ub1 = P.ub[1];
                if (!(x <= ub1)) {
                    // above the box, and so is every later action: park the lane on an all-covering cell
                    // with an infinite state cost, so that its Q (inf, or NaN when dt_cost is 0) never wins again
                    oob = true;
                    hi = PINF;
                    gxv = PINF;
                } else if (x < lb1) {
                    oob = true;    // still below the box
                } else {
                    int k;
                    double l, h;
                    if (c < 0) {   // first cell of this node: arithmetic guess, the table decides
                        k = min(max((int)((x - lb1) * P.inv_step[1]), 0), N1 - 2);
                        l = s_cell[k].x; h = s_cell[k + 1].x;
                        while (x < l && k > 0) { --k; h = l; l = s_cell[k].x; }
                    } else {
                        k = min(c + 1, N1 - 2);
                        l = s_cell[k].x; h = s_cell[k + 1].x;
                    }
                    while (x >= h && k < N1 - 2) { ++k; l = h; h = s_cell[k + 1].x; }
                    c = k; lo = l; hi = h; den = h - l; rinv = s_cell[k].y;
                    const double* __restrict__ r0 = row0 + k;
                    const double* __restrict__ r1 = row1 + k;
                    p00 = __ldg(r0) * omy0; p01 = __ldg(r0 + 1) * omy0;
                    p10 = __ldg(r1) * y0;   p11 = __ldg(r1 + 1) * y0;
                }
            }
            const double Q = evalq(x, gu);
            return oob ? INF : Q;
        };
        hi = live ? -PINF : PINF;   // live lanes: no cell yet, the first action locates it
#pragma unroll 2
        for (int a = g; a < A_pad; a += 2 * G) {
            const int b = a + G;
            const double2 actA = s_act[a], actB = s_act[b];
            const double xA = NODAMP ? (actA.x + dq) : ((Hinv * (actA.x - damp)) * dt + dq);
            const double xB = NODAMP ? (actB.x + dq) : ((Hinv * (actB.x - damp)) * dt + dq);
            double QA, QB;
            if (__any_sync(0xffffffffu, !(xB < hi))) {   // xA <= xB: one compare covers the pair
                QA = slow(xA, actA.y);
                QB = slow(xB, actB.y);
            } else {
                QA = evalq(xA, actA.y);
                QB = evalq(xB, actB.y);
            }
            if (QA < best) { best = QA; besta = a; }   // strict <, in action order: first index wins (np.argmin)
            if (QB < best) { best = QB; besta = b; }
        }
        if (besta >= A && besta != 0x7fffffff) besta = A - 1;   // a padded copy of the last action
    } else if constexpr (MONO == 2) {
        // Loop nest.  The inner loop runs the pairs of actions the cached cell still covers for EVERY lane of the warp
        // (one compare on the later action of a pair and one vote per pair, the cell state loop-invariant, two pairs per
        // pass): 35 FP64 + 11.5 other instructions per pair on the common path (static count; MONO = 1: 35 + 18).  The
        // outer loop handles the pair at which some lane leaves its cell.  Measured (ncu r02p): 19.4 FP64 + 16.5 other
        // instructions per warp-eval against 19.5 + 18.7 — and no change of the sweep time (note above the kernel).
        double gxv = gx;
        auto evalq = [&](double x, double gu) {
            const double y1 = exact_div(x - lo, den, rinv);
            const double omy1 = 1.0 - y1;
            double Jx = p00 * omy1;
            Jx = Jx + p01 * y1;
            Jx = Jx + p10 * omy1;
            Jx = Jx + p11 * y1;
            return (gxv + gu) * dt_cost + (ALPHA1 ? Jx : alpha * Jx);
        };
        // x reached the top of the cached cell (or there is none yet).  Common case (nine cell changes in ten at
        // cfg 2, and the lanes of a row change together): x lies in the NEXT cell — one LDS.128 brings its upper
        // level and 1/step, its lower corner products are the old upper ones (the same two factors, so the same
        // bits), two gathers bring the new upper pair.  Anything else: isavalidstate (system.py:198-205), then
        // walk the level table upwards — the interval scipy's search returns — and gather all four corners.
        const smem_addr_t cell0 = smem_addr_of(s_cell);
        auto advance = [&](double x) -> bool {
            if ((unsigned)c < (unsigned)(N1 - 2)) {
                const double2 nx = lds_double2<32>(cell0 + 16u * (unsigned)c);   // s_cell[c + 2] = {1/step, upper level} of cell c+1
                if (x < nx.y) {
                    ++c; lo = hi; hi = nx.y; rinv = nx.x; den = hi - lo;
                    p00 = p01; p10 = p11;
                    p01 = __ldg(row0 + c + 1) * omy0;
                    p11 = __ldg(row1 + c + 1) * y0;
                    return false;
                }
            }
            const double lb1 = P.lb[1], ub1 = P.ub[1];
            if (!(x <= ub1)) {
                // above the box, and so is every later action: park the lane on an all-covering cell
                // with an infinite state cost, so that its Q (inf, or NaN when dt_cost is 0) never wins again
                hi = PINF;
                gxv = PINF;
                return true;
            }
            if (x < lb1) return true;   // still below the box
            int k;
            double l, h;
            if (c < 0) {   // first cell of this node: arithmetic guess, the table decides
                k = min(max((int)((x - lb1) * P.inv_step[1]), 0), N1 - 2);
                l = s_cell[k].y; h = s_cell[k + 1].y;
                while (x < l && k > 0) { --k; h = l; l = s_cell[k].y; }
            } else {
                k = min(c + 1, N1 - 2);
                l = s_cell[k].y; h = s_cell[k + 1].y;
            }
            while (x >= h && k < N1 - 2) { ++k; l = h; h = s_cell[k + 1].y; }
            c = k; lo = l; hi = h; den = h - l; rinv = s_cell[k + 1].x;
            const double* __restrict__ r0 = row0 + k;
            const double* __restrict__ r1 = row1 + k;
            p00 = __ldg(r0) * omy0; p01 = __ldg(r0 + 1) * omy0;
            p10 = __ldg(r1) * y0;   p11 = __ldg(r1 + 1) * y0;
            return false;
        };
        auto slow = [&](double x, double gu) {
            bool oob = false;
            if (!(x < hi)) oob = advance(x);
            const double Q = evalq(x, gu);
            return oob ? INF : Q;
        };
        hi = live ? -PINF : PINF;   // live lanes: no cell yet, the first action locates it
        // The action records are walked by ADDRESS (one add per pair; the argmin remembers the address of its record).
        const smem_addr_t act0 = smem_addr_of(s_act);
        const smem_addr_t pend = act0 + 16u * (unsigned)A_pad;
        smem_addr_t pa = act0 + 16u * (unsigned)g, bestp = 0;
        while (pa < pend) {
            bool leave;
            // two pairs per pass, the end test after the second: a pass may run one pair into the padding (copies of
            // the last action: an equal Q never replaces an earlier index)
            do {
                {
                    const double2 actA = lds_double2<0>(pa), actB = lds_double2<16 * G>(pa);
                    const double xA = NODAMP ? (actA.x + dq) : ((Hinv * (actA.x - damp)) * dt + dq);
                    const double xB = NODAMP ? (actB.x + dq) : ((Hinv * (actB.x - damp)) * dt + dq);
                    leave = __any_sync(0xffffffffu, !(xB < hi));   // xA <= xB: one compare covers the pair
                    if (leave) break;
                    const double QA = evalq(xA, actA.y);
                    const double QB = evalq(xB, actB.y);
                    if (QA < best) { best = QA; bestp = pa; }   // strict <, in action order: first index wins (np.argmin)
                    if (QB < best) { best = QB; bestp = pa + 16 * G; }
                }
                {
                    const double2 actA = lds_double2<32 * G>(pa), actB = lds_double2<48 * G>(pa);
                    const double xA = NODAMP ? (actA.x + dq) : ((Hinv * (actA.x - damp)) * dt + dq);
                    const double xB = NODAMP ? (actB.x + dq) : ((Hinv * (actB.x - damp)) * dt + dq);
                    leave = __any_sync(0xffffffffu, !(xB < hi));
                    if (leave) { pa += 32 * G; break; }
                    const double QA = evalq(xA, actA.y);
                    const double QB = evalq(xB, actB.y);
                    if (QA < best) { best = QA; bestp = pa + 32 * G; }
                    if (QB < best) { best = QB; bestp = pa + 48 * G; }
                }
                pa += 64 * G;
            } while (pa < pend);
            if (!leave) break;
            {
                const double2 actA = lds_double2<0>(pa), actB = lds_double2<16 * G>(pa);
                const double xA = NODAMP ? (actA.x + dq) : ((Hinv * (actA.x - damp)) * dt + dq);
                const double xB = NODAMP ? (actB.x + dq) : ((Hinv * (actB.x - damp)) * dt + dq);
                const double QA = slow(xA, actA.y);
                const double QB = slow(xB, actB.y);
                if (QA < best) { best = QA; bestp = pa; }
                if (QB < best) { best = QB; bestp = pa + 16 * G; }
                pa += 32 * G;
            }
        }
        if (bestp) besta = (int)((bestp - act0) >> 4);
        if (besta >= A && besta != 0x7fffffff) besta = A - 1;   // a padded copy of the last action
    } else {
        const int A_up = (G > 1) ? ((A + G - 1) / G) * G : A;  // same trip count for every lane of the warp
    #pragma unroll 4
        for (int a0 = g; a0 < A_up; a0 += G) {
            const int a = (G > 1) ? min(a0, A - 1) : a0;
            const double2 act = s_act[a];
            const double xn1 = NODAMP ? (act.x + dq) : ((Hinv * (act.x - damp)) * dt + dq);
            const bool miss = !(xn1 >= lo && xn1 < hi);
            double Qa;
            if (__any_sync(0xffffffffu, miss)) {
                // some lane left its cached cell (rare, and lanes of a row leave together): box test of
                // isavalidstate (system.py:198-205; a NaN from a disallowed action fails it), then the
                // level table decides the cell exactly as scipy's search does
                bool oob = false;
                if (miss) {
                    const double lb1 = P.lb[1], ub1 = P.ub[1];
                    oob = !(xn1 >= lb1 && xn1 <= ub1);
                    if (!oob) {
                        int k = (c < 0) ? (int)((xn1 - lb1) * P.inv_step[1]) : c + (xn1 >= hi ? 1 : -1);
                        k = min(max(k, 0), N1 - 2);
                        double2 ck = s_cell[k];
                        double h = s_cell[k + 1].x;
                        if (xn1 >= h || xn1 < ck.x) {
                            k = min(max((int)((xn1 - lb1) * P.inv_step[1]), 0), N1 - 2);
                            double l = s_cell[k].x;
                            h = s_cell[k + 1].x;
                            while (xn1 < l && k > 0) { --k; h = l; l = s_cell[k].x; }
                            while (xn1 >= h && k < N1 - 2) { ++k; l = h; h = s_cell[k + 1].x; }
                            ck = s_cell[k];
                        }
                        c = k; lo = ck.x; hi = h; den = h - ck.x; rinv = ck.y;
                        const double* __restrict__ r0 = row0 + k;
                        const double* __restrict__ r1 = row1 + k;
                        p00 = __ldg(r0) * omy0; p01 = __ldg(r0 + 1) * omy0;
                        p10 = __ldg(r1) * y0;   p11 = __ldg(r1 + 1) * y0;
                    }
                }
                const double y1 = exact_div(xn1 - lo, den, rinv);
                const double omy1 = 1.0 - y1;
                double Jx = p00 * omy1;
                Jx = Jx + p01 * y1;
                Jx = Jx + p10 * omy1;
                Jx = Jx + p11 * y1;
                Qa = (gx + act.y) * dt_cost + (ALPHA1 ? Jx : alpha * Jx);
                if (oob) Qa = INF;
            } else {
                const double y1 = exact_div(xn1 - lo, den, rinv);
                const double omy1 = 1.0 - y1;
                // evaluate_linear_2d: value-first association
                double Jx = p00 * omy1;
                Jx = Jx + p01 * y1;
                Jx = Jx + p10 * omy1;
                Jx = Jx + p11 * y1;
                Qa = (gx + act.y) * dt_cost + (ALPHA1 ? Jx : alpha * Jx);
            }
            if ((G == 1 || a0 < A) && Qa < best) { best = Qa; besta = a; }
        }
    }
    if (!live) {  // every action leaves the box: Q = INF for all, argmin = 0
        best = (g == 0) ? P.INF : PINF;
        besta = (g == 0) ? 0 : 0x7fffffff;
    }
    if (G > 1) lane_group_argmin(best, besta, G);
    Stats3 st = stats_identity();
    if (active && g == 0) {
        if (besta == 0x7fffffff) besta = 0;  // no Q below +inf (cf.INF = inf): np.argmin of a constant row is 0
        Jo[node] = best;
        pi[node] = besta;
        const double d = best - Jn[node];
        st.jmax = best; st.dmax = d; st.dmin = d;
    }
    block_stats_finish(st, partials, counter, stats);
}

// =================================================================================================
// n = 4, 2-dof mechanical systems: 2-link arm form (pendulum.py:340 DoublePendulum,
// manipulator.py:795 TwoLinkManipulator) and cart-pole (cartpole.py:322)
// grid: 1-D, blockIdx.x = plane * chunks + chunk with the chunk of the (i2,i3) plane fastest, so that the blocks
// resident at any time share a few (i0,i1) planes of J in L2 (P.chunks = blocks per plane)
//
// Measured alternatives that did NOT pay (profiles/r01h_variants.txt): two actions per iteration with
// interleaved 16-corner blends (more registers, fewer resident warps, the L1 data pipe — ~65
// wavefronts per warp-eval for the 16 gathers — is the co-limiter, not FP64 latency), and folding the
// box test into the cell test (the out-of-box evals that dominate the 2-input systems got dearer).
// =================================================================================================
#ifndef MECH2_MIN_BLOCKS
#define MECH2_MIN_BLOCKS 4
#endif
template <int SYS, int G, bool ALPHA1>
__global__ void __launch_bounds__(SWEEP_THREADS, MECH2_MIN_BLOCKS)
sweep_mech2_kernel(const __grid_constant__ DevProblem P, const double* __restrict__ Jn, double* __restrict__ Jo,
                   long long* __restrict__ pi, unsigned long long* __restrict__ partials, unsigned int* counter,
                   double* __restrict__ stats) {
    extern __shared__ __align__(16) double smem[];
    const int N0 = P.dims[0], N1 = P.dims[1], N2 = P.dims[2], N3 = P.dims[3], A = P.A;
    const int N2p = (N2 + 1) & ~1, N3p = (N3 + 1) & ~1;
    double* s_lev2 = smem;              // [N2p]
    double* s_rinv2 = s_lev2 + N2p;     // [N2p]
    double* s_lev3 = s_rinv2 + N2p;     // [N3p]
    double* s_rinv3 = s_lev3 + N3p;     // [N3p]
    double2* s_act = (double2*)(s_rinv3 + N3p);  // [2A] {B.u[0], B.u[1]} (NaN when isavalidinput fails), {du'R du, -}
    stage(s_lev2, P.level[2], N2);
    stage(s_rinv2, P.rinv[2], N2 - 1);
    stage(s_lev3, P.level[3], N3);
    stage(s_rinv3, P.rinv[3], N3 - 1);
    for (int i = threadIdx.x; i < A; i += blockDim.x) {
        s_act[2 * i] = make_double2(__ldg(P.bu + 2 * i), __ldg(P.bu + 2 * i + 1));
        s_act[2 * i + 1] = make_double2(__ldg(P.gu + i), 0.0);
    }
    __syncthreads();

    // block-uniform part of the node index
    const int plane_sz = N2 * N3;                                  // checked on the host: < 2^31
    const unsigned chunks = (unsigned)P.chunks;
    const unsigned pl_local = blockIdx.x / chunks;
    const unsigned chunk = blockIdx.x - pl_local * chunks;
    const long long pl = P.plane_begin + (long long)pl_local;    // (i0,i1) pair, C order
    const int i0 = (int)(pl / N1);
    const int i1 = (int)(pl - (long long)i0 * N1);
    const int r = (int)((chunk * SWEEP_THREADS + threadIdx.x) / G);  // (i2,i3) within the plane
    const int g = threadIdx.x % G;
    const bool active = r < plane_sz;
    const long long node = pl * plane_sz + r;
    Stats3 st = stats_identity();
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int besta = 0x7fffffff;
    if (active) {
        const int i2 = r / N3;
        const int i3 = r - i2 * N3;
        const double q0 = __ldg(P.level[0] + i0), q1 = __ldg(P.level[1] + i1);
        const double dq0 = s_lev2[i2], dq1 = s_lev3[i3];
        const double dt = P.dt;

        // position rows of x_next are action independent: dq*dt + q
        const double xn0 = dq0 * dt + q0;
        const double xn1 = dq1 * dt + q1;
        const bool pos_ok = !(xn0 < P.lb[0] || xn0 > P.ub[0] || xn1 < P.lb[1] || xn1 > P.ub[1]);
        if (pos_ok) {
            const int c0 = find_cell(P.level[0], N0, xn0, P.lb[0], P.inv_step[0]);
            const int c1 = find_cell(P.level[1], N1, xn1, P.lb[1], P.inv_step[1]);
            const double lo0 = __ldg(P.level[0] + c0), hi0 = __ldg(P.level[0] + c0 + 1);
            const double lo1 = __ldg(P.level[1] + c1), hi1 = __ldg(P.level[1] + c1 + 1);
            const double y0 = (xn0 - lo0) / (hi0 - lo0);
            const double y1 = (xn1 - lo1) / (hi1 - lo1);
            // _evaluate_linear weight-first association: w = (((1*w0)*w1)*w2)*w3 (_rgi.py:543-546)
            const double w00 = (1.0 - y0) * (1.0 - y1), w01 = (1.0 - y0) * y1;
            const double w10 = y0 * (1.0 - y1), w11 = y0 * y1;
            const double* __restrict__ b00 = opaque(Jn + ((long long)c0 * N1 + c1) * plane_sz);
            const double* __restrict__ b01 = opaque(b00 + plane_sz);
            const double* __restrict__ b10 = opaque(b00 + (long long)N1 * plane_sz);
            const double* __restrict__ b11 = opaque(b10 + plane_sz);

            // ---- state-only dynamics terms: C(q,dq) dq, g(q), d(q,dq), inv(H(q)) ----
            const double* __restrict__ Hi = P.tab[0] + 4 * i1;
            const double H00 = __ldg(Hi), H01 = __ldg(Hi + 1), H10 = __ldg(Hi + 2), H11 = __ldg(Hi + 3);
            double cd0, cd1, g0, g1, d0, d1;
            if (SYS == PDP_SYS_TWOLINK) {
                const double h = __ldg(P.tab[1] + i1);
                const double C00 = (-h) * dq1, C10 = h * dq0, C01 = (-h) * (dq0 + dq1);
                cd0 = mv2(C00, C01, dq0, dq1);
                cd1 = mv2(C10, 0.0, dq0, dq1);
                const double* __restrict__ Gq = P.tab[2] + 2 * ((long long)i0 * N1 + i1);
                g0 = __ldg(Gq); g1 = __ldg(Gq + 1);
                d0 = mv2(P.par[0], 0.0, dq0, dq1);
                d1 = mv2(0.0, P.par[1], dq0, dq1);
            } else {  // CARTPOLE
                const double C01 = __ldg(P.tab[1] + i1) * dq1;
                cd0 = mv2(0.0, C01, dq0, dq1);
                cd1 = mv2(0.0, 0.0, dq0, dq1);
                g0 = 0.0; g1 = __ldg(P.tab[2] + i1);
                d0 = 0.0; d1 = 0.0;
            }

            // ---- state-only stage cost ----
            const double dx[4] = {q0 - P.xbar[0], q1 - P.xbar[1], dq0 - P.xbar[2], dq1 - P.xbar[3]};
            double gx = 1.0;
            if (P.cost_id == PDP_COST_QUADRATIC) gx = quad_form<4>(P.Q, dx);
    if (P.cost_id == PDP_COST_REACH) gx = 0.0;   // Reachability.g is 0 on every node inside the box (costfunction.py:468-481)
            const bool ontarget = P.ontarget_check && (norm2<4>(dx) < P.EPS);

            const double lb2 = P.lb[2], ub2 = P.ub[2], lb3 = P.lb[3], ub3 = P.ub[3];
            const double is2 = P.inv_step[2], is3 = P.inv_step[3];
            const double INF = P.INF, alpha = P.alpha;
            const double dt_cost = ontarget ? 0.0 : dt;
            CellCache k2, k3;
            cell_init(k2, s_lev2, s_rinv2, N2, i2);
            cell_init(k3, s_lev3, s_rinv3, N3, i3);
            for (int a = g; a < A; a += G) {
                const double2 bu = s_act[2 * a];
                // B u - C dq - g - d, left to right (mechanical.py:231).  For the cart-pole g[0], d[0]
                // and d[1] are literal zeros (cartpole.py:415-437) and x - 0.0 == x bit for bit.
                const double r0 = (SYS == PDP_SYS_CARTPOLE) ? (bu.x - cd0) : (((bu.x - cd0) - g0) - d0);
                const double r1 = (SYS == PDP_SYS_CARTPOLE) ? ((bu.y - cd1) - g1) : (((bu.y - cd1) - g1) - d1);
                const double ddq0 = mv2(H00, H01, r0, r1);
                const double ddq1 = mv2(H10, H11, r0, r1);
                const double xn2 = ddq0 * dt + dq0;
                const double xn3 = ddq1 * dt + dq1;
                double Qa = INF;
                if (xn2 >= lb2 && xn2 <= ub2 && xn3 >= lb3 && xn3 <= ub3) {
                    cell_seek(k2, s_lev2, s_rinv2, N2, xn2, lb2, is2);
                    cell_seek(k3, s_lev3, s_rinv3, N3, xn3, lb3, is3);
                    const double y2 = exact_div(xn2 - k2.lo, k2.hi - k2.lo, k2.rinv);
                    const double y3 = exact_div(xn3 - k3.lo, k3.hi - k3.lo, k3.rinv);
                    const double omy2 = 1.0 - y2, omy3 = 1.0 - y3;
                    const int o = k2.c * N3 + k3.c;
                    const int o2 = o + N3;
                    double Jx;
                    {   // corners in itertools.product order: axis 0 slowest, offset 0 before 1
                        const double wa = w00 * omy2, wb = w00 * y2;
                        Jx = 0.0 + __ldg(b00 + o) * (wa * omy3);  // value = 0 + term (_rgi.py:528,547)
                        Jx = Jx + __ldg(b00 + o + 1) * (wa * y3);
                        Jx = Jx + __ldg(b00 + o2) * (wb * omy3);
                        Jx = Jx + __ldg(b00 + o2 + 1) * (wb * y3);
                    }
                    {
                        const double wa = w01 * omy2, wb = w01 * y2;
                        Jx = Jx + __ldg(b01 + o) * (wa * omy3);
                        Jx = Jx + __ldg(b01 + o + 1) * (wa * y3);
                        Jx = Jx + __ldg(b01 + o2) * (wb * omy3);
                        Jx = Jx + __ldg(b01 + o2 + 1) * (wb * y3);
                    }
                    {
                        const double wa = w10 * omy2, wb = w10 * y2;
                        Jx = Jx + __ldg(b10 + o) * (wa * omy3);
                        Jx = Jx + __ldg(b10 + o + 1) * (wa * y3);
                        Jx = Jx + __ldg(b10 + o2) * (wb * omy3);
                        Jx = Jx + __ldg(b10 + o2 + 1) * (wb * y3);
                    }
                    {
                        const double wa = w11 * omy2, wb = w11 * y2;
                        Jx = Jx + __ldg(b11 + o) * (wa * omy3);
                        Jx = Jx + __ldg(b11 + o + 1) * (wa * y3);
                        Jx = Jx + __ldg(b11 + o2) * (wb * omy3);
                        Jx = Jx + __ldg(b11 + o2 + 1) * (wb * y3);
                    }
                    Qa = (gx + s_act[2 * a + 1].x) * dt_cost + (ALPHA1 ? Jx : alpha * Jx);
                }
                if (Qa < best) { best = Qa; besta = a; }
            }
        } else if (g == 0) {
            best = P.INF;
            besta = 0;
        }
    }
    if (G > 1) lane_group_argmin(best, besta, G);
    if (active && g == 0) {
        if (besta == 0x7fffffff) besta = 0;  // no Q below +inf (cf.INF = inf): np.argmin of a constant row is 0
        Jo[node] = best;
        pi[node] = besta;
        const double d = best - Jn[node];
        st.jmax = best; st.dmax = d; st.dmin = d;
    }
    block_stats_finish(st, partials, counter, stats);
}
