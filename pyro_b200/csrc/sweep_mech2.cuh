// Range-skipping fused Bellman sweep for the 2-dof systems (n = 4), sm_100a — the kernel the 4-D BASELINE
// configurations run (TwoLinkManipulator 101^4 x 21^2, CartPole 151^4 x 51, DoublePendulum 201^4 x 31^2).
//
// Same backup as sweep_mech2_kernel (sweep_fused.cuh), i.e. one sweep of
//   pyro/planning/dynamicprogramming.py:195-236 with x_next = f(x,u)*dt + x (discretizer.py:363),
//   f = [dq, inv(H)(B u - C dq - g - d)] (mechanical.py:222-263), strict box tests (system.py:198-215),
//   scipy's n-linear RegularGridInterpolator in its weight-first association (_rgi.py:520-549),
// bit for bit, but organised around what the measured workloads look like (scripts/valid_fraction.py):
//   * TwoLinkManipulator at its default bounds: 88 % of the (node, action) pairs leave the box — the
//     accelerations are tens of velocity cells per action step.  x_next[2], x_next[3] are affine in the
//     inner input, so the inner indices a1 that CAN stay inside the box form an interval; a linear estimate
//     widened by a safety margin brackets it and only that interval is evaluated (with the exact
//     arithmetic — the estimate only decides what is skipped, never a value).  Skipped pairs are Q = INF
//     (dynamicprogramming.py:228); np.argmin's first-index rule is kept by remembering the first skipped /
//     out-of-box index and merging INF at that index after the loop.
//   * What binds all three configurations is the L1 data pipe (ncu r02b: l1tex__data_pipe_lsu_wavefronts 91-99 % of
//     peak, FP64 pipe 27-47 %): ~56 wavefronts per warp-eval for the 16 LDG.64 gathers (2 is the floor of a 256-byte
//     warp access, unaligned segments and the lanes' different (c0,c1) planes make it 3.5) and, in the first version,
//     23 more for per-lane 32-byte cell records and per-lane action records in shared memory.  Hence: the action
//     loops are WARP-UNIFORM (every lane evaluates the same action at the same time, so the lanes' gathers are
//     neighbours in memory and the action records are broadcasts; a lane outside its own bracket idles), and the
//     level / reciprocal tables are plain arrays read with conflict-free LDS.64.
//   * The cell of an action-dependent axis is computed directly (arithmetic guess, verified against the level
//     table exactly as scipy's search decides it); an in-cell x_next needs no separate box test.  Keeping the
//     16 corner values of the current cell in registers for the cart-pole (a quarter of a cell per action) was
//     measured slower (r02a: 298 vs 254 ms at cfg4): the lanes of a warp change cells at different actions.
// Blocks are rasterised chunk-fastest (consecutive blocks = consecutive 128-node chunks of ONE (i0,i1) plane),
// so the ~600 resident blocks share a handful of J planes in L2 instead of 600 different ones
// (ncu r01b: 7.3x the compulsory DRAM reads with the plane-fastest order).
#pragma once
#include "pyrodp_device.cuh"

#ifndef SWEEP_THREADS
#define SWEEP_THREADS 128
#endif
// Resident blocks per SM asked of ptxas (register budget), measured on B200 (profiles/r02h_variants.jsonl, ms per sweep
// cfg3 / cfg4 / DoublePendulum 81^4): 4 blocks (128 registers) 119 / 243 / 384, 5 blocks (96, 40 B spilled) 122 / 249 / 459,
// 6 blocks (80, ~100 B spilled) 126 / 234 / 466.  The cart-pole kernel is the one that is not yet L1-saturated at four
// blocks (L1 data pipe 79 %, ncu r02g) and gains from the extra warps; the two-input kernel loses to its spills.
#ifndef MECH2R_MIN_BLOCKS
#define MECH2R_MIN_BLOCKS(SYS) ((SYS) == PDP_SYS_CARTPOLE ? 6 : 4)
#endif

// make a pointer opaque to the optimiser: ptr[int_index] then compiles to one IMAD.WIDE from a register-resident base
template <typename T>
__device__ __forceinline__ const T* opaque_ptr(const T* p) {
#ifdef __CUDA_ARCH__
    asm("" : "+l"(p));
#endif
    return p;
}

// Block shape (rows of the (i2,i3) plane per 128-node block), measured on B200 (profiles/r02k_tiles.jsonl, ms per sweep
// cfg3 / cfg4 / DoublePendulum 81^4 / cfg5): 1 row 113.8 / 236.2 / 373.9 / 17 079; 4 rows x 32 columns 113.1 / 245.2 / 358.6 / 14 670;
// 8 x 16: 109.2 / 238.1 / 337.1 / 14 159; 16 x 8: 109.4 / 240.9 / 343.5.  Adjacent i2 rows gather from the same base planes and
// columns and from k2 rows that differ by about one, so a tile's warps share L1 lines (an L1 miss costs extra data-pipe
// wavefronts, which is what binds the kernel); the cart-pole's hit rate is 91 % already and it only pays the ragged tiles.
// (256-thread blocks — 16 x 16 and 8 x 32 tiles — measured slower: r02l, DoublePendulum 81^4 363 / 378 ms against 351.)
#ifndef MECH2_DEFAULT_TILE_ROWS
#define MECH2_DEFAULT_TILE_ROWS(SYS) ((SYS) == PDP_SYS_TWOLINK ? 8 : 1)
#endif

// first out-of-box index bookkeeping: NONE = no INF entry in this node's Q row so far
#define MECH2_NONE 0x7fffffff

// 16-corner blend, corners in itertools.product order (axis 0 slowest, offset 0 before 1), weight-first:
// w = (((1*w0)*w1)*w2)*w3, value = 0 + v*w + ... (_rgi.py:528-547).  wXY = w0*w1 products of the node.
#define MECH2_BLEND(V)                                                                 \
    {                                                                                  \
        double wa = w00 * omy2, wb = w00 * y2;                                         \
        Jx = 0.0 + V(0) * (wa * omy3);                                                 \
        Jx = Jx + V(1) * (wa * y3);                                                    \
        Jx = Jx + V(2) * (wb * omy3);                                                  \
        Jx = Jx + V(3) * (wb * y3);                                                    \
        wa = w01 * omy2; wb = w01 * y2;                                                \
        Jx = Jx + V(4) * (wa * omy3);                                                  \
        Jx = Jx + V(5) * (wa * y3);                                                    \
        Jx = Jx + V(6) * (wb * omy3);                                                  \
        Jx = Jx + V(7) * (wb * y3);                                                    \
        wa = w10 * omy2; wb = w10 * y2;                                                \
        Jx = Jx + V(8) * (wa * omy3);                                                  \
        Jx = Jx + V(9) * (wa * y3);                                                    \
        Jx = Jx + V(10) * (wb * omy3);                                                 \
        Jx = Jx + V(11) * (wb * y3);                                                   \
        wa = w11 * omy2; wb = w11 * y2;                                                \
        Jx = Jx + V(12) * (wa * omy3);                                                 \
        Jx = Jx + V(13) * (wa * y3);                                                   \
        Jx = Jx + V(14) * (wb * omy3);                                                 \
        Jx = Jx + V(15) * (wb * y3);                                                   \
    }

// corner i of the 16: plane (i>>2) of the four (c0,c1) base planes, row (i>>1)&1 of {k2, k2+1}, column i&1.  One
// register-resident base pointer and 32-bit element offsets: each row pair is one IMAD.WIDE away (the four 64-bit
// base pointers of the first version cost eight registers, re-deriving them four integer instructions per row).
#define MECH2_CORNER_PTR(i) (b00 + (o + ((((i) >> 2) & 1) ? plane_sz : 0) + (((i) >> 3) ? n1ps : 0) + ((((i) >> 1) & 1) ? N3 : 0)) + ((i) & 1))

template <int SYS, bool ALPHA1>
__global__ void __launch_bounds__(SWEEP_THREADS, MECH2R_MIN_BLOCKS(SYS))
sweep_mech2_range_kernel(const __grid_constant__ DevProblem P, const double* __restrict__ Jn, double* __restrict__ Jo,
                         long long* __restrict__ pi, unsigned long long* __restrict__ partials, unsigned int* counter,
                         double* __restrict__ stats) {
    extern __shared__ __align__(16) double smem[];
    constexpr int CV = (SYS == PDP_SYS_TWOLINK) ? 1 : 0;   // row of B.u that varies along the inner action index
    const int N0 = P.dims[0], N1 = P.dims[1], N2 = P.dims[2], N3 = P.dims[3], A = P.A, A0 = P.A0, A1 = P.A1;
    const int N2p = (N2 + 1) & ~1, N3p = (N3 + 1) & ~1;
    double* s_lev2 = smem;                       // [N2p] levels of axis 2
    double* s_rinv2 = s_lev2 + N2p;              // [N2p] correctly rounded 1/(lev[k+1]-lev[k])
    double* s_lev3 = s_rinv2 + N2p;              // [N3p]
    double* s_rinv3 = s_lev3 + N3p;              // [N3p]
    // The action table is separable (mech2_plan.h): row CV of B.u is a function of the inner index, row 1-CV of the outer
    // one — A0 + A1 doubles instead of 2A (15 KB per block at 31 x 31 actions: shared memory is carved out of the L1, and
    // what binds this kernel is L1 miss fills).  du'R du stays a table per action.
    double* s_buv = s_rinv3 + N3p;               // [A1] B.u[CV] along the inner index
    double* s_buo = s_buv + ((A1 + 1) & ~1);     // [A0] B.u[1-CV] along the outer index
    double* s_gu = s_buo + ((A0 + 1) & ~1);      // [A] du'R du
    for (int i = threadIdx.x; i < N2; i += blockDim.x) { s_lev2[i] = __ldg(P.level[2] + i); s_rinv2[i] = __ldg(P.rinv[2] + i); }
    for (int i = threadIdx.x; i < N3; i += blockDim.x) { s_lev3[i] = __ldg(P.level[3] + i); s_rinv3[i] = __ldg(P.rinv[3] + i); }
    for (int i = threadIdx.x; i < A1; i += blockDim.x) s_buv[i] = __ldg(P.bu + 2 * i + CV);
    for (int i = threadIdx.x; i < A0; i += blockDim.x) s_buo[i] = __ldg(P.bu + 2 * (i * A1) + (1 - CV));
    for (int i = threadIdx.x; i < A; i += blockDim.x) s_gu[i] = __ldg(P.gu + i);
    __syncthreads();

    // block -> ((i0,i1) plane, chunk of the (i2,i3) plane), chunk fastest
    const int plane_sz = N2 * N3;                                  // checked on the host: < 2^31 / 16
    const unsigned chunks = (unsigned)P.chunks;
    const unsigned pl_local = blockIdx.x / chunks;
    const unsigned chunk = blockIdx.x - pl_local * chunks;
    const long long pl = P.plane_begin + (long long)pl_local;     // (i0,i1) pair, C order
    const int i0 = (int)(pl / N1);
    const int i1 = (int)(pl - (long long)i0 * N1);
    // Nodes of a block: 128 consecutive nodes of the plane (tile_rows == 1), or a tile of tile_rows adjacent i2 rows x
    // 128/tile_rows columns, one warp-row each: the warps of a block then gather from the same base planes, the same
    // columns and adjacent k2 rows, i.e. they share L1 lines (ncu: an L1 miss costs extra data-pipe wavefronts).
    int i2, i3;
    bool active;
    if (P.tile_rows > 1) {
        const int tc = SWEEP_THREADS / P.tile_rows;                 // columns per tile (32: one warp per row; 16 / 8: two / four rows per warp)
        const unsigned tiles_c = (unsigned)((N3 + tc - 1) / tc);
        const unsigned tile_r = chunk / tiles_c, tile_c = chunk - tile_r * tiles_c;
        i2 = (int)tile_r * P.tile_rows + (int)threadIdx.x / tc;
        i3 = (int)tile_c * tc + (int)threadIdx.x % tc;
        active = i2 < N2 && i3 < N3;
    } else {
        const int r_raw = (int)(chunk * SWEEP_THREADS + threadIdx.x);
        active = r_raw < plane_sz;
        const int rr = min(r_raw, plane_sz - 1);
        i2 = rr / N3;
        i3 = rr - i2 * N3;
    }
    // lanes outside the plane shadow a valid node: the action loops are warp-uniform (votes inside), their result is discarded
    i2 = min(i2, N2 - 1);
    i3 = min(i3, N3 - 1);
    const int r = i2 * N3 + i3;                                   // (i2,i3) within the plane
    const long long node = pl * plane_sz + r;
    const double PINF = __longlong_as_double(0x7ff0000000000000LL);

    const double q0 = __ldg(P.level[0] + i0), q1 = __ldg(P.level[1] + i1);
    const double dq0 = s_lev2[i2], dq1 = s_lev3[i3];
    const double dt = P.dt;

    // position rows of x_next are action independent: dq*dt + q
    const double xn0 = dq0 * dt + q0;
    const double xn1 = dq1 * dt + q1;
    const bool pos_ok = !(xn0 < P.lb[0] || xn0 > P.ub[0] || xn1 < P.lb[1] || xn1 > P.ub[1]);
    const bool live = active && pos_ok;     // this lane evaluates actions
    // (for a lane that is not live the cells below are clamped garbage; its bracket is empty, nothing is read through them)
    const int c0 = find_cell(P.level[0], N0, xn0, P.lb[0], P.inv_step[0]);
    const int c1 = find_cell(P.level[1], N1, xn1, P.lb[1], P.inv_step[1]);
    const double lo0 = __ldg(P.level[0] + c0), hi0 = __ldg(P.level[0] + c0 + 1);
    const double lo1 = __ldg(P.level[1] + c1), hi1 = __ldg(P.level[1] + c1 + 1);
    const double y0 = (xn0 - lo0) / (hi0 - lo0);
    const double y1 = (xn1 - lo1) / (hi1 - lo1);
    // _evaluate_linear weight-first association: w = (((1*w0)*w1)*w2)*w3 (_rgi.py:543-546)
    const double w00 = (1.0 - y0) * (1.0 - y1), w01 = (1.0 - y0) * y1;
    const double w10 = y0 * (1.0 - y1), w11 = y0 * y1;
    const double* __restrict__ b00 = opaque_ptr(Jn + ((long long)c0 * N1 + c1) * plane_sz);
    const int n1ps = N1 * plane_sz;     // checked on the host: 2 * N1 * N2 * N3 < 2^31

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
    const double dt_cost = ontarget ? 0.0 : dt;
    const double alpha = P.alpha;

    // ---- bracket of the inner indices whose x_next[2], x_next[3] can lie inside the box ----
    // x_next[k] = (hv_k * r_v + ho_k * r_o) * dt + dq_k with r_v ~ B.u[CV](a1) - off_v, so the index where
    // x_next[k] = e is  (e - dq_k)/dt * S_k + Bk - (ho_k * S_k) * r_o,  S_k = inv_ustep / hv_k.
    // An estimate only: it selects the evaluated interval; the margin mg covers its rounding.
    const double hv2 = CV ? H01 : H00, hv3 = CV ? H11 : H10;
    const double ho2 = CV ? H00 : H01, ho3 = CV ? H10 : H11;
    double C2lo = -1e300, C2hi = 1e300, HA2 = 0.0, C3lo = -1e300, C3hi = 1e300, HA3 = 0.0;
    double mg = 1e-3;   // safety margin of the bracket, in index units (the estimate's own rounding is ~1e-9 of its terms)
    {
        const double off_v = CV ? ((cd1 + g1) + d1) : ((cd0 + g0) + d0);
        const double Bk = (off_v - P.uv_first) * P.uv_inv_step;
        const double rdt = 1.0 / dt;
        if (fabs(hv2) > 1e-6 * (fabs(H00) + fabs(H01))) {
            const double S = P.uv_inv_step / hv2;
            const double ca = ((P.lb[2] - dq0) * rdt) * S + Bk, cb = ((P.ub[2] - dq0) * rdt) * S + Bk;
            C2lo = fmin(ca, cb); C2hi = fmax(ca, cb); HA2 = ho2 * S;
            mg = mg + 1e-9 * (fabs(ca) + fabs(cb));
        }
        if (fabs(hv3) > 1e-6 * (fabs(H10) + fabs(H11))) {
            const double S = P.uv_inv_step / hv3;
            const double ca = ((P.lb[3] - dq1) * rdt) * S + Bk, cb = ((P.ub[3] - dq1) * rdt) * S + Bk;
            C3lo = fmin(ca, cb); C3hi = fmax(ca, cb); HA3 = ho3 * S;
            mg = mg + 1e-9 * (fabs(ca) + fabs(cb));
        }
    }

    const double lb2 = P.lb[2], lb3 = P.lb[3], is2 = P.inv_step[2], is3 = P.inv_step[3];
    const double gb2 = -(lb2 * is2), gb3 = -(lb3 * is3);   // cell guess = (int)fma(x, is, gb): a starting point only
    int first_inf = MECH2_NONE;     // lowest action index whose Q is INF (skipped or out of the box)
    double best = PINF;
    int besta = MECH2_NONE;

    for (int a0 = 0; a0 < A0; ++a0) {
        const int base = a0 * A1;
        // the outer residual: row (1-CV) of B u - C dq - g - d, left to right (mechanical.py:231).  For the
        // cart-pole g[0], d[0], d[1] are literal zeros (cartpole.py:415-437) and x - 0.0 == x bit for bit.
        const double bu_o = s_buo[a0];
        double r_o;
        if (SYS == PDP_SYS_TWOLINK) r_o = ((bu_o - cd0) - g0) - d0;       // CV = 1: outer row 0
        else r_o = (bu_o - cd1) - g1;                                    // CV = 0: outer row 1 (B.u[1] = +-0)
        int lo = A1, hi = 0;    // a lane that is not live has an empty bracket
        if (live) {
            const double t2 = HA2 * r_o, t3 = HA3 * r_o;
            const double lo_d = fmax(C2lo - t2, C3lo - t3) - mg;
            const double hi_d = fmin(C2hi - t2, C3hi - t3) + mg;
            lo = 0; hi = A1;
            if (lo_d > 0.0) lo = min(__double2int_ru(lo_d), A1);
            if (hi_d < (double)(A1 - 1)) hi = max(__double2int_rd(hi_d) + 1, 0);
            if (lo > 0) first_inf = min(first_inf, base);
            else if (hi < A1) first_inf = min(first_inf, base + hi);
        }
        // Wide brackets (most actions stay in the box): the warp walks the UNION of its lanes' brackets in step, every
        // lane on the same action — the gathers of the lanes are neighbours in memory and the action records are
        // broadcasts.  Narrow brackets (TwoLinkManipulator at its default bounds: two or three of 21): each lane walks
        // its own, the trip count is the widest lane's instead of the union's.  A warp-uniform choice per outer action.
        const int lo_w = __reduce_min_sync(0xffffffffu, lo), hi_w = __reduce_max_sync(0xffffffffu, hi);
        const int wmax = __reduce_max_sync(0xffffffffu, hi - lo);
        const bool in_step = 4 * (hi_w - lo_w) <= 5 * wmax;
        const int start = in_step ? lo_w : lo;
        const int trip = in_step ? (hi_w - lo_w) : wmax;
#ifdef MECH2_COUNT_EVALS   // test instrumentation (CPU emulation only): how many pairs the bracket let through
        mech2_evals_done += (hi > lo) ? (hi - lo) : 0;
#endif
        for (int it = 0; it < trip; ++it) {
            const int a1 = start + it;
            if (a1 < lo || a1 >= hi) continue;  // outside this lane's bracket: Q = INF, already accounted for
            const int a = base + a1;
            const double bu_v = s_buv[a1];      // a broadcast when the warp is in step
            const double gua = s_gu[a];
            double r0, r1;
            if (SYS == PDP_SYS_TWOLINK) { r0 = r_o; r1 = ((bu_v - cd1) - g1) - d1; }
            else { r0 = bu_v - cd0; r1 = r_o; }
            const double ddq0 = mv2(H00, H01, r0, r1);
            const double ddq1 = mv2(H10, H11, r0, r1);
            const double x2 = ddq0 * dt + dq0;
            const double x3 = ddq1 * dt + dq1;
            // direct cell: arithmetic guess, the level table verifies it (an in-cell x is inside the box)
            int k2 = min(max((int)fma(x2, is2, gb2), 0), N2 - 2);
            int k3 = min(max((int)fma(x3, is3, gb3), 0), N3 - 2);
            double l2 = s_lev2[k2], h2 = s_lev2[k2 + 1], l3 = s_lev3[k3], h3 = s_lev3[k3 + 1];
            if (!(x2 >= l2 && x2 < h2 && x3 >= l3 && x3 < h3)) {
                // isavalidstate (system.py:198-205), then the level table decides the cell as scipy's search does
                if (!(x2 >= lb2 && x2 <= P.ub[2] && x3 >= lb3 && x3 <= P.ub[3])) { first_inf = min(first_inf, a); continue; }
                while (x2 < l2 && k2 > 0) { --k2; h2 = l2; l2 = s_lev2[k2]; }
                while (x2 >= h2 && k2 < N2 - 2) { ++k2; l2 = h2; h2 = s_lev2[k2 + 1]; }
                while (x3 < l3 && k3 > 0) { --k3; h3 = l3; l3 = s_lev3[k3]; }
                while (x3 >= h3 && k3 < N3 - 2) { ++k3; l3 = h3; h3 = s_lev3[k3 + 1]; }
            }
            const double y2 = exact_div(x2 - l2, h2 - l2, s_rinv2[k2]);
            const double y3 = exact_div(x3 - l3, h3 - l3, s_rinv3[k3]);
            const double omy2 = 1.0 - y2, omy3 = 1.0 - y3;
            const int o = k2 * N3 + k3;
            double Jx;
#define MECH2_V(i) __ldg(MECH2_CORNER_PTR(i))
            MECH2_BLEND(MECH2_V)
#undef MECH2_V
            const double Qa = (gx + gua) * dt_cost + (ALPHA1 ? Jx : alpha * Jx);
            if (Qa < best) { best = Qa; besta = a; }   // strict <, ascending a: first index wins (np.argmin)
        }
    }
    Stats3 st = stats_identity();
    if (active) {
        if (pos_ok) {
            // merge the INF entries of the row: np.argmin takes the lowest index among equal minima
            if (first_inf != MECH2_NONE) {
                const double INF = P.INF;
                if (INF < best || (INF == best && first_inf < besta)) { best = INF; besta = first_inf; }
            }
            if (!(best < PINF)) besta = 0;   // nothing below +inf (cf.INF = inf): argmin of a constant row is 0
        } else {
            best = P.INF;                    // every action leaves the box through a position row
            besta = 0;
        }
        Jo[node] = best;
        pi[node] = besta;
        const double d = best - Jn[node];
        st.jmax = best; st.dmax = d; st.dmin = d;
    }
    block_stats_finish(st, partials, counter, stats);
}
