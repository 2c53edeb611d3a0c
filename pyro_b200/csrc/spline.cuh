// Bicubic-spline variant of the table-mode sweep: DynamicProgramming2DRectBivariateSpline
// (pyro/planning/dynamicprogramming.py:578-614) interpolates J_next with
//     RectBivariateSpline(x_level[0], x_level[1], J_grid, kx=3, ky=3)          (discretizer.py:591-612, s = 0)
// instead of the RegularGridInterpolator, then Q = G + alpha * J_interpol(x_next_table), J = min, pi = argmin.
//
// The algorithm lives in scipy's FITPACK wrapper (third party, scipy 1.18.1 in this image; not under /root/reference):
//   regrid / fpregr   s = 0: interpolating spline on the knots  t = {x0 x4, x[2..m-3], x[m-1] x4}  (not-a-knot: the second
//                     and the second-to-last sample are not knots), coefficients = solution of  A_x C A_y' = Z  with the
//                     collocation matrices A[i][j] = B_j(x_i)  (FITPACK reaches it by Givens QR; same unique solution)
//   bispeu / fpbisp   evaluation: the argument is CLAMPED to [x0, x[m-1]] (so states outside the grid extrapolate with
//                     the boundary value — unlike RGI's fill value 0), knot interval by upward search, the four non-zero
//                     cubic B-splines per axis by the de Boor-Cox recurrence (fpbspl), 4 x 4 coefficient sum
// Restated here: the fit as two sweeps of banded forward/back substitution on the device (LU factors of A without
// pivoting — collocation matrices of B-splines are totally positive — built once on the host from the level tables),
// the evaluation inside the sweep kernel.  Floating-point parity (tests: J <= 1e-9 relative against the unmodified
// reference class; pi equal except where the reference's own Q values tie to that tolerance), not bit parity: the
// elimination order differs from FITPACK's.
#pragma once
#include <vector>
#include "pyrodp_device.cuh"

struct SplineDev {
    const double* knots[2];   // [m + 4] per axis
    const double* lu[2];      // [5][m] per axis: l1 (sub-diagonal), l2 (second sub-diagonal), d (diagonal of U), u1, u2
    const double* rden[2];    // [m + 4][6] per axis: reciprocals of the six knot differences fpbspl divides by on interval l
    double* coef;             // [m0][m1] B-spline coefficients of J_next, rewritten before every sweep
    int m[2];
};

// fpbspl for k = 3: the four cubic B-splines that are non-zero on [t[l], t[l+1]) at x (0-based knot index l)
__host__ __device__ inline void bspl3(const double* t, int l, double x, double h[4]) {
    double hh[3];
    h[0] = 1.0;
    for (int j = 1; j <= 3; ++j) {
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 1; i <= j; ++i) {
            const int li = l + i, lj = li - j;
            const double f = hh[i - 1] / (t[li] - t[lj]);
            h[i - 1] = h[i - 1] + f * (t[li] - x);
            h[i] = f * (x - t[lj]);
        }
    }
}

// The same recurrence with the six divisors of interval l tabulated as reciprocals (host, spline_plan_axis): twelve FP64
// divisions per evaluation were half of the sweep kernel's instructions (ncu r02r).  One extra rounding per factor —
// this path is floating-point parity anyway.
__host__ __device__ inline void bspl3r(const double* t, const double* r6, int l, double x, double h[4]) {
    double hh[3];
    h[0] = 1.0;
    int q = 0;
    for (int j = 1; j <= 3; ++j) {
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 1; i <= j; ++i, ++q) {
            const int li = l + i, lj = li - j;
            const double f = hh[i - 1] * r6[q];
            h[i - 1] = h[i - 1] + f * (t[li] - x);
            h[i] = f * (x - t[lj]);
        }
    }
}

// fpbisp's interval search on the not-a-knot knot vector of m samples: 3 <= l <= m-1 with t[l] <= x < t[l+1]
// (x already clamped to [t[3], t[m]]; the last interval is closed on the right)
__host__ __device__ inline int spline_interval(const double* t, int m, double x, int guess) {
    int l = guess < 3 ? 3 : (guess > m - 1 ? m - 1 : guess);
    while (l > 3 && x < t[l]) --l;
    while (l < m - 1 && x >= t[l + 1]) ++l;
    return l;
}

// Host side: knots and banded LU (kl = ku = 2, no pivoting) of the collocation matrix of one axis.  false when m < 4
// (RectBivariateSpline needs more samples than the degree) or a pivot vanishes.
inline bool spline_plan_axis(const double* x, int m, std::vector<double>& knots, std::vector<double>& lu, std::vector<double>& rden) {
    if (m < 4) return false;
    knots.assign((size_t)m + 4, 0.0);
    for (int j = 0; j < 4; ++j) { knots[j] = x[0]; knots[m + j] = x[m - 1]; }
    for (int j = 4; j < m; ++j) knots[j] = x[j - 2];
    rden.assign(((size_t)m + 4) * 6, 0.0);
    for (int l = 3; l <= m - 1; ++l) {
        int q = 0;
        for (int j = 1; j <= 3; ++j)
            for (int i = 1; i <= j; ++i, ++q) rden[(size_t)l * 6 + q] = 1.0 / (knots[l + i] - knots[l + i - j]);
    }
    // ab[i][c], c = j - i + 2 in 0..4
    std::vector<double> ab((size_t)m * 5, 0.0);
    for (int i = 0; i < m; ++i) {
        const int l = spline_interval(knots.data(), m, x[i], i + 2);
        double h[4];
        bspl3(knots.data(), l, x[i], h);
        for (int r = 0; r < 4; ++r) {
            const int j = l - 3 + r, c = j - i + 2;
            if (c < 0 || c > 4) {
                if (h[r] != 0.0) return false;   // outside the band: must be a structural zero
                continue;
            }
            ab[(size_t)i * 5 + c] = h[r];
        }
    }
    lu.assign((size_t)5 * m, 0.0);
    double* l1 = lu.data(), *l2 = l1 + m, *d = l2 + m, *u1 = d + m, *u2 = u1 + m;
    for (int k = 0; k < m; ++k) {
        const double piv = ab[(size_t)k * 5 + 2];
        if (!(piv != 0.0)) return false;
        for (int i = k + 1; i <= k + 2 && i < m; ++i) {
            const double f = ab[(size_t)i * 5 + (k - i + 2)] / piv;
            ab[(size_t)i * 5 + (k - i + 2)] = f;
            for (int j = k + 1; j <= k + 2 && j < m; ++j) ab[(size_t)i * 5 + (j - i + 2)] -= f * ab[(size_t)k * 5 + (j - k + 2)];
        }
    }
    for (int i = 0; i < m; ++i) {
        l1[i] = i >= 1 ? ab[(size_t)i * 5 + 1] : 0.0;
        l2[i] = i >= 2 ? ab[(size_t)i * 5 + 0] : 0.0;
        d[i] = ab[(size_t)i * 5 + 2];
        u1[i] = i + 1 < m ? ab[(size_t)i * 5 + 3] : 0.0;
        u2[i] = i + 2 < m ? ab[(size_t)i * 5 + 4] : 0.0;
    }
    return true;
}

// ---- fit: C = A_x^{-1} Z A_y^{-T} as two sweeps of banded substitutions ---------------------------------------------
// axis 0: one thread per column j (consecutive threads = consecutive addresses); reads Z, writes C
__global__ void __launch_bounds__(128)
spline_fit_axis0_kernel(const double* __restrict__ Z, const __grid_constant__ SplineDev S) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int m0 = S.m[0], m1 = S.m[1];
    if (j >= m1) return;
    const double* __restrict__ l1 = S.lu[0], *l2 = l1 + m0, *d = l2 + m0, *u1 = d + m0, *u2 = u1 + m0;
    double* __restrict__ C = S.coef;
    double y1 = 0.0, y2 = 0.0;   // y[i-1], y[i-2]
    for (int i = 0; i < m0; ++i) {
        const double y = Z[(long long)i * m1 + j] - __ldg(l1 + i) * y1 - __ldg(l2 + i) * y2;
        C[(long long)i * m1 + j] = y;
        y2 = y1; y1 = y;
    }
    double c1 = 0.0, c2 = 0.0;   // c[i+1], c[i+2]
    for (int i = m0 - 1; i >= 0; --i) {
        const double c = (C[(long long)i * m1 + j] - __ldg(u1 + i) * c1 - __ldg(u2 + i) * c2) / __ldg(d + i);
        C[(long long)i * m1 + j] = c;
        c2 = c1; c1 = c;
    }
}
// axis 1: one thread per row i, in place (strided accesses; the grid of a 2-D problem is a few MB and sits in L2)
__global__ void __launch_bounds__(128)
spline_fit_axis1_kernel(const __grid_constant__ SplineDev S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m0 = S.m[0], m1 = S.m[1];
    if (i >= m0) return;
    const double* __restrict__ l1 = S.lu[1], *l2 = l1 + m1, *d = l2 + m1, *u1 = d + m1, *u2 = u1 + m1;
    double* __restrict__ row = S.coef + (long long)i * m1;
    double y1 = 0.0, y2 = 0.0;
    for (int j = 0; j < m1; ++j) {
        const double y = row[j] - __ldg(l1 + j) * y1 - __ldg(l2 + j) * y2;
        row[j] = y;
        y2 = y1; y1 = y;
    }
    double c1 = 0.0, c2 = 0.0;
    for (int j = m1 - 1; j >= 0; --j) {
        const double c = (row[j] - __ldg(u1 + j) * c1 - __ldg(u2 + j) * c2) / __ldg(d + j);
        row[j] = c;
        c2 = c1; c1 = c;
    }
}

// bispeu at one point: clamp, knot intervals, 4 + 4 basis values, 16 coefficients (fpbisp's summation order)
__device__ __forceinline__ double spline_eval(const DevProblem& P, const SplineDev& S, const double* x) {
    double w[2][4];
    int l[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const double* __restrict__ t = S.knots[a];
        const int m = S.m[a];
        double arg = x[a];
        const double tb = t[3], te = t[m];
        if (arg < tb) arg = tb;
        if (arg > te) arg = te;
        const int cell = (int)((arg - tb) * P.inv_step[a]);   // level cell c -> knot interval c + 2 (3 for the first two cells)
        l[a] = spline_interval(t, m, arg, cell + 2);
        bspl3r(t, S.rden[a] + 6 * l[a], l[a], arg, w[a]);
    }
    const double* __restrict__ c = S.coef + (long long)(l[0] - 3) * S.m[1] + (l[1] - 3);
    double sp = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sp = sp + c[(long long)i * S.m[1] + j] * w[0][i] * w[1][j];
    return sp;
}

// dynamicprogramming.py:598-614 on device tables: the structure of sweep_lut_kernel<2, G> with the spline interpolant
template <int G>
__global__ void __launch_bounds__(SWEEP_THREADS)
sweep_lut_spline_kernel(const __grid_constant__ DevProblem P, const __grid_constant__ SplineDev S, const double* __restrict__ Jn,
                        double* __restrict__ Jo, long long* __restrict__ pi, const double* __restrict__ xnext,
                        const double* __restrict__ Gtab, unsigned long long* __restrict__ partials, unsigned int* counter,
                        double* __restrict__ stats) {
    const int lane_in_group = threadIdx.x % G;
    const int A = P.A;
    const long long total = P.node_end - P.node_begin;
    const long long gstride = (long long)gridDim.x * blockDim.x / G;
    const long long iters = (total + gstride - 1) / gstride;
    Stats3 st = stats_identity();
    for (long long it = 0; it < iters; ++it) {
        const long long node = P.node_begin + it * gstride + ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
        const long long slot = node - P.slab_node_begin;
        const bool active = node < P.node_end;
        double best = __longlong_as_double(0x7ff0000000000000LL);
        int besta = 0x7fffffff;
        if (active) {
            const double* __restrict__ xrow = xnext + slot * (long long)A * 2;
            const double* __restrict__ grow = Gtab + slot * (long long)A;
            for (int a = lane_in_group; a < A; a += G) {
                const double x[2] = {__ldcs(xrow + 2LL * a), __ldcs(xrow + 2LL * a + 1)};
                const double Jx = spline_eval(P, S, x);
                const double Qa = __ldcs(grow + a) + P.alpha * Jx;
                if (Qa < best) { best = Qa; besta = a; }
            }
        }
        lane_group_argmin(best, besta, G);
        if (active && lane_in_group == 0) {
            if (besta == 0x7fffffff) besta = 0;
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
