/*
 * dp_oracle.c — plain-C restatement of the reference's grid-DP Bellman backup.
 *
 * TEST INFRASTRUCTURE ONLY.  Linked/loaded only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py, as the checker or the baseline being timed.
 * Nothing under pyro_b200/ may use it.
 *
 * Parity status: PINNED through oracle/np_oracle.py and the tier-0 fixtures in tests/golden/
 * (tests/test_oracle.py checks this file against both, bit for bit).
 *
 * It evaluates, for a node range, the literal per-pair formulation of
 *   pyro/planning/dynamicprogramming.py:195-236  (base-class compute_backward_step)
 * with
 *   x_next = f(x,u)*dt + x                      pyro/planning/discretizer.py:363
 *   f = [dq, inv(H)(B u - C dq - g - d)]        pyro/dynamic/mechanical.py:222-263
 *   validity: strict box tests                  pyro/dynamic/system.py:198-215
 *   J(x_next): scipy RegularGridInterpolator linear, fill 0 (scipy/interpolate/_rgi.py:375-483,
 *              520-549; find_interval_ascending binary search; 2-D value-first, N-D weight-first)
 *   Q = g(x,u)*dt + alpha*J ; INF when invalid   dynamicprogramming.py:223-233
 *   J = min Q, pi = argmin Q (first index)       dynamicprogramming.py:235-236
 * Transcendental / LAPACK terms come in as per-level tables (same descriptor as include/pyrodp.h)
 * because they must carry NumPy's bits, not libm's.  np.dot's fused-multiply-add association
 * (OpenBLAS) is restated with fma(); everything else is separate IEEE operations
 * (compile with -ffp-contract=off).
 *
 * Deliberately written differently from the CUDA kernels: generic n-D loops, a real binary
 * search and a real division, so that agreement is evidence rather than shared code.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pyrodp.h"

void orc_vfma(int64_t n, const double* a, const double* b, const double* c, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = fma(a[i], b[i], c[i]);
}

/* row of np.dot(M 2x2, v) */
static inline double mv2(double m0, double m1, double v0, double v1) { return fma(m0, v0, m1 * v1); }

/* scipy find_interval_ascending (extrapolate=1) for lb <= x <= ub: x[i] <= xval < x[i+1], last cell closed */
static int find_interval(const double* lev, int n, double x) {
    if (x == lev[n - 1]) return n - 2;
    int low = 0, high = n - 2;
    if (x < lev[low + 1]) high = low;
    while (low < high) {
        int mid = (high + low) / 2;
        if (x < lev[mid]) high = mid;
        else if (x >= lev[mid + 1]) low = mid + 1;
        else { low = mid; break; }
    }
    return low;
}

/* RegularGridInterpolator linear; returns 0 and sets *oob when outside [level[0], level[-1]] on any axis */
static double rgi_linear(const pdp_problem* p, const double* J, const double* x, int* oob) {
    const int n = p->n;
    int idx[PDP_MAX_N];
    double y[PDP_MAX_N];
    int64_t stride[PDP_MAX_N];
    *oob = 0;
    for (int d = 0; d < n; ++d)
        if (x[d] < p->x_level[d][0] || x[d] > p->x_level[d][p->dims[d] - 1]) *oob = 1;
    if (*oob) return 0.0;
    int64_t s = 1;
    for (int d = n - 1; d >= 0; --d) { stride[d] = s; s *= p->dims[d]; }
    for (int d = 0; d < n; ++d) {
        const double* lev = p->x_level[d];
        idx[d] = find_interval(lev, p->dims[d], x[d]);
        y[d] = (x[d] - lev[idx[d]]) / (lev[idx[d] + 1] - lev[idx[d]]);
    }
    if (n == 2) {
        const double* v = J + (int64_t)idx[0] * stride[0] + idx[1];
        double r = v[0] * (1 - y[0]) * (1 - y[1]);
        r = r + v[1] * (1 - y[0]) * y[1];
        r = r + v[stride[0]] * y[0] * (1 - y[1]);
        r = r + v[stride[0] + 1] * y[0] * y[1];
        return r;
    }
    double value = 0.0;
    for (int corner = 0; corner < (1 << n); ++corner) {
        double w = 1.0;
        int64_t off = 0;
        for (int d = 0; d < n; ++d) {
            int bit = (corner >> (n - 1 - d)) & 1;
            w = w * (bit ? y[d] : (1 - y[d]));
            off += (int64_t)(idx[d] + bit) * stride[d];
        }
        value = value + J[off] * w;
    }
    return value;
}

/* np.dot(dx.T, np.dot(W, dx)) and np.linalg.norm(dx) with the measured OpenBLAS association */
static double quad_form(const double* W, const double* d, int n) {
    double w[PDP_MAX_N];
    if (n == 2) {
        w[0] = mv2(W[0], W[1], d[0], d[1]);
        w[1] = mv2(W[2], W[3], d[0], d[1]);
        return fma(d[1], w[1], d[0] * w[0]);
    }
    for (int i = 0; i < 4; ++i) {
        const double* r = W + 4 * i;
        w[i] = (r[0] * d[0] + r[2] * d[2]) + (r[1] * d[1] + r[3] * d[3]);
    }
    return fma(d[3], w[3], fma(d[2], w[2], fma(d[1], w[1], d[0] * w[0])));
}
static double norm_l2(const double* d, int n) {
    double acc = d[0] * d[0];
    for (int i = 1; i < n; ++i) acc = fma(d[i], d[i], acc);
    return sqrt(acc);
}

/* dx = f(x,u) for node multi-index ix[] and action a */
static void f_eval(const pdp_problem* p, const int* ix, const double* x, int a, double* dx) {
    const int dof = p->n / 2;
    const double* bu = p->bu + (int64_t)a * dof;
    if (p->system_id == PDP_SYS_PENDULUM) {
        double g = p->sys_tab[0][ix[0]];
        double d = p->sys_par[1] * x[1];
        double rhs = ((bu[0] - 0.0 * x[1]) - g) - d;
        dx[0] = x[1];
        dx[1] = p->sys_par[0] * rhs;
        return;
    }
    const double dq0 = x[2], dq1 = x[3];
    const double* Hi = p->sys_tab[0] + 4 * ix[1];
    double cd0, cd1, g0, g1, d0, d1;
    if (p->system_id == PDP_SYS_TWOLINK) {
        double h = p->sys_tab[1][ix[1]];
        double C00 = -h * dq1, C10 = h * dq0, C01 = -h * (dq0 + dq1);
        cd0 = mv2(C00, C01, dq0, dq1);
        cd1 = mv2(C10, 0.0, dq0, dq1);
        const double* G = p->sys_tab[2] + 2 * ((int64_t)ix[0] * p->dims[1] + ix[1]);
        g0 = G[0]; g1 = G[1];
        d0 = mv2(p->sys_par[0], 0.0, dq0, dq1);
        d1 = mv2(0.0, p->sys_par[1], dq0, dq1);
    } else { /* CARTPOLE */
        double C01 = p->sys_tab[1][ix[1]] * dq1;
        cd0 = mv2(0.0, C01, dq0, dq1);
        cd1 = mv2(0.0, 0.0, dq0, dq1);
        g0 = 0.0; g1 = p->sys_tab[2][ix[1]];
        d0 = 0.0; d1 = 0.0;
    }
    double r0 = ((bu[0] - cd0) - g0) - d0;
    double r1 = ((bu[1] - cd1) - g1) - d1;
    dx[0] = dq0; dx[1] = dq1;
    dx[2] = mv2(Hi[0], Hi[1], r0, r1);
    dx[3] = mv2(Hi[2], Hi[3], r0, r1);
}

/* One Bellman backup for nodes [lo, hi).  Writes J_out[s-lo], pi_out[s-lo].  Returns 0 / -1. */
int orc_sweep_fused(const pdp_problem* p, const double* J_next, int64_t lo, int64_t hi, double* J_out, int64_t* pi_out,
                    int n_threads) {
    if (p->system_id == PDP_SYS_LUT) return -1;
    const int n = p->n;
    int64_t A = 1;
    for (int d = 0; d < p->m; ++d) A *= p->udims[d];
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
    for (int64_t s = lo; s < hi; ++s) {
        int ix[PDP_MAX_N];
        double x[PDP_MAX_N], dxb[PDP_MAX_N];
        int64_t r = s;
        for (int d = n - 1; d >= 0; --d) { ix[d] = (int)(r % p->dims[d]); r /= p->dims[d]; x[d] = p->x_level[d][ix[d]]; }
        for (int d = 0; d < n; ++d) dxb[d] = x[d] - p->xbar[d];
        double gx = (p->cost_id == PDP_COST_QUADRATIC) ? quad_form(p->Q, dxb, n) : 1.0;
        if (p->cost_id == PDP_COST_REACH) gx = 0.0;   /* Reachability.g on an in-box node (costfunction.py:468-481) */
        int ontarget = p->ontarget_check && (norm_l2(dxb, n) < p->EPS);
        double best = INFINITY;
        int64_t besta = 0;
        for (int64_t a = 0; a < A; ++a) {
            double Q = p->INF;
            if (p->act_ok[a]) {
                double f[PDP_MAX_N], xn[PDP_MAX_N];
                f_eval(p, ix, x, (int)a, f);
                int ok = 1;
                for (int d = 0; d < n; ++d) {
                    xn[d] = f[d] * p->dt + x[d];
                    if (xn[d] < p->x_lb[d] || xn[d] > p->x_ub[d]) ok = 0;
                }
                if (ok) {
                    int oob;
                    double Jx = rgi_linear(p, J_next, xn, &oob);
                    double g = ontarget ? 0.0 : (gx + p->gu[a]);
                    Q = g * p->dt + p->alpha * Jx;
                }
            }
            if (Q < best) { best = Q; besta = a; }
        }
        J_out[s - lo] = best;
        pi_out[s - lo] = besta;
    }
    return 0;
}

/* LUT formulation (dynamicprogramming.py:564-570) for nodes [lo,hi): x_next (K,A,n), G (K,A) */
int orc_sweep_lut(const pdp_problem* p, const double* J_next, const double* x_next, const double* G, int64_t lo, int64_t hi,
                  double* J_out, int64_t* pi_out, int n_threads) {
    const int n = p->n;
    int64_t A = 1;
    for (int d = 0; d < p->m; ++d) A *= p->udims[d];
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
    for (int64_t s = lo; s < hi; ++s) {
        double best = INFINITY;
        int64_t besta = 0;
        for (int64_t a = 0; a < A; ++a) {
            int oob;
            const int64_t k = (s - lo) * A + a;
            double Jx = rgi_linear(p, J_next, x_next + k * n, &oob);
            double Q = G[k] + p->alpha * Jx;
            if (Q < best) { best = Q; besta = a; }
        }
        J_out[s - lo] = best;
        pi_out[s - lo] = besta;
    }
    return 0;
}

/* terminal cost h(x) (dynamicprogramming.py:159-171; costfunction.py:139-149) */
int orc_terminal(const pdp_problem* p, double* J_out) {
    int64_t N = 1;
    for (int d = 0; d < p->n; ++d) N *= p->dims[d];
    for (int64_t s = 0; s < N; ++s) {
        double dxb[PDP_MAX_N];
        int64_t r = s;
        for (int d = p->n - 1; d >= 0; --d) { int i = (int)(r % p->dims[d]); r /= p->dims[d]; dxb[d] = p->x_level[d][i] - p->xbar[d]; }
        double h = 0.0;
        if (p->cost_id == PDP_COST_QUADRATIC) {
            h = quad_form(p->S, dxb, p->n);
            if (p->ontarget_check && norm_l2(dxb, p->n) < p->EPS) h = 0.0;
        }
        if (p->cost_id == PDP_COST_REACH) h = (norm_l2(dxb, p->n) < p->EPS) ? 0.0 : p->INF;   /* costfunction.py:442-465 */
        J_out[s] = h;
    }
    return 0;
}
