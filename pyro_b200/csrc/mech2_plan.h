// Host-side analysis of the action table of the 2-dof fused systems (plain C++, no CUDA): decides whether the
// range-skipping kernel sweep_mech2_range_kernel (sweep_mech2.cuh) may run and with which parameters.
// Shared by pyrodp.cu and, as test infrastructure, by the CPU emulation in tests/emu/.
//
// The kernel needs the action list a = a0*A1 + a1 (C order of u_grid_dim, discretizer.py:263-302) to be
//   * separable: the row of B.u that varies along the inner index a1 ("cv") depends on a1 only and the other
//     row on a0 only (B = I for the 2-link arm, B = [1,0]' for the cart-pole: mechanical.py:231),
//   * an (almost) uniform ascending ladder along a1 — np.linspace input levels (discretizer.py:150-163) —
//     so that the set of a1 whose x_next can lie inside the box is bracketed by a linear estimate,
//   * free of disallowed actions (isavalidinput true everywhere, system.py:208-215).
// Anything else runs the order-agnostic kernel sweep_mech2_kernel (sweep_fused.cuh).
#pragma once
#include <math.h>

struct Mech2Plan {
    int ok;                 // the range kernel may be used
    int A0, A1, cv;         // outer / inner action counts, row of B.u that varies with a1
    double uv_first;        // B.u[cv] of a1 = 0
    double uv_inv_step;     // 1 / (ladder step of B.u[cv])
};

// bu: [A][2] with NaN rows for disallowed actions (as uploaded); udims: u_grid_dim; twolink: 2-input arm, else cart-pole
static inline Mech2Plan mech2_plan(int twolink, const int* udims, const double* bu, long long A, int all_act_ok, double dt) {
    Mech2Plan p = {0, 1, (int)A, twolink ? 1 : 0, 0.0, 0.0};
    if (!all_act_ok || !(dt > 0.0) || A < 2) return p;
    if (twolink) { p.A0 = udims[0]; p.A1 = udims[1]; }
    if ((long long)p.A0 * p.A1 != A || p.A1 < 2) return p;
    const int cv = p.cv, co = 1 - cv;
    for (int a0 = 0; a0 < p.A0; ++a0)
        for (int a1 = 0; a1 < p.A1; ++a1) {
            const double* row = bu + 2 * ((long long)a0 * p.A1 + a1);
            if (!(row[cv] == bu[2 * a1 + cv])) return p;                        // inner row: a function of a1 only
            if (!(row[co] == bu[2 * ((long long)a0 * p.A1) + co])) return p;    // outer row: a function of a0 only (+-0 compare equal)
        }
    const double first = bu[cv], last = bu[2 * (p.A1 - 1) + cv];
    const double step = (last - first) / (double)(p.A1 - 1);
    if (!(step > 0.0) || !isfinite(step)) return p;
    for (int a1 = 0; a1 < p.A1; ++a1)
        if (!(fabs(bu[2 * a1 + cv] - (first + a1 * step)) <= 1e-9 * (last - first))) return p;
    p.uv_first = first;
    p.uv_inv_step = 1.0 / step;
    p.ok = 1;
    return p;
}
