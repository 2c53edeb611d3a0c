// Batches of closed-loop Euler rollouts under the tabulated policy: the validation step the reference's examples run
// after value iteration (`cl_sys = ctl + sys; cl_sys.compute_trajectory(tf, n, 'euler')`), for many initial states at once.
//
//   pyro/analysis/simulation.py:298-324   x[i+1] = f(x[i], u[i]) * dt + x[i], dt = tf / (npts - 1)
//   pyro/control/controller.py:326-355    ClosedLoopSystem.f: y = plant.h(x) = x, u = controller.c(y, r, t), dx = plant.f(x, u, t)
//   pyro/planning/dynamicprogramming.py:27-107  LookUpTableController: u[k] = RGI(levels, u_k grid, 'linear',
//                                         bounds_error=False, fill_value=0)(x), u_k[s] = input_from_action_id[pi[s], k]
//   pyro/dynamic/mechanical.py:222-263    f = [dq, inv(H)(B u - C dq - g - d)] with the model's H, C, g, d
//
// One thread per trajectory; the policy is read through pi (the 2^n corner nodes' actions, then the action table), so no
// u_k tables are materialised.  Unlike the sweep, the state is NOT on a grid level: sin / cos and the 2x2 inverse are
// evaluated on the device, so this path is floating-point parity (tests: <= 1e-9 relative against trajectories of the
// unmodified reference), not bit parity.  Sequential in time by nature: a batch of B trajectories of npts points costs
// npts dependent steps; the kernel is latency-bound and only pays off for batches (policy validation over a set of
// initial states), which the reference would loop over in Python at ~50 us per step.
#pragma once
#include "pyrodp_device.cuh"

#define PDP_ROLLOUT_PHYS 16

// plant.f(x, u) for the fused systems from their raw physical parameters (layout: include/pyrodp.h, pdp_rollout)
template <int N>
__device__ __forceinline__ void plant_f(int system_id, const double* __restrict__ ph, const double* x, const double* u, double* dx) {
    if (N == 2) {
        // SinglePendulum (pendulum.py:52-150): H = m1 lc1^2 + I1, C = 0, B = 1, g = m1 g lc1 sin q, d = d1 dq
        const double m1 = ph[0], lc1 = ph[1], I1 = ph[2], grav = ph[3], d1 = ph[4];
        const double H = m1 * (lc1 * lc1) + I1;
        const double g = m1 * grav * lc1 * sin(x[0]);
        const double d = d1 * x[1];
        const double rhs = ((u[0] - 0.0 * x[1]) - g) - d;
        dx[0] = x[1];
        dx[1] = (1.0 / H) * rhs;
        return;
    }
    const double q1 = x[1], dq0 = x[N - 2], dq1 = x[N - 1];
    double H00, H01, H11, r0, r1;
    if (system_id == PDP_SYS_TWOLINK) {
        // DoublePendulum / TwoLinkManipulator (pendulum.py:400-493 == manipulator.py:897-992)
        const double m1 = ph[0], l1 = ph[1], lc1 = ph[2], I1 = ph[3], m2 = ph[4], lc2 = ph[5], I2 = ph[6], grav = ph[7];
        const double d1 = ph[8], d2 = ph[9];
        const double c2 = cos(q1), s2 = sin(q1);
        H00 = m1 * (lc1 * lc1) + I1 + m2 * (l1 * l1 + lc2 * lc2 + 2 * l1 * lc2 * c2) + I2;
        H01 = m2 * (lc2 * lc2) + m2 * l1 * lc2 * c2 + I2;
        H11 = m2 * (lc2 * lc2) + I2;
        const double h = m2 * l1 * lc2 * s2;
        const double C00 = -h * dq1, C10 = h * dq0, C01 = -h * (dq0 + dq1);
        const double s1 = sin(x[0]), s12 = sin(x[0] + q1);
        const double g1c = (m1 * lc1 + m2 * l1) * grav, g2c = m2 * lc2 * grav;
        const double g0 = -g1c * s1 - g2c * s12, g1 = -g2c * s12;
        r0 = ((u[0] - (C00 * dq0 + C01 * dq1)) - g0) - d1 * dq0;
        r1 = ((u[1] - C10 * dq0) - g1) - d2 * dq1;
    } else {
        // CartPole (cartpole.py:335-437): B = [1, 0]', C[0,1] = -m2 lcg sin(th) dth, g[1] = m2 g lcg sin(th)
        const double m1 = ph[0], m2 = ph[1], lcg = ph[2], grav = ph[3];
        const double c = cos(q1), s = sin(q1);
        H00 = m1 + m2;
        H01 = m2 * lcg * c;
        H11 = m2 * (lcg * lcg);
        const double C01 = (-m2 * lcg * s) * dq1;
        r0 = u[0] - C01 * dq1;
        r1 = 0.0 - m2 * grav * lcg * s;
    }
    const double det = H00 * H11 - H01 * H01;
    dx[0] = dq0;
    dx[1] = dq1;
    dx[N - 2] = (H11 * r0 - H01 * r1) / det;
    dx[N - 1] = (H00 * r1 - H01 * r0) / det;
}

// LookUpTableController.c(x): per input axis the linear RegularGridInterpolator of u_k over the state grid, 0 outside
// the grid (bounds_error=False, fill_value=0).  Interval / distance / blend conventions as in the sweep:
// 2-D value-first, N-D weight-first in itertools.product corner order.
template <int N>
__device__ __forceinline__ void policy_lookup(const DevProblem& P, const long long* __restrict__ pi, const double* x, double* u) {
    const int m = P.m;
    int c[N];
    double y[N];
    bool inside = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const double lo = __ldg(P.level[d]), hi = __ldg(P.level[d] + P.dims[d] - 1);
        if (x[d] < lo || x[d] > hi) inside = false;
    }
    for (int k = 0; k < m; ++k) u[k] = 0.0;
    if (!inside) return;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        if (x[d] != x[d]) {   // NaN in -> NaN out
            for (int k = 0; k < m; ++k) u[k] = x[d];
            return;
        }
        c[d] = find_cell(P.level[d], P.dims[d], x[d], P.lb[d], P.inv_step[d]);
        const double l = __ldg(P.level[d] + c[d]), h = __ldg(P.level[d] + c[d] + 1);
        y[d] = (x[d] - l) / (h - l);
    }
    long long base = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) base += (long long)c[d] * P.stride[d];
    for (int k = 0; k < m; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int corner = 0; corner < (1 << N); ++corner) {
            long long node = base;
            double w = 1.0;
#pragma unroll
            for (int d = 0; d < N; ++d) {
                const int bit = (corner >> (N - 1 - d)) & 1;   // axis 0 slowest
                node += bit ? P.stride[d] : 0;
                w = w * (bit ? y[d] : 1.0 - y[d]);
            }
            const double v = __ldg(P.u_flat + pi[node - P.slab_node_begin] * m + k);
            if (N == 2) {
                // evaluate_linear_2d: ((v * w0) * w1), terms added left to right
                const double t = (v * ((corner >> 1) ? y[0] : 1.0 - y[0])) * ((corner & 1) ? y[1] : 1.0 - y[1]);
                acc = corner == 0 ? t : acc + t;
            } else {
                acc = acc + v * w;
            }
        }
        u[k] = acc;
    }
}

// x_out / u_out: kept sample j = point j*stride, layout [n_keep][n or m][B] (trajectory fastest: coalesced stores)
template <int N>
__global__ void __launch_bounds__(128)
rollout_kernel(const __grid_constant__ DevProblem P, const long long* __restrict__ pi, const double* __restrict__ phys,
               const double* __restrict__ x0, long long B, int npts, double dt, int stride, double* __restrict__ x_out,
               double* __restrict__ u_out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double ph[PDP_ROLLOUT_PHYS];
#pragma unroll
    for (int i = 0; i < PDP_ROLLOUT_PHYS; ++i) ph[i] = __ldg(phys + i);
    double x[N], u[2], dx[N];
#pragma unroll
    for (int d = 0; d < N; ++d) x[d] = x0[b * N + d];
    const int m = P.m;
    for (int i = 0; i < npts; ++i) {
        policy_lookup<N>(P, pi, x, u);
        if (i % stride == 0) {
            const long long j = i / stride;
#pragma unroll
            for (int d = 0; d < N; ++d) x_out[(j * N + d) * B + b] = x[d];
            if (u_out)
                for (int k = 0; k < m; ++k) u_out[(j * m + k) * B + b] = u[k];
        }
        if (i + 1 < npts) {
            plant_f<N>(P.system_id, ph, x, u, dx);
#pragma unroll
            for (int d = 0; d < N; ++d) x[d] = dx[d] * dt + x[d];
        }
    }
}
