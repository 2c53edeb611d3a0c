/*
 * pyrodp.h — C ABI of the B200 grid dynamic-programming (value-iteration) engine.
 *
 * This is the drop-in boundary for ONE hot path of SherbyRobotics/pyro: the Bellman sweep of
 * pyro.planning.dynamicprogramming.DynamicProgramming over
 * pyro.planning.discretizer.GridDynamicSystem.  pyro has no FFI; its extension point is
 * subclassing DynamicProgramming and overriding the three per-sweep hooks
 * (pyro/planning/dynamicprogramming.py:175 initialize_backward_step, :195/:557
 * compute_backward_step, :240 finalize_backward_step).  The functions below are exactly what
 * such a subclass binds through ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - plain C, no CUDA/torch types in any signature; device pointers and streams travel as void*.
 *   - every function returns 0 on success, a negative PDP_E* code on failure;
 *     pdp_last_error() gives the message (handle may be NULL for creation errors).
 *   - host arrays are borrowed for the duration of the call only; the handle owns device memory.
 *   - node ids are C order of x_grid_dim (last axis fastest), action ids C order of u_grid_dim,
 *     as pyro/planning/discretizer.py:167-310 enumerates them.
 *   - one handle = one caller thread = one GPU (the current CUDA device at pdp_create).
 *   - there is NO CPU fallback: if no CUDA device is usable pdp_create fails with PDP_ECUDA.
 */
#ifndef PYRODP_H
#define PYRODP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDP_ABI_VERSION 1
#define PDP_MAX_N 4 /* state dims supported by pyro's grid (discretizer.py:183-245: n in {2,3,4}) */
#define PDP_MAX_M 2 /* input dims supported (discretizer.py:271-306: m in {1,2}) */

/* error codes */
#define PDP_OK 0
#define PDP_EINVAL (-1)  /* bad argument / size mismatch   -> Python ValueError            */
#define PDP_ENOTSUP (-2) /* unsupported dims / system       -> Python NotImplementedError   */
#define PDP_ECUDA (-3)   /* CUDA runtime failure (sticky)   -> Python RuntimeError          */
#define PDP_ESTATE (-4)  /* call order violated (e.g. sweep before set_J)                   */

/* system_id: which closed-form f(x,u) the fused kernel evaluates on the fly */
#define PDP_SYS_LUT 0      /* no fused dynamics: sweeps use uploaded x_next/G tables (dynamicprogramming.py:557-570) */
#define PDP_SYS_PENDULUM 1 /* 1-dof MechanicalSystem, n=2,m=1   (pyro/dynamic/pendulum.py:16 SinglePendulum)       */
#define PDP_SYS_TWOLINK 2  /* 2-link arm form, n=4,m=2 (pendulum.py:340 DoublePendulum, manipulator.py:795 TwoLinkManipulator) */
#define PDP_SYS_CARTPOLE 3 /* cart + pole, n=4,m=1              (pyro/dynamic/cartpole.py:322 CartPole)              */

/* cost_id: which stage cost g(x,u) / terminal cost h(x) */
/* interpolant of J_next in table mode (pdp_set_interpolant) */
#define PDP_INTERP_LINEAR 0  /* RegularGridInterpolator 'linear' (discretizer.py:570-587): every mode's default          */
#define PDP_INTERP_SPLINE3 1 /* RectBivariateSpline kx=ky=3, s=0 (discretizer.py:591-612): n = 2, table mode only        */

#define PDP_COST_QUADRATIC 1 /* pyro/analysis/costfunction.py:100-204 QuadraticCostFunction */
#define PDP_COST_TIME 2      /* pyro/analysis/costfunction.py:287-334 TimeCostFunction      */
#define PDP_COST_REACH 3     /* pyro/analysis/costfunction.py:421-481 Reachability with the system's own box isavalidstate and the
                                default norm test: g = 0 on every (in-box) node, h = 0 if ||x - xbar|| < EPS else INF.
                                (QuadraticCostFunctionWithDomainCheck, :339-415, with the box isavalidstate IS PDP_COST_QUADRATIC
                                on grid nodes: a node never fails the box test.)  Custom callbacks run in LUT mode. */

/*
 * Problem descriptor (POD).  Everything a sweep needs, read by the host shim from the pyro
 * objects lazily at the first sweep (grid_sys, grid_sys.sys, cf, dp.alpha).
 *
 * Transcendentals and LAPACK results are NOT recomputed on the device: the shim tabulates them
 * per grid level with the very NumPy calls the reference makes (np.sin, np.cos, np.linalg.inv),
 * so those bits are identical by construction (hard part 1).
 *
 * sys_tab / sys_par layout per system_id (all float64, host pointers, copied at pdp_create):
 *  PENDULUM : sys_tab[0][dims[0]]        = gravity torque g(q_i)   (pendulum.py:126-137)
 *             sys_par[0] = inv(H)[0,0] (mechanical.py:231), sys_par[1] = d1 (pendulum.py:141-150)
 *  TWOLINK  : sys_tab[0][dims[1]*4]      = inv(H(q1_j)) row-major  (pendulum.py:400-417 / manipulator.py:897-918)
 *             sys_tab[1][dims[1]]        = h(q1_j)=m2*l1*lc2*sin q1 (pendulum.py:434 / manipulator.py:933)
 *             sys_tab[2][dims[0]*dims[1]*2] = g(q0_i,q1_j)          (pendulum.py:465-473 / manipulator.py:964-972)
 *             sys_par[0] = d1, sys_par[1] = d2                      (pendulum.py:477-493 / manipulator.py:978-992)
 *  CARTPOLE : sys_tab[0][dims[1]*4]      = inv(H(theta_j))          (cartpole.py:369-384)
 *             sys_tab[1][dims[1]]        = -m2*lcg*sin(theta_j)     (cartpole.py:399; times theta_dot on device)
 *             sys_tab[2][dims[1]]        = m2*g*lcg*sin(theta_j)    (cartpole.py:426)
 * bu[A*dof]  = np.dot(B, u_a) per action (mechanical.py:231), dof = n/2.
 * gu[A]      = du^T R du per action (costfunction.py:191), 0 for the time cost.
 * act_ok[A]  = isavalidinput(u_a) box test (pyro/dynamic/system.py:208-215), 1 = allowed.
 */
typedef struct pdp_problem {
    int32_t abi_version; /* PDP_ABI_VERSION */
    int32_t n;           /* state dimension, 2..4 */
    int32_t m;           /* input dimension, 1..2 */
    int32_t system_id;   /* PDP_SYS_* */
    int32_t cost_id;     /* PDP_COST_* */
    int32_t ontarget_check; /* costfunction.py:134 */
    int32_t slab_begin;  /* axis-0 planes [slab_begin, slab_end) are computed by this handle;   */
    int32_t slab_end;    /* 0, dims[0] on a single GPU (multi-GPU: SURVEY.md 8e)                  */
    int32_t alloc_planes; /* 0: the handle holds its slab plus the halo its backups can read (the whole
                             grid when the slab is the whole grid, or in LUT mode).  > 0: it holds the
                             whole grid in buffers of alloc_planes planes, e.g. W*ceil(dims[0]/W), so an
                             equal-count in-place all-gather of whole slabs fits (fallback exchange)  */
    int32_t reserved0;
    int32_t dims[PDP_MAX_N];  /* x_grid_dim (discretizer.py:91)  */
    int32_t udims[PDP_MAX_M]; /* u_grid_dim (discretizer.py:92)  */
    const double* x_level[PDP_MAX_N]; /* np.linspace levels verbatim (discretizer.py:142) */
    const double* u_level[PDP_MAX_M]; /* (discretizer.py:158) */
    double x_lb[PDP_MAX_N], x_ub[PDP_MAX_N]; /* sys.x_lb / x_ub used by isavalidstate (system.py:198-205) */
    double dt;    /* grid_sys.dt  */
    double alpha; /* dp.alpha (dynamicprogramming.py:130) */
    double INF;   /* cf.INF (costfunction.py:32) */
    double EPS;   /* cf.EPS (costfunction.py:33) */
    double Q[PDP_MAX_N * PDP_MAX_N]; /* row-major n x n stage weights  (costfunction.py:129) */
    double S[PDP_MAX_N * PDP_MAX_N]; /* row-major n x n terminal weights (costfunction.py:131) */
    double xbar[PDP_MAX_N];          /* cf.xbar */
    double sys_par[8];
    const double* sys_tab[4];
    int64_t sys_tab_len[4];
    const double* bu;      /* [A * n/2] */
    const double* gu;      /* [A] */
    const uint8_t* act_ok; /* [A] */
} pdp_problem;

typedef struct pdp_handle pdp_handle;

/* per-sweep convergence statistics, what finalize_backward_step prints/returns
 * (dynamicprogramming.py:247-261): max J, max(J - J_next), min(J - J_next), over this handle's slab */
typedef struct pdp_stats {
    double j_max;
    double delta_max;
    double delta_min;
} pdp_stats;

/* ---- lifecycle ------------------------------------------------------------------------- */
int pdp_abi_version(void);
int pdp_create(const pdp_problem* p, pdp_handle** out);
int pdp_destroy(pdp_handle* h);
const char* pdp_last_error(const pdp_handle* h);
/* Use an externally owned cudaStream_t for all work of this handle (default: a private stream). */
int pdp_set_stream(pdp_handle* h, void* cuda_stream);

/* ---- cost-to-go state ---------------------------------------------------------------------
 * J is the latest cost-to-go, pi the latest policy (int64), J_next the previous J
 * (dynamicprogramming.py:181-185).  The getters return THE HANDLE'S SLAB: (slab_end-slab_begin) *
 * prod(dims[1:]) values starting at plane slab_begin — on a single GPU that is the reference's
 * (N,) array.  pdp_set_J always takes the full (N,) array and keeps the planes the handle holds. */
/* replaces DynamicProgramming.evaluate_terminal_cost (dynamicprogramming.py:159-171): J = h(x, tf), pi = 0, on device */
int pdp_eval_terminal_cost(pdp_handle* h);
/* upload a full J (N doubles, host), e.g. terminal cost computed by the caller or load_J_next (:489-499) */
int pdp_set_J(pdp_handle* h, const double* J_host);
int pdp_get_J(pdp_handle* h, double* J_host);       /* slab doubles  */
int pdp_get_J_next(pdp_handle* h, double* J_host);  /* slab doubles  */
int pdp_get_pi(pdp_handle* h, int64_t* pi_host);    /* slab int64    */
/* `count` values starting at GLOBAL node id node_begin of J (which = 0), J_next (1; doubles; any node of the planes the
 * handle holds, halo included) or pi (2; int64; inside the handle's slab): J[s0:s1] / pi[s0:s1] of the reference's arrays without moving the whole grid —
 * what sampled parity checks and look-ups on 10^9-node grids need */
int pdp_get_range(pdp_handle* h, int32_t which, int64_t node_begin, int64_t count, void* out_host);

/* ---- the hot path ---------------------------------------------------------------------------
 * replaces initialize_backward_step + compute_backward_step + the reductions of
 * finalize_backward_step (dynamicprogramming.py:175-261), n_sweeps times back to back on the
 * device.  stats_out (host, may be NULL) receives n_sweeps entries.  Blocking. */
int pdp_sweep(pdp_handle* h, int32_t n_sweeps, pdp_stats* stats_out);

/* The same, split in two for callers that must not block between sweeps (benchmarks, pipelines):
 * pdp_sweep_enqueue queues one sweep (and, with a communicator attached, its halo exchange) on the
 * handle's stream; pdp_sweep_collect waits, all-reduces the statistics over the ranks and returns
 * the triples of the sweeps enqueued since the last collect (oldest first, at most max_out). */
int pdp_sweep_enqueue(pdp_handle* h);
int pdp_sweep_collect(pdp_handle* h, pdp_stats* stats_out, int32_t max_out, int32_t* n_out);

/* One sweep with HOST arrays on both sides — the reference's own calling convention, where J_next goes
 * in as a NumPy array and J, pi come out as NumPy arrays (dynamicprogramming.py:181-236): upload of
 * J_next (N doubles), backup, download of J (N doubles) and pi (N int64), pipelined over chunks of
 * axis-0 planes so that with pinned host buffers both PCIe directions overlap the kernels (and the whole
 * pipeline replays as one CUDA graph).  As with pdp_set_J / the getters, J_next_host is the FULL (N,) array
 * and J_host / pi_host receive the handle's slab; a slab handle uploads the planes it holds (slab + halo),
 * needs no exchange for this one sweep, and must refresh its halo (pdp_exchange_current or pdp_set_J) before
 * device-resident sweeps continue.  On a single GPU the handle is afterwards as after pdp_set_J + pdp_sweep(1). */
int pdp_sweep_host(pdp_handle* h, const double* J_next_host, double* J_host, int64_t* pi_host, pdp_stats* stats_out);

/* pdp_sweep_host for a rank of a sharded run: J_held_host holds ONLY the planes [alloc_begin, alloc_end) the handle
 * keeps (pdp_slab_layout), so no rank needs the full (N,) array in host memory */
int pdp_sweep_host_local(pdp_handle* h, const double* J_held_host, double* J_host, int64_t* pi_host, pdp_stats* stats_out);

/* LUT mode (system_id == PDP_SYS_LUT): the generic, bit-exact path for arbitrary user systems.
 * x_next: (N_slab, A, n) float64 as discretizer.py:349, G: (N_slab, A) float64 as
 * dynamicprogramming.py:523 (INF already folded in).  Uploaded once, then pdp_sweep() runs
 * dynamicprogramming.py:564-570 on the device. */
int pdp_set_lut(pdp_handle* h, const double* x_next_host, const double* G_host);

/* DynamicProgramming2DRectBivariateSpline (dynamicprogramming.py:578-614): the table sweep with J_next interpolated by the
 * interpolating bicubic spline scipy's RectBivariateSpline(x_level[0], x_level[1], J_grid, kx=3, ky=3) builds (FITPACK
 * regrid, not-a-knot knots; arguments outside the grid are clamped to its edge) instead of the RegularGridInterpolator.
 * Table-mode handles (PDP_SYS_LUT) of 2-D grids holding the whole grid; every pdp_sweep then refits the spline to J_next on
 * the device (two banded-substitution kernels) before the backup.  Floating-point parity with the reference (<= 1e-9). */
int pdp_set_interpolant(pdp_handle* h, int32_t which);

/* ---- step before the sweep, for callers that want the reference's dense tables (discretizer.py:342-376
 * compute_xnext_table, dynamicprogramming.py:517-553 compute_cost_lookuptable) of a fused system without the O(N*A)
 * Python loops: nodes [node_begin, node_begin+count), any output may be NULL.
 * x_next: (count, A, n) float64; x_next_isok: (count, A) uint8; G: (count, A) float64 = g*dt or INF */
int pdp_build_tables(pdp_handle* h, int64_t node_begin, int64_t count, double* x_next_host, uint8_t* x_next_isok_host, double* G_host);

/* ---- step after the sweep: policy -> input tables (discretizer.py:616-633 get_input_from_policy),
 * u_k[s] = input_from_action_id[pi[s], k], computed on the device, slab doubles to host */
int pdp_get_input_from_policy(pdp_handle* h, int32_t k, double* uk_host);
/* clean_infeasible_set (dynamicprogramming.py:322-334) on the device */
int pdp_clean_infeasible_set(pdp_handle* h, double tol, int64_t default_action);

/* ---- step after the sweep: batches of closed-loop Euler rollouts under the handle's current policy — what the
 * reference's examples run to validate a policy (`cl_sys = ctl + sys; cl_sys.compute_trajectory(tf, n, 'euler')`:
 * simulation.py:298-324 with ClosedLoopSystem.f controller.py:326-355 and LookUpTableController.c
 * dynamicprogramming.py:85-107), for B initial states at once, one device thread per trajectory:
 *     u[i] = RGI_linear(u_k tables of pi, fill 0 outside the grid)(x[i]);  x[i+1] = f(x[i], u[i]) * dt + x[i]
 * Fused systems only (the plant's f is evaluated on the device at arbitrary states: floating-point parity, not bit
 * parity), whole-grid handles only.  phys[16] = raw physical parameters of the plant:
 *   PENDULUM {m1, lc1, I1, gravity, d1}   TWOLINK {m1, l1, lc1, I1, m2, lc2, I2, gravity, d1, d2}   CARTPOLE {m1, m2, lcg, gravity}
 * x0_host (B, n).  Point i is kept when i % stride == 0: n_keep = (npts-1)/stride + 1.  Outputs are laid out
 * [n_keep][n][B] / [n_keep][m][B] (trajectory index fastest); u_out_host may be NULL. */
int pdp_rollout(pdp_handle* h, const double* phys, const double* x0_host, int64_t B, int32_t npts, double dt, int32_t stride,
                double* x_out_host, double* u_out_host);

/* ---- multi-GPU plumbing (one process per GPU; the host layer does the exchange) -------------
 * One asynchronous sweep of this handle's slab on its stream, no host sync.  The new J of the slab
 * is written into the "new" buffer; the caller then exchanges the halo planes of that buffer with
 * the neighbouring ranks (NCCL send/recv, or an all-gather of whole slabs) and calls
 * pdp_commit_sweep() to swap J/J_next.  pdp_sweep_planes_async does the same for a sub-range of the
 * slab's planes (boundary planes first, so their exchange overlaps the interior); each concurrent
 * range uses its own stat_set in 0..3 and writes its statistics triple to stats[3*stat_set]. */
/* Native exchange: attach an NCCL communicator (libnccl.so.2 through dlopen — the copy the process
 * already loaded, e.g. torch's) and pdp_sweep / pdp_sweep_enqueue run slab sweep + exchange +
 * statistics all-reduce themselves: boundary planes first, grouped ncclSend/ncclRecv with ranks
 * r-1 / r+1 on a side stream under the interior planes (exchange_mode 1), or an in-place
 * ncclAllGather of whole slabs (exchange_mode 2).  id128 = the 128-byte ncclUniqueId made by
 * pdp_nccl_unique_id on rank 0 and distributed by the caller. */
int pdp_nccl_unique_id(void* id128);
int pdp_comm_init(pdp_handle* h, int32_t rank, int32_t world, const void* id128, int32_t exchange_mode, int32_t overlap);
int pdp_exchange_current(pdp_handle* h);
/* Peer-memory halo exchange (upgrade of halo mode): every rank exports CUDA IPC handles of its two J buffers
 * and of a flag word pair (pdp_peer_export, 200 bytes), the caller hands each rank the exports of ranks r-1
 * and r+1 (NULL at the ends), and from then on a sweep stores the planes its neighbours read straight into
 * the neighbours' buffers over NVLink from a device kernel, publishes a sequence number with a system-scope
 * fence, and waits on the device for the neighbours' numbers before the next sweep: no NCCL call, no side
 * stream, no host step per sweep.  The communicator stays attached for the statistics all-reduce. */
int pdp_peer_export(pdp_handle* h, void* out200);
int pdp_peer_attach(pdp_handle* h, const void* lower200, const void* upper200);
/* Caller-driven exchange (any transport): */
int pdp_sweep_async(pdp_handle* h);
int pdp_sweep_planes_async(pdp_handle* h, int32_t plane_begin, int32_t plane_end, int32_t stat_set);
int pdp_commit_sweep(pdp_handle* h);
/* wait for the handle's stream and copy the four statistics triples (12 doubles) to the host */
int pdp_read_stats(pdp_handle* h, double* stats_host);
/* layout[8] = {slab_begin, slab_end, alloc_begin, alloc_end, halo_lo, halo_hi, dims[0], lanes_per_node}:
 * the J buffers hold planes [alloc_begin, alloc_end); a backup of plane i reads planes
 * [i - halo_lo, i + halo_hi] at most (computed from the levels, dt and bounds at pdp_create). */
int pdp_slab_layout(const pdp_handle* h, int32_t layout[8]);
/* the halo alone, from the descriptor (host arithmetic only, needs no device) */
int pdp_compute_halo(const pdp_problem* p, int32_t* halo_lo, int32_t* halo_hi);
/* device pointers: J (current) and J_new (being written), element 0 = first node of plane alloc_begin,
 * pdp_nodes_padded() doubles each; pi (slab int64, element 0 = first node of plane slab_begin);
 * stats (4 triples {j_max, delta_max, delta_min}, one per stat_set, of the last async launches) */
int pdp_device_buffers(pdp_handle* h, void** J_cur, void** J_new, void** pi, void** stats);
int64_t pdp_nodes(const pdp_handle* h);        /* N = prod(dims) */
int64_t pdp_nodes_padded(const pdp_handle* h); /* N_pad: allocation size of the J buffers */
int64_t pdp_actions(const pdp_handle* h);      /* A = prod(udims) */
/* number of sweep-kernel launches issued by this handle so far (bench.py gpu_launches) */
int64_t pdp_launch_count(const pdp_handle* h);
/* name of the sweep kernel this handle launches, e.g. "sweep_mech2_range_kernel<TWOLINK,direct> G=1" */
int pdp_kernel_info(const pdp_handle* h, char* out, int32_t len);
/* device time in ms of the last pdp_sweep() call, measured with CUDA events on the handle's stream */
double pdp_last_sweep_ms(const pdp_handle* h);

/* ---- single-process multi-GPU: one host thread drives n slab handles -------------------------------
 * pdp_multi_create cuts axis 0 into n_parts balanced slabs, part i on CUDA device devices[i] (NULL: round-robin over
 * all visible devices; a device may repeat — several slabs on one GPU).  p is the whole-grid descriptor (slab fields
 * ignored).  Per sweep every part computes its boundary planes first, stores them into its neighbours' halo planes
 * with peer copies (NVLink between GPUs) on a side stream, and computes its interior meanwhile; no NCCL, no
 * launcher.  The pdp_multi_* calls mirror their single-handle counterparts on the FULL (N,) arrays. */
typedef struct pdp_multi pdp_multi;
int pdp_device_count(void);
int pdp_multi_create(const pdp_problem* p, int32_t n_parts, const int32_t* devices, pdp_multi** out);
int pdp_multi_destroy(pdp_multi* m);
const char* pdp_multi_last_error(const pdp_multi* m);
int32_t pdp_multi_parts(const pdp_multi* m);
pdp_handle* pdp_multi_part(const pdp_multi* m, int32_t i);       /* the slab handle of part i (diagnostics, getters) */
int32_t pdp_multi_part_device(const pdp_multi* m, int32_t i);
int64_t pdp_multi_launch_count(const pdp_multi* m);
int pdp_multi_eval_terminal_cost(pdp_multi* m);
int pdp_multi_set_J(pdp_multi* m, const double* J_full);
int pdp_multi_get(pdp_multi* m, int32_t which, void* out_full);  /* which: 0 J, 1 J_next (doubles), 2 pi (int64) */
int pdp_multi_sweep(pdp_multi* m, int32_t n_sweeps, pdp_stats* stats_out);
int pdp_multi_sweep_enqueue(pdp_multi* m);
int pdp_multi_sweep_collect(pdp_multi* m, pdp_stats* stats_out, int32_t max_out, int32_t* n_out);
int pdp_multi_get_input_from_policy(pdp_multi* m, int32_t k, double* uk_full);
int pdp_multi_clean_infeasible_set(pdp_multi* m, double tol, int64_t default_action);

#ifdef __cplusplus
}
#endif
#endif /* PYRODP_H */
