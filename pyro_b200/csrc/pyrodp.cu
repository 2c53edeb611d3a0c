// pyrodp.cu — sm_100a Bellman-sweep kernels + the C ABI declared in include/pyrodp.h.
//
// Reference path being replaced (SherbyRobotics/pyro):
//   pyro/planning/dynamicprogramming.py:175-261  (initialize/compute/finalize_backward_step)
//   pyro/planning/dynamicprogramming.py:557-570  (LUT variant: RGI(x_next_table), G + alpha*J, min/argmin)
//   pyro/planning/discretizer.py:342-376         (x_next = f(x,u)*dt + x, isavalidstate)
//   scipy RegularGridInterpolator linear path     (_rgi.py:375-483, 520-549; _rgi_cython find_indices / evaluate_linear_2d)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/pyrodp.h"
#include "pyrodp_device.cuh"

// =================================================================================================
// Kernels
// =================================================================================================

#include "sweep_fused.cuh"

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
        c[d] = find_cell(P.level[d], P.dims[d], x[d], P.lb[d], P.inv_step[d]);
        const double lo = __ldg(P.level[d] + c[d]), hi = __ldg(P.level[d] + c[d] + 1);
        y[d] = (x[d] - lo) / (hi - lo);
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
    const int lane_in_group = threadIdx.x % G;
    const long long slot = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;  // node within slab
    const long long node = P.node_begin + slot;
    const int A = P.A;
    Stats3 st = stats_identity();
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
        Jo[node] = best;
        pi[node] = besta;
        const double d = best - Jn[node];
        st.jmax = best; st.dmax = d; st.dmin = d;
    }
    block_stats_finish(st, partials, counter, stats);
}

// ---- terminal cost (dynamicprogramming.py:159-171) -------------------------------------------------
template <int N>
__global__ void terminal_cost_kernel(const __grid_constant__ DevProblem P, double* __restrict__ J, long long* __restrict__ pi) {
    const long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= P.N) return;
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
    J[node] = h;
    if (node >= P.node_begin && node < P.node_end) pi[node] = 0;
}

__global__ void fill_pi_kernel(long long* pi, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pi[i] = 0;
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

// exposed for tests: exact_div against IEEE division on the device
__global__ void exact_div_test_kernel(const double* a, const double* den, double* q_fast, double* q_ieee, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double r = 1.0 / den[i];
        q_fast[i] = exact_div(a[i], den[i], r);
        q_ieee[i] = a[i] / den[i];
    }
}

// =================================================================================================
// Host side: handle + C ABI
// =================================================================================================

struct pdp_handle {
    DevProblem P{};
    int device = 0;
    long long N = 0, N_pad = 0, plane = 0;
    int A = 0;
    double* dJ[2] = {nullptr, nullptr};  // cur = dJ[cur_idx], new = dJ[1-cur_idx]
    int cur_idx = 0;
    long long* dpi = nullptr;
    double* dstats = nullptr;     // [stats_cap][3]
    int stats_cap = 0;
    unsigned long long* dpartials = nullptr;  // [STATS_SLOTS][3] order-preserving keys (block_stats_finish)
    unsigned int* dcounter = nullptr;
    std::vector<void*> owned;     // small device tables
    double* d_xnext = nullptr;
    double* d_G = nullptr;
    bool have_J = false, have_lut = false, pending = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long long launches = 0;
    double last_ms = 0.0;
    int sticky = 0;
    std::string err;
    size_t smem_bytes = 0;
    int lanes_per_node = 1;       // G of the fused kernels
    int force_lanes = 0;          // test hook (PYRODP_LANES): pin G to 1, 4 or 16
    void* fused = nullptr;        // selected fused kernel instantiation
    dim3 grid{1, 1, 1};
};

static thread_local std::string g_err;

static int fail(pdp_handle* h, int code, const std::string& msg) {
    if (h) {
        h->err = msg;
        if (code == PDP_ECUDA) h->sticky = code;
    }
    g_err = msg;
    return code;
}

#define CUDA_TRY(h, expr)                                                                       \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(h, PDP_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)

template <typename T>
static int upload(pdp_handle* h, const T* src, size_t count, const T** dst) {
    void* d = nullptr;
    CUDA_TRY(h, cudaMalloc(&d, (count ? count : 1) * sizeof(T)));
    h->owned.push_back(d);
    if (count) CUDA_TRY(h, cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T*)d;
    return PDP_OK;
}

extern "C" int pdp_abi_version(void) { return PDP_ABI_VERSION; }

extern "C" const char* pdp_last_error(const pdp_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

static int expected_tab_len(const pdp_problem* p, int t, long long* len) {
    const long long d0 = p->dims[0], d1 = p->dims[1];
    *len = 0;
    switch (p->system_id) {
        case PDP_SYS_PENDULUM: if (t == 0) *len = d0; break;
        case PDP_SYS_TWOLINK:
            if (t == 0) *len = d1 * 4; else if (t == 1) *len = d1; else if (t == 2) *len = d0 * d1 * 2;
            break;
        case PDP_SYS_CARTPOLE:
            if (t == 0) *len = d1 * 4; else if (t == 1) *len = d1; else if (t == 2) *len = d1;
            break;
        default: break;
    }
    return 0;
}

typedef void (*fused_kernel_t)(const DevProblem, const double*, double*, long long*, unsigned long long*, unsigned int*, double*);

template <int G, bool A1>
static fused_kernel_t fused_for(int system_id) {
    switch (system_id) {
        case PDP_SYS_PENDULUM: return sweep_pendulum_kernel<G, A1>;
        case PDP_SYS_TWOLINK: return sweep_mech2_kernel<PDP_SYS_TWOLINK, G, A1>;
        case PDP_SYS_CARTPOLE: return sweep_mech2_kernel<PDP_SYS_CARTPOLE, G, A1>;
    }
    return nullptr;
}

// G lanes per node: 1 when the slab alone fills the GPU, else 4 or 16 so that small grids still
// spread over the 148 SMs (the shuffle argmin keeps np.argmin's first-index rule).
static int select_fused_kernel(pdp_handle* h, const pdp_problem* p) {
    DevProblem& P = h->P;
    const long long slab_nodes = P.node_end - P.node_begin;
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
    const long long want_threads = (long long)sm_count * 2048;  // one full wave of resident threads
    int G = 1;
    if (slab_nodes * 1 < want_threads && P.A >= 4) G = 4;
    if (slab_nodes * 4 < want_threads && P.A >= 16) G = 16;
    if (h->force_lanes == 1 || h->force_lanes == 4 || h->force_lanes == 16) G = h->force_lanes;
    const bool a1 = P.alpha_is_one != 0;
    fused_kernel_t k = nullptr;
    if (G == 1) k = a1 ? fused_for<1, true>(P.system_id) : fused_for<1, false>(P.system_id);
    else if (G == 4) k = a1 ? fused_for<4, true>(P.system_id) : fused_for<4, false>(P.system_id);
    else k = a1 ? fused_for<16, true>(P.system_id) : fused_for<16, false>(P.system_id);
    if (!k) return fail(h, PDP_ENOTSUP, "no fused kernel for this system");
    h->lanes_per_node = G;
    h->fused = (void*)k;
    const size_t A = (size_t)P.A;
    if (P.system_id == PDP_SYS_PENDULUM) {
        const size_t n1p = (size_t)((P.dims[1] + 1) & ~1);
        h->smem_bytes = (2 * n1p + 2 * A) * sizeof(double) + 16;
        const long long threads = slab_nodes * G;
        const long long blocks = (threads + SWEEP_THREADS - 1) / SWEEP_THREADS;
        if (blocks > 0x7fffffffLL) return fail(h, PDP_ENOTSUP, "grid too large for one launch");
        h->grid = dim3((unsigned)(blocks > 0 ? blocks : 1), 1, 1);
    } else {
        const size_t n2p = (size_t)((P.dims[2] + 1) & ~1), n3p = (size_t)((P.dims[3] + 1) & ~1);
        h->smem_bytes = (2 * n2p + 2 * n3p + 4 * A) * sizeof(double) + 16;
        const long long plane_sz = (long long)P.dims[2] * P.dims[3];
        const long long chunks = (plane_sz * G + SWEEP_THREADS - 1) / SWEEP_THREADS;
        const long long planes = (long long)(p->slab_end - p->slab_begin) * P.dims[1];
        if (plane_sz > 0x7fffffffLL / 16 || chunks > 65535 || planes > 0x7fffffffLL)
            return fail(h, PDP_ENOTSUP, "grid too large for one launch (dims[2]*dims[3] too big)");
        h->grid = dim3((unsigned)(planes > 0 ? planes : 1), (unsigned)chunks, 1);
    }
    if (h->smem_bytes > 227 * 1024) return fail(h, PDP_ENOTSUP, "level/action tables exceed shared memory (227 KB)");
    cudaError_t ce = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (ce != cudaSuccess) return fail(h, PDP_ECUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ce));
    return PDP_OK;
}

extern "C" int pdp_create(const pdp_problem* p, pdp_handle** out) {
    if (!p || !out) return fail(nullptr, PDP_EINVAL, "pdp_create: null argument");
    *out = nullptr;
    if (p->abi_version != PDP_ABI_VERSION) return fail(nullptr, PDP_EINVAL, "pdp_create: ABI version mismatch");
    if (p->n < 2 || p->n > 4) return fail(nullptr, PDP_ENOTSUP, "pdp_create: state dimension must be 2, 3 or 4 (discretizer.py:243-245)");
    if (p->m < 1 || p->m > 2) return fail(nullptr, PDP_ENOTSUP, "pdp_create: input dimension must be 1 or 2 (discretizer.py:304-306)");
    switch (p->system_id) {
        case PDP_SYS_LUT: break;
        case PDP_SYS_PENDULUM: if (p->n != 2 || p->m != 1) return fail(nullptr, PDP_EINVAL, "PENDULUM needs n=2, m=1"); break;
        case PDP_SYS_TWOLINK: if (p->n != 4 || p->m != 2) return fail(nullptr, PDP_EINVAL, "TWOLINK needs n=4, m=2"); break;
        case PDP_SYS_CARTPOLE: if (p->n != 4 || p->m != 1) return fail(nullptr, PDP_EINVAL, "CARTPOLE needs n=4, m=1"); break;
        default: return fail(nullptr, PDP_ENOTSUP, "pdp_create: unknown system_id");
    }
    if (p->system_id != PDP_SYS_LUT && p->cost_id != PDP_COST_QUADRATIC && p->cost_id != PDP_COST_TIME)
        return fail(nullptr, PDP_ENOTSUP, "pdp_create: unknown cost_id");
    long long N = 1, A = 1;
    for (int d = 0; d < p->n; ++d) {
        if (p->dims[d] < 2) return fail(nullptr, PDP_EINVAL, "pdp_create: every state axis needs >= 2 levels");
        if (!p->x_level[d]) return fail(nullptr, PDP_EINVAL, "pdp_create: x_level pointer is null");
        N *= p->dims[d];
    }
    for (int d = 0; d < p->m; ++d) {
        if (p->udims[d] < 1) return fail(nullptr, PDP_EINVAL, "pdp_create: every input axis needs >= 1 level");
        if (!p->u_level[d]) return fail(nullptr, PDP_EINVAL, "pdp_create: u_level pointer is null");
        A *= p->udims[d];
    }
    if (A > 0x7fffffff) return fail(nullptr, PDP_EINVAL, "pdp_create: too many actions");
    if (p->slab_begin < 0 || p->slab_end > p->dims[0] || p->slab_begin > p->slab_end)
        return fail(nullptr, PDP_EINVAL, "pdp_create: slab range outside axis 0");
    if (p->system_id != PDP_SYS_LUT && (!p->bu || !p->gu || !p->act_ok))
        return fail(nullptr, PDP_EINVAL, "pdp_create: bu / gu / act_ok tables are required for fused systems");
    for (int t = 0; t < 4; ++t) {
        long long want;
        expected_tab_len(p, t, &want);
        if (want && (!p->sys_tab[t] || p->sys_tab_len[t] != want))
            return fail(nullptr, PDP_EINVAL, "pdp_create: sys_tab[" + std::to_string(t) + "] has wrong length");
    }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PDP_ECUDA, std::string("pdp_create: no usable CUDA device (") + cudaGetErrorString(e) +
                                            "); this engine has no CPU fallback");

    pdp_handle* h = new pdp_handle();
    auto bail = [&](int code) { pdp_destroy(h); return code; };
    if (cudaGetDevice(&h->device) != cudaSuccess) { g_err = "cudaGetDevice failed"; return bail(PDP_ECUDA); }

    DevProblem& P = h->P;
    P.n = p->n; P.m = p->m; P.dof = p->n / 2; P.A = (int)A;
    P.system_id = p->system_id; P.cost_id = p->cost_id; P.ontarget_check = p->ontarget_check;
    P.alpha_is_one = (p->alpha == 1.0);
    P.dt = p->dt; P.alpha = p->alpha; P.INF = p->INF; P.EPS = p->EPS;
    long long stride = 1;
    for (int d = p->n - 1; d >= 0; --d) { P.stride[d] = stride; stride *= p->dims[d]; }
    h->N = N; h->A = (int)A; P.N = N;
    h->plane = N / p->dims[0];
    // J buffers may be padded (alloc_planes) so an in-place equal-count all-gather fits
    long long pad_planes = p->alloc_planes > 0 ? p->alloc_planes : p->dims[0];
    if (pad_planes < p->dims[0]) { g_err = "pdp_create: alloc_planes < dims[0]"; return bail(PDP_EINVAL); }
    h->N_pad = pad_planes * h->plane;
    P.node_begin = (long long)p->slab_begin * h->plane;
    P.node_end = (long long)p->slab_end * h->plane;
    P.plane_begin = (long long)p->slab_begin * (p->n >= 2 ? p->dims[1] : 1);
    P.all_act_ok = 1;
    if (p->system_id != PDP_SYS_LUT)
        for (long long a = 0; a < A; ++a) if (!p->act_ok[a]) P.all_act_ok = 0;
    if (const char* env = getenv("PYRODP_LANES")) h->force_lanes = atoi(env);

    int rc;
    for (int d = 0; d < p->n; ++d) {
        P.dims[d] = p->dims[d];
        P.lb[d] = p->x_lb[d]; P.ub[d] = p->x_ub[d];
        P.inv_step[d] = (double)(p->dims[d] - 1) / (p->x_ub[d] - p->x_lb[d]);
        if (p->x_level[d][0] != p->x_lb[d] || p->x_level[d][p->dims[d] - 1] != p->x_ub[d]) {
            g_err = "pdp_create: x_level end points must equal x_lb/x_ub (np.linspace, discretizer.py:142)";
            return bail(PDP_EINVAL);
        }
        std::vector<double> rinv(p->dims[d]);
        for (int i = 0; i + 1 < p->dims[d]; ++i) {
            const double den = p->x_level[d][i + 1] - p->x_level[d][i];
            if (!(den > 0.0)) { g_err = "pdp_create: x_level must be strictly increasing"; return bail(PDP_EINVAL); }
            rinv[i] = 1.0 / den;
        }
        rinv[p->dims[d] - 1] = 0.0;
        if ((rc = upload(h, p->x_level[d], p->dims[d], &P.level[d])) != PDP_OK) return bail(rc);
        if ((rc = upload(h, rinv.data(), rinv.size(), &P.rinv[d])) != PDP_OK) return bail(rc);
    }
    memcpy(P.Q, p->Q, sizeof(P.Q)); memcpy(P.S, p->S, sizeof(P.S));
    memcpy(P.xbar, p->xbar, sizeof(P.xbar)); memcpy(P.par, p->sys_par, sizeof(P.par));
    for (int t = 0; t < 4; ++t) {
        long long want; expected_tab_len(p, t, &want);
        if (want && (rc = upload(h, p->sys_tab[t], (size_t)want, &P.tab[t])) != PDP_OK) return bail(rc);
    }
    // input_from_action_id (discretizer.py:263-302), C order of u_grid_dim
    {
        std::vector<double> u_flat((size_t)A * p->m);
        for (long long a = 0; a < A; ++a) {
            if (p->m == 1) u_flat[a] = p->u_level[0][a];
            else { u_flat[2 * a] = p->u_level[0][a / p->udims[1]]; u_flat[2 * a + 1] = p->u_level[1][a % p->udims[1]]; }
        }
        if ((rc = upload(h, u_flat.data(), u_flat.size(), &P.u_flat)) != PDP_OK) return bail(rc);
    }
    if (p->system_id != PDP_SYS_LUT) {
        // isavalidinput (system.py:208-215) is folded into the B.u table: a disallowed action carries
        // NaN, its x_next is NaN, fails the box test and gets Q = INF exactly as dynamicprogramming.py:233
        std::vector<double> bu(p->bu, p->bu + (size_t)A * P.dof);
        for (long long a = 0; a < A; ++a)
            if (!p->act_ok[a]) for (int d = 0; d < P.dof; ++d) bu[(size_t)a * P.dof + d] = __builtin_nan("");
        if ((rc = upload(h, bu.data(), bu.size(), &P.bu)) != PDP_OK) return bail(rc);
        if ((rc = upload(h, p->gu, (size_t)A, &P.gu)) != PDP_OK) return bail(rc);
        if ((rc = upload(h, (const unsigned char*)p->act_ok, (size_t)A, &P.act_ok)) != PDP_OK) return bail(rc);
    }

    auto cu = [&](cudaError_t ce, const char* what) -> bool {
        if (ce != cudaSuccess) { g_err = std::string(what) + ": " + cudaGetErrorString(ce); return false; }
        return true;
    };
    if (!cu(cudaMalloc(&h->dJ[0], h->N_pad * sizeof(double)), "cudaMalloc J0")) return bail(PDP_ECUDA);
    if (!cu(cudaMalloc(&h->dJ[1], h->N_pad * sizeof(double)), "cudaMalloc J1")) return bail(PDP_ECUDA);
    if (!cu(cudaMalloc(&h->dpi, h->N * sizeof(long long)), "cudaMalloc pi")) return bail(PDP_ECUDA);
    h->stats_cap = 256;
    if (!cu(cudaMalloc(&h->dstats, h->stats_cap * 3 * sizeof(double)), "cudaMalloc stats")) return bail(PDP_ECUDA);
    if (!cu(cudaMalloc(&h->dpartials, 3 * STATS_SLOTS * sizeof(unsigned long long)), "cudaMalloc stats slots")) return bail(PDP_ECUDA);
    if (!cu(cudaMemset(h->dpartials, 0, 3 * STATS_SLOTS * sizeof(unsigned long long)), "cudaMemset stats slots")) return bail(PDP_ECUDA);
    if (!cu(cudaMalloc(&h->dcounter, sizeof(unsigned int)), "cudaMalloc counter")) return bail(PDP_ECUDA);
    if (!cu(cudaMemset(h->dcounter, 0, sizeof(unsigned int)), "cudaMemset counter")) return bail(PDP_ECUDA);
    if (!cu(cudaMemset(h->dpi, 0, h->N * sizeof(long long)), "cudaMemset pi")) return bail(PDP_ECUDA);
    if (!cu(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return bail(PDP_ECUDA);
    h->own_stream = true;
    if (!cu(cudaEventCreate(&h->ev0), "cudaEventCreate")) return bail(PDP_ECUDA);
    if (!cu(cudaEventCreate(&h->ev1), "cudaEventCreate")) return bail(PDP_ECUDA);

    // fused kernels: lanes per node, grid shape, kernel instantiation, dynamic shared memory
    if (p->system_id != PDP_SYS_LUT) {
        int rcsel = select_fused_kernel(h, p);
        if (rcsel != PDP_OK) return bail(rcsel);
    }
    *out = h;
    return PDP_OK;
}

extern "C" int pdp_destroy(pdp_handle* h) {
    if (!h) return PDP_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void* p : h->owned) cudaFree(p);
    cudaFree(h->dJ[0]); cudaFree(h->dJ[1]); cudaFree(h->dpi); cudaFree(h->dstats);
    cudaFree(h->dpartials); cudaFree(h->dcounter); cudaFree(h->d_xnext); cudaFree(h->d_G);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PDP_OK;
}

#define CHECK_HANDLE(h)                                                          \
    do {                                                                         \
        if (!(h)) return fail(nullptr, PDP_EINVAL, "null handle");               \
        if ((h)->sticky) return (h)->sticky;                                     \
        CUDA_TRY(h, cudaSetDevice((h)->device));                                 \
    } while (0)

extern "C" int pdp_set_stream(pdp_handle* h, void* cuda_stream) {
    CHECK_HANDLE(h);
    if (h->own_stream && h->stream) {
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        CUDA_TRY(h, cudaStreamDestroy(h->stream));
    }
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    return PDP_OK;
}

extern "C" int64_t pdp_nodes(const pdp_handle* h) { return h ? h->N : 0; }
extern "C" int64_t pdp_nodes_padded(const pdp_handle* h) { return h ? h->N_pad : 0; }
extern "C" int64_t pdp_actions(const pdp_handle* h) { return h ? h->A : 0; }
extern "C" int64_t pdp_launch_count(const pdp_handle* h) { return h ? h->launches : 0; }
extern "C" double pdp_last_sweep_ms(const pdp_handle* h) { return h ? h->last_ms : 0.0; }

extern "C" int pdp_eval_terminal_cost(pdp_handle* h) {
    CHECK_HANDLE(h);
    if (h->P.system_id == PDP_SYS_LUT) return fail(h, PDP_ENOTSUP, "pdp_eval_terminal_cost: LUT mode has no cost model; use pdp_set_J");
    const int threads = 256;
    const long long blocks = (h->N + threads - 1) / threads;
    double* J = h->dJ[h->cur_idx];
    if (h->P.n == 2) terminal_cost_kernel<2><<<(unsigned)blocks, threads, 0, h->stream>>>(h->P, J, h->dpi);
    else if (h->P.n == 4) terminal_cost_kernel<4><<<(unsigned)blocks, threads, 0, h->stream>>>(h->P, J, h->dpi);
    else return fail(h, PDP_ENOTSUP, "pdp_eval_terminal_cost: n must be 2 or 4");
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_J = true;
    return PDP_OK;
}

extern "C" int pdp_set_J(pdp_handle* h, const double* J_host) {
    CHECK_HANDLE(h);
    if (!J_host) return fail(h, PDP_EINVAL, "pdp_set_J: null pointer");
    CUDA_TRY(h, cudaMemcpyAsync(h->dJ[h->cur_idx], J_host, h->N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_J = true;
    return PDP_OK;
}

static int copy_out(pdp_handle* h, void* dst, const void* src, size_t bytes) {
    CUDA_TRY(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PDP_OK;
}

extern "C" int pdp_get_J(pdp_handle* h, double* J_host) {
    CHECK_HANDLE(h);
    if (!J_host) return fail(h, PDP_EINVAL, "pdp_get_J: null pointer");
    if (!h->have_J) return fail(h, PDP_ESTATE, "pdp_get_J: no cost-to-go yet");
    return copy_out(h, J_host, h->dJ[h->cur_idx], h->N * sizeof(double));
}
extern "C" int pdp_get_J_next(pdp_handle* h, double* J_host) {
    CHECK_HANDLE(h);
    if (!J_host) return fail(h, PDP_EINVAL, "pdp_get_J_next: null pointer");
    if (!h->have_J) return fail(h, PDP_ESTATE, "pdp_get_J_next: no cost-to-go yet");
    return copy_out(h, J_host, h->dJ[1 - h->cur_idx], h->N * sizeof(double));
}
extern "C" int pdp_get_pi(pdp_handle* h, int64_t* pi_host) {
    CHECK_HANDLE(h);
    if (!pi_host) return fail(h, PDP_EINVAL, "pdp_get_pi: null pointer");
    return copy_out(h, pi_host, h->dpi, h->N * sizeof(long long));
}

extern "C" int pdp_set_lut(pdp_handle* h, const double* x_next_host, const double* G_host) {
    CHECK_HANDLE(h);
    if (h->P.system_id != PDP_SYS_LUT) return fail(h, PDP_ESTATE, "pdp_set_lut: handle was not created with PDP_SYS_LUT");
    if (!x_next_host || !G_host) return fail(h, PDP_EINVAL, "pdp_set_lut: null pointer");
    const size_t slab = (size_t)(h->P.node_end - h->P.node_begin);
    const size_t nx = slab * h->A * h->P.n, ng = slab * h->A;
    if (!h->d_xnext) CUDA_TRY(h, cudaMalloc(&h->d_xnext, (nx ? nx : 1) * sizeof(double)));
    if (!h->d_G) CUDA_TRY(h, cudaMalloc(&h->d_G, (ng ? ng : 1) * sizeof(double)));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_xnext, x_next_host, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_G, G_host, ng * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_lut = true;
    return PDP_OK;
}

template <int N>
static void launch_lut(pdp_handle* h, int G, unsigned blocks, const double* Jn, double* Jo, double* stats) {
#define LUT_CASE(g)                                                                                                  \
    case g:                                                                                                          \
        sweep_lut_kernel<N, g><<<blocks, SWEEP_THREADS, 0, h->stream>>>(h->P, Jn, Jo, h->dpi, h->d_xnext, h->d_G,  \
                                                                        h->dpartials, h->dcounter, stats);          \
        break;
    switch (G) {
        LUT_CASE(1) LUT_CASE(2) LUT_CASE(4) LUT_CASE(8) LUT_CASE(16) LUT_CASE(32)
    }
#undef LUT_CASE
}

// one sweep on the stream: reads dJ[cur], writes dJ[1-cur] (slab only), pi (slab), stats[3]
static int launch_sweep(pdp_handle* h, double* stats) {
    const DevProblem& P = h->P;
    const long long slab_nodes = P.node_end - P.node_begin;
    const double* Jn = h->dJ[h->cur_idx];
    double* Jo = h->dJ[1 - h->cur_idx];
    if (slab_nodes <= 0) {
        const double ident[3] = {-__builtin_inf(), -__builtin_inf(), __builtin_inf()};
        CUDA_TRY(h, cudaMemcpyAsync(stats, ident, sizeof(ident), cudaMemcpyHostToDevice, h->stream));
        return PDP_OK;
    }
    if (P.system_id == PDP_SYS_LUT) {
        if (!h->have_lut) return fail(h, PDP_ESTATE, "pdp_sweep: LUT mode needs pdp_set_lut first");
        int G = 1;
        while (G < 32 && G < P.A) G <<= 1;
        const long long threads = slab_nodes * G;
        const unsigned blocks = (unsigned)((threads + SWEEP_THREADS - 1) / SWEEP_THREADS);
        if (P.n == 2) launch_lut<2>(h, G, blocks, Jn, Jo, stats);
        else if (P.n == 3) launch_lut<3>(h, G, blocks, Jn, Jo, stats);
        else launch_lut<4>(h, G, blocks, Jn, Jo, stats);
    } else {
        ((fused_kernel_t)h->fused)<<<h->grid, SWEEP_THREADS, h->smem_bytes, h->stream>>>(P, Jn, Jo, h->dpi, h->dpartials, h->dcounter, stats);
    }
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return PDP_OK;
}

extern "C" int pdp_sweep_async(pdp_handle* h) {
    CHECK_HANDLE(h);
    if (!h->have_J) return fail(h, PDP_ESTATE, "pdp_sweep: no cost-to-go yet (pdp_set_J / pdp_eval_terminal_cost)");
    if (h->pending) return fail(h, PDP_ESTATE, "pdp_sweep_async: previous sweep not committed");
    int rc = launch_sweep(h, h->dstats);
    if (rc != PDP_OK) return rc;
    h->pending = true;
    return PDP_OK;
}

extern "C" int pdp_commit_sweep(pdp_handle* h) {
    CHECK_HANDLE(h);
    if (!h->pending) return fail(h, PDP_ESTATE, "pdp_commit_sweep: nothing pending");
    h->cur_idx = 1 - h->cur_idx;
    h->pending = false;
    return PDP_OK;
}

extern "C" int pdp_device_buffers(pdp_handle* h, void** J_cur, void** J_new, void** pi, void** stats) {
    CHECK_HANDLE(h);
    if (J_cur) *J_cur = h->dJ[h->cur_idx];
    if (J_new) *J_new = h->dJ[1 - h->cur_idx];
    if (pi) *pi = h->dpi;
    if (stats) *stats = h->dstats;
    return PDP_OK;
}

extern "C" int pdp_sweep(pdp_handle* h, int32_t n_sweeps, pdp_stats* stats_out) {
    CHECK_HANDLE(h);
    if (n_sweeps < 0) return fail(h, PDP_EINVAL, "pdp_sweep: n_sweeps < 0");
    if (!h->have_J) return fail(h, PDP_ESTATE, "pdp_sweep: no cost-to-go yet (pdp_set_J / pdp_eval_terminal_cost)");
    if (h->pending) return fail(h, PDP_ESTATE, "pdp_sweep: an async sweep is pending");
    if (h->P.node_end - h->P.node_begin != h->N)
        return fail(h, PDP_ESTATE, "pdp_sweep: handle owns a slab only; drive it with pdp_sweep_async + exchange + pdp_commit_sweep");
    if (n_sweeps > h->stats_cap) {
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        CUDA_TRY(h, cudaFree(h->dstats));
        h->dstats = nullptr;
        h->stats_cap = n_sweeps;
        CUDA_TRY(h, cudaMalloc(&h->dstats, (size_t)h->stats_cap * 3 * sizeof(double)));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    for (int k = 0; k < n_sweeps; ++k) {
        int rc = launch_sweep(h, h->dstats + 3 * k);
        if (rc != PDP_OK) return rc;
        h->cur_idx = 1 - h->cur_idx;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    if (stats_out && n_sweeps > 0)
        CUDA_TRY(h, cudaMemcpyAsync(stats_out, h->dstats, (size_t)n_sweeps * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    return PDP_OK;
}

extern "C" int pdp_get_input_from_policy(pdp_handle* h, int32_t k, double* uk_host) {
    CHECK_HANDLE(h);
    if (k < 0 || k >= h->P.m) return fail(h, PDP_EINVAL, "pdp_get_input_from_policy: input axis out of range");
    if (!uk_host) return fail(h, PDP_EINVAL, "pdp_get_input_from_policy: null pointer");
    double* scratch = h->dJ[1 - h->cur_idx];  // J_next is scratch only until the next sweep overwrites it anyway
    double* tmp = nullptr;
    CUDA_TRY(h, cudaMalloc(&tmp, h->N * sizeof(double)));
    (void)scratch;
    const int threads = 256;
    input_from_policy_kernel<<<(unsigned)((h->N + threads - 1) / threads), threads, 0, h->stream>>>(h->dpi, h->P.u_flat, h->P.m, k, tmp, h->N);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(uk_host, tmp, h->N * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(h, PDP_ECUDA, std::string("pdp_get_input_from_policy: ") + cudaGetErrorString(e));
    return PDP_OK;
}

extern "C" int pdp_clean_infeasible_set(pdp_handle* h, double tol, int64_t default_action) {
    CHECK_HANDLE(h);
    if (default_action < 0 || default_action >= h->A) return fail(h, PDP_EINVAL, "pdp_clean_infeasible_set: default action out of range");
    const int threads = 256;
    clean_infeasible_kernel<<<(unsigned)((h->N + threads - 1) / threads), threads, 0, h->stream>>>(
        h->dJ[h->cur_idx], h->dpi, h->P.INF - tol, h->P.INF, (long long)default_action, h->N);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PDP_OK;
}

// test hook (not part of the reference-facing ABI): exact_div vs IEEE division on the device
extern "C" int pdp_test_exact_div(const double* a, const double* den, double* q_fast, double* q_ieee, int64_t n) {
    double *da, *dd, *df, *di;
    size_t bytes = (size_t)n * sizeof(double);
    if (cudaMalloc(&da, bytes) != cudaSuccess) return fail(nullptr, PDP_ECUDA, "pdp_test_exact_div: cudaMalloc failed");
    cudaMalloc(&dd, bytes); cudaMalloc(&df, bytes); cudaMalloc(&di, bytes);
    cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(dd, den, bytes, cudaMemcpyHostToDevice);
    exact_div_test_kernel<<<(unsigned)((n + 255) / 256), 256>>>(da, dd, df, di, n);
    cudaMemcpy(q_fast, df, bytes, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaMemcpy(q_ieee, di, bytes, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(dd); cudaFree(df); cudaFree(di);
    return e == cudaSuccess ? PDP_OK : fail(nullptr, PDP_ECUDA, cudaGetErrorString(e));
}
