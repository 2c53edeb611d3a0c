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
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/pyrodp.h"
#include "pyrodp_device.cuh"

// =================================================================================================
// Kernels
// =================================================================================================

#include "sweep_fused.cuh"
#include "sweep_mech2.cuh"
#include "mech2_plan.h"

#include "table_kernels.cuh"
#include "rollout.cuh"
#include "spline.cuh"

// exposed for tests: exact_div against IEEE division on the device
__global__ void exact_div_test_kernel(const double* a, const double* den, double* q_fast, double* q_ieee, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double r = 1.0 / den[i];
        q_fast[i] = exact_div(a[i], den[i], r);
        q_ieee[i] = a[i] / den[i];
    }
}

// ---- statistics plumbing for range launches / ranks ------------------------------------------------
// fold `nsets` triples {jmax, dmax, dmin} into dst = {jmax, dmax, -dmin} (all-reduce with MAX), and back
__global__ void stats_fold_kernel(const double* __restrict__ sets, int nsets, double* __restrict__ dst) {
    double a = sets[0], b = sets[1], c = sets[2];
    for (int i = 1; i < nsets; ++i) {
        a = fmax(a, sets[3 * i]); b = fmax(b, sets[3 * i + 1]); c = fmin(c, sets[3 * i + 2]);
    }
    dst[0] = a; dst[1] = b; dst[2] = -c;
}
// two triples that are not adjacent (interior + upper boundary of the lowest part)
__global__ void stats_fold2_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ dst) {
    dst[0] = fmax(a[0], b[0]); dst[1] = fmax(a[1], b[1]); dst[2] = -fmin(a[2], b[2]);
}
__global__ void stats_unfold_kernel(double* __restrict__ st, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st[3 * i + 2] = -st[3 * i + 2];
}

// ---- peer-memory halo exchange (one process per GPU, buffers opened through CUDA IPC) -----------------
// After a rank's slab planes of the new J are written, the planes its neighbours read as their halo are
// stored straight into the neighbours' own J buffers over NVLink (peer pointers), followed by a
// system-scope fence and one sequence-number flag per neighbour.  Before its next sweep a rank waits
// (on the device) until both neighbours' flags carry the previous exchange's number.  No NCCL call,
// no side stream, no host involvement per sweep.
__global__ void halo_push_kernel(const double* __restrict__ src_lo, double* __restrict__ dst_lo, long long n_lo,
                                 const double* __restrict__ src_hi, double* __restrict__ dst_hi, long long n_hi,
                                 unsigned int* ticket, unsigned int* peer_flag_lo, unsigned int* peer_flag_hi, unsigned int seq) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_lo; i += stride) dst_lo[i] = src_lo[i];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_hi; i += stride) dst_hi[i] = src_hi[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {   // every block's stores are fenced: publish
            __threadfence_system();
            if (peer_flag_lo) *(volatile unsigned int*)peer_flag_lo = seq;
            if (peer_flag_hi) *(volatile unsigned int*)peer_flag_hi = seq;
            __threadfence_system();
            *ticket = 0u;
        }
    }
}
// flags[0] is written by the rank below, flags[1] by the rank above.  Bounded spin: a peer that never
// arrives raises *err instead of hanging the device.
__global__ void halo_wait_kernel(const unsigned int* flags, unsigned int want, int has_lo, int has_hi, unsigned int* err,
                                 long long timeout_cycles) {
    const volatile unsigned int* f = flags;
    const long long t0 = clock64();
    while ((has_lo && (int)(f[0] - want) < 0) || (has_hi && (int)(f[1] - want) < 0)) {
        if (clock64() - t0 > timeout_cycles) { *err = 1u; break; }
        __nanosleep(100);
    }
    __threadfence_system();
}

// test hook: occupy the stream for about `cycles` SM clocks
__global__ void delay_kernel(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) __nanosleep(200);
}

// =================================================================================================
// Host side: handle + C ABI
// =================================================================================================

#define PDP_STAT_SETS 4  // independent stats scratch sets, so range launches may overlap on different streams

struct pdp_handle {
    DevProblem P{};
    int device = 0;
    long long N = 0, plane = 0;   // nodes of the whole grid, nodes per axis-0 plane
    int A = 0;
    int n0 = 0;                   // dims[0]
    int slab_begin = 0, slab_end = 0;    // planes computed by this handle
    int alloc_begin = 0, alloc_end = 0;  // planes held in the J buffers (slab + halo, or everything)
    int halo_lo = 0, halo_hi = 0;        // planes below / above a node that its backup can read
    long long alloc_planes_cap = 0;      // planes allocated (>= alloc_end - alloc_begin; padded all-gather layout)
    double* dJ[2] = {nullptr, nullptr};  // cur = dJ[cur_idx], new = dJ[1-cur_idx]; element 0 = plane alloc_begin
    int cur_idx = 0;
    long long* dpi = nullptr;     // slab planes only; element 0 = plane slab_begin
    double* dstats = nullptr;     // [stats_cap][3] history of enqueued sweeps (pdp_sweep / pdp_sweep_enqueue)
    int stats_cap = 0;
    int enqueued = 0;             // sweeps enqueued since the last collect
    double* dstats_sets = nullptr;  // [PDP_STAT_SETS][3] triples of the range launches
    unsigned long long* dslots = nullptr;  // [PDP_STAT_SETS][STATS_SLOTS][3] order-preserving keys
    unsigned int* dcounter = nullptr;      // [PDP_STAT_SETS]
    std::vector<void*> owned;     // small device tables
    double* d_xnext = nullptr;
    double* d_G = nullptr;
    bool have_J = false, have_lut = false, pending = false;
    bool halo_stale = false;      // pdp_sweep_host on a slab handle: halo planes of J need pdp_exchange_current / pdp_set_J
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
    int policy_blocks = 0;        // one resident wave of sweep_policy_kernel blocks
    bool pend_mono = false;       // pendulum: x_next[1] is non-decreasing along the action list (see sweep_fused.cuh, MONO)
    bool spline = false;          // LUT mode, n = 2: bicubic-spline interpolant of J_next (pdp_set_interpolant, spline.cuh)
    SplineDev S{};
    bool force_generic = false;   // test hook (PYRODP_GENERIC=1): use the order-agnostic action loop anyway
    int pend_loop = 1;            // MONO variant of the pendulum kernel: 1 = pair loop (shipped), 2 = loop nest (PYRODP_PEND_LOOP=2; 11 % fewer
                                  // non-FP64 instructions, same sweep time: profiles/r02p_pendulum_loops_ab.jsonl)
    int mech2_mode = 0;           // 4-D fused systems: 0 order-agnostic kernel, 1 range-skipping kernel
    int force_mech2 = -1;         // test / A-B hook (PYRODP_MECH2=generic|range)
    int test_interior_delay_us = 0;   // test hook (PYRODP_TEST_INTERIOR_DELAY_US): spin before the interior planes of a sharded sweep
    Mech2Plan plan{};             // action-table structure the range kernel relies on
    void* fused = nullptr;        // selected fused kernel instantiation
    // multi-GPU (one process per GPU): NCCL communicator, side stream for the halo exchange
    void* comm = nullptr;
    int rank = 0, world = 1;
    int exchange_mode = 0;        // 0 none, 1 halo send/recv with ranks r-1 / r+1, 2 in-place all-gather of whole slabs
    int overlap = 1;              // boundary planes first, exchange under the interior planes
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_boundary = nullptr, ev_comm = nullptr;
    long long exchanges = 0;
    // peer-memory halo exchange (exchange_mode 3)
    unsigned int* dflags = nullptr;       // [0] written by the rank below, [1] by the rank above, [2] push ticket, [3] wait error
    double* peer_J[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [below/above][buffer index], opened through CUDA IPC
    unsigned int* peer_flags[2] = {nullptr, nullptr};
    int peer_alloc_begin[2] = {0, 0};
    unsigned int peer_seq = 0;            // number of exchanges pushed so far
    // pdp_sweep_host: copy streams, per-chunk events and statistics
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaStream_t side_stream[2] = {nullptr, nullptr};  // chunk kernels rotate over {stream, side_stream[0], side_stream[1]}
    cudaEvent_t ev_side[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_up, ev_done;
    cudaEvent_t ev_start = nullptr;
    cudaEvent_t ev_join[2] = {nullptr, nullptr};
    double* dchunk_stats = nullptr;
    double* h_chunk_stats = nullptr;     // pinned host copy of the per-chunk statistics
    struct HostGraph { const void* jin; void* jout; void* piout; int cur_idx; int chunks; cudaGraphExec_t exec; };
    std::vector<HostGraph> host_graphs;  // captured pdp_sweep_host pipelines, keyed by buffers, J parity and chunking

    long long slab_nodes() const { return (long long)(slab_end - slab_begin) * plane; }
    long long alloc_nodes() const { return (long long)(alloc_end - alloc_begin) * plane; }
    // virtual bases: index with the global node id
    double* Jv(int which) const { return dJ[which] - (long long)alloc_begin * plane; }
    long long* piv() const { return dpi - (long long)slab_begin * plane; }
};

static thread_local std::string g_err;

// NCCL entry points, resolved at run time (see nccl_load)
typedef struct { char internal[128]; } pdp_nccl_id;
static struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(pdp_nccl_id*) = nullptr;
    int (*CommInitRank)(void**, int, pdp_nccl_id, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
} g_nccl;
enum { PDP_NCCL_F64 = 8, PDP_NCCL_MAX = 2 };  // ncclFloat64, ncclMax (nccl.h)


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
static fused_kernel_t fused_for(int system_id, bool nodamp, int mono) {
    switch (system_id) {
        case PDP_SYS_PENDULUM:
            if (mono == 2) return nodamp ? sweep_pendulum_kernel<G, A1, true, 2> : sweep_pendulum_kernel<G, A1, false, 2>;
            if (mono == 1) return nodamp ? sweep_pendulum_kernel<G, A1, true, 1> : sweep_pendulum_kernel<G, A1, false, 1>;
            return nodamp ? sweep_pendulum_kernel<G, A1, true, 0> : sweep_pendulum_kernel<G, A1, false, 0>;
        case PDP_SYS_TWOLINK: return sweep_mech2_kernel<PDP_SYS_TWOLINK, G, A1>;
        case PDP_SYS_CARTPOLE: return sweep_mech2_kernel<PDP_SYS_CARTPOLE, G, A1>;
    }
    return nullptr;
}

// G lanes per node: 1 when the slab alone fills the GPU, else 4 or 16 so that small grids still
// spread over the 148 SMs (the shuffle argmin keeps np.argmin's first-index rule).
static int select_fused_kernel(pdp_handle* h) {
    DevProblem& P = h->P;
    const long long slab_nodes = h->slab_nodes();
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
    const long long want_threads = (long long)sm_count * 2048;  // one full wave of resident threads
    int G = 1;
    if (slab_nodes * 1 < want_threads && P.A >= 4) G = 4;
    if (slab_nodes * 4 < want_threads && P.A >= 16) G = 16;
    if (h->force_lanes == 1 || h->force_lanes == 4 || h->force_lanes == 16) G = h->force_lanes;
    const bool a1 = P.alpha_is_one != 0;
    fused_kernel_t k = nullptr;
    const bool nd = (P.system_id == PDP_SYS_PENDULUM) && P.par[1] == 0.0;  // d1 == 0: no damping term
    const int mono = (h->pend_mono && !h->force_generic) ? h->pend_loop : 0;
    if (G == 1) k = a1 ? fused_for<1, true>(P.system_id, nd, mono) : fused_for<1, false>(P.system_id, nd, mono);
    else if (G == 4) k = a1 ? fused_for<4, true>(P.system_id, nd, mono) : fused_for<4, false>(P.system_id, nd, mono);
    else k = a1 ? fused_for<16, true>(P.system_id, nd, mono) : fused_for<16, false>(P.system_id, nd, mono);
    if (!k) return fail(h, PDP_ENOTSUP, "no fused kernel for this system");
    const size_t A = (size_t)P.A;
    // 4-D systems, one lane per node: the range-skipping kernel when the action table has the structure it relies on
    h->mech2_mode = 0;
    if (P.system_id != PDP_SYS_PENDULUM && G == 1 && h->plan.ok && !h->force_generic && h->force_mech2 != 0) {
        const size_t n2p = (size_t)((P.dims[2] + 1) & ~1), n3p = (size_t)((P.dims[3] + 1) & ~1);
        const size_t smem = (2 * n2p + 2 * n3p + (size_t)((h->plan.A0 + 1) & ~1) + (size_t)((h->plan.A1 + 1) & ~1) + A) * sizeof(double) + 16;
        const bool offsets_fit = 2LL * P.dims[1] * P.dims[2] * P.dims[3] < 0x7fffffffLL;   // 32-bit corner offsets of the range kernel
        if (smem <= 200 * 1024 && offsets_fit) {
            if (P.system_id == PDP_SYS_TWOLINK) k = a1 ? sweep_mech2_range_kernel<PDP_SYS_TWOLINK, true> : sweep_mech2_range_kernel<PDP_SYS_TWOLINK, false>;
            else k = a1 ? sweep_mech2_range_kernel<PDP_SYS_CARTPOLE, true> : sweep_mech2_range_kernel<PDP_SYS_CARTPOLE, false>;
            h->mech2_mode = 1;
            h->smem_bytes = smem;
        }
    }
    h->lanes_per_node = G;
    h->fused = (void*)k;
    if (P.system_id == PDP_SYS_PENDULUM) {
        const size_t n1p = (size_t)((P.dims[1] + 1) & ~1);
        h->smem_bytes = (2 * n1p + 2 * (A + 4 * (size_t)G)) * sizeof(double) + 16;  // {level, 1/step} records + padded action records
        if (h->N > 0x7fffffffLL) return fail(h, PDP_ENOTSUP, "2-D grids are limited to 2^31-1 nodes");
        if (((long long)P.dims[1] * G + SWEEP_THREADS - 1) / SWEEP_THREADS > 65535)
            return fail(h, PDP_ENOTSUP, "grid too large for one launch (dims[1] too big)");
    } else {
        const size_t n2p = (size_t)((P.dims[2] + 1) & ~1), n3p = (size_t)((P.dims[3] + 1) & ~1);
        if (!h->mech2_mode) h->smem_bytes = (2 * n2p + 2 * n3p + 4 * A) * sizeof(double) + 16;
        const long long plane_sz = (long long)P.dims[2] * P.dims[3];
        const long long chunks = (plane_sz * G + SWEEP_THREADS - 1) / SWEEP_THREADS;
        if (plane_sz > 0x7fffffffLL / 16 || chunks > 0x7fffffffLL / 16)
            return fail(h, PDP_ENOTSUP, "grid too large for one launch (dims[2]*dims[3] too big)");
        P.chunks = (int)chunks;
        P.tile_rows = 1;
        if (h->mech2_mode) {   // range kernel: blocks may be tiles of tile_rows x (128 / tile_rows) nodes (PYRODP_TILE_ROWS, default below)
            int tr = MECH2_DEFAULT_TILE_ROWS(P.system_id);
            if (const char* env = getenv("PYRODP_TILE_ROWS")) tr = atoi(env);
            if (tr != 1 && tr != 2 && tr != 4 && tr != 8 && tr != 16) tr = 1;
            P.tile_rows = tr;
            if (tr > 1) {
                const long long tc = SWEEP_THREADS / tr;
                P.chunks = (int)(((P.dims[2] + tr - 1) / tr) * ((P.dims[3] + tc - 1) / tc));
            }
        }
    }
    if (h->smem_bytes > 227 * 1024) return fail(h, PDP_ENOTSUP, "level/action tables exceed shared memory (227 KB)");
    cudaError_t ce = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (ce != cudaSuccess) return fail(h, PDP_ECUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ce));
    // Shared memory is carved out of the 256 KB L1: ask for no more than the resident blocks need, the rest stays cache
    // (the 4-D kernels are bound by L1 miss fills).  A hint; the driver rounds it up to a supported split.
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k, SWEEP_THREADS, h->smem_bytes) == cudaSuccess && occ > 0) {
        const double need = (double)occ * (double)(h->smem_bytes + 1024);   // + the per-block reservation
        int pct = (int)(need / (228.0 * 1024.0) * 100.0) + 1;
        pct = std::min(std::max(pct, 1), 100);
        if (cudaFuncSetAttribute((const void*)k, cudaFuncAttributePreferredSharedMemoryCarveout, pct) != cudaSuccess) cudaGetLastError();
    } else {
        cudaGetLastError();
    }
    return PDP_OK;
}

// Halo of the axis-0 slab decomposition: how many planes below / above its own plane a node's
// backup can read.  Axis 0 is a position (MechanicalSystem: x = [q, dq], mechanical.py:238-263), so
// x_next[0] = dq0*dt + q0 depends on (i0, index of dq0) only: enumerate those pairs with the same two
// IEEE operations the kernels use and let the level table decide the cell, exactly as they do.
static void compute_halo(const pdp_problem* p, int* halo_lo, int* halo_hi) {
    const int n0 = p->dims[0], vax = p->n / 2, nv = p->dims[vax];
    const double* lev0 = p->x_level[0];
    const double* levv = p->x_level[vax];
    int lo = 0, hi = 1;
    for (int i0 = 0; i0 < n0; ++i0)
        for (int iv = 0; iv < nv; ++iv) {
            volatile double prod = levv[iv] * p->dt;   // volatile: two roundings, never an fma
            const double xn0 = prod + lev0[i0];
            if (xn0 < p->x_lb[0] || xn0 > p->x_ub[0]) continue;
            int c = (int)(std::upper_bound(lev0, lev0 + n0, xn0) - lev0) - 1;
            c = std::min(std::max(c, 0), n0 - 2);
            lo = std::max(lo, i0 - c);
            hi = std::max(hi, c + 1 - i0);
        }
    *halo_lo = lo;
    *halo_hi = hi;
}

extern "C" int pdp_compute_halo(const pdp_problem* p, int32_t* halo_lo, int32_t* halo_hi) {
    if (!p || !halo_lo || !halo_hi) return fail(nullptr, PDP_EINVAL, "pdp_compute_halo: null argument");
    if (p->n < 2 || p->n > 4 || p->system_id == PDP_SYS_LUT) { *halo_lo = *halo_hi = p->dims[0]; return PDP_OK; }
    int lo, hi;
    compute_halo(p, &lo, &hi);
    *halo_lo = lo; *halo_hi = hi;
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
    if (p->system_id != PDP_SYS_LUT && p->cost_id != PDP_COST_QUADRATIC && p->cost_id != PDP_COST_TIME && p->cost_id != PDP_COST_REACH)
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
    for (int d = 0; d < p->n; ++d) {
        if (p->x_level[d][0] != p->x_lb[d] || p->x_level[d][p->dims[d] - 1] != p->x_ub[d])
            return fail(nullptr, PDP_EINVAL, "pdp_create: x_level end points must equal x_lb/x_ub (np.linspace, discretizer.py:142)");
        for (int i = 0; i + 1 < p->dims[d]; ++i)
            if (!(p->x_level[d][i + 1] - p->x_level[d][i] > 0.0))
                return fail(nullptr, PDP_EINVAL, "pdp_create: x_level must be strictly increasing");
    }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PDP_ECUDA, std::string("pdp_create: no usable CUDA device (") + cudaGetErrorString(e) +
                                            "); this engine has no CPU fallback");

    pdp_handle* h = new pdp_handle();
    auto bail = [&](int code) { pdp_destroy(h); return code; };
    if (cudaGetDevice(&h->device) != cudaSuccess) { g_err = "cudaGetDevice failed"; return bail(PDP_ECUDA); }
    if (const char* env = getenv("PYRODP_LANES")) h->force_lanes = atoi(env);
    if (const char* env = getenv("PYRODP_GENERIC")) h->force_generic = atoi(env) != 0;
    if (const char* env = getenv("PYRODP_PEND_LOOP")) h->pend_loop = (atoi(env) == 2) ? 2 : 1;
    if (const char* env = getenv("PYRODP_TEST_INTERIOR_DELAY_US")) h->test_interior_delay_us = atoi(env);
    if (const char* env = getenv("PYRODP_MECH2"))
        h->force_mech2 = !strcmp(env, "generic") ? 0 : !strcmp(env, "range") ? 1 : -1;

    DevProblem& P = h->P;
    P.n = p->n; P.m = p->m; P.dof = p->n / 2; P.A = (int)A;
    P.system_id = p->system_id; P.cost_id = p->cost_id; P.ontarget_check = p->ontarget_check;
    P.alpha_is_one = (p->alpha == 1.0);
    P.dt = p->dt; P.alpha = p->alpha; P.INF = p->INF; P.EPS = p->EPS;
    long long stride = 1;
    for (int d = p->n - 1; d >= 0; --d) { P.stride[d] = stride; stride *= p->dims[d]; }
    h->N = N; h->A = (int)A; P.N = N;
    h->n0 = p->dims[0];
    h->plane = N / p->dims[0];
    h->slab_begin = p->slab_begin; h->slab_end = p->slab_end;

    // ---- which planes of J this handle holds ---------------------------------------------------
    //  * whole grid on one GPU: everything.
    //  * a slab, alloc_planes == 0: slab + halo (fused systems), the 180 GB-per-GPU layout;
    //    LUT mode keeps everything (an arbitrary x_next_table has no a-priori halo).
    //  * a slab, alloc_planes > 0: everything, padded to alloc_planes planes so that an in-place
    //    equal-count all-gather of whole slabs fits (used when the halo exceeds a neighbour's slab).
    const bool partial = (p->slab_begin != 0 || p->slab_end != p->dims[0]);
    h->halo_lo = h->halo_hi = p->dims[0];
    if (p->system_id != PDP_SYS_LUT) compute_halo(p, &h->halo_lo, &h->halo_hi);
    if (!partial || p->alloc_planes > 0 || p->system_id == PDP_SYS_LUT) {
        h->alloc_begin = 0; h->alloc_end = p->dims[0];
        h->alloc_planes_cap = p->alloc_planes > 0 ? p->alloc_planes : p->dims[0];
        if (h->alloc_planes_cap < p->dims[0]) { g_err = "pdp_create: alloc_planes < dims[0]"; return bail(PDP_EINVAL); }
    } else {
        h->alloc_begin = std::max(0, p->slab_begin - h->halo_lo);
        h->alloc_end = std::min((int)p->dims[0], p->slab_end + h->halo_hi);
        h->alloc_planes_cap = h->alloc_end - h->alloc_begin;
    }
    P.node_begin = (long long)p->slab_begin * h->plane;
    P.node_end = (long long)p->slab_end * h->plane;
    P.slab_node_begin = P.node_begin;
    P.plane_begin = (long long)p->slab_begin * p->dims[1];

    int rc;
    for (int d = 0; d < p->n; ++d) {
        P.dims[d] = p->dims[d];
        P.lb[d] = p->x_lb[d]; P.ub[d] = p->x_ub[d];
        P.inv_step[d] = (double)(p->dims[d] - 1) / (p->x_ub[d] - p->x_lb[d]);
        std::vector<double> rinv(p->dims[d]);
        for (int i = 0; i + 1 < p->dims[d]; ++i) rinv[i] = 1.0 / (p->x_level[d][i + 1] - p->x_level[d][i]);
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
    P.all_act_ok = 1;
    if (p->system_id != PDP_SYS_LUT) {
        // isavalidinput (system.py:208-215) is folded into the B.u table: a disallowed action carries
        // NaN, its x_next is NaN, fails the box test and gets Q = INF exactly as dynamicprogramming.py:233
        std::vector<double> bu(p->bu, p->bu + (size_t)A * P.dof);
        for (long long a = 0; a < A; ++a)
            if (!p->act_ok[a]) {
                P.all_act_ok = 0;
                for (int d = 0; d < P.dof; ++d) bu[(size_t)a * P.dof + d] = __builtin_nan("");
            }
        if (p->system_id == PDP_SYS_PENDULUM) {
            // x_next[1] = ((B.u - g - d) * inv(H)) * dt + dq is a chain of monotone IEEE operations of B.u when
            // inv(H) > 0 and dt > 0: ascending B.u (and no disallowed action) => non-decreasing x_next[1]
            bool asc = P.all_act_ok && p->sys_par[0] > 0.0 && p->dt > 0.0;
            for (long long a = 1; a < A && asc; ++a) asc = bu[a] >= bu[a - 1];
            h->pend_mono = asc;
        }
        if (p->system_id == PDP_SYS_TWOLINK || p->system_id == PDP_SYS_CARTPOLE) {
            h->plan = mech2_plan(p->system_id == PDP_SYS_TWOLINK, p->udims, bu.data(), A, P.all_act_ok, p->dt);
            P.A0 = h->plan.A0; P.A1 = h->plan.A1;
            P.uv_first = h->plan.uv_first; P.uv_inv_step = h->plan.uv_inv_step;
        }
        if ((rc = upload(h, bu.data(), bu.size(), &P.bu)) != PDP_OK) return bail(rc);
        if ((rc = upload(h, p->gu, (size_t)A, &P.gu)) != PDP_OK) return bail(rc);
        if ((rc = upload(h, (const unsigned char*)p->act_ok, (size_t)A, &P.act_ok)) != PDP_OK) return bail(rc);
    }

    auto cu = [&](cudaError_t ce, const char* what) -> bool {
        if (ce != cudaSuccess) { g_err = std::string(what) + ": " + cudaGetErrorString(ce); return false; }
        return true;
    };
    const size_t jbytes = (size_t)std::max<long long>(h->alloc_planes_cap * h->plane, 1) * sizeof(double);
    const size_t pibytes = (size_t)std::max<long long>(h->slab_nodes(), 1) * sizeof(long long);
    if (!cu(cudaMalloc(&h->dJ[0], jbytes), "cudaMalloc J0")) return bail(PDP_ECUDA);
    if (!cu(cudaMalloc(&h->dJ[1], jbytes), "cudaMalloc J1")) return bail(PDP_ECUDA);
    if (!cu(cudaMalloc(&h->dpi, pibytes), "cudaMalloc pi")) return bail(PDP_ECUDA);
    h->stats_cap = 256;
    if (!cu(cudaMalloc(&h->dstats, h->stats_cap * 3 * sizeof(double)), "cudaMalloc stats")) return bail(PDP_ECUDA);
    if (!cu(cudaMalloc(&h->dstats_sets, PDP_STAT_SETS * 3 * sizeof(double)), "cudaMalloc stats sets")) return bail(PDP_ECUDA);
    const size_t slot_bytes = (size_t)PDP_STAT_SETS * 3 * STATS_SLOTS * sizeof(unsigned long long);
    if (!cu(cudaMalloc(&h->dslots, slot_bytes), "cudaMalloc stats slots")) return bail(PDP_ECUDA);
    if (!cu(cudaMemset(h->dslots, 0, slot_bytes), "cudaMemset stats slots")) return bail(PDP_ECUDA);
    if (!cu(cudaMalloc(&h->dcounter, PDP_STAT_SETS * sizeof(unsigned int)), "cudaMalloc counter")) return bail(PDP_ECUDA);
    if (!cu(cudaMemset(h->dcounter, 0, PDP_STAT_SETS * sizeof(unsigned int)), "cudaMemset counter")) return bail(PDP_ECUDA);
    if (!cu(cudaMemset(h->dpi, 0, pibytes), "cudaMemset pi")) return bail(PDP_ECUDA);
    if (!cu(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return bail(PDP_ECUDA);
    h->own_stream = true;
    if (!cu(cudaEventCreate(&h->ev0), "cudaEventCreate")) return bail(PDP_ECUDA);
    if (!cu(cudaEventCreate(&h->ev1), "cudaEventCreate")) return bail(PDP_ECUDA);

    // fused kernels: lanes per node, kernel instantiation, dynamic shared memory
    if (p->system_id != PDP_SYS_LUT) {
        int rcsel = select_fused_kernel(h);
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
    cudaFree(h->dJ[0]); cudaFree(h->dJ[1]); cudaFree(h->dpi); cudaFree(h->dstats); cudaFree(h->dstats_sets);
    cudaFree(h->dslots); cudaFree(h->dcounter); cudaFree(h->d_xnext); cudaFree(h->d_G);
    for (int side = 0; side < 2; ++side) {
        for (int b = 0; b < 2; ++b) if (h->peer_J[side][b]) cudaIpcCloseMemHandle(h->peer_J[side][b]);
        if (h->peer_flags[side]) cudaIpcCloseMemHandle(h->peer_flags[side]);
    }
    cudaFree(h->dflags);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    if (h->ev_boundary) cudaEventDestroy(h->ev_boundary);
    if (h->ev_comm) cudaEventDestroy(h->ev_comm);
    for (cudaEvent_t e : h->ev_up) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_done) cudaEventDestroy(e);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    for (int i = 0; i < 2; ++i) {
        if (h->side_stream[i]) cudaStreamDestroy(h->side_stream[i]);
        if (h->ev_side[i]) cudaEventDestroy(h->ev_side[i]);
    }
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    cudaFree(h->dchunk_stats);
    if (h->h_chunk_stats) cudaFreeHost(h->h_chunk_stats);
    for (auto& g : h->host_graphs) cudaGraphExecDestroy(g.exec);
    for (int i = 0; i < 2; ++i) if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
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
extern "C" int64_t pdp_nodes_padded(const pdp_handle* h) { return h ? h->alloc_planes_cap * h->plane : 0; }
extern "C" int64_t pdp_actions(const pdp_handle* h) { return h ? h->A : 0; }
extern "C" int64_t pdp_launch_count(const pdp_handle* h) { return h ? h->launches : 0; }
extern "C" double pdp_last_sweep_ms(const pdp_handle* h) { return h ? h->last_ms : 0.0; }

extern "C" int pdp_slab_layout(const pdp_handle* h, int32_t out[8]) {
    if (!h || !out) return fail(nullptr, PDP_EINVAL, "pdp_slab_layout: null argument");
    out[0] = h->slab_begin; out[1] = h->slab_end;
    out[2] = h->alloc_begin; out[3] = h->alloc_end;
    out[4] = h->halo_lo; out[5] = h->halo_hi;
    out[6] = h->n0; out[7] = h->lanes_per_node;
    return PDP_OK;
}

extern "C" int pdp_eval_terminal_cost(pdp_handle* h) {
    CHECK_HANDLE(h);
    if (h->P.system_id == PDP_SYS_LUT) return fail(h, PDP_ENOTSUP, "pdp_eval_terminal_cost: LUT mode has no cost model; use pdp_set_J");
    const int threads = 256;
    const long long first = (long long)h->alloc_begin * h->plane, last = (long long)h->alloc_end * h->plane;
    const long long blocks = (last - first + threads - 1) / threads;
    if (blocks > 0) {
        double* J = h->Jv(h->cur_idx);
        if (h->P.n == 2) terminal_cost_kernel<2><<<(unsigned)blocks, threads, 0, h->stream>>>(h->P, J, h->piv(), first, last);
        else if (h->P.n == 4) terminal_cost_kernel<4><<<(unsigned)blocks, threads, 0, h->stream>>>(h->P, J, h->piv(), first, last);
        else return fail(h, PDP_ENOTSUP, "pdp_eval_terminal_cost: n must be 2 or 4");
        CUDA_TRY(h, cudaGetLastError());
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_J = true;
    return PDP_OK;
}

// J_host is always the FULL grid (N doubles); the handle copies the planes it holds.
extern "C" int pdp_set_J(pdp_handle* h, const double* J_host) {
    CHECK_HANDLE(h);
    if (!J_host) return fail(h, PDP_EINVAL, "pdp_set_J: null pointer");
    if (h->alloc_nodes() > 0)
        CUDA_TRY(h, cudaMemcpyAsync(h->dJ[h->cur_idx], J_host + (long long)h->alloc_begin * h->plane,
                                    h->alloc_nodes() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_J = true;
    h->halo_stale = false;
    return PDP_OK;
}

static int copy_out(pdp_handle* h, void* dst, const void* src, size_t bytes) {
    if (bytes) CUDA_TRY(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PDP_OK;
}

// The getters return THIS HANDLE'S SLAB (slab_nodes values, plane slab_begin first); on a single
// GPU the slab is the whole grid, i.e. the reference's (N,) arrays.
extern "C" int pdp_get_J(pdp_handle* h, double* J_host) {
    CHECK_HANDLE(h);
    if (!J_host) return fail(h, PDP_EINVAL, "pdp_get_J: null pointer");
    if (!h->have_J) return fail(h, PDP_ESTATE, "pdp_get_J: no cost-to-go yet");
    return copy_out(h, J_host, h->Jv(h->cur_idx) + h->P.slab_node_begin, h->slab_nodes() * sizeof(double));
}
extern "C" int pdp_get_J_next(pdp_handle* h, double* J_host) {
    CHECK_HANDLE(h);
    if (!J_host) return fail(h, PDP_EINVAL, "pdp_get_J_next: null pointer");
    if (!h->have_J) return fail(h, PDP_ESTATE, "pdp_get_J_next: no cost-to-go yet");
    return copy_out(h, J_host, h->Jv(1 - h->cur_idx) + h->P.slab_node_begin, h->slab_nodes() * sizeof(double));
}
extern "C" int pdp_get_pi(pdp_handle* h, int64_t* pi_host) {
    CHECK_HANDLE(h);
    if (!pi_host) return fail(h, PDP_EINVAL, "pdp_get_pi: null pointer");
    return copy_out(h, pi_host, h->dpi, h->slab_nodes() * sizeof(long long));
}

// Node ranges of the latest J / J_next / pi (global node ids inside this handle's slab): sampled parity checks of
// grids whose full arrays are tens of gigabytes.  which: 0 = J, 1 = J_next, 2 = pi (int64).
extern "C" int pdp_get_range(pdp_handle* h, int32_t which, int64_t node_begin, int64_t count, void* out_host) {
    CHECK_HANDLE(h);
    if (!out_host || count < 0) return fail(h, PDP_EINVAL, "pdp_get_range: bad argument");
    if (which < 0 || which > 2) return fail(h, PDP_EINVAL, "pdp_get_range: which must be 0 (J), 1 (J_next) or 2 (pi)");
    if (which != 2 && !h->have_J) return fail(h, PDP_ESTATE, "pdp_get_range: no cost-to-go yet");
    // J / J_next: any node of the planes the handle holds (slab + halo); pi: the slab only
    const long long lo = which == 2 ? h->P.slab_node_begin : (long long)h->alloc_begin * h->plane;
    const long long hi = which == 2 ? h->P.slab_node_begin + h->slab_nodes() : (long long)h->alloc_end * h->plane;
    if (node_begin < lo || node_begin + count > hi)
        return fail(h, PDP_EINVAL, "pdp_get_range: node range outside the planes this handle holds");
    if (which == 2) return copy_out(h, out_host, h->piv() + node_begin, (size_t)count * sizeof(long long));
    return copy_out(h, out_host, h->Jv(which == 0 ? h->cur_idx : 1 - h->cur_idx) + node_begin, (size_t)count * sizeof(double));
}

// which sweep kernel this handle launches (for benchmark records and tests)
extern "C" int pdp_kernel_info(const pdp_handle* h, char* out, int32_t len) {
    if (!h || !out || len < 1) return fail(nullptr, PDP_EINVAL, "pdp_kernel_info: bad argument");
    std::string name;
    const DevProblem& P = h->P;
    if (P.system_id == PDP_SYS_LUT) name = P.A == 1 ? "sweep_policy_kernel" : (h->spline ? "sweep_lut_spline_kernel" : "sweep_lut_kernel");
    else if (P.system_id == PDP_SYS_PENDULUM) name = std::string("sweep_pendulum_kernel<") + (h->pend_mono && !h->force_generic ? (h->pend_loop == 2 ? "mono, loop nest" : "mono, pair loop") : "generic") + ">";
    else {
        const char* sys = P.system_id == PDP_SYS_TWOLINK ? "TWOLINK" : "CARTPOLE";
        if (h->mech2_mode) name = std::string("sweep_mech2_range_kernel<") + sys + ">";
        else name = std::string("sweep_mech2_kernel<") + sys + ">";
    }
    name += " G=" + std::to_string(h->lanes_per_node);
    snprintf(out, (size_t)len, "%s", name.c_str());
    return PDP_OK;
}

extern "C" int pdp_set_lut(pdp_handle* h, const double* x_next_host, const double* G_host) {
    CHECK_HANDLE(h);
    if (h->P.system_id != PDP_SYS_LUT) return fail(h, PDP_ESTATE, "pdp_set_lut: handle was not created with PDP_SYS_LUT");
    if (!x_next_host || !G_host) return fail(h, PDP_EINVAL, "pdp_set_lut: null pointer");
    const size_t slab = (size_t)h->slab_nodes();
    const size_t nx = slab * h->A * h->P.n, ng = slab * h->A;
    if (!h->d_xnext) CUDA_TRY(h, cudaMalloc(&h->d_xnext, (nx ? nx : 1) * sizeof(double)));
    if (!h->d_G) CUDA_TRY(h, cudaMalloc(&h->d_G, (ng ? ng : 1) * sizeof(double)));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_xnext, x_next_host, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_G, G_host, ng * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_lut = true;
    return PDP_OK;
}

// DynamicProgramming2DRectBivariateSpline (dynamicprogramming.py:578-614): which interpolant of J_next the table sweep uses
extern "C" int pdp_set_interpolant(pdp_handle* h, int32_t which) {
    CHECK_HANDLE(h);
    if (which == PDP_INTERP_LINEAR) { h->spline = false; return PDP_OK; }
    if (which != PDP_INTERP_SPLINE3) return fail(h, PDP_EINVAL, "pdp_set_interpolant: unknown interpolant");
    const DevProblem& P = h->P;
    if (P.system_id != PDP_SYS_LUT || P.n != 2 || P.A < 2)
        return fail(h, PDP_ENOTSUP, "pdp_set_interpolant: the bicubic spline needs a table-mode handle of a 2-D grid (discretizer.py:600-612)");
    if (h->slab_begin != 0 || h->slab_end != P.dims[0]) return fail(h, PDP_ENOTSUP, "pdp_set_interpolant: the spline is fitted on the whole grid; one handle");
    if (!h->S.coef) {
        for (int d = 0; d < 2; ++d) {
            const int m = P.dims[d];
            std::vector<double> lev((size_t)m), knots, lu, rden;
            CUDA_TRY(h, cudaMemcpy(lev.data(), P.level[d], (size_t)m * sizeof(double), cudaMemcpyDeviceToHost));
            if (!spline_plan_axis(lev.data(), m, knots, lu, rden))
                return fail(h, PDP_ENOTSUP, "pdp_set_interpolant: a cubic spline needs at least 4 levels per axis");
            int rc;
            if ((rc = upload(h, knots.data(), knots.size(), &h->S.knots[d])) != PDP_OK) return rc;
            if ((rc = upload(h, lu.data(), lu.size(), &h->S.lu[d])) != PDP_OK) return rc;
            if ((rc = upload(h, rden.data(), rden.size(), &h->S.rden[d])) != PDP_OK) return rc;
            h->S.m[d] = m;
        }
        void* c = nullptr;
        CUDA_TRY(h, cudaMalloc(&c, (size_t)P.dims[0] * P.dims[1] * sizeof(double)));
        h->owned.push_back(c);
        h->S.coef = (double*)c;
    }
    h->spline = true;
    return PDP_OK;
}

template <int N>
static void launch_lut(pdp_handle* h, cudaStream_t stream, const DevProblem& P, int G, unsigned blocks, const double* Jn, double* Jo,
                       unsigned long long* slots, unsigned int* counter, double* stats) {
#define LUT_CASE(g)                                                                                                  \
    case g:                                                                                                          \
        sweep_lut_kernel<N, g><<<blocks, SWEEP_THREADS, 0, stream>>>(P, Jn, Jo, h->piv(), h->d_xnext, h->d_G,    \
                                                                        slots, counter, stats);                     \
        break;
    switch (G) {
        LUT_CASE(1) LUT_CASE(2) LUT_CASE(4) LUT_CASE(8) LUT_CASE(16) LUT_CASE(32)
    }
#undef LUT_CASE
}

// One backup of axis-0 planes [p0,p1) (a sub-range of the slab) on the handle's stream:
// reads J[cur] (slab + halo), writes J[1-cur] and pi on those planes, stats triple -> `stats`.
static int launch_planes(pdp_handle* h, int p0, int p1, int stat_set, double* stats, cudaStream_t stream = nullptr) {
    if (!stream) stream = h->stream;
    DevProblem P = h->P;
    P.node_begin = (long long)p0 * h->plane;
    P.node_end = (long long)p1 * h->plane;
    P.plane_begin = (long long)p0 * P.dims[1];
    const long long nodes = P.node_end - P.node_begin;
    const double* Jn = h->Jv(h->cur_idx);
    double* Jo = h->Jv(1 - h->cur_idx);
    unsigned long long* slots = h->dslots + (size_t)stat_set * 3 * STATS_SLOTS;
    unsigned int* counter = h->dcounter + stat_set;
    if (nodes <= 0) {
        const double ident[3] = {-__builtin_inf(), -__builtin_inf(), __builtin_inf()};
        CUDA_TRY(h, cudaMemcpyAsync(stats, ident, sizeof(ident), cudaMemcpyHostToDevice, stream));
        return PDP_OK;
    }
    if (P.system_id == PDP_SYS_LUT) {
        if (!h->have_lut) return fail(h, PDP_ESTATE, "pdp_sweep: LUT mode needs pdp_set_lut first");
        if (P.A == 1) {   // policy evaluation: the streaming kernel
            typedef void (*policy_kernel_t)(const DevProblem, const double*, double*, long long*, const double*, const double*,
                                            unsigned long long*, unsigned int*, double*);
            policy_kernel_t pk = P.n == 2 ? (policy_kernel_t)sweep_policy_kernel<2, POLICY_U2>
                               : P.n == 3 ? (policy_kernel_t)sweep_policy_kernel<3, POLICY_U4> : (policy_kernel_t)sweep_policy_kernel<4, POLICY_U4>;
            if (h->policy_blocks == 0) {
                int per_sm = 0, sm_count = 148;
                cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
                CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)pk, 256, 0));
                h->policy_blocks = std::max(per_sm, 1) * sm_count;
            }
            const long long blocks = std::min<long long>((nodes + 256 - 1) / 256, h->policy_blocks);
            pk<<<(unsigned)blocks, 256, 0, stream>>>(P, Jn, Jo, h->piv(), h->d_xnext, h->d_G, slots, counter, stats);
            CUDA_TRY(h, cudaGetLastError());
            h->launches += 1;
            return PDP_OK;
        }
        int G = 1;
        while (G < 32 && G < P.A) G <<= 1;
        int sm_count = 148;
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
        // about one resident wave (12 blocks of 128 threads per SM at 40-56 registers) strides over the nodes
        const long long blocks = std::min<long long>((nodes * G + SWEEP_THREADS - 1) / SWEEP_THREADS, (long long)sm_count * 12);
        if (h->spline) {
            // DynamicProgramming2DRectBivariateSpline: fit the interpolating bicubic spline of J_next (two sweeps of banded
            // substitutions), then the table sweep with the spline evaluation
            spline_fit_axis0_kernel<<<(unsigned)((P.dims[1] + 127) / 128), 128, 0, stream>>>(Jn, h->S);
            spline_fit_axis1_kernel<<<(unsigned)((P.dims[0] + 127) / 128), 128, 0, stream>>>(h->S);
            h->launches += 2;
            switch (G) {
#define SPL_CASE(g) case g: sweep_lut_spline_kernel<g><<<(unsigned)blocks, SWEEP_THREADS, 0, stream>>>(P, h->S, Jn, Jo, h->piv(), h->d_xnext, h->d_G, slots, counter, stats); break;
                SPL_CASE(1) SPL_CASE(2) SPL_CASE(4) SPL_CASE(8) SPL_CASE(16) SPL_CASE(32)
#undef SPL_CASE
            }
            CUDA_TRY(h, cudaGetLastError());
            h->launches += 1;
            return PDP_OK;
        }
        if (P.n == 2) launch_lut<2>(h, stream, P, G, (unsigned)blocks, Jn, Jo, slots, counter, stats);
        else if (P.n == 3) launch_lut<3>(h, stream, P, G, (unsigned)blocks, Jn, Jo, slots, counter, stats);
        else launch_lut<4>(h, stream, P, G, (unsigned)blocks, Jn, Jo, slots, counter, stats);
    } else {
        const int G = h->lanes_per_node;
        dim3 grid;
        if (P.system_id == PDP_SYS_PENDULUM) {
            P.plane_begin = p0;  // blockIdx.x = row of the plane range, blockIdx.y = chunk of the row
            grid = dim3((unsigned)(p1 - p0), (unsigned)(((long long)P.dims[1] * G + SWEEP_THREADS - 1) / SWEEP_THREADS), 1);
        } else {
            const long long pairs = (long long)(p1 - p0) * P.dims[1];
            if (pairs * P.chunks > 0x7fffffffLL) return fail(h, PDP_ENOTSUP, "grid too large for one launch");
            grid = dim3((unsigned)(pairs * P.chunks), 1, 1);   // chunk of the (i2,i3) plane fastest
        }
        ((fused_kernel_t)h->fused)<<<grid, SWEEP_THREADS, h->smem_bytes, stream>>>(P, Jn, Jo, h->piv(), slots, counter, stats);
    }
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return PDP_OK;
}

extern "C" int pdp_sweep_async(pdp_handle* h) {
    CHECK_HANDLE(h);
    if (!h->have_J) return fail(h, PDP_ESTATE, "pdp_sweep: no cost-to-go yet (pdp_set_J / pdp_eval_terminal_cost)");
    if (h->pending) return fail(h, PDP_ESTATE, "pdp_sweep_async: previous sweep not committed");
    int rc = launch_planes(h, h->slab_begin, h->slab_end, 0, h->dstats_sets);
    if (rc != PDP_OK) return rc;
    h->pending = true;
    return PDP_OK;
}

extern "C" int pdp_sweep_planes_async(pdp_handle* h, int32_t plane_begin, int32_t plane_end, int32_t stat_set) {
    CHECK_HANDLE(h);
    if (!h->have_J) return fail(h, PDP_ESTATE, "pdp_sweep: no cost-to-go yet (pdp_set_J / pdp_eval_terminal_cost)");
    if (plane_begin < h->slab_begin || plane_end > h->slab_end || plane_begin > plane_end)
        return fail(h, PDP_EINVAL, "pdp_sweep_planes_async: plane range outside this handle's slab");
    if (stat_set < 0 || stat_set >= PDP_STAT_SETS) return fail(h, PDP_EINVAL, "pdp_sweep_planes_async: stat_set out of range");
    int rc = launch_planes(h, plane_begin, plane_end, stat_set, h->dstats_sets + 3 * stat_set);
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

extern "C" int pdp_read_stats(pdp_handle* h, double* stats_host) {
    CHECK_HANDLE(h);
    if (!stats_host) return fail(h, PDP_EINVAL, "pdp_read_stats: null pointer");
    return copy_out(h, stats_host, h->dstats_sets, (size_t)PDP_STAT_SETS * 3 * sizeof(double));
}

extern "C" int pdp_device_buffers(pdp_handle* h, void** J_cur, void** J_new, void** pi, void** stats) {
    CHECK_HANDLE(h);
    if (J_cur) *J_cur = h->dJ[h->cur_idx];
    if (J_new) *J_new = h->dJ[1 - h->cur_idx];
    if (pi) *pi = h->dpi;
    if (stats) *stats = h->dstats_sets;
    return PDP_OK;
}

// ---- multi-GPU: NCCL through dlopen (the process-wide libnccl.so.2, i.e. the one torch already loaded) ----
static int nccl_load() {
    if (g_nccl.lib) return PDP_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(nullptr, PDP_ECUDA, std::string("cannot load libnccl.so.2: ") + dlerror());
#define NCCL_SYM(field, name)                                                              \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                                           \
    if (!g_nccl.field) return fail(nullptr, PDP_ECUDA, std::string("libnccl.so.2 lacks ") + name);
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId") NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy") NCCL_SYM(Send, "ncclSend") NCCL_SYM(Recv, "ncclRecv")
    NCCL_SYM(GroupStart, "ncclGroupStart") NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(AllReduce, "ncclAllReduce") NCCL_SYM(AllGather, "ncclAllGather")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    g_nccl.lib = lib;
    return PDP_OK;
}

#define NCCL_TRY(h, expr)                                                                            \
    do {                                                                                             \
        int _r = (expr);                                                                             \
        if (_r != 0) return fail(h, PDP_ECUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)

extern "C" int pdp_nccl_unique_id(void* id128) {
    if (!id128) return fail(nullptr, PDP_EINVAL, "pdp_nccl_unique_id: null pointer");
    int rc = nccl_load();
    if (rc != PDP_OK) return rc;
    NCCL_TRY(nullptr, g_nccl.GetUniqueId((pdp_nccl_id*)id128));
    return PDP_OK;
}

extern "C" int pdp_comm_init(pdp_handle* h, int32_t rank, int32_t world, const void* id128, int32_t exchange_mode,
                             int32_t overlap) {
    CHECK_HANDLE(h);
    if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(h, PDP_EINVAL, "pdp_comm_init: bad rank / world / id");
    if (exchange_mode != 1 && exchange_mode != 2) return fail(h, PDP_EINVAL, "pdp_comm_init: exchange_mode must be 1 (halo) or 2 (all-gather)");
    if (h->comm) return fail(h, PDP_ESTATE, "pdp_comm_init: communicator already attached");
    if (exchange_mode == 2 && (h->alloc_begin != 0 || h->alloc_end != h->n0 || h->alloc_planes_cap % world != 0 ||
                               (long long)h->slab_begin != std::min<long long>((long long)rank * (h->alloc_planes_cap / world), h->n0)))
        return fail(h, PDP_EINVAL, "pdp_comm_init: all-gather mode needs the whole grid in buffers of W*ceil(dims[0]/W) planes "
                                   "and slab r = planes [r*P, (r+1)*P)");
    if (exchange_mode == 1 && world > 1) {
        const bool lo_ok = rank == 0 || (h->alloc_begin == h->slab_begin - h->halo_lo);
        const bool hi_ok = rank == world - 1 || (h->alloc_end == h->slab_end + h->halo_hi);
        if (!lo_ok || !hi_ok || h->slab_end - h->slab_begin < std::max(h->halo_lo, h->halo_hi))
            return fail(h, PDP_EINVAL, "pdp_comm_init: halo mode needs slab + halo buffers and a slab at least as thick as the halo");
    }
    int rc = nccl_load();
    if (rc != PDP_OK) return fail(h, rc, g_err);
    pdp_nccl_id id;
    memcpy(&id, id128, sizeof(id));
    NCCL_TRY(h, g_nccl.CommInitRank(&h->comm, world, id, rank));
    h->rank = rank; h->world = world; h->exchange_mode = exchange_mode; h->overlap = overlap;
    int prio_lo = 0, prio_hi = 0;
    CUDA_TRY(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CUDA_TRY(h, cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, prio_hi));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_boundary, cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
    return PDP_OK;
}

// Complete buffer `which` on this rank after its slab planes were (re)written: halo planes from the
// neighbouring ranks (grouped send/recv), or the in-place all-gather of whole slabs.
static int exchange(pdp_handle* h, int which, cudaStream_t st) {
    if (!h->comm || h->world == 1) return PDP_OK;
    double* base = h->dJ[which];  // element 0 = plane alloc_begin
    auto at = [&](int plane) { return base + (long long)(plane - h->alloc_begin) * h->plane; };
    h->exchanges += 1;
    if (h->exchange_mode == 3) {
        // my lowest halo_hi planes are the upper halo of the rank below; my highest halo_lo planes the lower halo of the rank above
        const bool lo = h->rank > 0, hi = h->rank < h->world - 1;
        const long long n_lo = lo ? (long long)h->halo_hi * h->plane : 0, n_hi = hi ? (long long)h->halo_lo * h->plane : 0;
        double* dst_lo = lo ? h->peer_J[0][which] + (long long)(h->slab_begin - h->peer_alloc_begin[0]) * h->plane : nullptr;
        double* dst_hi = hi ? h->peer_J[1][which] + (long long)(h->slab_end - h->halo_lo - h->peer_alloc_begin[1]) * h->plane : nullptr;
        h->peer_seq += 1;
        const long long n = std::max(n_lo, n_hi);
        const int blocks = (int)std::min<long long>(std::max<long long>((n + 1023) / 1024, 1), 592);
        halo_push_kernel<<<blocks, 256, 0, st>>>(at(h->slab_begin), dst_lo, n_lo, at(h->slab_end - h->halo_lo), dst_hi, n_hi,
                                                 h->dflags + 2, lo ? h->peer_flags[0] + 1 : nullptr, hi ? h->peer_flags[1] + 0 : nullptr,
                                                 h->peer_seq);
        CUDA_TRY(h, cudaGetLastError());
        return PDP_OK;
    }
    if (h->exchange_mode == 2) {
        const size_t cnt = (size_t)(h->alloc_planes_cap / h->world) * h->plane;
        NCCL_TRY(h, g_nccl.AllGather(base + (size_t)h->rank * cnt, base, cnt, PDP_NCCL_F64, h->comm, st));
        return PDP_OK;
    }
    const size_t nlo = (size_t)h->halo_lo * h->plane, nhi = (size_t)h->halo_hi * h->plane;
    NCCL_TRY(h, g_nccl.GroupStart());
    if (h->rank > 0) {  // rank r-1 reads my lowest halo_hi planes; I read its highest halo_lo planes
        NCCL_TRY(h, g_nccl.Send(at(h->slab_begin), nhi, PDP_NCCL_F64, h->rank - 1, h->comm, st));
        NCCL_TRY(h, g_nccl.Recv(at(h->slab_begin - h->halo_lo), nlo, PDP_NCCL_F64, h->rank - 1, h->comm, st));
    }
    if (h->rank < h->world - 1) {
        NCCL_TRY(h, g_nccl.Send(at(h->slab_end - h->halo_lo), nlo, PDP_NCCL_F64, h->rank + 1, h->comm, st));
        NCCL_TRY(h, g_nccl.Recv(at(h->slab_end), nhi, PDP_NCCL_F64, h->rank + 1, h->comm, st));
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    return PDP_OK;
}

// One sweep of this rank's slab + exchange, enqueued without host synchronisation; the folded
// statistics {jmax, dmax, -dmin} of the slab go to dst[0..2].
static int sharded_sweep_enqueue(pdp_handle* h, double* dst) {
    const int b = h->slab_begin, e = h->slab_end, lo = h->halo_lo, hi = h->halo_hi;
    double* sets = h->dstats_sets;
    int rc;
    if (h->exchange_mode == 3) {
        // the halo planes of J[cur] were stored by the neighbours in exchange number peer_seq: wait for them on the device
        halo_wait_kernel<<<1, 1, 0, h->stream>>>(h->dflags, h->peer_seq, h->rank > 0, h->rank < h->world - 1, h->dflags + 3,
                                                 20000000000LL /* ~10 s */);
        CUDA_TRY(h, cudaGetLastError());
    }
    // Peer-store mode: the neighbours store their next sweep's planes into THIS rank's J[cur] halo as soon as they have
    // this rank's flag, which is published after the boundary kernels.  The interior kernel must therefore never read
    // a halo plane: with halo_lo != halo_hi (asymmetric velocity bounds) a boundary of halo_hi planes at the low end
    // would leave interior planes that still reach below the slab — so both boundaries are max(lo, hi) planes thick.
    // (NCCL mode orders the receive after the interior through ev_boundary / the stream, and keeps the thin boundaries.)
    int blo = hi, bhi = lo;   // planes computed first at the low / high end of the slab
    if (h->exchange_mode == 3) blo = bhi = std::max(lo, hi);
    if ((h->exchange_mode == 1 || h->exchange_mode == 3) && h->overlap && b + blo < e - bhi) {
        // Boundary planes + their exchange on the high-priority side stream, interior planes on the
        // main stream: the two kernels share the SMs (no serialised tail), the boundary blocks are
        // scheduled first, and the exchange (NCCL send/recv, or the peer-store kernel) runs under the
        // interior planes — the neighbours' flags are up long before their next sweep asks for them.
        CUDA_TRY(h, cudaEventRecord(h->ev_boundary, h->stream));            // previous sweep (and its exchange) done
        CUDA_TRY(h, cudaStreamWaitEvent(h->comm_stream, h->ev_boundary, 0));
        if ((rc = launch_planes(h, b, b + blo, 1, sets + 3, h->comm_stream)) != PDP_OK) return rc;
        if ((rc = launch_planes(h, e - bhi, e, 2, sets + 6, h->comm_stream)) != PDP_OK) return rc;
        if ((rc = exchange(h, 1 - h->cur_idx, h->comm_stream)) != PDP_OK) return rc;
        CUDA_TRY(h, cudaEventRecord(h->ev_comm, h->comm_stream));
        if (h->test_interior_delay_us > 0)   // test hook: hold the interior back so that a racing neighbour would be caught
            delay_kernel<<<1, 1, 0, h->stream>>>((long long)h->test_interior_delay_us * 2000LL);
        if ((rc = launch_planes(h, b + blo, e - bhi, 0, sets)) != PDP_OK) return rc;
        CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_comm, 0));
        stats_fold_kernel<<<1, 1, 0, h->stream>>>(sets, 3, dst);
    } else {
        if ((rc = launch_planes(h, b, e, 0, sets)) != PDP_OK) return rc;
        if ((rc = exchange(h, 1 - h->cur_idx, h->stream)) != PDP_OK) return rc;
        stats_fold_kernel<<<1, 1, 0, h->stream>>>(sets, 1, dst);
    }
    CUDA_TRY(h, cudaGetLastError());
    h->cur_idx = 1 - h->cur_idx;
    return PDP_OK;
}

// Enqueue one full sweep (all of this handle's planes, plus the exchange when a communicator is
// attached) without host synchronisation; its statistics go to the next slot of the history.
extern "C" int pdp_sweep_enqueue(pdp_handle* h) {
    CHECK_HANDLE(h);
    if (!h->have_J) return fail(h, PDP_ESTATE, "pdp_sweep: no cost-to-go yet (pdp_set_J / pdp_eval_terminal_cost)");
    if (h->pending) return fail(h, PDP_ESTATE, "pdp_sweep: an async sweep is pending");
    if (h->halo_stale) return fail(h, PDP_ESTATE, "pdp_sweep: the halo planes are stale after pdp_sweep_host on a slab; call pdp_exchange_current or pdp_set_J");
    const bool sharded = h->comm && h->world > 1;
    if (!sharded && h->slab_nodes() != h->N)
        return fail(h, PDP_ESTATE, "pdp_sweep: handle owns a slab only; attach a communicator (pdp_comm_init) or drive it with "
                                   "pdp_sweep_async + exchange + pdp_commit_sweep");
    if (h->enqueued >= h->stats_cap) {  // grow the history, keeping what is there
        double* bigger = nullptr;
        const int cap = h->stats_cap * 2;
        CUDA_TRY(h, cudaMalloc(&bigger, (size_t)cap * 3 * sizeof(double)));
        CUDA_TRY(h, cudaMemcpyAsync(bigger, h->dstats, (size_t)h->stats_cap * 3 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        CUDA_TRY(h, cudaFree(h->dstats));
        h->dstats = bigger;
        h->stats_cap = cap;
    }
    double* dst = h->dstats + 3 * h->enqueued;
    int rc;
    if (sharded) {
        rc = sharded_sweep_enqueue(h, dst);
    } else {
        rc = launch_planes(h, h->slab_begin, h->slab_end, 0, dst);
        if (rc == PDP_OK) h->cur_idx = 1 - h->cur_idx;
    }
    if (rc != PDP_OK) return rc;
    h->enqueued += 1;
    return PDP_OK;
}

// Wait for the enqueued sweeps, reduce their statistics over the ranks (one all-reduce of 3 doubles
// per sweep) and copy up to max_out of them (oldest first) to the host.
extern "C" int pdp_sweep_collect(pdp_handle* h, pdp_stats* stats_out, int32_t max_out, int32_t* n_out) {
    CHECK_HANDLE(h);
    const int n = h->enqueued;
    if (n_out) *n_out = std::min(n, (int)max_out);
    if (n > 0 && h->comm && h->world > 1) {
        NCCL_TRY(h, g_nccl.AllReduce(h->dstats, h->dstats, (size_t)3 * n, PDP_NCCL_F64, PDP_NCCL_MAX, h->comm, h->stream));
        stats_unfold_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(h->dstats, n);
        CUDA_TRY(h, cudaGetLastError());
    }
    const int m = std::min(n, (int)max_out);
    if (stats_out && m > 0)
        CUDA_TRY(h, cudaMemcpyAsync(stats_out, h->dstats, (size_t)m * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    unsigned int wait_err = 0;
    if (h->exchange_mode == 3)
        CUDA_TRY(h, cudaMemcpyAsync(&wait_err, h->dflags + 3, sizeof(wait_err), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->enqueued = 0;
    if (wait_err) return fail(h, PDP_ECUDA, "peer halo exchange: a neighbouring rank did not deliver its planes within the time limit");
    return PDP_OK;
}

extern "C" int pdp_sweep(pdp_handle* h, int32_t n_sweeps, pdp_stats* stats_out) {
    CHECK_HANDLE(h);
    if (n_sweeps < 0) return fail(h, PDP_EINVAL, "pdp_sweep: n_sweeps < 0");
    if (h->enqueued) return fail(h, PDP_ESTATE, "pdp_sweep: collect the enqueued sweeps first (pdp_sweep_collect)");
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    for (int k = 0; k < n_sweeps; ++k) {
        int rc = pdp_sweep_enqueue(h);
        if (rc != PDP_OK) { h->enqueued = 0; return rc; }
    }
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    int rc = pdp_sweep_collect(h, stats_out, n_sweeps, nullptr);
    if (rc != PDP_OK) return rc;
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    return PDP_OK;
}

// ---- one sweep with HOST arrays on both sides (the reference's own calling convention: J_next is a
// NumPy array going in, J and pi are NumPy arrays coming out, dynamicprogramming.py:181-236) ----------
// The grid is cut into chunks of axis-0 planes and the three stages are pipelined on three streams:
// upload of J_next planes -> backup of a chunk as soon as the planes it can read (chunk + halo) have
// arrived -> download of the chunk's J and pi while the next chunks compute.  With pinned host
// buffers the PCIe transfers in both directions overlap the kernels; pageable buffers work but
// serialise.  Afterwards the handle's state is as after pdp_set_J(J_next) + pdp_sweep(1).
#define PDP_HOST_MAX_CHUNKS 64

// Enqueue the whole pipeline of one host-array sweep; every stream it uses has rejoined h->stream when it
// returns, and the per-chunk statistics land in the handle's pinned host buffer.  Runs either directly or
// under stream capture (then nothing executes and the work becomes a CUDA graph).
static int host_pipeline_enqueue(pdp_handle* h, const double* J_next_host, long long host_plane0, double* J_host, int64_t* pi_host, int C) {
    // upload chunks partition the planes the handle holds (slab + halo), backup chunks the planes it computes
    std::vector<int> ub(C + 1), sb(C + 1);
    for (int i = 0; i <= C; ++i) {
        ub[i] = h->alloc_begin + (int)((long long)i * (h->alloc_end - h->alloc_begin) / C);
        sb[i] = h->slab_begin + (int)((long long)i * (h->slab_end - h->slab_begin) / C);
    }
    double* Jc = h->dJ[h->cur_idx];       // element 0 = plane alloc_begin
    double* Jw = h->dJ[1 - h->cur_idx];
    // uploads start once earlier work of the handle's stream (which may read J[cur]) is done
    CUDA_TRY(h, cudaEventRecord(h->ev_start, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->h2d_stream, h->ev_start, 0));
    CUDA_TRY(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_start, 0));
    // The chunk backups are independent of each other (all read J[cur], each writes its own planes), so they
    // rotate over three streams with their own statistics scratch: a chunk that cannot fill the 148 SMs
    // shares them with the next one instead of leaving a partial wave behind.
    cudaStream_t cs[3] = {h->stream, h->side_stream[0], h->side_stream[1]};
    const int K = std::min(3, C);
    for (int i = 1; i < K; ++i) CUDA_TRY(h, cudaStreamWaitEvent(cs[i], h->ev_start, 0));
    for (int i = 0; i < C; ++i) {
        const size_t cnt = (size_t)(ub[i + 1] - ub[i]) * h->plane;
        if (cnt) CUDA_TRY(h, cudaMemcpyAsync(Jc + (size_t)(ub[i] - h->alloc_begin) * h->plane, J_next_host + (size_t)(ub[i] - host_plane0) * h->plane,
                                             cnt * sizeof(double), cudaMemcpyHostToDevice, h->h2d_stream));
        CUDA_TRY(h, cudaEventRecord(h->ev_up[i], h->h2d_stream));
    }
    int waited[3] = {-1, -1, -1};  // uploads [0..waited] are already ordered before that compute stream
    for (int i = 0; i < C; ++i) {
        // the chunk's backups read planes < sb[i+1] + halo_hi (and >= sb[i] - halo_lo: uploaded earlier, uploads are in order)
        const int top = std::min(h->alloc_end, sb[i + 1] + h->halo_hi) - 1;
        int need = 0;
        while (need + 1 < C && ub[need + 1] <= top) ++need;
        const int si = i % K;
        if (need > waited[si]) { CUDA_TRY(h, cudaStreamWaitEvent(cs[si], h->ev_up[need], 0)); waited[si] = need; }
        int rc = launch_planes(h, sb[i], sb[i + 1], si, h->dchunk_stats + 3 * i, cs[si]);
        if (rc != PDP_OK) return rc;
        CUDA_TRY(h, cudaEventRecord(h->ev_done[i], cs[si]));
        CUDA_TRY(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_done[i], 0));
        const size_t off = (size_t)(sb[i] - h->slab_begin) * h->plane, cnt = (size_t)(sb[i + 1] - sb[i]) * h->plane;
        if (cnt) {
            CUDA_TRY(h, cudaMemcpyAsync(J_host + off, Jw + (size_t)(sb[i] - h->alloc_begin) * h->plane, cnt * sizeof(double),
                                        cudaMemcpyDeviceToHost, h->d2h_stream));
            CUDA_TRY(h, cudaMemcpyAsync(pi_host + off, h->dpi + off, cnt * sizeof(long long), cudaMemcpyDeviceToHost, h->d2h_stream));
        }
    }
    // join: side compute streams, the upload stream (it has no successor otherwise) and the download stream
    for (int i = 1; i < K; ++i) {
        CUDA_TRY(h, cudaEventRecord(h->ev_side[i - 1], cs[i]));
        CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_side[i - 1], 0));
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->h_chunk_stats, h->dchunk_stats, (size_t)C * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_join[0], h->h2d_stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_join[0], 0));
    CUDA_TRY(h, cudaEventRecord(h->ev_join[1], h->d2h_stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_join[1], 0));
    return PDP_OK;
}

static bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int sweep_host_impl(pdp_handle* h, const double* J_next_host, long long host_plane0, double* J_host, int64_t* pi_host,
                           pdp_stats* stats_out);

extern "C" int pdp_sweep_host(pdp_handle* h, const double* J_next_host, double* J_host, int64_t* pi_host, pdp_stats* stats_out) {
    return sweep_host_impl(h, J_next_host, 0, J_host, pi_host, stats_out);
}

// The same with a host input that holds ONLY the planes this handle keeps (slab + halo, pdp_slab_layout: planes
// [alloc_begin, alloc_end)): what a rank of a sharded run passes, so that no rank needs the full (N,) array in host memory.
extern "C" int pdp_sweep_host_local(pdp_handle* h, const double* J_held_host, double* J_host, int64_t* pi_host, pdp_stats* stats_out) {
    return sweep_host_impl(h, J_held_host, h ? h->alloc_begin : 0, J_host, pi_host, stats_out);
}

static int sweep_host_impl(pdp_handle* h, const double* J_next_host, long long host_plane0, double* J_host, int64_t* pi_host,
                           pdp_stats* stats_out) {
    CHECK_HANDLE(h);
    if (!J_next_host || !J_host || !pi_host) return fail(h, PDP_EINVAL, "pdp_sweep_host: null pointer");
    if (h->slab_end <= h->slab_begin) return fail(h, PDP_ESTATE, "pdp_sweep_host: this handle computes no planes");
    if (h->enqueued || h->pending) return fail(h, PDP_ESTATE, "pdp_sweep_host: collect / commit the outstanding sweeps first");
    if (h->P.system_id == PDP_SYS_LUT && !h->have_lut) return fail(h, PDP_ESTATE, "pdp_sweep: LUT mode needs pdp_set_lut first");
    if (h->spline) return fail(h, PDP_ENOTSUP, "pdp_sweep_host: the spline interpolant is fitted on the whole J_next before a sweep; use pdp_set_J + pdp_sweep");
    int C = 8;
    if (const char* env = getenv("PYRODP_HOST_CHUNKS")) C = atoi(env);
    C = std::max(1, std::min(std::min(C, PDP_HOST_MAX_CHUNKS), h->slab_end - h->slab_begin));
    if (!h->h2d_stream) {
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
        CUDA_TRY(h, cudaMalloc(&h->dchunk_stats, PDP_HOST_MAX_CHUNKS * 3 * sizeof(double)));
        CUDA_TRY(h, cudaMallocHost(&h->h_chunk_stats, PDP_HOST_MAX_CHUNKS * 3 * sizeof(double)));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(h, cudaStreamCreateWithFlags(&h->side_stream[i], cudaStreamNonBlocking));
            CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_side[i], cudaEventDisableTiming));
            CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
        }
    }
    while ((int)h->ev_up.size() < C) {
        cudaEvent_t a = nullptr, b = nullptr;
        CUDA_TRY(h, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        h->ev_up.push_back(a);
        CUDA_TRY(h, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        h->ev_done.push_back(b);
    }
    // With pinned host buffers the pipeline (≈ 8 x {2 copies, 1 kernel, 4 event operations} + joins) is captured
    // once per {buffers, J parity} into a CUDA graph and replayed with ONE launch per call: the host cost of a
    // step no longer depends on the chunk count.  Pageable buffers (or PYRODP_HOST_GRAPH=0) enqueue directly.
    bool use_graph = is_pinned_host(J_next_host) && is_pinned_host(J_host) && is_pinned_host(pi_host);
    if (const char* env = getenv("PYRODP_HOST_GRAPH")) use_graph = use_graph && atoi(env) != 0;
    h->have_J = true;
    if (use_graph) {
        pdp_handle::HostGraph* g = nullptr;
        for (auto& c : h->host_graphs)
            if (c.jin == J_next_host && c.jout == J_host && c.piout == pi_host && c.cur_idx == h->cur_idx && c.chunks == C) g = &c;
        if (!g) {
            const long long launches0 = h->launches;
            cudaGraph_t graph = nullptr;
            int rc = PDP_OK;
            // (the legacy default stream cannot be captured: then the direct path runs)
            cudaError_t ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
            if (ce == cudaSuccess) {
                rc = host_pipeline_enqueue(h, J_next_host, host_plane0, J_host, pi_host, C);
                ce = cudaStreamEndCapture(h->stream, &graph);
            }
            h->launches = launches0;
            if (rc != PDP_OK || ce != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                h->sticky = 0;              // a failed capture leaves the device healthy: fall back to the direct path
                use_graph = false;
            } else {
                cudaGraphExec_t exec = nullptr;
                ce = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) { cudaGetLastError(); use_graph = false; }
                else {
                    if (h->host_graphs.size() >= 8) {   // bounded cache: drop the oldest
                        cudaGraphExecDestroy(h->host_graphs.front().exec);
                        h->host_graphs.erase(h->host_graphs.begin());
                    }
                    h->host_graphs.push_back({J_next_host, J_host, pi_host, h->cur_idx, C, exec});
                    g = &h->host_graphs.back();
                }
            }
        }
        if (use_graph) {
            CUDA_TRY(h, cudaGraphLaunch(g->exec, h->stream));
            h->launches += C;
        }
    }
    if (!use_graph) {
        int rc = host_pipeline_enqueue(h, J_next_host, host_plane0, J_host, pi_host, C);
        if (rc != PDP_OK) return rc;
    }
    h->cur_idx = 1 - h->cur_idx;
    // a slab handle has just rewritten its own planes only: the halo copies of the new J are its neighbours' to give
    h->halo_stale = h->slab_nodes() != h->N;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (stats_out) {
        const double* cst = h->h_chunk_stats;
        pdp_stats s = {cst[0], cst[1], cst[2]};
        for (int i = 1; i < C; ++i) {
            s.j_max = std::max(s.j_max, cst[3 * i]);
            s.delta_max = std::max(s.delta_max, cst[3 * i + 1]);
            s.delta_min = std::min(s.delta_min, cst[3 * i + 2]);
        }
        *stats_out = s;
    }
    return PDP_OK;
}

// exchange the halo planes of the CURRENT J (after pdp_set_J with rank-local data or pdp_clean_infeasible_set)
extern "C" int pdp_exchange_current(pdp_handle* h) {
    CHECK_HANDLE(h);
    int rc = exchange(h, h->cur_idx, h->stream);
    if (rc != PDP_OK) return rc;
    if (h->comm && h->world > 1) h->halo_stale = false;
    if (h->exchange_mode == 3 && h->world > 1) {
        halo_wait_kernel<<<1, 1, 0, h->stream>>>(h->dflags, h->peer_seq, h->rank > 0, h->rank < h->world - 1, h->dflags + 3,
                                                 20000000000LL);
        CUDA_TRY(h, cudaGetLastError());
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PDP_OK;
}

// ---- peer-memory halo exchange: export / attach -------------------------------------------------------
// out200 = {cudaIpcMemHandle_t J buffer 0, J buffer 1, flags (3 x 64 bytes), int32 alloc_begin, int32 pad}
extern "C" int pdp_peer_export(pdp_handle* h, void* out200) {
    CHECK_HANDLE(h);
    if (!out200) return fail(h, PDP_EINVAL, "pdp_peer_export: null pointer");
    if (!h->dflags) {
        CUDA_TRY(h, cudaMalloc(&h->dflags, 4 * sizeof(unsigned int)));
        CUDA_TRY(h, cudaMemset(h->dflags, 0, 4 * sizeof(unsigned int)));
    }
    cudaIpcMemHandle_t hd[3];
    CUDA_TRY(h, cudaIpcGetMemHandle(&hd[0], h->dJ[0]));
    CUDA_TRY(h, cudaIpcGetMemHandle(&hd[1], h->dJ[1]));
    CUDA_TRY(h, cudaIpcGetMemHandle(&hd[2], h->dflags));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    memcpy(out200, hd, sizeof(hd));
    int32_t tail[2] = {h->alloc_begin, 0};
    memcpy((char*)out200 + sizeof(hd), tail, sizeof(tail));
    return PDP_OK;
}

// lower200 / upper200: what pdp_peer_export produced on rank-1 / rank+1 (NULL at the ends).  Needs the halo
// layout (slab + halo buffers) and a communicator for the statistics all-reduce (pdp_comm_init, halo mode).
extern "C" int pdp_peer_attach(pdp_handle* h, const void* lower200, const void* upper200) {
    CHECK_HANDLE(h);
    if (!h->comm || h->exchange_mode != 1) return fail(h, PDP_ESTATE, "pdp_peer_attach: call pdp_comm_init in halo mode first");
    if (!h->dflags) return fail(h, PDP_ESTATE, "pdp_peer_attach: call pdp_peer_export first");
    if ((h->rank > 0) != (lower200 != nullptr) || (h->rank < h->world - 1) != (upper200 != nullptr))
        return fail(h, PDP_EINVAL, "pdp_peer_attach: neighbour handles do not match the rank's position");
    const void* src[2] = {lower200, upper200};
    for (int side = 0; side < 2; ++side) {
        if (!src[side]) continue;
        cudaIpcMemHandle_t hd[3];
        int32_t tail[2];
        memcpy(hd, src[side], sizeof(hd));
        memcpy(tail, (const char*)src[side] + sizeof(hd), sizeof(tail));
        CUDA_TRY(h, cudaIpcOpenMemHandle((void**)&h->peer_J[side][0], hd[0], cudaIpcMemLazyEnablePeerAccess));
        CUDA_TRY(h, cudaIpcOpenMemHandle((void**)&h->peer_J[side][1], hd[1], cudaIpcMemLazyEnablePeerAccess));
        CUDA_TRY(h, cudaIpcOpenMemHandle((void**)&h->peer_flags[side], hd[2], cudaIpcMemLazyEnablePeerAccess));
        h->peer_alloc_begin[side] = tail[0];
    }
    h->exchange_mode = 3;
    return PDP_OK;
}

// Dense look-up tables of a node range, built on the device (the step before the sweep for callers that want the
// reference's tables: discretizer.py:342-376, dynamicprogramming.py:517-553).  Host outputs, any may be NULL:
// x_next (count*A*n doubles), x_next_isok (count*A bytes), G (count*A doubles).
extern "C" int pdp_build_tables(pdp_handle* h, int64_t node_begin, int64_t count, double* x_next_host, uint8_t* x_ok_host, double* G_host) {
    CHECK_HANDLE(h);
    if (h->P.system_id == PDP_SYS_LUT) return fail(h, PDP_ENOTSUP, "pdp_build_tables: needs a fused system (LUT-mode handles are given their tables)");
    if (!h->P.all_act_ok) return fail(h, PDP_ENOTSUP, "pdp_build_tables: a disallowed action is folded into the B.u table; build the tables on the host");
    if (node_begin < 0 || count < 0 || node_begin + count > h->N) return fail(h, PDP_EINVAL, "pdp_build_tables: node range outside the grid");
    const long long pairs = count * (long long)h->A;
    if (pairs == 0) return PDP_OK;
    if (pairs > (1LL << 31) * 256) return fail(h, PDP_EINVAL, "pdp_build_tables: range too large for one call");
    double *dx = nullptr, *dG = nullptr;
    unsigned char* dok = nullptr;
    cudaError_t e = cudaSuccess;
    if (x_next_host) e = cudaMalloc(&dx, (size_t)pairs * h->P.n * sizeof(double));
    if (e == cudaSuccess && x_ok_host) e = cudaMalloc(&dok, (size_t)pairs);
    if (e == cudaSuccess && G_host) e = cudaMalloc(&dG, (size_t)pairs * sizeof(double));
    if (e == cudaSuccess) {
        const unsigned blocks = (unsigned)((pairs + 255) / 256);
        if (h->P.n == 2) build_tables_kernel<2><<<blocks, 256, 0, h->stream>>>(h->P, node_begin, count, dx, dok, dG);
        else build_tables_kernel<4><<<blocks, 256, 0, h->stream>>>(h->P, node_begin, count, dx, dok, dG);
        e = cudaGetLastError();
        h->launches += 1;
    }
    if (e == cudaSuccess && dx) e = cudaMemcpyAsync(x_next_host, dx, (size_t)pairs * h->P.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && dok) e = cudaMemcpyAsync(x_ok_host, dok, (size_t)pairs, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && dG) e = cudaMemcpyAsync(G_host, dG, (size_t)pairs * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dx); cudaFree(dok); cudaFree(dG);
    if (e != cudaSuccess) return fail(h, PDP_ECUDA, std::string("pdp_build_tables: ") + cudaGetErrorString(e));
    return PDP_OK;
}

// u_k of this handle's slab (slab_nodes doubles)
extern "C" int pdp_get_input_from_policy(pdp_handle* h, int32_t k, double* uk_host) {
    CHECK_HANDLE(h);
    if (k < 0 || k >= h->P.m) return fail(h, PDP_EINVAL, "pdp_get_input_from_policy: input axis out of range");
    if (!uk_host) return fail(h, PDP_EINVAL, "pdp_get_input_from_policy: null pointer");
    const long long n = h->slab_nodes();
    if (n == 0) return PDP_OK;
    double* tmp = nullptr;
    CUDA_TRY(h, cudaMalloc(&tmp, n * sizeof(double)));
    const int threads = 256;
    input_from_policy_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, h->stream>>>(h->dpi, h->P.u_flat, h->P.m, k, tmp, n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(uk_host, tmp, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(h, PDP_ECUDA, std::string("pdp_get_input_from_policy: ") + cudaGetErrorString(e));
    return PDP_OK;
}

extern "C" int pdp_clean_infeasible_set(pdp_handle* h, double tol, int64_t default_action) {
    CHECK_HANDLE(h);
    if (default_action < 0 || default_action >= h->A) return fail(h, PDP_EINVAL, "pdp_clean_infeasible_set: default action out of range");
    const long long n = h->slab_nodes();
    if (n == 0) return PDP_OK;
    const int threads = 256;
    clean_infeasible_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, h->stream>>>(
        h->Jv(h->cur_idx) + h->P.slab_node_begin, h->dpi, h->P.INF - tol, h->P.INF, (long long)default_action, n);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PDP_OK;
}

// Batches of closed-loop Euler rollouts under the handle's current policy (rollout.cuh)
extern "C" int pdp_rollout(pdp_handle* h, const double* phys, const double* x0_host, int64_t B, int32_t npts, double dt,
                           int32_t stride, double* x_out_host, double* u_out_host) {
    CHECK_HANDLE(h);
    const DevProblem& P = h->P;
    if (P.system_id == PDP_SYS_LUT) return fail(h, PDP_ENOTSUP, "pdp_rollout: needs a fused system (the plant's f is evaluated on the device)");
    if (h->slab_begin != 0 || h->slab_end != P.dims[0]) return fail(h, PDP_ENOTSUP, "pdp_rollout: the handle must hold the whole grid's policy");
    if (!phys || !x0_host || !x_out_host) return fail(h, PDP_EINVAL, "pdp_rollout: null pointer");
    if (B <= 0 || npts < 1 || stride < 1 || !(dt > 0.0)) return fail(h, PDP_EINVAL, "pdp_rollout: B, npts, stride, dt must be positive");
    const int n = P.n, m = P.m;
    const long long keep = (npts - 1) / stride + 1;
    double *dphys = nullptr, *dx0 = nullptr, *dx = nullptr, *du = nullptr;
    cudaError_t e = cudaMalloc(&dphys, PDP_ROLLOUT_PHYS * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&dx0, (size_t)B * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&dx, (size_t)keep * n * B * sizeof(double));
    if (e == cudaSuccess && u_out_host) e = cudaMalloc(&du, (size_t)keep * m * B * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dphys, phys, PDP_ROLLOUT_PHYS * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dx0, x0_host, (size_t)B * n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        const unsigned blocks = (unsigned)((B + 127) / 128);
        if (n == 2) rollout_kernel<2><<<blocks, 128, 0, h->stream>>>(P, h->dpi, dphys, dx0, B, npts, dt, stride, dx, du);
        else rollout_kernel<4><<<blocks, 128, 0, h->stream>>>(P, h->dpi, dphys, dx0, B, npts, dt, stride, dx, du);
        h->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(x_out_host, dx, (size_t)keep * n * B * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && du) e = cudaMemcpyAsync(u_out_host, du, (size_t)keep * m * B * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dphys); cudaFree(dx0); cudaFree(dx); cudaFree(du);
    if (e == cudaErrorMemoryAllocation) {   // a batch that does not fit is the caller's argument, not a broken handle
        cudaGetLastError();
        return fail(h, PDP_EINVAL, "pdp_rollout: the kept points (n_keep x (n + m) x B doubles) do not fit the device memory; use a larger stride or fewer trajectories");
    }
    if (e != cudaSuccess) return fail(h, PDP_ECUDA, std::string("pdp_rollout: ") + cudaGetErrorString(e));
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

#include "multi.cuh"
