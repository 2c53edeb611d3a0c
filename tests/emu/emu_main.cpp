// CPU run of the product's fused sweep kernels (TEST INFRASTRUCTURE ONLY; see cuda_emu.h).
//   build: tests/emu/build.sh  ->  tests/emu/libpyrodp_emu.so
// Entry point: emu_sweep(const pdp_problem*, J_next, J, pi, stats, lanes, force_generic) — the same descriptor the
// CUDA library takes (include/pyrodp.h), one sweep of the whole grid.
#include "cuda_emu.h"

#include <algorithm>
#include <string>

#include "../../include/pyrodp.h"

#include <stdexcept>

emu_uint3 threadIdx, blockIdx, blockDim, gridDim;
EmuBlock* emu_block = nullptr;
double smem[32 * 1024] __attribute__((aligned(16)));   // the dynamic shared memory of the running block (256 KB)

// ---- coroutines: a minimal x86-64 System V context switch (callee-saved registers + stack pointer) --------------
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace {
struct Co {
    void* sp = nullptr;
    bool done = false;
    EmuGroup* wait_group = nullptr;     // parked at a barrier of this group ...
    unsigned long long wait_gen = 0;    // ... until its generation moves past this value
    emu_uint3 tid{};
};
constexpr size_t kStack = 256 * 1024;
std::vector<Co> g_co;
std::vector<char> g_stacks;
void* g_sched_sp = nullptr;
int g_cur = -1;
const std::function<void()>* g_body = nullptr;

void co_yield_to_scheduler() { emu_switch(&g_co[g_cur].sp, g_sched_sp); }

void co_entry() {
    (*g_body)();
    g_co[g_cur].done = true;
    co_yield_to_scheduler();
    __builtin_trap();   // a finished coroutine is never resumed
}
}  // namespace

void emu_barrier(EmuGroup& g) {
    if (++g.arrived == g.size) {        // last to arrive: release everybody, go on
        g.arrived = 0;
        ++g.generation;
        return;
    }
    Co& me = g_co[g_cur];
    me.wait_group = &g;
    me.wait_gen = g.generation;
    co_yield_to_scheduler();            // resumed by the scheduler once the generation has moved
    me.wait_group = nullptr;
}

void emu_launch(emu_uint3 grid, emu_uint3 block, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads % 32) throw std::runtime_error("emulator: block size must be a multiple of 32");
    blockDim = block;
    gridDim = grid;
    g_body = &body;
    g_stacks.assign((size_t)nthreads * kStack + 64, 0);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                EmuBlock blk;
                blk.all.size = nthreads;
                blk.warps.resize(nthreads / 32);
                for (auto& w : blk.warps) w.size = 32;
                emu_block = &blk;
                blockIdx = {bx, by, bz};
                g_co.assign(nthreads, Co());
                for (int t = 0; t < nthreads; ++t) {
                    uintptr_t top = ((uintptr_t)(g_stacks.data() + (size_t)(t + 1) * kStack)) & ~(uintptr_t)15;
                    void** sp = (void**)(top - 64);          // 6 callee-saved registers, entry address, dummy return
                    for (int i = 0; i < 8; ++i) sp[i] = nullptr;
                    sp[6] = (void*)&co_entry;
                    g_co[t].sp = sp;
                    g_co[t].tid = {(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                }
                int remaining = nthreads;
                while (remaining > 0) {
                    bool progressed = false;
                    for (int t = 0; t < nthreads; ++t) {
                        Co& c = g_co[t];
                        if (c.done) continue;
                        if (c.wait_group && c.wait_group->generation == c.wait_gen) continue;   // still parked
                        g_cur = t;
                        threadIdx = c.tid;
                        emu_switch(&g_sched_sp, c.sp);
                        progressed = true;
                        if (c.done) --remaining;
                    }
                    if (!progressed) throw std::runtime_error("emulator: deadlock (a barrier some thread never reaches)");
                }
            }
    emu_block = nullptr;
    g_body = nullptr;
}

#include "gen/pyrodp_device.cuh"
#include "gen/sweep_fused.cuh"
#include "gen/table_kernels.cuh"
#define MECH2_COUNT_EVALS
long long mech2_evals_done = 0;
extern "C" long long emu_mech2_evals(int reset) { long long v = mech2_evals_done; if (reset) mech2_evals_done = 0; return v; }
#include "gen/sweep_mech2.cuh"
#include "gen/mech2_plan.h"
#include "gen/rollout.cuh"
#include "gen/spline.cuh"

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

// The device problem exactly as pdp_create lays it out (pyro_b200/csrc/pyrodp.cu), with host pointers.
struct HostProblem {
    DevProblem P{};
    std::vector<std::vector<double>> keep;
    bool mono = false;
    Mech2Plan plan{};
    const double* hold(const double* src, size_t n) { keep.emplace_back(src, src + n); return keep.back().data(); }
};

static void fill(const pdp_problem* p, HostProblem& H) {
    DevProblem& P = H.P;
    long long N = 1, A = 1;
    for (int d = 0; d < p->n; ++d) N *= p->dims[d];
    for (int d = 0; d < p->m; ++d) A *= p->udims[d];
    P.n = p->n; P.m = p->m; P.dof = p->n / 2; P.A = (int)A;
    P.system_id = p->system_id; P.cost_id = p->cost_id; P.ontarget_check = p->ontarget_check;
    P.alpha_is_one = (p->alpha == 1.0);
    P.dt = p->dt; P.alpha = p->alpha; P.INF = p->INF; P.EPS = p->EPS;
    long long stride = 1;
    for (int d = p->n - 1; d >= 0; --d) { P.stride[d] = stride; stride *= p->dims[d]; }
    P.N = N;
    const long long plane = N / p->dims[0];
    P.node_begin = 0; P.node_end = N; P.slab_node_begin = 0; P.plane_begin = 0;
    (void)plane;
    for (int d = 0; d < p->n; ++d) {
        P.dims[d] = p->dims[d];
        P.lb[d] = p->x_lb[d]; P.ub[d] = p->x_ub[d];
        P.inv_step[d] = (double)(p->dims[d] - 1) / (p->x_ub[d] - p->x_lb[d]);
        std::vector<double> rinv(p->dims[d]);
        for (int i = 0; i + 1 < p->dims[d]; ++i) rinv[i] = 1.0 / (p->x_level[d][i + 1] - p->x_level[d][i]);
        rinv[p->dims[d] - 1] = 0.0;
        P.level[d] = H.hold(p->x_level[d], p->dims[d]);
        P.rinv[d] = H.hold(rinv.data(), rinv.size());
    }
    memcpy(P.Q, p->Q, sizeof(P.Q)); memcpy(P.S, p->S, sizeof(P.S));
    memcpy(P.xbar, p->xbar, sizeof(P.xbar)); memcpy(P.par, p->sys_par, sizeof(P.par));
    for (int t = 0; t < 4; ++t)
        if (p->sys_tab[t] && p->sys_tab_len[t] > 0) P.tab[t] = H.hold(p->sys_tab[t], (size_t)p->sys_tab_len[t]);
    P.all_act_ok = 1;
    if (p->system_id == PDP_SYS_LUT) return;
    std::vector<double> bu(p->bu, p->bu + (size_t)A * P.dof);
    for (long long a = 0; a < A; ++a)
        if (!p->act_ok[a]) {
            P.all_act_ok = 0;
            for (int d = 0; d < P.dof; ++d) bu[(size_t)a * P.dof + d] = __builtin_nan("");
        }
    if (p->system_id == PDP_SYS_PENDULUM) {
        bool asc = P.all_act_ok && p->sys_par[0] > 0.0 && p->dt > 0.0;
        for (long long a = 1; a < A && asc; ++a) asc = bu[a] >= bu[a - 1];
        H.mono = asc;
    }
    if (p->system_id == PDP_SYS_TWOLINK || p->system_id == PDP_SYS_CARTPOLE) {
        H.plan = mech2_plan(p->system_id == PDP_SYS_TWOLINK, p->udims, bu.data(), A, P.all_act_ok, p->dt);
        P.A0 = H.plan.A0; P.A1 = H.plan.A1;
        P.uv_first = H.plan.uv_first; P.uv_inv_step = H.plan.uv_inv_step;
    }
    P.bu = H.hold(bu.data(), bu.size());
    P.gu = H.hold(p->gu, (size_t)A);
    std::vector<double> u_flat((size_t)A * p->m);   // input_from_action_id, as pdp_create builds it
    for (long long a = 0; a < A; ++a) {
        if (p->m == 1) u_flat[a] = p->u_level[0][a];
        else { u_flat[2 * a] = p->u_level[0][a / p->udims[1]]; u_flat[2 * a + 1] = p->u_level[1][a % p->udims[1]]; }
    }
    P.u_flat = H.hold(u_flat.data(), u_flat.size());
}

// One backup of axis-0 planes [p0, p1) — the launch pyrodp.cu's launch_planes() makes for a rank's slab (or for a
// boundary / interior sub-range of it).  J_next, J and pi are indexed by the global node id.
extern "C" int emu_sweep_planes(const pdp_problem* p, const double* J_next, double* J, long long* pi, double* stats3, int lanes,
                                int force_generic, int p0, int p1) {
    if (!p || p->system_id == PDP_SYS_LUT) return -1;
    try {
        HostProblem H;
        fill(p, H);
        DevProblem& P = H.P;
        const int G = lanes;
        const bool a1 = P.alpha_is_one != 0, nd = (P.system_id == PDP_SYS_PENDULUM) && P.par[1] == 0.0;
        // force_generic: 0 = what the library selects, 1 = order-agnostic kernels, 2 = the pendulum kernel's loop nest (PYRODP_PEND_LOOP=2)
        const int mono = (H.mono && force_generic != 1) ? (force_generic == 2 ? 2 : 1) : 0;
        fused_kernel_t k = nullptr;
        if (G == 1) k = a1 ? fused_for<1, true>(P.system_id, nd, mono) : fused_for<1, false>(P.system_id, nd, mono);
        else if (G == 4) k = a1 ? fused_for<4, true>(P.system_id, nd, mono) : fused_for<4, false>(P.system_id, nd, mono);
        else if (G == 16) k = a1 ? fused_for<16, true>(P.system_id, nd, mono) : fused_for<16, false>(P.system_id, nd, mono);
        if (!k || p0 < 0 || p1 > P.dims[0] || p0 >= p1) return -2;
        const long long plane = P.N / P.dims[0];
        P.node_begin = (long long)p0 * plane;
        P.node_end = (long long)p1 * plane;
        emu_uint3 grid, block = {SWEEP_THREADS, 1, 1};
        if (P.system_id == PDP_SYS_PENDULUM) {
            P.plane_begin = p0;
            grid = {(unsigned)(p1 - p0), (unsigned)(((long long)P.dims[1] * G + SWEEP_THREADS - 1) / SWEEP_THREADS), 1};
        } else {
            P.plane_begin = (long long)p0 * P.dims[1];
            P.chunks = (int)(((long long)P.dims[2] * P.dims[3] * G + SWEEP_THREADS - 1) / SWEEP_THREADS);
            grid = {(unsigned)((p1 - p0) * P.dims[1] * P.chunks), 1, 1};
            P.tile_rows = 1;
            if (G == 1 && H.plan.ok && force_generic != 1) {   // pyrodp.cu select_fused_kernel
                int tr = MECH2_DEFAULT_TILE_ROWS(P.system_id);
                if (const char* env = getenv("PYRODP_TILE_ROWS")) tr = atoi(env);
                if (tr != 1 && tr != 2 && tr != 4 && tr != 8 && tr != 16) tr = 1;
                P.tile_rows = tr;
                if (tr > 1) {
                    const int tc = SWEEP_THREADS / tr;
                    P.chunks = ((P.dims[2] + tr - 1) / tr) * ((P.dims[3] + tc - 1) / tc);
                }
                grid = {(unsigned)((p1 - p0) * P.dims[1] * P.chunks), 1, 1};
                if (P.system_id == PDP_SYS_TWOLINK) k = a1 ? sweep_mech2_range_kernel<PDP_SYS_TWOLINK, true> : sweep_mech2_range_kernel<PDP_SYS_TWOLINK, false>;
                else k = a1 ? sweep_mech2_range_kernel<PDP_SYS_CARTPOLE, true> : sweep_mech2_range_kernel<PDP_SYS_CARTPOLE, false>;
            }
        }
        std::vector<unsigned long long> slots(3 * STATS_SLOTS, 0);
        unsigned int counter = 0;
        emu_launch(grid, block, [&]() { k(P, J_next, J, pi, slots.data(), &counter, stats3); });
        return 0;
    } catch (const std::exception&) {
        return -3;
    }
}

extern "C" int emu_sweep(const pdp_problem* p, const double* J_next, double* J, long long* pi, double* stats3, int lanes,
                         int force_generic) {
    return p ? emu_sweep_planes(p, J_next, J, pi, stats3, lanes, force_generic, 0, p->dims[0]) : -1;
}

// LUT mode (dynamicprogramming.py:557-570) and its one-column special case, policy evaluation (:743-752), launched
// as pyrodp.cu launches them: the streaming kernel when A == 1, else the generic kernel with G = min(32, 2^ceil(log2 A)).
extern "C" int emu_lut_sweep(const pdp_problem* p, const double* J_next, const double* x_next, const double* Gtab, double* J,
                             long long* pi, double* stats3, int grid_blocks) {
    if (!p) return -1;
    try {
        HostProblem H;
        fill(p, H);
        DevProblem& P = H.P;
        std::vector<unsigned long long> slots(3 * STATS_SLOTS, 0);
        unsigned int counter = 0;
        const long long nodes = P.N;
        if (P.A == 1) {
            const long long blocks = std::min<long long>((nodes + 255) / 256, grid_blocks);
            emu_uint3 grid = {(unsigned)blocks, 1, 1}, block = {256, 1, 1};
            emu_launch(grid, block, [&]() {
                if (P.n == 2) sweep_policy_kernel<2, POLICY_U2>(P, J_next, J, pi, x_next, Gtab, slots.data(), &counter, stats3);
                else if (P.n == 3) sweep_policy_kernel<3, POLICY_U4>(P, J_next, J, pi, x_next, Gtab, slots.data(), &counter, stats3);
                else sweep_policy_kernel<4, POLICY_U4>(P, J_next, J, pi, x_next, Gtab, slots.data(), &counter, stats3);
            });
            return 0;
        }
        int G = 1;
        while (G < 32 && G < P.A) G <<= 1;
        const long long blocks = std::min<long long>((nodes * G + SWEEP_THREADS - 1) / SWEEP_THREADS, grid_blocks);
        emu_uint3 grid = {(unsigned)blocks, 1, 1}, block = {SWEEP_THREADS, 1, 1};
#define EMU_LUT(NN, GG) sweep_lut_kernel<NN, GG>(P, J_next, J, pi, x_next, Gtab, slots.data(), &counter, stats3)
#define EMU_LUT_N(NN)                                                                                        \
    switch (G) { case 1: EMU_LUT(NN, 1); break; case 2: EMU_LUT(NN, 2); break; case 4: EMU_LUT(NN, 4); break;  \
                 case 8: EMU_LUT(NN, 8); break; case 16: EMU_LUT(NN, 16); break; default: EMU_LUT(NN, 32); }
        emu_launch(grid, block, [&]() {
            if (P.n == 2) { EMU_LUT_N(2) } else if (P.n == 3) { EMU_LUT_N(3) } else { EMU_LUT_N(4) }
        });
        return 0;
    } catch (const std::exception&) {
        return -3;
    }
}

// evaluate_terminal_cost (dynamicprogramming.py:159-171)
extern "C" int emu_terminal(const pdp_problem* p, double* J, long long* pi) {
    if (!p || p->system_id == PDP_SYS_LUT) return -1;
    try {
        HostProblem H;
        fill(p, H);
        DevProblem& P = H.P;
        emu_uint3 grid = {(unsigned)((P.N + 255) / 256), 1, 1}, block = {256, 1, 1};
        emu_launch(grid, block, [&]() {
            if (P.n == 2) terminal_cost_kernel<2>(P, J, pi, 0, P.N);
            else terminal_cost_kernel<4>(P, J, pi, 0, P.N);
        });
        return 0;
    } catch (const std::exception&) {
        return -3;
    }
}

// pdp_rollout's kernel (closed-loop Euler trajectories under a policy), outputs in the device layout [n_keep][n|m][B]
extern "C" int emu_rollout(const pdp_problem* p, const long long* pi, const double* phys, const double* x0, long long B,
                           int npts, double dt, int stride, double* x_out, double* u_out) {
    if (!p || p->system_id == PDP_SYS_LUT) return -1;
    try {
        HostProblem H;
        fill(p, H);
        DevProblem& P = H.P;
        emu_uint3 grid = {(unsigned)((B + 127) / 128), 1, 1}, block = {128, 1, 1};
        emu_launch(grid, block, [&]() {
            if (P.n == 2) rollout_kernel<2>(P, pi, phys, x0, B, npts, dt, stride, x_out, u_out);
            else rollout_kernel<4>(P, pi, phys, x0, B, npts, dt, stride, x_out, u_out);
        });
        return 0;
    } catch (const std::exception&) {
        return -3;
    }
}

// pdp_set_interpolant(PDP_INTERP_SPLINE3) + one sweep: the plan, the two fit kernels and the spline table sweep as
// launch_planes() issues them.  coef_out (optional): the fitted B-spline coefficients.
extern "C" int emu_spline_sweep(const pdp_problem* p, const double* J_next, const double* x_next, const double* Gtab, double* J,
                                long long* pi, double* stats3, int grid_blocks, double* coef_out) {
    if (!p || p->system_id != PDP_SYS_LUT || p->n != 2) return -1;
    try {
        HostProblem H;
        fill(p, H);
        DevProblem& P = H.P;
        SplineDev S{};
        std::vector<double> knots[2], lu[2], rden[2], coef((size_t)P.dims[0] * P.dims[1]);
        for (int d = 0; d < 2; ++d) {
            if (!spline_plan_axis(P.level[d], P.dims[d], knots[d], lu[d], rden[d])) return -2;
            S.knots[d] = knots[d].data(); S.lu[d] = lu[d].data(); S.rden[d] = rden[d].data(); S.m[d] = P.dims[d];
        }
        S.coef = coef.data();
        emu_uint3 b128 = {128, 1, 1};
        emu_launch(emu_uint3{(unsigned)((P.dims[1] + 127) / 128), 1, 1}, b128, [&]() { spline_fit_axis0_kernel(J_next, S); });
        emu_launch(emu_uint3{(unsigned)((P.dims[0] + 127) / 128), 1, 1}, b128, [&]() { spline_fit_axis1_kernel(S); });
        if (coef_out) memcpy(coef_out, coef.data(), coef.size() * sizeof(double));
        std::vector<unsigned long long> slots(3 * STATS_SLOTS, 0);
        unsigned int counter = 0;
        int G = 1;
        while (G < 32 && G < P.A) G <<= 1;
        const long long blocks = std::min<long long>((P.N * G + SWEEP_THREADS - 1) / SWEEP_THREADS, grid_blocks);
        emu_uint3 grid = {(unsigned)blocks, 1, 1}, block = {SWEEP_THREADS, 1, 1};
#define EMU_SPL(GG) sweep_lut_spline_kernel<GG>(P, S, J_next, J, pi, x_next, Gtab, slots.data(), &counter, stats3)
        emu_launch(grid, block, [&]() {
            switch (G) { case 1: EMU_SPL(1); break; case 2: EMU_SPL(2); break; case 4: EMU_SPL(4); break;
                         case 8: EMU_SPL(8); break; case 16: EMU_SPL(16); break; default: EMU_SPL(32); }
        });
        return 0;
    } catch (const std::exception&) {
        return -3;
    }
}

// pdp_build_tables' kernel: the reference's dense tables of nodes [node0, node0 + count) of a fused system
extern "C" int emu_build_tables(const pdp_problem* p, long long node0, long long count, double* x_next, unsigned char* x_ok, double* G) {
    if (!p || p->system_id == PDP_SYS_LUT) return -1;
    try {
        HostProblem H;
        fill(p, H);
        DevProblem& P = H.P;
        std::vector<unsigned char> act_ok(p->act_ok, p->act_ok + P.A);
        P.act_ok = act_ok.data();
        const long long pairs = count * P.A;
        emu_uint3 grid = {(unsigned)((pairs + 255) / 256), 1, 1}, block = {256, 1, 1};
        emu_launch(grid, block, [&]() {
            if (P.n == 2) build_tables_kernel<2>(P, node0, count, x_next, x_ok, G);
            else build_tables_kernel<4>(P, node0, count, x_next, x_ok, G);
        });
        return 0;
    } catch (const std::exception&) {
        return -3;
    }
}
