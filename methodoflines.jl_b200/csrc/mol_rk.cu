// Explicit Runge-Kutta drivers on top of the fused RHS kernels.
//
// Replaces OrdinaryDiffEq's perform_step! loop for the explicit methods the reference's tests use
// (Tsit5 48x, SSPRK33 17x, Euler 8x).  Stage vectors never leave HBM:
//   * the stage input  u + dt*sum_j a_sj k_j  is formed inside the RHS kernel's loader
//     (MOL_NIN inputs, no separate axpy pass, no tmp array),
//   * Tsit5's stage 6 writes u+ and the partial error estimate instead of k6 (PRE epilogue), stage 7 runs on
//     the single array u+ and accumulates the scaled error norm sum (utilde/(abstol+max(|u|,|u+|)*reltol))^2
//     with warp shuffles (FIN epilogue): 30 array passes per step,
//   * FSAL: k7 of an accepted step is k1 of the next.
// The step controller (PI, OrdinaryDiffEq defaults) runs on the host from one 8-byte readback for large problems, as a
// one-thread kernel behind the last sweep for mid-size ones (attempts queued as a captured graph, solve_queued below), and
// inside the persistent single-CTA solver kernel for small ones (solve_persistent).
// saveat never clips a step: saved states come from dense output (Tsit5 interpolant / cubic Hermite).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "mol_internal.h"

using namespace mol;

namespace {

// ---- small precompiled helpers (PDE independent) ------------------------------------------------
struct CombArgs {
    const double* a[8];
    double c[8];
    int n;
};

// out[i] = sum_j c[j]*a[j][i]; 128-bit accesses, grid-stride.  In-place on a[0] is safe (pointwise).
__global__ void __launch_bounds__(256) mol_combine_kernel(CombArgs A, double* __restrict__ out, int64_t len) {
    const int64_t n2 = len / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        double2 s = make_double2(0.0, 0.0);
#pragma unroll 8
        for (int j = 0; j < A.n; ++j) {
            const double2 v = reinterpret_cast<const double2*>(A.a[j])[i];
            s.x = fma(A.c[j], v.x, s.x);
            s.y = fma(A.c[j], v.y, s.y);
        }
        reinterpret_cast<double2*>(out)[i] = s;
    }
    if ((len & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int j = 0; j < A.n; ++j) s = fma(A.c[j], A.a[j][len - 1], s);
        out[len - 1] = s;
    }
}

// the same combination with 64-bit accesses: arrays that are not 16-byte aligned (a save slot at an odd multiple of an
// odd state length)
__global__ void __launch_bounds__(256) mol_combine_scalar_kernel(CombArgs A, double* __restrict__ out, int64_t len) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
#pragma unroll 8
        for (int j = 0; j < A.n; ++j) s = fma(A.c[j], A.a[j][i], s);
        out[i] = s;
    }
}

// Sum of 32 threads' strided shares of p[0..n) in a fixed order (one warp; xor-shuffle tree): the error norms are formed
// from per-CTA partial sums WITHOUT atomics, so an adaptive solve takes the same steps, bit for bit, every time it runs.
__device__ __forceinline__ double mol_det_sum32(const double* __restrict__ p, int n) {
    double s = 0.0;
    for (int i = (int)threadIdx.x; i < n; i += 32) s += p[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
__global__ void __launch_bounds__(32) mol_err_reduce_kernel(const double* __restrict__ parts, int n, double* out) {
    const double s = mol_det_sum32(parts, n);
    if (threadIdx.x == 0) *out = s;
}

// parts[blockIdx.x] = this CTA's share of sum_i ((ca*a_i + cb*b_i) / (abstol + |u_i|*reltol))^2   (Hairer initial-step norms)
__global__ void __launch_bounds__(256) mol_wrms_kernel(const double* __restrict__ a, const double* __restrict__ b, double ca,
                                                       double cb, const double* __restrict__ u, double abstol, double reltol,
                                                       int64_t len, double* parts) {
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        double v = ca * a[i];
        if (b) v = fma(cb, b[i], v);
        const double r = v / (abstol + fabs(u[i]) * reltol);
        s = fma(r, r, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) parts[blockIdx.x] = v;
    }
}

// ---- queued adaptive Tsit5: step control on the device ------------------------------------------------------------------
// The host-driven loop pays one 8-byte read-back and a stream synchronisation per attempt, and issues its seven launches
// only after it (96 us per step at 512^2, where the sweeps themselves need 30).  For mid-size problems the controller runs
// as a one-thread kernel behind the last sweep instead: it turns the accumulated error norm into accept / reject, the next
// step size and the next time, and the sweeps of the NEXT attempt -- MOL_DEVDT kernel variants, queued by the host long
// before -- read {t, dt, skip} from this block.  One attempt = six sweeps + controller + "advance" (u <- u+, k1 <- k7 when the
// step was accepted), captured once as a CUDA graph and launched in batches; the host looks at the block once per batch.
// All doubles: the MOL_DEVDT kernels address the first three as double[3], and one copy moves the block.
struct RkCtl {
    double t, dt, skip;          // attempt being computed: start time, (clipped) step size; skip != 0: sweeps do nothing
    double dtprop;               // the controller's step size before clipping to t1
    double qold;                 // PI controller memory
    double t1, ttol;
    double save_next;            // next save time (+inf: none); an accepted step that reaches it puts the solve on hold
    double accepted, hold;       // outcome of the last attempt; hold: the host saves before `advance` overwrites u and k1
    double t_prev, dt_used;      // the step just accepted, t_prev -> t (the host's dense output needs it)
    double it, maxiters, naccept, nreject;
    double status;               // 0 running, 1 reached t1, 2 maxiters, 3 step size underflow / NaN, 4 on hold for a save
    double nglobal;
    double eest;                 // scaled error norm of the last attempt (diagnostics)
};

// Same arithmetic as the host loop in mol_rk_solve (pi_accept_factor / pi_reject_factor, clipping, ttol snapping).
// One warp: the 32 threads first add up the CTAs' partial error sums in a fixed order, thread 0 then decides.
__global__ void __launch_bounds__(32) mol_tsit5_control_kernel(RkCtl* c, const double* __restrict__ parts, int nparts) {
    if (c->skip != 0.0) {
        if (threadIdx.x == 0) c->accepted = 0.0;
        return;
    }
    const double errsum = mol_det_sum32(parts, nparts);
    if (threadIdx.x != 0) return;
    const double eest = sqrt(errsum / c->nglobal);
    c->eest = eest;
    c->it += 1.0;
    c->accepted = 0.0;
    const double dtu = c->dt, t = c->t, t1 = c->t1;
    if (!(eest == eest)) { c->status = 3.0; c->skip = 1.0; return; }
    if (eest <= 1.0) {
        const double q = eest > 0 ? fmax(0.1, fmin(5.0, pow(eest, 7.0 / 50) / pow(c->qold, 2.0 / 25) / 0.9)) : 0.1;
        c->qold = fmax(eest, 1e-4);
        const bool clipped = dtu < c->dtprop;
        const double tnew = (fabs((t + dtu) - t1) <= c->ttol) ? t1 : t + dtu;
        c->t_prev = t;
        c->dt_used = dtu;
        c->t = tnew;
        c->accepted = 1.0;
        c->naccept += 1.0;
        if (!clipped || tnew < t1) c->dtprop = dtu / q;
        if (c->save_next <= tnew + c->ttol) { c->hold = 1.0; c->skip = 1.0; c->status = 4.0; }
        else if (!(tnew < t1)) { c->skip = 1.0; c->status = 1.0; }
    } else {
        c->nreject += 1.0;
        c->dtprop = dtu / fmin(5.0, pow(eest, 7.0 / 50) / 0.9);
        if (c->dtprop < 1e-14 * fmax(1.0, fabs(t))) { c->status = 3.0; c->skip = 1.0; return; }
    }
    if (c->status == 0.0 && c->it >= c->maxiters && c->t < t1) { c->status = 2.0; c->skip = 1.0; }
    c->dt = fmin(c->dtprop, t1 - c->t);
}

// u <- u+ and k1 <- k7 after an accepted attempt (FSAL).  force: the host's call after it has served a hold.
__global__ void __launch_bounds__(256) mol_tsit5_advance_kernel(const RkCtl* c, double* __restrict__ u, const double* __restrict__ unew,
                                                                double* __restrict__ k1, const double* __restrict__ k7, int64_t len,
                                                                int force) {
    if (!force && (c->accepted == 0.0 || c->hold != 0.0)) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        u[i] = unew[i];
        k1[i] = k7[i];
    }
}

const double T5_C[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
const double T5_A[7][6] = {
    {0},
    {0.161},
    {-0.008480655492356989, 0.335480655492357},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
const double T5_BT[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                         0.5823571654525552, -0.45808210592918697, 0.015151515151515152};

// Free 4th-order interpolant of the Tsit5 pair (Tsitouras 2011, the dense output OrdinaryDiffEq uses for saveat):
// u(t + theta dt) = u + dt sum_i b_i(theta) k_i.  b_i(1) reproduces row 7 of the tableau (checked in tests).
void tsit5_dense_weights(double th, double b[7]) {
    const double th2 = th * th;
    b[0] = -1.0530884977290216 * th * (th - 1.3299890189751412) * (th2 - 1.4364028541716351 * th + 0.7139816917074209);
    b[1] = 0.1017 * th2 * (th2 - 2.1966568338249754 * th + 1.2949852507374631);
    b[2] = 2.490627285651252793 * th2 * (th2 - 2.38535645472061657 * th + 1.57803468208092486);
    b[3] = -16.54810288924490272 * (th - 1.21712927295533244) * (th - 0.61620406037800089) * th2;
    b[4] = 47.37952196281928122 * (th - 1.203071208372362603) * (th - 0.658047292653547382) * th2;
    b[5] = -34.87065786149660974 * (th - 1.2) * (th - 0.666666666666666667) * th2;
    b[6] = 2.5 * (th - 1.0) * (th - 0.6) * th2;
}

}  // namespace

struct mol_rk {
    mol_plan* plan = nullptr;
    int alg = MOL_ALG_TSIT5;
    double abstol = 1e-6, reltol = 1e-3;
    int64_t n = 0;               // unknowns held by this rank
    int64_t n_global = 0;        // unknowns of the whole problem (error norms are global RMS values)
    double* k[7] = {nullptr};
    double* alt = nullptr;       // second state buffer (ping-pong)
    double* d_err = nullptr;     // one double: the (all-reduced) error sum
    double* d_parts = nullptr;   // MOL_FIN_SLOTS doubles: per-CTA partial sums of an error norm (no atomics: reproducible)
    int nparts = 0;              // slots the last FIN evaluation wrote
    double* h_err = nullptr;     // pinned
    bool fsal_valid = false;
    double qold = 1e-4;
    int64_t nf = 0;
    const double* last_u = nullptr;   // array / time the last accepted step reached (FSAL continuation check)
    double last_t = 0.0;
    double* spare = nullptr;          // persistent solver: one more work array, save times and the result block
    double* d_saveat = nullptr;
    int saveat_cap = 0;
    double* d_out = nullptr;
    double* h_out = nullptr;          // pinned
    // queued adaptive Tsit5 (device-side step control): control block, its pinned mirror, a capturable stream, the
    // captured attempt and the state array it was captured for
    RkCtl* d_ctl = nullptr;
    RkCtl* h_ctl = nullptr;
    cudaStream_t qs = nullptr;
    cudaEvent_t qev = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    std::vector<double> graph_sig;    // everything the capture baked in: array addresses, tolerances, parameter values
};

// Problems up to this many unknowns are integrated by the persistent single-CTA kernel (one launch per solve, step
// control on the device; kernels/mol_generic.cuh MOL_KERNEL_SOLVE) instead of the host-driven loop, whose seven
// launches and one read-back per step (45 us) dwarf the arithmetic of a small problem.  Measured on a B200 per Tsit5
// step: 99 unknowns (config 1) 8.7 us vs 45.3 us; 128-node WENO5 / SSPRK33 20.2 vs 35.9 us; 2048 unknowns (32^2
// Brusselator) 84.6 vs 71.2 us -- one SM no longer wins there, hence the threshold.  MOL_RK_PERSISTENT=0 disables it,
// MOL_RK_PERSISTENT_MAX=<n> moves the threshold.
static int64_t persistent_max_unknowns() {
    const char* e = getenv("MOL_RK_PERSISTENT");
    if (e && *e == '0') return 0;
    const char* m = getenv("MOL_RK_PERSISTENT_MAX");
    return (m && *m) ? atoll(m) : 1024;
}

// Adaptive Tsit5 on problems above that threshold and up to this many unknowns runs with the controller on the device
// (RkCtl above).  The upper bound keeps the two extra array copies per accepted step (u <- u+, k1 <- k7) inside L2 and the
// saved synchronisation worth more than they cost: at 4096^2 (3.4e7 unknowns) a step takes 1.5 ms and the host's 10 us
// no longer matter.  MOL_RK_QUEUED=0 disables it, MOL_RK_QUEUED_MAX=<n> moves the bound.
static int64_t queued_max_unknowns() {
    const char* e = getenv("MOL_RK_QUEUED");
    if (e && *e == '0') return 0;
    const char* m = getenv("MOL_RK_QUEUED_MAX");
    return (m && *m) ? atoll(m) : ((int64_t)1 << 21);
}

static int cuda_fail(cudaError_t e, const char* what) {
    return fail(MOL_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

extern "C" int mol_rk_init(mol_plan* plan, int alg, double abstol, double reltol, mol_rk** out) {
    if (!plan || !out) return fail(MOL_E_ARG, "null argument");
    if (plan->device < 0) return fail(MOL_E_NOCUDA, "plan was created compile-only; there is no CPU fallback");
    if (alg < MOL_ALG_EULER || alg > MOL_ALG_TSIT5) return fail(MOL_E_ARG, "unknown algorithm");
    mol_rk* rk = new mol_rk();
    rk->plan = plan;
    rk->alg = alg;
    rk->abstol = abstol;
    rk->reltol = reltol;
    rk->n = (int64_t)mol_plan_state_len(plan);
    rk->n_global = plan->P.nstate;
    // (Euler needs one stage vector; a second one holds f(u1) for a Hermite-interpolated save point)
    int nk = alg == MOL_ALG_TSIT5 ? 7 : (alg == MOL_ALG_RK4 ? 4 : (alg == MOL_ALG_SSPRK33 ? 3 : 2));
    const bool small = !plan->dist.on && rk->n <= persistent_max_unknowns();
    if (small) nk = 7;                      // the persistent kernel addresses seven stage arrays whatever the method
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < nk && e == cudaSuccess; ++i) e = cudaMalloc(&rk->k[i], rk->n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&rk->alt, rk->n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&rk->d_err, 8);
    if (e == cudaSuccess) e = cudaMalloc(&rk->d_parts, (size_t)MOL_FIN_SLOTS * 8);
    if (e == cudaSuccess) e = cudaMallocHost(&rk->h_err, 8);
    if (small) {
        if (e == cudaSuccess) e = cudaMalloc(&rk->spare, rk->n * 8);
        if (e == cudaSuccess) e = cudaMalloc(&rk->d_out, 8 * 8);
        if (e == cudaSuccess) e = cudaMallocHost(&rk->h_out, 8 * 8);
        for (int i = 0; i < nk && e == cudaSuccess; ++i) e = cudaMemset(rk->k[i], 0, rk->n * 8);   // inputs with coefficient 0
        if (e == cudaSuccess) e = cudaMemset(rk->alt, 0, rk->n * 8);
        if (e == cudaSuccess) e = cudaMemset(rk->spare, 0, rk->n * 8);
    }
    if (e != cudaSuccess) { mol_rk_destroy(rk); return cuda_fail(e, "mol_rk_init allocation"); }
    // slab mode: every resident stage vector carries its own ghost planes (exchanged once per rewrite)
    int rc = MOL_OK;
    for (int i = 0; i < nk && rc == MOL_OK; ++i) rc = mol_dist_register(plan, rk->k[i]);
    if (rc == MOL_OK) rc = mol_dist_register(plan, rk->alt);
    if (rc != MOL_OK) { mol_rk_destroy(rk); return rc; }
    // every kernel variant of this integrator, compiled on several host threads at once (a variant that fails here is
    // simply compiled -- and reported -- when a step first needs it)
    (void)mol_plan_precompile(plan, alg);
    if (alg == MOL_ALG_TSIT5 && !small) {       // the step controller's block (both adaptive loops run the controller kernel)
        e = cudaMalloc(&rk->d_ctl, sizeof(RkCtl));
        if (e == cudaSuccess) e = cudaMallocHost(&rk->h_ctl, sizeof(RkCtl));
        if (e != cudaSuccess) { mol_rk_destroy(rk); return cuda_fail(e, "mol_rk_init (step controller)"); }
    }
    if (alg == MOL_ALG_TSIT5 && !small && !plan->dist.on && rk->n <= queued_max_unknowns()) {
        e = cudaStreamCreateWithFlags(&rk->qs, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&rk->qev, cudaEventDisableTiming);
        if (e != cudaSuccess) { mol_rk_destroy(rk); return cuda_fail(e, "mol_rk_init (queued solve)"); }
        (void)mol_plan_precompile(plan, alg | MOL_ALG_DEVDT);
    }
    *out = rk;
    return MOL_OK;
}

extern "C" int mol_rk_destroy(mol_rk* rk) {
    if (!rk) return MOL_OK;
    for (int i = 0; i < 7; ++i)
        if (rk->k[i]) { mol_dist_unregister(rk->plan, rk->k[i]); cudaFree(rk->k[i]); }
    if (rk->alt) { mol_dist_unregister(rk->plan, rk->alt); cudaFree(rk->alt); }
    if (rk->d_err) cudaFree(rk->d_err);
    if (rk->d_parts) cudaFree(rk->d_parts);
    if (rk->h_err) cudaFreeHost(rk->h_err);
    if (rk->spare) cudaFree(rk->spare);
    if (rk->d_saveat) cudaFree(rk->d_saveat);
    if (rk->d_out) cudaFree(rk->d_out);
    if (rk->h_out) cudaFreeHost(rk->h_out);
    if (rk->gexec) cudaGraphExecDestroy(rk->gexec);
    if (rk->graph) cudaGraphDestroy(rk->graph);
    if (rk->d_ctl) cudaFree(rk->d_ctl);
    if (rk->h_ctl) cudaFreeHost(rk->h_ctl);
    if (rk->qev) cudaEventDestroy(rk->qev);
    if (rk->qs) cudaStreamDestroy(rk->qs);
    delete rk;
    return MOL_OK;
}

extern "C" int mol_rk_set_params(mol_rk* rk, const double* p) {
    if (!rk || !p) return fail(MOL_E_ARG, "null argument");
    for (int i = 0; i < rk->plan->P.nparam; ++i) rk->plan->params[i] = p[i];
    rk->fsal_valid = false;
    return MOL_OK;
}

static int rhs_plain(mol_rk* rk, const double* u, double* out, double t, cudaStream_t st) {
    MolRhsIn in;
    in.nin = 1;
    in.a[0] = u;
    in.c[0] = 1.0;
    MolRhsEpi epi;
    rk->nf++;
    return mol_rhs_launch(rk->plan, in, out, t, epi, st);
}

static int combine(mol_rk* rk, int n, const double* const* a, const double* c, double* out, cudaStream_t st) {
    CombArgs A;
    A.n = n;
    for (int j = 0; j < n; ++j) { A.a[j] = a[j]; A.c[j] = c[j]; }
    for (int j = n; j < 8; ++j) { A.a[j] = nullptr; A.c[j] = 0; }
    bool aligned = reinterpret_cast<uintptr_t>(out) % 16 == 0;
    for (int j = 0; j < n; ++j) aligned = aligned && reinterpret_cast<uintptr_t>(a[j]) % 16 == 0;
    int grid = (int)std::min<int64_t>((rk->n / 2 + 255) / 256 + 1, (int64_t)rk->plan->sm_count * 8);
    if (aligned) mol_combine_kernel<<<grid, 256, 0, st>>>(A, out, rk->n);
    else mol_combine_scalar_kernel<<<grid, 256, 0, st>>>(A, out, rk->n);
    rk->plan->launches++;
    dist_mark_stale(rk->plan, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MOL_OK : cuda_fail(e, "mol_combine_kernel");
}

// fixed-step methods in Butcher form; the final combination is applied in place on u
static int step_fixed(mol_rk* rk, double* u, double t, double dt, cudaStream_t st) {
    int rc;
    MolRhsEpi noepi;
    auto stage = [&](int nin, const double* const* arrs, const double* coefs, double* out, double ts) {
        MolRhsIn in;
        in.nin = nin;
        for (int j = 0; j < nin; ++j) { in.a[j] = arrs[j]; in.c[j] = coefs[j]; }
        rk->nf++;
        return mol_rhs_launch(rk->plan, in, out, ts, noepi, st);
    };
    if (rk->alg == MOL_ALG_EULER) {
        if ((rc = rhs_plain(rk, u, rk->k[0], t, st))) return rc;
        const double* a[2] = {u, rk->k[0]};
        double c[2] = {1.0, dt};
        return combine(rk, 2, a, c, u, st);
    }
    if (rk->alg == MOL_ALG_SSPRK33) {
        // Shu-Osher SSPRK33 == Butcher (c = 0, 1, 1/2; a21 = 1; a31 = a32 = 1/4; b = 1/6, 1/6, 2/3)
        if ((rc = rhs_plain(rk, u, rk->k[0], t, st))) return rc;
        { const double* a[2] = {u, rk->k[0]}; double c[2] = {1.0, dt};
          if ((rc = stage(2, a, c, rk->k[1], t + dt))) return rc; }
        { const double* a[3] = {u, rk->k[0], rk->k[1]}; double c[3] = {1.0, dt / 4, dt / 4};
          if ((rc = stage(3, a, c, rk->k[2], t + dt / 2))) return rc; }
        const double* a[4] = {u, rk->k[0], rk->k[1], rk->k[2]};
        double c[4] = {1.0, dt / 6, dt / 6, 2 * dt / 3};
        return combine(rk, 4, a, c, u, st);
    }
    // RK4
    if ((rc = rhs_plain(rk, u, rk->k[0], t, st))) return rc;
    { const double* a[2] = {u, rk->k[0]}; double c[2] = {1.0, dt / 2};
      if ((rc = stage(2, a, c, rk->k[1], t + dt / 2))) return rc; }
    { const double* a[2] = {u, rk->k[1]}; double c[2] = {1.0, dt / 2};
      if ((rc = stage(2, a, c, rk->k[2], t + dt / 2))) return rc; }
    { const double* a[2] = {u, rk->k[2]}; double c[2] = {1.0, dt};
      if ((rc = stage(2, a, c, rk->k[3], t + dt))) return rc; }
    const double* a[5] = {u, rk->k[0], rk->k[1], rk->k[2], rk->k[3]};
    double c[5] = {1.0, dt / 6, dt / 3, dt / 3, dt / 6};
    return combine(rk, 5, a, c, u, st);
}

// one Tsit5 attempt from (u,t) with step dt: writes u+ into unew, k7 into k[6]; returns EEst (eest == nullptr: no
// read-back, the error sum stays in d_err).
// Stage inputs are formed on load (no axpy passes).  k6 is never stored: the stage-6 sweep (PRE epilogue) writes
// u+ = u + dt sum_j a_7j k_j and the partial error estimate e6 = dt sum_{j<=6} btilde_j k_j (into k[5]'s storage)
// while u, k1..k5 are in registers; the stage-7 sweep (FIN epilogue) then reads the single array u+ through the
// TMA path, stores k7 (the next step's k1, FSAL) and accumulates the scaled error norm from e6, u and u+.
// Array passes per step: (2+1) + (3+1) + (4+1) + (5+1) + (6+2) + (3+1) = 30.
static int tsit5_attempt(mol_rk* rk, const double* u, double* unew, double t, double dt, double* eest, cudaStream_t st) {
    int rc;
    if (!rk->fsal_valid) {
        if ((rc = rhs_plain(rk, u, rk->k[0], t, st))) return rc;
        rk->fsal_valid = true;
    }
    MolRhsEpi noepi;
    for (int s = 1; s <= 4; ++s) {
        MolRhsIn in;
        in.nin = s + 1;
        in.a[0] = u;
        in.c[0] = 1.0;
        for (int j = 0; j < s; ++j) { in.a[j + 1] = rk->k[j]; in.c[j + 1] = dt * T5_A[s][j]; }
        rk->nf++;
        if ((rc = mol_rhs_launch(rk->plan, in, rk->k[s], t + T5_C[s] * dt, noepi, st))) return rc;
    }
    {   // stage 6 (PRE)
        MolRhsIn in;
        in.nin = 6;
        in.a[0] = u;
        in.c[0] = 1.0;
        MolRhsEpi epi;
        epi.mode = MOL_EPI_PRE;
        epi.comb = unew;
        epi.eout = rk->k[5];
        epi.cb[0] = 1.0;
        epi.ce[0] = 0.0;
        for (int j = 0; j < 5; ++j) {
            in.a[j + 1] = rk->k[j];
            in.c[j + 1] = dt * T5_A[5][j];
            epi.cb[j + 1] = dt * T5_A[6][j];
            epi.ce[j + 1] = dt * T5_BT[j];
        }
        epi.cbk = dt * T5_A[6][5];
        epi.cek = dt * T5_BT[5];
        rk->nf++;
        if ((rc = mol_rhs_launch(rk->plan, in, nullptr, t + T5_C[5] * dt, epi, st))) return rc;
    }
    {   // stage 7 (FIN)
        MolRhsIn in;
        in.nin = 1;
        in.a[0] = unew;
        in.c[0] = 1.0;
        MolRhsEpi epi;
        epi.mode = MOL_EPI_FIN;
        epi.e = rk->k[5];
        epi.u0 = u;
        epi.ek = dt * T5_BT[6];
        epi.abstol = rk->abstol;
        epi.reltol = rk->reltol;
        epi.err = rk->d_parts;           // one slot per CTA of every launch of this evaluation
        rk->nf++;
        if ((rc = mol_rhs_launch(rk->plan, in, rk->k[6], t + dt, epi, st))) return rc;
        rk->nparts = rk->plan->fin_slot;
    }
    // adaptive solve on one device: the controller kernel adds the parts up itself
    if (!eest && !rk->plan->dist.on) return MOL_OK;
    mol_err_reduce_kernel<<<1, 32, 0, st>>>(rk->d_parts, rk->nparts, rk->d_err);
    rk->plan->launches++;
    if ((rc = dist_allreduce_sum(rk->plan, rk->d_err, 1, st))) return rc;
    if (!eest) return MOL_OK;        // adaptive solve on slabs: the controller kernel takes the all-reduced sum from d_err
    cudaMemcpyAsync(rk->h_err, rk->d_err, 8, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "tsit5 step");
    *eest = std::sqrt(*rk->h_err / (double)rk->n_global);
    return MOL_OK;
}

static int wrms(mol_rk* rk, const double* a, const double* b, double ca, double cb, const double* u, double* out,
                cudaStream_t st) {
    int grid = (int)std::min<int64_t>((rk->n + 255) / 256, (int64_t)rk->plan->sm_count * 8);
    mol_wrms_kernel<<<grid, 256, 0, st>>>(a, b, ca, cb, u, rk->abstol, rk->reltol, rk->n, rk->d_parts);
    mol_err_reduce_kernel<<<1, 32, 0, st>>>(rk->d_parts, grid, rk->d_err);
    rk->plan->launches += 2;
    int rc = dist_allreduce_sum(rk->plan, rk->d_err, 1, st);
    if (rc != MOL_OK) return rc;
    cudaMemcpyAsync(rk->h_err, rk->d_err, 8, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "wrms");
    *out = std::sqrt(*rk->h_err / (double)rk->n_global);
    return MOL_OK;
}

// Hairer-Norsett-Wanner starting step (what OrdinaryDiffEq's initdt does; 2 RHS calls)
static int initial_dt(mol_rk* rk, const double* u, double t, double* dt_out, cudaStream_t st) {
    int rc;
    double d0, d1, d2;
    if ((rc = rhs_plain(rk, u, rk->k[0], t, st))) return rc;
    rk->fsal_valid = true;
    if ((rc = wrms(rk, u, nullptr, 1.0, 0.0, u, &d0, st))) return rc;
    if ((rc = wrms(rk, rk->k[0], nullptr, 1.0, 0.0, u, &d1, st))) return rc;
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    {   // u1 = u + dt0*f0 fused into the second RHS evaluation's loader
        MolRhsIn in;
        in.nin = 2;
        in.a[0] = u; in.c[0] = 1.0;
        in.a[1] = rk->k[0]; in.c[1] = dt0;
        MolRhsEpi noepi;
        rk->nf++;
        if ((rc = mol_rhs_launch(rk->plan, in, rk->k[1], t + dt0, noepi, st))) return rc;
    }
    if ((rc = wrms(rk, rk->k[1], rk->k[0], 1.0, -1.0, u, &d2, st))) return rc;
    d2 /= dt0;
    const double m = std::max(d1, d2);
    const double dt1 = (m <= 1e-15) ? std::max(1e-6, dt0 * 1e-3) : std::pow(10.0, -(2.0 + std::log10(m)) / 5.0);
    *dt_out = std::min(100 * dt0, dt1);
    return MOL_OK;
}

// PI controller of OrdinaryDiffEq for an accepted step (stepsize_controller! / step_accept_controller!, PIController:
// beta1 = 7/50, beta2 = 2/25, gamma = 9/10, qmin = 1/5, qmax = 10).  The "steady" dead band [qsteady_min, qsteady_max] in
// which dt is kept is [1, 1] for explicit algorithms (6/5 is the default of the adaptive IMPLICIT ones), i.e. empty.
static double pi_accept_factor(mol_rk* rk, double eest) {
    const double q = eest > 0 ? std::max(0.1, std::min(5.0, std::pow(eest, 7.0 / 50) / std::pow(rk->qold, 2.0 / 25) / 0.9)) : 0.1;
    rk->qold = std::max(eest, 1e-4);
    return q;
}
static double pi_reject_factor(double eest) { return std::min(5.0, std::pow(eest, 7.0 / 50) / 0.9); }

// Dense output of the Tsit5 step (u, t) -> (unew, t + dt) that tsit5_attempt just computed, at t + theta dt.
// k6 is never stored (PRE epilogue), but unew = u + dt sum_j a_7j k_j contains it:
//   dt k6 = (unew - u - dt sum_{j<=5} a_7j k_j) / a_76,
// so the interpolant is one 8-array combination of u, unew, k1..k5, k7.
static int combine(mol_rk* rk, int n, const double* const* a, const double* c, double* out, cudaStream_t st);
static int tsit5_dense(mol_rk* rk, const double* u, const double* unew, double dt, double theta, double* out, cudaStream_t st) {
    double b[7];
    tsit5_dense_weights(theta, b);
    const double g = b[5] / T5_A[6][5];
    const double* a[8] = {u, unew, rk->k[0], rk->k[1], rk->k[2], rk->k[3], rk->k[4], rk->k[6]};
    double c[8] = {1.0 - g, g, 0, 0, 0, 0, 0, dt * b[6]};
    for (int j = 0; j < 5; ++j) c[2 + j] = dt * (b[j] - g * T5_A[6][j]);
    return combine(rk, 8, a, c, out, st);
}

// Cubic Hermite interpolant between (u0, f0) and (u1, f1): what OrdinaryDiffEq falls back to for saveat inside a
// step of the methods without their own dense output (Euler, SSPRK33, RK4).
static int hermite_dense(mol_rk* rk, const double* u0, const double* u1, const double* f0, const double* f1, double dt,
                         double th, double* out, cudaStream_t st) {
    const double* a[4] = {u0, u1, f0, f1};
    const double w = th * (th - 1.0);
    double c[4] = {(1.0 - th) - w * (1.0 - 2.0 * th), th + w * (1.0 - 2.0 * th), w * (th - 1.0) * dt, w * th * dt};
    return combine(rk, 4, a, c, out, st);
}

extern "C" int mol_rk_reinit(mol_rk* rk) {
    if (!rk) return fail(MOL_E_ARG, "null argument");
    rk->fsal_valid = false;
    rk->qold = 1e-4;
    rk->last_u = nullptr;
    return MOL_OK;
}

// One step from u_in to u_out (two different arrays: the caller ping-pongs, no state copy).  A rejected adaptive
// step leaves u_out undefined and *t_io unchanged.  FSAL is reused only when this call continues the previous one
// (u_in is the array the last accepted step wrote, at the time it reached); otherwise k1 is re-evaluated.
extern "C" int mol_rk_step_to(mol_rk* rk, const double* u_in, double* u_out, double* t_io, double* dt_io, int adaptive,
                              mol_step_stats* out, void* stream) {
    if (!rk || !u_in || !u_out || !t_io || !dt_io) return fail(MOL_E_ARG, "null argument");
    if (u_in == u_out) return fail(MOL_E_ARG, "mol_rk_step_to needs two different arrays (use mol_rk_step for an in-place step)");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nf0 = rk->nf;
    double t = *t_io, dt = *dt_io;
    int rc;
    if ((rc = mol_dist_register(rk->plan, u_in))) return rc;
    if ((rc = mol_dist_register(rk->plan, u_out))) return rc;
    if (u_in != rk->last_u || t != rk->last_t) {        // not a continuation: stale FSAL stage and ghost planes
        rk->fsal_valid = false;
        dist_mark_stale(rk->plan, u_in);
    }
    mol_step_stats s = {t, dt, 0.0, 1, 0};
    if (rk->alg != MOL_ALG_TSIT5) {
        cudaMemcpyAsync(u_out, u_in, rk->n * 8, cudaMemcpyDeviceToDevice, st);      // the fixed-step forms update in place
        dist_mark_stale(rk->plan, u_out);
        if ((rc = step_fixed(rk, u_out, t, dt, st))) return rc;
        s.t = t + dt;
    } else {
        double eest = 0.0;
        if ((rc = tsit5_attempt(rk, u_in, u_out, t, dt, &eest, st))) return rc;
        s.eest = eest;
        if (!adaptive || eest <= 1.0) {
            std::swap(rk->k[0], rk->k[6]);       // FSAL
            s.t = t + dt;
            if (adaptive) s.dt_next = dt / pi_accept_factor(rk, eest);
        } else {
            s.accepted = 0;
            s.dt_next = dt / pi_reject_factor(eest);
        }
    }
    if (s.accepted) { rk->last_u = u_out; rk->last_t = s.t; }
    s.nf = (int)(rk->nf - nf0);
    *t_io = s.t;
    *dt_io = s.dt_next;
    if (out) *out = s;
    return MOL_OK;
}

// In-place variant: u is overwritten with the new state (one extra state copy per accepted step; prefer
// mol_rk_step_to or mol_rk_solve, which ping-pong).  The caller may have rewritten u between calls: unless this
// call continues the previous one at the time it reached, the FSAL stage is re-evaluated (see mol_rk_reinit).
extern "C" int mol_rk_step(mol_rk* rk, double* u, double* t_io, double* dt_io, int adaptive, mol_step_stats* out,
                           void* stream) {
    if (!rk || !u || !t_io || !dt_io) return fail(MOL_E_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (rk->alg != MOL_ALG_TSIT5) {
        const int64_t nf0 = rk->nf;
        if ((rc = mol_dist_register(rk->plan, u))) return rc;
        dist_mark_stale(rk->plan, u);
        if ((rc = step_fixed(rk, u, *t_io, *dt_io, st))) return rc;
        mol_step_stats s = {*t_io + *dt_io, *dt_io, 0.0, 1, (int)(rk->nf - nf0)};
        *t_io = s.t;
        if (out) *out = s;
        return MOL_OK;
    }
    if (u != rk->last_u || *t_io != rk->last_t) rk->fsal_valid = false;
    rk->last_u = u;                 // tsit5 reads u, writes alt: present it to mol_rk_step_to as a continuation
    rk->last_t = *t_io;
    dist_mark_stale(rk->plan, u);
    mol_step_stats s;
    if ((rc = mol_rk_step_to(rk, u, rk->alt, t_io, dt_io, adaptive, &s, stream))) return rc;
    if (s.accepted) {
        cudaMemcpyAsync(u, rk->alt, rk->n * 8, cudaMemcpyDeviceToDevice, st);
        dist_mark_stale(rk->plan, u);
        rk->last_u = u;
    }
    if (out) *out = s;
    return MOL_OK;
}

// layout of MolSolveArgs in kernels/mol_generic.cuh
struct SolveArgs {
    double* u;
    double* w[9];
    double* save;
    const double* saveat;
    double* out;
    double t0, t1, dt0, abstol, reltol;
    long long maxiters, n, nglobal;
    int nsave, alg, adaptive, pad;
};

static int solve_persistent(mol_rk* rk, double* u_dev, double t0, double t1, double dt0, int adaptive, const double* saveat,
                            int nsave, double* save_dev, int64_t maxiters, mol_solve_stats* out, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    if (nsave > rk->saveat_cap) {
        if (rk->d_saveat) cudaFree(rk->d_saveat);
        rk->d_saveat = nullptr;
        e = cudaMalloc(&rk->d_saveat, (size_t)nsave * 8);
        if (e != cudaSuccess) return cuda_fail(e, "save times");
        rk->saveat_cap = nsave;
    }
    if (nsave > 0) e = cudaMemcpyAsync(rk->d_saveat, saveat, (size_t)nsave * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "save times");
    SolveArgs A;
    A.u = u_dev;
    for (int i = 0; i < 7; ++i) A.w[i] = rk->k[i];
    A.w[7] = rk->alt;
    A.w[8] = rk->spare;
    A.save = save_dev;
    A.saveat = rk->d_saveat;
    A.out = rk->d_out;
    A.t0 = t0; A.t1 = t1; A.dt0 = dt0; A.abstol = rk->abstol; A.reltol = rk->reltol;
    A.maxiters = maxiters; A.n = rk->n; A.nglobal = rk->n_global;
    A.nsave = nsave; A.alg = rk->alg; A.adaptive = (rk->alg == MOL_ALG_TSIT5 && adaptive) ? 1 : 0; A.pad = 0;
    int rc = mol_plan_solve_small(rk->plan, &A, sizeof A, t0, st);
    if (rc != MOL_OK) return rc;
    e = cudaMemcpyAsync(rk->h_out, rk->d_out, 7 * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "mol_rk_solve (persistent kernel)");
    mol_solve_stats S;
    S.t_final = rk->h_out[0];
    S.dt_last = rk->h_out[1];
    S.nf = (int64_t)rk->h_out[2];
    S.naccept = (int64_t)rk->h_out[3];
    S.nreject = (int64_t)rk->h_out[4];
    S.retcode = (int)rk->h_out[5];
    rk->nf += S.nf;
    rk->fsal_valid = false;
    rk->last_u = nullptr;
    if (S.retcode == 0 && (int)rk->h_out[6] != nsave) return fail(MOL_E_ARG, "internal: a saveat point was not produced");
    if (out) *out = S;
    return MOL_OK;
}

// One Tsit5 attempt with the step size on the device: six sweeps (MOL_DEVDT variants, unscaled coefficients), the
// controller, the conditional advance.  u is both the start of the step and, after an accepted one, its end.
static int queue_attempt(mol_rk* rk, double* u, cudaStream_t qs) {
    int rc;
    const double* ctl = reinterpret_cast<const double*>(rk->d_ctl);
    MolRhsEpi noepi;
    for (int s = 1; s <= 4; ++s) {
        MolRhsIn in;
        in.nin = s + 1;
        in.a[0] = u;
        in.c[0] = 1.0;
        for (int j = 0; j < s; ++j) { in.a[j + 1] = rk->k[j]; in.c[j + 1] = T5_A[s][j]; }
        in.ctl = ctl;
        if ((rc = mol_rhs_launch(rk->plan, in, rk->k[s], T5_C[s], noepi, qs))) return rc;
    }
    {   // stage 6 (PRE): u+ into alt, partial error estimate into k[5]
        MolRhsIn in;
        in.nin = 6;
        in.a[0] = u;
        in.c[0] = 1.0;
        in.ctl = ctl;
        MolRhsEpi epi;
        epi.mode = MOL_EPI_PRE;
        epi.comb = rk->alt;
        epi.eout = rk->k[5];
        epi.cb[0] = 1.0;
        epi.ce[0] = 0.0;
        for (int j = 0; j < 5; ++j) {
            in.a[j + 1] = rk->k[j];
            in.c[j + 1] = T5_A[5][j];
            epi.cb[j + 1] = T5_A[6][j];
            epi.ce[j + 1] = T5_BT[j];
        }
        epi.cbk = T5_A[6][5];
        epi.cek = T5_BT[5];
        if ((rc = mol_rhs_launch(rk->plan, in, nullptr, T5_C[5], epi, qs))) return rc;
    }
    {   // stage 7 (FIN): k7 and the error norm
        MolRhsIn in;
        in.nin = 1;
        in.a[0] = rk->alt;
        in.c[0] = 1.0;
        in.ctl = ctl;
        MolRhsEpi epi;
        epi.mode = MOL_EPI_FIN;
        epi.e = rk->k[5];
        epi.u0 = u;
        epi.ek = T5_BT[6];
        epi.abstol = rk->abstol;
        epi.reltol = rk->reltol;
        epi.err = rk->d_parts;
        if ((rc = mol_rhs_launch(rk->plan, in, rk->k[6], 1.0, epi, qs))) return rc;
        rk->nparts = rk->plan->fin_slot;
    }
    mol_tsit5_control_kernel<<<1, 32, 0, qs>>>(rk->d_ctl, rk->d_parts, rk->nparts);
    const int grid = (int)std::min<int64_t>((rk->n + 255) / 256, (int64_t)rk->plan->sm_count * 8);
    mol_tsit5_advance_kernel<<<grid, 256, 0, qs>>>(rk->d_ctl, u, rk->alt, rk->k[0], rk->k[6], rk->n, 0);
    rk->plan->launches += 2;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MOL_OK : cuda_fail(e, "queued Tsit5 attempt");
}

// Adaptive Tsit5 with the controller on the device (see RkCtl).  Same step sequence as the host-driven loop below up to
// the last bit of pow() (device vs host libm); same dense-output saves, served by the host while the solve is on hold.
static int solve_queued(mol_rk* rk, double* u_dev, double t0, double t1, double dt0, const double* saveat, int nsave,
                        double* save_dev, int64_t maxiters, mol_solve_stats* out, cudaStream_t st) {
    cudaStream_t qs = rk->qs;
    cudaError_t e = cudaEventRecord(rk->qev, st);                  // the caller's stream order carries over to the private one
    if (e == cudaSuccess) e = cudaStreamWaitEvent(qs, rk->qev, 0);
    if (e != cudaSuccess) return cuda_fail(e, "mol_rk_solve (queued)");
    const double ttol = 1e-14 * std::max(1.0, std::max(std::fabs(t0), std::fabs(t1)));
    const int64_t nf0 = rk->nf;
    rk->fsal_valid = false;
    rk->qold = 1e-4;
    rk->last_u = nullptr;
    mol_solve_stats S = {t0, dt0, 0, 0, 0, 0};
    int rc;
    int isave = 0;
    auto save_copy = [&](const double* cur) {
        cudaMemcpyAsync(save_dev + (int64_t)isave * rk->n, cur, rk->n * 8, cudaMemcpyDeviceToDevice, qs);
        ++isave;
    };
    while (isave < nsave && saveat[isave] <= t0 + ttol) save_copy(u_dev);
    double dt = dt0;
    if (t0 < t1) {
        if (dt <= 0) { if ((rc = initial_dt(rk, u_dev, t0, &dt, qs))) return rc; }      // leaves k1 = f(u0, t0)
        else if ((rc = rhs_plain(rk, u_dev, rk->k[0], t0, qs))) return rc;
        RkCtl& H = *rk->h_ctl;
        H = RkCtl();
        H.t = t0;
        H.dtprop = dt;
        H.dt = std::min(dt, t1 - t0);
        H.qold = 1e-4;
        H.t1 = t1;
        H.ttol = ttol;
        H.save_next = isave < nsave ? saveat[isave] : INFINITY;
        H.maxiters = (double)maxiters;
        H.nglobal = (double)rk->n_global;
        e = cudaMemcpyAsync(rk->d_ctl, rk->h_ctl, sizeof(RkCtl), cudaMemcpyHostToDevice, qs);
        if (e != cudaSuccess) return cuda_fail(e, "mol_rk_solve (queued)");
        int batch = 8;
        if (const char* b = getenv("MOL_RK_QUEUED_BATCH")) batch = std::max(1, atoi(b));
        // what a captured attempt has baked into its kernel arguments (k1 / k7 trade places in the host-driven paths)
        std::vector<double> sig;
        auto addr = [](const void* p) { return (double)reinterpret_cast<uintptr_t>(p); };
        for (const void* p : {(const void*)u_dev, (const void*)rk->alt, (const void*)rk->k[0], (const void*)rk->k[5], (const void*)rk->k[6]})
            sig.push_back(addr(p));
        sig.push_back(rk->abstol);
        sig.push_back(rk->reltol);
        for (int i = 0; i < rk->plan->P.nparam; ++i) sig.push_back(rk->plan->params[i]);
        const bool stale = !rk->gexec || sig.size() != rk->graph_sig.size() ||
                           memcmp(sig.data(), rk->graph_sig.data(), sig.size() * sizeof(double)) != 0;
        bool first = true;
        for (;;) {
            int queued = 0;
            if (first) {
                // the first attempt runs eagerly (it also loads every kernel variant), then the same sequence is captured
                if ((rc = queue_attempt(rk, u_dev, qs))) return rc;
                ++queued;
                first = false;
                if (stale) {
                    if (rk->gexec) { cudaGraphExecDestroy(rk->gexec); rk->gexec = nullptr; }
                    if (rk->graph) { cudaGraphDestroy(rk->graph); rk->graph = nullptr; }
                    const int64_t l0 = rk->plan->launches;
                    e = cudaStreamBeginCapture(qs, cudaStreamCaptureModeThreadLocal);
                    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamBeginCapture");
                    rc = queue_attempt(rk, u_dev, qs);
                    e = cudaStreamEndCapture(qs, &rk->graph);
                    rk->plan->launches = l0;
                    if (rc != MOL_OK) return rc;
                    if (e == cudaSuccess) e = cudaGraphInstantiate(&rk->gexec, rk->graph, 0);
                    if (e != cudaSuccess) return cuda_fail(e, "capturing the Tsit5 attempt");
                    rk->graph_sig = sig;
                }
            }
            for (; queued < batch; ++queued) {
                if ((e = cudaGraphLaunch(rk->gexec, qs)) != cudaSuccess) return cuda_fail(e, "cudaGraphLaunch");
                rk->plan->launches += 8;
            }
            e = cudaMemcpyAsync(rk->h_ctl, rk->d_ctl, sizeof(RkCtl), cudaMemcpyDeviceToHost, qs);
            if (e == cudaSuccess) e = cudaStreamSynchronize(qs);
            if (e != cudaSuccess) return cuda_fail(e, "mol_rk_solve (queued)");
            if (getenv("MOL_RK_DEBUG"))
                fprintf(stderr, "[mol rk queued] it %.0f t %.17g dt_next %.17g dtprop %.17g eest %.17g accepted %.0f status %.0f\n", H.it, H.t, H.dt,
                        H.dtprop, H.eest, H.accepted, H.status);
            if (H.status == 4.0) {          // on hold: the step t_prev -> t (u_dev -> alt) covers save points
                const double tp = H.t_prev, tn = H.t, dtu = H.dt_used;
                while (isave < nsave && saveat[isave] <= tn + ttol) {
                    const double ts = saveat[isave];
                    if (std::fabs(ts - tn) <= ttol) save_copy(rk->alt);
                    else {
                        if ((rc = tsit5_dense(rk, u_dev, rk->alt, dtu, (ts - tp) / dtu, save_dev + (int64_t)isave * rk->n, qs))) return rc;
                        ++isave;
                    }
                }
                const int grid = (int)std::min<int64_t>((rk->n + 255) / 256, (int64_t)rk->plan->sm_count * 8);
                mol_tsit5_advance_kernel<<<grid, 256, 0, qs>>>(rk->d_ctl, u_dev, rk->alt, rk->k[0], rk->k[6], rk->n, 1);
                rk->plan->launches++;
                H.save_next = isave < nsave ? saveat[isave] : INFINITY;
                H.hold = 0.0;
                H.accepted = 0.0;
                if (tn < t1 && H.it < H.maxiters) { H.status = 0.0; H.skip = 0.0; }
                else H.status = tn < t1 ? 2.0 : 1.0;
                e = cudaMemcpyAsync(rk->d_ctl, rk->h_ctl, sizeof(RkCtl), cudaMemcpyHostToDevice, qs);
                if (e != cudaSuccess) return cuda_fail(e, "mol_rk_solve (queued)");
            }
            if (H.status != 0.0) break;
        }
        rk->nf += 6 * (int64_t)H.it;
        S.naccept = (int64_t)H.naccept;
        S.nreject = (int64_t)H.nreject;
        S.retcode = H.status == 1.0 ? 0 : (H.status == 2.0 ? 1 : 2);
        S.dt_last = H.dtprop;
        S.t_final = H.t;
    }
    e = cudaStreamSynchronize(qs);
    if (e != cudaSuccess) return cuda_fail(e, "mol_rk_solve (queued)");
    if (S.retcode == 0 && isave != nsave) return fail(MOL_E_ARG, "internal: a saveat point was not produced");
    S.nf = rk->nf - nf0;
    rk->fsal_valid = false;
    if (out) *out = S;
    return MOL_OK;
}

// Integrate t0 -> t1.  t1 is a stop time (the last step is shortened to land on it, as OrdinaryDiffEq does); save
// points are NOT: states at saveat[] come from dense output inside the step that covers them (Tsit5: its own
// 4th-order interpolant; Euler / SSPRK33 / RK4: cubic Hermite, one extra RHS evaluation per step that contains a
// save point), so `saveat` does not change the step sequence.  saveat must be non-decreasing inside [t0, t1].
extern "C" int mol_rk_solve(mol_rk* rk, double* u_dev, double t0, double t1, double dt0, int adaptive, const double* saveat,
                            int nsave, double* save_dev, int64_t maxiters, mol_solve_stats* out, void* stream) {
    if (!rk || !u_dev) return fail(MOL_E_ARG, "null argument");
    if (!(t1 >= t0)) return fail(MOL_E_ARG, "mol_rk_solve integrates forward: t1 >= t0");
    if (nsave > 0 && (!saveat || !save_dev)) return fail(MOL_E_ARG, "saveat / save_dev missing");
    const double ttol = 1e-14 * std::max(1.0, std::max(std::fabs(t0), std::fabs(t1)));
    for (int i = 0; i < nsave; ++i) {
        if (!(saveat[i] >= t0 - ttol && saveat[i] <= t1 + ttol)) return fail(MOL_E_ARG, "saveat point outside [t0, t1]");
        if (i > 0 && saveat[i] < saveat[i - 1]) return fail(MOL_E_ARG, "saveat must be non-decreasing");
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (maxiters <= 0) maxiters = 1000000;
    mol_solve_stats S = {t0, dt0, 0, 0, 0, 0};
    if (rk->spare && !rk->plan->dist.on) {           // small problem: the whole solve in one launch
        if ((rk->alg != MOL_ALG_TSIT5 || !adaptive) && dt0 <= 0) return fail(MOL_E_ARG, "fixed-step integration needs dt > 0");
        return solve_persistent(rk, u_dev, t0, t1, dt0, adaptive, saveat, nsave, save_dev, maxiters, out, st);
    }
    if (rk->qs && rk->alg == MOL_ALG_TSIT5 && adaptive && !rk->plan->dist.on)         // mid-size problem: attempts queued, no host in the loop
        return solve_queued(rk, u_dev, t0, t1, dt0, saveat, nsave, save_dev, maxiters, out, st);
    const int64_t nf0 = rk->nf;
    rk->fsal_valid = false;
    rk->qold = 1e-4;
    rk->last_u = nullptr;
    int rc;
    if ((rc = mol_dist_register(rk->plan, u_dev))) return rc;
    dist_mark_stale(rk->plan, u_dev);
    double t = t0;
    int isave = 0;
    auto save_copy = [&](const double* cur) {
        cudaMemcpyAsync(save_dev + (int64_t)isave * rk->n, cur, rk->n * 8, cudaMemcpyDeviceToDevice, st);
        ++isave;
    };
    while (isave < nsave && saveat[isave] <= t0 + ttol) save_copy(u_dev);
    const bool tsit5 = rk->alg == MOL_ALG_TSIT5;
    double* cur = u_dev;
    double* alt = rk->alt;
    // save points inside the Tsit5 step (cur, t) -> (alt, tnew) that was just accepted (before the buffers swap)
    auto tsit5_saves = [&](double tnew, double dtu) -> int {
        while (isave < nsave && saveat[isave] <= tnew + ttol) {
            const double ts = saveat[isave];
            if (std::fabs(ts - tnew) <= ttol) save_copy(alt);
            else {
                int r = tsit5_dense(rk, cur, alt, dtu, (ts - t) / dtu, save_dev + (int64_t)isave * rk->n, st);
                if (r != MOL_OK) return r;
                ++isave;
            }
        }
        return MOL_OK;
    };
    if (!tsit5 || !adaptive) {
        if (dt0 <= 0) return fail(MOL_E_ARG, "fixed-step integration needs dt > 0");
        const double span = t1 - t0;
        int64_t nsteps = (int64_t)std::ceil(span / dt0 - 1e-9);
        if (nsteps < 0) nsteps = 0;
        if (nsteps > maxiters) { nsteps = maxiters; S.retcode = 1; }
        for (int64_t i = 0; i < nsteps; ++i) {
            const bool last = (i == nsteps - 1) && S.retcode == 0;
            const double tnew = last ? t1 : t0 + (double)(i + 1) * dt0;
            const double dtu = tnew - t;
            if (tsit5) {                      // ping-pong between the caller's state and the spare one
                double eest;
                if ((rc = tsit5_attempt(rk, cur, alt, t, dtu, &eest, st))) return rc;
                if ((rc = tsit5_saves(tnew, dtu))) return rc;
                std::swap(cur, alt);
                std::swap(rk->k[0], rk->k[6]);
            } else {
                const bool inside = isave < nsave && saveat[isave] < tnew - ttol;       // a save point strictly inside
                if (inside) cudaMemcpyAsync(alt, cur, rk->n * 8, cudaMemcpyDeviceToDevice, st);      // keep u0
                if ((rc = step_fixed(rk, cur, t, dtu, st))) return rc;                  // k[0] = f(u0, t) afterwards
                if (inside) {
                    if ((rc = rhs_plain(rk, cur, rk->k[1], tnew, st))) return rc;        // f(u1, tnew)
                    while (isave < nsave && saveat[isave] < tnew - ttol) {
                        if ((rc = hermite_dense(rk, alt, cur, rk->k[0], rk->k[1], dtu, (saveat[isave] - t) / dtu,
                                                save_dev + (int64_t)isave * rk->n, st)))
                            return rc;
                        ++isave;
                    }
                }
                while (isave < nsave && saveat[isave] <= tnew + ttol) save_copy(cur);
            }
            t = tnew;
            S.naccept++;
        }
        S.dt_last = dt0;
    } else {
        // Host-driven adaptive loop (large problems, slabs): per attempt the six sweeps with host-scaled coefficients, then
        // the SAME one-thread controller kernel the queued solve uses (identical arithmetic, hence identical step sequences
        // in both modes) and one read-back of its block instead of the bare error sum.
        double dt = dt0;
        if (dt <= 0 && t < t1 && (rc = initial_dt(rk, cur, t, &dt, st))) return rc;
        if (t < t1) {
            RkCtl& H = *rk->h_ctl;
            H = RkCtl();
            H.t = t;
            H.dtprop = dt;
            H.dt = std::min(dt, t1 - t);
            H.qold = 1e-4;
            H.t1 = t1;
            H.ttol = ttol;
            H.save_next = INFINITY;              // (save points are served right here, after every accepted step)
            H.maxiters = (double)maxiters;
            H.nglobal = (double)rk->n_global;
            cudaError_t ce = cudaMemcpyAsync(rk->d_ctl, rk->h_ctl, sizeof(RkCtl), cudaMemcpyHostToDevice, st);
            if (ce != cudaSuccess) return cuda_fail(ce, "mol_rk_solve");
            const bool dbg = getenv("MOL_RK_DEBUG") != nullptr;
            for (;;) {
                const double dtu = H.dt;
                if ((rc = tsit5_attempt(rk, cur, alt, t, dtu, nullptr, st))) return rc;
                if (rk->plan->dist.on) mol_tsit5_control_kernel<<<1, 32, 0, st>>>(rk->d_ctl, rk->d_err, 1);
                else mol_tsit5_control_kernel<<<1, 32, 0, st>>>(rk->d_ctl, rk->d_parts, rk->nparts);
                rk->plan->launches++;
                ce = cudaMemcpyAsync(rk->h_ctl, rk->d_ctl, sizeof(RkCtl), cudaMemcpyDeviceToHost, st);
                if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
                if (ce != cudaSuccess) return cuda_fail(ce, "tsit5 step");
                if (dbg) fprintf(stderr, "[mol rk host] it %.0f t %.17g dt_next %.17g dtprop %.17g eest %.17g accepted %.0f status %.0f\n", H.it,
                                 H.t, H.dt, H.dtprop, H.eest, H.accepted, H.status);
                if (H.accepted != 0.0) {
                    if ((rc = tsit5_saves(H.t, dtu))) return rc;
                    t = H.t;
                    std::swap(cur, alt);
                    std::swap(rk->k[0], rk->k[6]);
                }
                if (H.status != 0.0) break;
            }
            S.naccept = (int64_t)H.naccept;
            S.nreject = (int64_t)H.nreject;
            S.retcode = H.status == 1.0 ? 0 : (H.status == 2.0 ? 1 : 2);
            dt = H.dtprop;
        }
        S.dt_last = dt;
    }
    if (cur != u_dev) cudaMemcpyAsync(u_dev, cur, rk->n * 8, cudaMemcpyDeviceToDevice, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "mol_rk_solve");
    if (S.retcode == 0 && isave != nsave) return fail(MOL_E_ARG, "internal: a saveat point was not produced");
    S.t_final = t;
    S.nf = rk->nf - nf0;
    if (out) *out = S;
    return MOL_OK;
}
