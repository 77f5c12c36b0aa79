// mol_generic.cuh — table-driven RHS kernel: one thread per interior node of a box region.
//
// Universal path: any grid (uniform / non-uniform), any approximation order, one-sided frame
// rows, every boundary rule.  Stencil rows come from per-node tables (the reference picks a row per
// point at discretize time, centered_difference.jl:16-27 / upwind_difference.jl:8-26; here the
// same choice is data).  Used for the frame around the tiled core and for non-uniform grids.
// Loads go straight to global memory (L1/L2 serve the tap reuse); x is the fastest index so a warp
// reads 32 consecutive doubles per tap.
#pragma once

template <int V>
struct MolGenericVars {
    static __device__ __forceinline__ void run(const MolIn& in, const MolCtx& c, int i0, int i1, int i2,
                                               double* __restrict__ out, const MolEpi* epi, double& errsum) {
        bool inside = (i0 >= MOL_ILO(V, 0)) && (i0 <= MOL_IHI(V, 0));
        if (MOL_NDIM >= 2) inside = inside && (i1 >= MOL_ILO(V, 1)) && (i1 <= MOL_IHI(V, 1));
        if (MOL_NDIM >= 3) inside = inside && (i2 >= MOL_ILO(V, 2)) && (i2 <= MOL_IHI(V, 2));
        if (inside) {
            const double du = mol_eq_generic<V>(in, c, i0, i1, i2);
            const mol_i64 f = mol_flat<V>(c, i0, i1, i2);
#if MOL_EPI_PRE
            double v, p, q;
            mol_load3(in, *epi, f, v, p, q);
            epi->comb[f] = fma(epi->cbk, du, p);
            epi->eout[f] = fma(epi->cek, du, q);
#else
            out[f] = du;
#endif
#if MOL_EPI_FIN
            mol_fin_point(*epi, __ldg(epi->e + f), __ldg(epi->u0 + f), du, mol_load(in, f), errsum);
#endif
        }
        MolGenericVars<V + 1>::run(in, c, i0, i1, i2, out, epi, errsum);
    }
};
template <>
struct MolGenericVars<MOL_NVAR> {
    static __device__ __forceinline__ void run(const MolIn&, const MolCtx&, int, int, int, double*, const MolEpi*, double&) {}
};

// up to MOL_MAX_BOXES box regions per launch (the frame around the tiled core is 2*ndim boxes, the
// slab edges 2): `start[k]` = number of nodes in boxes 0..k-1
#define MOL_MAX_BOXES 8
struct MolBoxes { int n; int pad; MolBox b[MOL_MAX_BOXES]; mol_i64 start[MOL_MAX_BOXES + 1]; };

extern "C" __global__ void __launch_bounds__(256)
mol_rhs_generic(MolIn in, MolCtx c, MolBoxes B, double* __restrict__ out
#if MOL_EPI
                , MolEpi epi
#endif
) {
#if MOL_DEVDT
#if MOL_EPI
    if (!mol_devdt_apply(in, c, &epi)) return;
#else
    if (!mol_devdt_apply(in, c, nullptr)) return;
#endif
#endif
    const mol_i64 total = B.start[B.n];
#if MOL_EPI
    double errsum = 0.0;
#endif
    for (mol_i64 g0 = (mol_i64)blockIdx.x * blockDim.x + threadIdx.x; g0 < total;
         g0 += (mol_i64)gridDim.x * blockDim.x) {
        int k = 0;
        while (k + 1 < B.n && g0 >= B.start[k + 1]) ++k;
        const MolBox& box = B.b[k];
        const mol_i64 g = g0 - B.start[k];
        const int e0 = box.hi[0] - box.lo[0] + 1;
        const int e1 = (MOL_NDIM >= 2) ? box.hi[1] - box.lo[1] + 1 : 1;
        const int i0 = box.lo[0] + (int)(g % e0);
        const int i1 = (MOL_NDIM >= 2) ? box.lo[1] + (int)((g / e0) % e1) : 1;
        const int i2 = (MOL_NDIM >= 3) ? box.lo[2] + (int)(g / ((mol_i64)e0 * e1)) : 1;
#if MOL_EPI
        MolGenericVars<0>::run(in, c, i0, i1, i2, out, &epi, errsum);
#else
        double dummy = 0.0;
        MolGenericVars<0>::run(in, c, i0, i1, i2, out, nullptr, dummy);
#endif
    }
#if MOL_EPI_FIN
    __shared__ double red[8];
    errsum = mol_warp_sum(errsum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = errsum;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        v = mol_warp_sum(v);
        if (threadIdx.x == 0 && epi.err) epi.err[blockIdx.x] = v;      // this CTA's slot: fixed node -> thread map, no atomics
    }
#endif
}

// ---- persistent explicit Runge-Kutta solver for small problems (MOL_KERNEL_SOLVE, compiled with MOL_NIN = 7) ----------
// A whole solve(prob, Tsit5() / SSPRK33() / RK4() / Euler(); saveat, adaptive) in ONE launch of ONE CTA: stages, the
// embedded error norm, the PI step controller and dense-output saves all run on the device, with __syncthreads between
// sweeps.  The host-driven loop (csrc/mol_rk.cu) costs seven launches and one read-back per step -- 60-100 us per step
// whatever the problem size -- which is the whole cost of the reference's config 1 (1-D heat, 99 unknowns); here a step
// of such a problem costs a few microseconds.  Same methods, same controller constants, same save semantics as
// mol_rk.cu (OrdinaryDiffEq defaults; saveat by dense output).  Used by mol_rk_solve for single-device plans whose
// state fits one CTA's reach (mol_rk.cu: kPersistentMaxUnknowns).
#if MOL_KERNEL_SOLVE
struct MolSolveArgs {
    double* u;              // state, in/out
    double* w[9];           // k1..k7, the second state buffer, one spare
    double* save;           // nsave states
    const double* saveat;   // device copy of the save times
    double* out;            // [t_final, dt_last, nf, naccept, nreject, retcode]
    double t0, t1, dt0, abstol, reltol;
    long long maxiters, n, nglobal;
    int nsave, alg, adaptive, pad;
};

__constant__ double mol_t5c[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
__constant__ double mol_t5a[7][6] = {
    {0},
    {0.161},
    {-0.008480655492356989, 0.335480655492357},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
__constant__ double mol_t5bt[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                                   -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 0.015151515151515152};

__device__ __forceinline__ double mol_block_sum(double v, double* red) {
    v = mol_warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += red[k];      // same order in every thread: identical bits
    return s;
}

// one RHS sweep over the boxes: out = f(sum_j c_j a_j, t); the CTA's threads stride over the nodes
// (not inlined: the solver calls it from a dozen places, and a dozen copies of the equations neither fit the instruction
// cache nor the register file -- measured: 126 us per sweep of a 128-node WENO problem when inlined)
__device__ __noinline__ void mol_solve_rhs(const MolIn& in, const MolCtx& c0, double t, const MolBox& box, double* out) {
    MolCtx c = c0;
    c.t = t;
    const int e0 = box.hi[0] - box.lo[0] + 1;
    const int e1 = (MOL_NDIM >= 2) ? box.hi[1] - box.lo[1] + 1 : 1;
    const int e2 = (MOL_NDIM >= 3) ? box.hi[2] - box.lo[2] + 1 : 1;
    const mol_i64 total = (mol_i64)e0 * e1 * e2;
    for (mol_i64 g = threadIdx.x; g < total; g += blockDim.x) {
        const int i0 = box.lo[0] + (int)(g % e0);
        const int i1 = (MOL_NDIM >= 2) ? box.lo[1] + (int)((g / e0) % e1) : 1;
        const int i2 = (MOL_NDIM >= 3) ? box.lo[2] + (int)(g / ((mol_i64)e0 * e1)) : 1;
        double dummy = 0.0;
        MolGenericVars<0>::run(in, c, i0, i1, i2, out, nullptr, dummy);
    }
    __syncthreads();
}

// in = {a0 + sum_j cj aj}: arrays beyond `n` alias a0 with coefficient 0 (MOL_NIN is a compile-time 7).  ONE MolIn lives
// in the kernel's frame and is refilled per sweep (a temporary per call site costs 112 B of stack each).
__device__ __forceinline__ void mol_solve_set(MolIn& in, const double* a0, int n, double* const* k, const double* cf) {
    in.a[0] = a0;
    in.c[0] = 1.0;
    for (int j = 1; j < MOL_NIN; ++j) {
        in.a[j] = (j <= n) ? k[j - 1] : a0;
        in.c[j] = (j <= n) ? cf[j - 1] : 0.0;
    }
}
#define MOL_SOLVE_RHS(a0, n, t_, out_) do { mol_solve_set(in, a0, n, k, cf); mol_solve_rhs(in, c, t_, B, out_); } while (0)

extern "C" __global__ void __launch_bounds__(256) mol_solve_small(MolCtx c, MolBoxes Bs, MolSolveArgs A) {
    __shared__ double red[32];
    const MolBox B = Bs.b[0];            // one bounding box of every variable's interior
    const mol_i64 n = A.n;
    double* u = A.u;
    double* un = A.w[7];
    double* k[7] = {A.w[0], A.w[1], A.w[2], A.w[3], A.w[4], A.w[5], A.w[6]};
    MolIn in;
    double cf[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double t = A.t0, dt = A.dt0, qold = 1e-4;
    const double ttol = 1e-14 * fmax(1.0, fmax(fabs(A.t0), fabs(A.t1)));
    long long nf = 0, nacc = 0, nrej = 0, it = 0;
    int isave = 0, retcode = 0;
    auto wrms = [&](const double* a, const double* b, double ca, double cb) {      // Hairer norms of the initial step
        double s = 0.0;
        for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x) {
            double v = ca * a[i];
            if (b) v = fma(cb, b[i], v);
            const double r = v / (A.abstol + fabs(u[i]) * A.reltol);
            s = fma(r, r, s);
        }
        return sqrt(mol_block_sum(s, red) / (double)A.nglobal);
    };
    auto copy_save = [&](const double* src) {
        for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x) A.save[(mol_i64)isave * n + i] = src[i];
        ++isave;
    };
    while (isave < A.nsave && A.saveat[isave] <= A.t0 + ttol) copy_save(u);
    const bool tsit5 = A.alg == 4;
    bool have_k1 = false;
    if (tsit5 && A.adaptive && dt <= 0.0 && t < A.t1) {      // Hairer-Norsett-Wanner starting step (OrdinaryDiffEq initdt)
        MOL_SOLVE_RHS(u, 0, t, k[0]);
        nf++;
        have_k1 = true;
        const double d0 = wrms(u, nullptr, 1.0, 0.0), d1 = wrms(k[0], nullptr, 1.0, 0.0);
        const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        cf[0] = h0;
        MOL_SOLVE_RHS(u, 1, t + h0, k[1]);
        nf++;
        const double d2 = wrms(k[1], k[0], 1.0, -1.0) / h0;
        const double m = fmax(d1, d2);
        const double h1 = (m <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(10.0, -(2.0 + log10(m)) / 5.0);
        dt = fmin(100 * h0, h1);
    }
    if (tsit5) {
        const long long nfixed = A.adaptive ? 0 : (long long)ceil((A.t1 - A.t0) / A.dt0 - 1e-9);
        long long step = 0;
        while (t < A.t1 && it < A.maxiters) {
            ++it;
            double tnew, dtu;
            if (A.adaptive) {
                dtu = fmin(dt, A.t1 - t);
                tnew = (fabs((t + dtu) - A.t1) <= ttol) ? A.t1 : t + dtu;
            } else {
                tnew = (step == nfixed - 1) ? A.t1 : A.t0 + (double)(step + 1) * A.dt0;
                dtu = tnew - t;
            }
            if (!have_k1) {
                MOL_SOLVE_RHS(u, 0, t, k[0]);
                nf++;
                have_k1 = true;
            }
            for (int s = 1; s <= 5; ++s) {
                for (int j = 0; j < s; ++j) cf[j] = dtu * mol_t5a[s][j];
                MOL_SOLVE_RHS(u, s, t + mol_t5c[s] * dtu, k[s]);
            }
            double err = 0.0;
            for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x) {
                double acc = 0.0, e = 0.0;
                for (int j = 0; j < 6; ++j) { acc = fma(dtu * mol_t5a[6][j], k[j][i], acc); e = fma(dtu * mol_t5bt[j], k[j][i], e); }
                un[i] = u[i] + acc;
                k[6][i] = e;                                  // parked: the partial error estimate (k7 overwrites it below)
            }
            __syncthreads();
            // k7 goes to the spare array first: the error needs the parked partial sums of k[6]
            MOL_SOLVE_RHS(un, 0, tnew, A.w[8]);
            nf += 6;
            for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x) {
                const double k7 = A.w[8][i];
                const double e = fma(dtu * mol_t5bt[6], k7, k[6][i]);
                const double r = e / (A.abstol + fmax(fabs(u[i]), fabs(un[i])) * A.reltol);
                err = fma(r, r, err);
                k[6][i] = k7;
            }
            const double eest = sqrt(mol_block_sum(err, red) / (double)A.nglobal);
            __syncthreads();
            if (A.adaptive && !(eest == eest)) { retcode = 2; break; }
            if (!A.adaptive || eest <= 1.0) {
                while (isave < A.nsave && A.saveat[isave] <= tnew + ttol) {        // dense output inside the step
                    const double ts = A.saveat[isave];
                    if (fabs(ts - tnew) <= ttol) copy_save(un);
                    else {
                        const double th = (ts - t) / dtu, th2 = th * th;
                        double b[7];
                        b[0] = -1.0530884977290216 * th * (th - 1.3299890189751412) * (th2 - 1.4364028541716351 * th + 0.7139816917074209);
                        b[1] = 0.1017 * th2 * (th2 - 2.1966568338249754 * th + 1.2949852507374631);
                        b[2] = 2.490627285651252793 * th2 * (th2 - 2.38535645472061657 * th + 1.57803468208092486);
                        b[3] = -16.54810288924490272 * (th - 1.21712927295533244) * (th - 0.61620406037800089) * th2;
                        b[4] = 47.37952196281928122 * (th - 1.203071208372362603) * (th - 0.658047292653547382) * th2;
                        b[5] = -34.87065786149660974 * (th - 1.2) * (th - 0.666666666666666667) * th2;
                        b[6] = 2.5 * (th - 1.0) * (th - 0.6) * th2;
                        for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x) {
                            double acc = 0.0;
                            for (int j = 0; j < 7; ++j) acc = fma(dtu * b[j], k[j][i], acc);
                            A.save[(mol_i64)isave * n + i] = u[i] + acc;
                        }
                        ++isave;
                    }
                }
                __syncthreads();
                if (A.adaptive) {
                    const double q = eest > 0 ? fmax(0.1, fmin(5.0, pow(eest, 7.0 / 50) / pow(qold, 2.0 / 25) / 0.9)) : 0.1;
                    qold = fmax(eest, 1e-4);
                    const bool clipped = dtu < dt;
                    if (!clipped || tnew < A.t1) dt = dtu / q;
                }
                t = tnew;
                double* tmp = u; u = un; un = tmp;            // ping-pong the state
                tmp = k[0]; k[0] = k[6]; k[6] = tmp;          // FSAL
                nacc++;
                step++;
            } else {
                nrej++;
                dt = dtu / fmin(5.0, pow(eest, 7.0 / 50) / 0.9);
                if (dt < 1e-14 * fmax(1.0, fabs(t))) { retcode = 2; break; }
            }
        }
        if (retcode == 0 && t < A.t1) retcode = 1;
    } else {
        // Euler / SSPRK33 / RK4 with a fixed step (the last one shortened to land on t1); save points strictly inside a
        // step by cubic Hermite interpolation between its end points (u0 is kept in `un`, f(u1) goes to k[4])
        const long long nsteps = (long long)ceil((A.t1 - A.t0) / A.dt0 - 1e-9);
        for (long long step = 0; step < nsteps && it < A.maxiters; ++step, ++it) {
            const double tnew = (step == nsteps - 1) ? A.t1 : A.t0 + (double)(step + 1) * A.dt0;
            const double h = tnew - t;
            const bool inside = isave < A.nsave && A.saveat[isave] < tnew - ttol;
            if (inside) {
                for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x) un[i] = u[i];
                __syncthreads();
            }
            MOL_SOLVE_RHS(u, 0, t, k[0]);
            if (A.alg == 1) {
                for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x) u[i] = fma(h, k[0][i], u[i]);
                nf += 1;
            } else if (A.alg == 2) {
                cf[0] = h;
                MOL_SOLVE_RHS(u, 1, t + h, k[1]);
                cf[0] = h / 4; cf[1] = h / 4;
                MOL_SOLVE_RHS(u, 2, t + h / 2, k[2]);
                for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x)
                    u[i] = u[i] + (h / 6) * k[0][i] + (h / 6) * k[1][i] + (2 * h / 3) * k[2][i];
                nf += 3;
            } else {
                cf[0] = h / 2;
                MOL_SOLVE_RHS(u, 1, t + h / 2, k[1]);
                cf[0] = 0.0; cf[1] = h / 2;
                MOL_SOLVE_RHS(u, 2, t + h / 2, k[2]);
                cf[1] = 0.0; cf[2] = h;
                MOL_SOLVE_RHS(u, 3, t + h, k[3]);
                for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x)
                    u[i] = u[i] + (h / 6) * k[0][i] + (h / 3) * k[1][i] + (h / 3) * k[2][i] + (h / 6) * k[3][i];
                nf += 4;
            }
            __syncthreads();
            if (inside) {
                MOL_SOLVE_RHS(u, 0, tnew, k[4]);
                nf++;
                while (isave < A.nsave && A.saveat[isave] < tnew - ttol) {
                    const double th = (A.saveat[isave] - t) / h, w = th * (th - 1.0);
                    const double c0 = (1.0 - th) - w * (1.0 - 2.0 * th), c1 = th + w * (1.0 - 2.0 * th), c2 = w * (th - 1.0) * h,
                                 c3 = w * th * h;
                    for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x)
                        A.save[(mol_i64)isave * n + i] = c0 * un[i] + c1 * u[i] + c2 * k[0][i] + c3 * k[4][i];
                    ++isave;
                }
            }
            while (isave < A.nsave && A.saveat[isave] <= tnew + ttol) copy_save(u);
            __syncthreads();
            t = tnew;
            nacc++;
        }
        if (t < A.t1) retcode = 1;
        dt = A.dt0;
    }
    if (u != A.u) {
        for (mol_i64 i = threadIdx.x; i < n; i += blockDim.x) A.u[i] = u[i];
    }
    if (threadIdx.x == 0) {
        A.out[0] = t; A.out[1] = dt; A.out[2] = (double)nf; A.out[3] = (double)nacc; A.out[4] = (double)nrej;
        A.out[5] = (double)retcode; A.out[6] = (double)isave;
    }
}
#endif  // MOL_KERNEL_SOLVE

// ---- solution unpacking: flat unknown vector -> full grid (interface/solution/timedep.jl:30-72) -------------------
// The reference rebuilds u over the WHOLE grid of every dependent variable when a solution is indexed
// (sol[u(t,x)]): unknowns come from the state vector, boundary-face nodes from the eliminated boundary
// equations (`observed`), and nodes outside the interior in two or more dimensions are the "invalid corner
// points set to 0" of generate_corner_eqs! (generate_bc_eqs.jl:396-416).  Here: one thread per grid node, the
// face nodes through the same ghost rules / periodic wrap the stencils use (mol_node).
// out: variable-major, x fastest, MOL_N0 x MOL_N1 x MOL_N2 nodes per variable.
#if MOL_KERNEL_UNPACK
template <int V>
struct MolUnpackVars {
    static __device__ __forceinline__ void run(const MolIn& in, const MolCtx& c, int i0, int i1, int i2, mol_i64 g,
                                               double* __restrict__ out) {
        int nout = (i0 < MOL_ILO(V, 0) || i0 > MOL_IHI(V, 0)) ? 1 : 0;
        if (MOL_NDIM >= 2) nout += (i1 < MOL_ILO(V, 1) || i1 > MOL_IHI(V, 1)) ? 1 : 0;
        if (MOL_NDIM >= 3) nout += (i2 < MOL_ILO(V, 2) || i2 > MOL_IHI(V, 2)) ? 1 : 0;
        out[(mol_i64)V * MOL_N0 * MOL_N1 * MOL_N2 + g] = (nout >= 2) ? 0.0 : mol_node<V>(in, c, i0, i1, i2);
        MolUnpackVars<V + 1>::run(in, c, i0, i1, i2, g, out);
    }
};
template <>
struct MolUnpackVars<MOL_NVAR> {
    static __device__ __forceinline__ void run(const MolIn&, const MolCtx&, int, int, int, mol_i64, double*) {}
};

extern "C" __global__ void __launch_bounds__(256) mol_unpack_full(MolIn in, MolCtx c, double* __restrict__ out) {
    const mol_i64 total = (mol_i64)MOL_N0 * MOL_N1 * MOL_N2;
    for (mol_i64 g = (mol_i64)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (mol_i64)gridDim.x * blockDim.x) {
        const int i0 = 1 + (int)(g % MOL_N0);
        const int i1 = 1 + (int)((g / MOL_N0) % MOL_N1);
        const int i2 = 1 + (int)(g / ((mol_i64)MOL_N0 * MOL_N1));
        MolUnpackVars<0>::run(in, c, i0, i1, i2, g, out);
    }
}
#endif

// ---- Jacobian-vector product, table-driven (kernels/mol_jvp.cuh) -----------------------------------------
// out[f] = d/d(eps) f_f(u + eps v) at eps = 0 for every unknown f of the listed boxes: the generated equations on dual
// numbers, value part from in.a[0] (= u), tangent part from jv.v (= v).
#if MOL_KERNEL_JVP
template <int V>
struct MolJvpVars {
    static __device__ __forceinline__ void run(const MolIn& in, const MolJv& jv, const MolCtx& c, int i0, int i1, int i2,
                                               double* __restrict__ out) {
        bool inside = (i0 >= MOL_ILO(V, 0)) && (i0 <= MOL_IHI(V, 0));
        if (MOL_NDIM >= 2) inside = inside && (i1 >= MOL_ILO(V, 1)) && (i1 <= MOL_IHI(V, 1));
        if (MOL_NDIM >= 3) inside = inside && (i2 >= MOL_ILO(V, 2)) && (i2 <= MOL_IHI(V, 2));
        if (inside) out[mol_flat<V>(c, i0, i1, i2)] = mol_eq_jvp<V>(in, jv, c, i0, i1, i2).d;
        MolJvpVars<V + 1>::run(in, jv, c, i0, i1, i2, out);
    }
};
template <>
struct MolJvpVars<MOL_NVAR> {
    static __device__ __forceinline__ void run(const MolIn&, const MolJv&, const MolCtx&, int, int, int, double*) {}
};

extern "C" __global__ void __launch_bounds__(256)
mol_jvp_generic(MolIn in, MolJv jv, MolCtx c, MolBoxes B, double* __restrict__ out) {
    const mol_i64 total = B.start[B.n];
    for (mol_i64 g0 = (mol_i64)blockIdx.x * blockDim.x + threadIdx.x; g0 < total;
         g0 += (mol_i64)gridDim.x * blockDim.x) {
        int k = 0;
        while (k + 1 < B.n && g0 >= B.start[k + 1]) ++k;
        const MolBox& box = B.b[k];
        const mol_i64 g = g0 - B.start[k];
        const int e0 = box.hi[0] - box.lo[0] + 1;
        const int e1 = (MOL_NDIM >= 2) ? box.hi[1] - box.lo[1] + 1 : 1;
        const int i0 = box.lo[0] + (int)(g % e0);
        const int i1 = (MOL_NDIM >= 2) ? box.lo[1] + (int)((g / e0) % e1) : 1;
        const int i2 = (MOL_NDIM >= 3) ? box.lo[2] + (int)(g / ((mol_i64)e0 * e1)) : 1;
        MolJvpVars<0>::run(in, jv, c, i0, i1, i2, out);
    }
}
#endif
