// mol_device.cuh — hand-written device runtime shared by every generated stencil program.
//
// Compiled at plan-creation time by NVRTC for sm_100a together with a generated prelude
// (MOL_* constants), generated ghost rules / pointwise equations, and the kernel files
// mol_generic.cuh / mol_tiled.cuh.  No CUDA headers are needed: everything below is plain
// CUDA C++ plus inline PTX.
//
// Naming follows the reference's domain: nodes (1-based grid indices, like CartesianIndex in
// src/discretization/discretize_vars.jl), interior box (interior_map.jl:105-115), taps
// (centered_difference.jl:16-27), ghost rules (generate_bc_eqs.jl), periodic wrap
// (interface_boundary.jl:33-42).
#pragma once

typedef unsigned long long mol_u64;
typedef long long mol_i64;

#ifndef MOL_NIN
#define MOL_NIN 1          // number of input arrays combined on load (RK stage fusion)
#endif
#ifndef MOL_WENO_RATIO
#define MOL_WENO_RATIO 1   // 1: division-free nonlinear weights in mol_weno5_uniform (0: one reciprocal per weight)
#endif
#ifndef MOL_DEVDT
#define MOL_DEVDT 0        // 1: step size, time and a skip flag come from a device-resident control block (MolIn::ctl)
#endif

// ---- state input: value(idx) = sum_j c[j] * a[j][idx]  (u + dt*sum a_sj k_j, fused on load) ----
struct MolIn {
    const double* a[MOL_NIN];
    double c[MOL_NIN];
#if MOL_DIST
    // slab decomposition: ghost planes of every input array, received from the two
    // neighbouring ranks: MOL_HALO planes below the slab / above the slab, var-major.
    const double* hlo[MOL_NIN];
    const double* hhi[MOL_NIN];
#endif
#if MOL_DEVDT
    // Device-side step control (csrc/mol_rk.cu, queued adaptive solve): {t, dt, skip} written by the controller kernel
    // of the previous attempt.  The host then passes UNSCALED Runge-Kutta coefficients: c[j >= 1] = a_sj, MolCtx::t = c_s,
    // epilogue coefficients without their factor dt; mol_devdt_apply() below completes them.
    const double* ctl;
#endif
};

// The persistent solver kernel (MOL_KERNEL_SOLVE, kernels/mol_generic.cuh) rewrites its stage arrays while it runs, so
// its state loads must not take the read-only (ld.global.nc) path.
#ifndef MOL_KERNEL_SOLVE
#define MOL_KERNEL_SOLVE 0
#endif
#if MOL_KERNEL_SOLVE
#define MOL_STATE_LD(p) (*(p))
#else
#define MOL_STATE_LD(p) __ldg(p)
#endif
__device__ __forceinline__ double mol_load(const MolIn& in, mol_i64 idx) {
#if MOL_NIN == 1
    return MOL_STATE_LD(in.a[0] + idx);
#else
    double s = in.c[0] * MOL_STATE_LD(in.a[0] + idx);
#pragma unroll
    for (int j = 1; j < MOL_NIN; ++j) s = fma(in.c[j], MOL_STATE_LD(in.a[j] + idx), s);
    return s;
#endif
}

#if MOL_DIST
// same combination on the ghost planes (lower != 0: planes below the slab)
__device__ __forceinline__ double mol_load_halo(const MolIn& in, bool lower, mol_i64 idx) {
#if MOL_NIN == 1
    return __ldg((lower ? in.hlo[0] : in.hhi[0]) + idx);
#else
    double s = in.c[0] * __ldg((lower ? in.hlo[0] : in.hhi[0]) + idx);
#pragma unroll
    for (int j = 1; j < MOL_NIN; ++j) s = fma(in.c[j], __ldg((lower ? in.hlo[j] : in.hhi[j]) + idx), s);
    return s;
#endif
}
#endif

// ---- per-launch context -------------------------------------------------------------------------
struct MolCtx {
    double t;
    double p[MOL_NPARAM > 0 ? MOL_NPARAM : 1];
    const double* grid[3];      // node coordinates per dimension, grid[j][node-1]
    const double* tabw;         // stencil-row weights, all operators concatenated
    const int*    tabs;         // per row: {first tap node, number of taps}
    // slab decomposition along the last dimension: this rank owns nodes
    // [loc_lo, loc_hi] of that dimension (planes outside come from MolIn::hlo/hhi) and stores
    // `vstride` doubles per variable.
    int loc_lo, loc_hi;
    mol_i64 vstride;
};

// generated: ghost rule for variable V along dimension D at node (i0,i1,i2) outside the interior
template <int V, int D>
__device__ double mol_ghost(const MolIn& in, const MolCtx& c, int i0, int i1, int i2);

// Value of discretised variable V at node (i0,i1,i2) (1-based).  Interior nodes are loaded from
// the state vector(s); nodes outside the interior box are resolved one dimension at a time:
// periodic dimensions wrap by +-(n-1) (u[1] == u[n], interface_boundary.jl:33-42), the others go
// through the generated ghost rule (Dirichlet value, affine Neumann/Robin solve, extrapolation).
template <int V>
__device__ __forceinline__ double mol_node(const MolIn& in, const MolCtx& c, int i0, int i1, int i2) {
#if !(MOL_DIST && MOL_NDIM == 1)
    if (i0 < MOL_ILO(V, 0) || i0 > MOL_IHI(V, 0)) {
        if (MOL_PER(V, 0)) i0 += (i0 <= 1) ? (MOL_N0 - 1) : -(MOL_N0 - 1);
        else return mol_ghost<V, 0>(in, c, i0, i1, i2);
    }
#endif
#if MOL_NDIM >= 2 && !(MOL_DIST && MOL_NDIM == 2)
    if (i1 < MOL_ILO(V, 1) || i1 > MOL_IHI(V, 1)) {
        if (MOL_PER(V, 1)) i1 += (i1 <= 1) ? (MOL_N1 - 1) : -(MOL_N1 - 1);
        else return mol_ghost<V, 1>(in, c, i0, i1, i2);
    }
#endif
#if MOL_NDIM >= 3 && !MOL_DIST
    if (i2 < MOL_ILO(V, 2) || i2 > MOL_IHI(V, 2)) {
        if (MOL_PER(V, 2)) i2 += (i2 <= 1) ? (MOL_N2 - 1) : -(MOL_N2 - 1);
        else return mol_ghost<V, 2>(in, c, i0, i1, i2);
    }
#endif
#if MOL_DIST
    {   // the last dimension is split across ranks: planes outside the slab are ghost planes from the
        // neighbouring rank (ring across a periodic seam), except beyond a non-periodic domain edge,
        // where the rank that owns the edge applies the boundary rule itself
        const int il = (MOL_NDIM == 1) ? i0 : (MOL_NDIM == 2 ? i1 : i2);
        if (il < c.loc_lo || il > c.loc_hi) {
            if (!MOL_PER(V, MOL_NDIM - 1) && (il < MOL_ILO(V, MOL_NDIM - 1) || il > MOL_IHI(V, MOL_NDIM - 1)))
                return mol_ghost<V, MOL_NDIM - 1>(in, c, i0, i1, i2);
            mol_i64 inplane = (MOL_NDIM >= 2) ? (i0 - MOL_ILO(V, 0)) : 0;
#if MOL_NDIM >= 3
            inplane += (mol_i64)(i1 - MOL_ILO(V, 1)) * MOL_EXT(V, 0);
#endif
            const bool lower = il < c.loc_lo;
            const int r = lower ? il - (c.loc_lo - MOL_HALO) : il - (c.loc_hi + 1);
            return mol_load_halo(in, lower, (mol_i64)V * MOL_HALO * MOL_PLANE_MAX + (mol_i64)r * MOL_PLANE(V) + inplane);
        }
    }
#endif
    mol_i64 flat = MOL_VOFF(V, c) + (i0 - MOL_LLO0(V, c));
#if MOL_NDIM == 2
    flat += (mol_i64)(i1 - MOL_LLO(V, c)) * MOL_EXT(V, 0);
#elif MOL_NDIM == 3
    flat += (mol_i64)(i1 - MOL_ILO(V, 1)) * MOL_EXT(V, 0)
          + (mol_i64)(i2 - MOL_LLO(V, c)) * MOL_EXT(V, 0) * MOL_EXT(V, 1);
#endif
    return mol_load(in, flat);
}

// flat index of an interior node of variable V in the state vector
template <int V>
__device__ __forceinline__ mol_i64 mol_flat(const MolCtx& c, int i0, int i1, int i2) {
    mol_i64 flat = MOL_VOFF(V, c) + (i0 - MOL_LLO0(V, c));
#if MOL_NDIM == 2
    flat += (mol_i64)(i1 - MOL_LLO(V, c)) * MOL_EXT(V, 0);
#elif MOL_NDIM == 3
    flat += (mol_i64)(i1 - MOL_ILO(V, 1)) * MOL_EXT(V, 0)
          + (mol_i64)(i2 - MOL_LLO(V, c)) * MOL_EXT(V, 0) * MOL_EXT(V, 1);
#endif
    return flat;
}

// ---- table-driven linear stencil row (generic path): sum_k w[row][k] * V(node start+k along DIM) --
template <int V, int DIM>
__device__ __forceinline__ double mol_lin_g(const MolIn& in, const MolCtx& c, int woff, int soff,
                                            int L, int row, int i0, int i1, int i2) {
    const int* sr = c.tabs + soff + 2 * row;
    const int start = __ldg(sr), nt = __ldg(sr + 1);
    const double* w = c.tabw + woff + (mol_i64)row * L;
    double acc = 0.0;
    for (int k = 0; k < nt; ++k) {
        int j0 = i0, j1 = i1, j2 = i2;
        if (DIM == 0) j0 = start + k; else if (DIM == 1) j1 = start + k; else j2 = start + k;
        acc = fma(__ldg(w + k), mol_node<V>(in, c, j0, j1, j2), acc);
    }
    return acc;
}

// ---- mixed derivative Dx Dy u (2nd_order_mixed_deriv.jl:5-22): sum_kx sum_ky wx[kx] wy[ky] V(node + x tap + y tap), both
// rows being those of the centred first-derivative operators at this node.  A tap that lies outside the interior in two
// non-periodic dimensions is a corner node, which the reference defines as 0 (generate_corner_eqs!, generate_bc_eqs.jl:396-416).
template <int V>
__device__ __forceinline__ bool mol_is_corner(int i0, int i1, int i2) {
    int out = 0;
    out += (!MOL_PER(V, 0) && (i0 < MOL_ILO(V, 0) || i0 > MOL_IHI(V, 0))) ? 1 : 0;
#if MOL_NDIM >= 2
    out += (!MOL_PER(V, 1) && (i1 < MOL_ILO(V, 1) || i1 > MOL_IHI(V, 1))) ? 1 : 0;
#endif
#if MOL_NDIM >= 3
    out += (!MOL_PER(V, 2) && (i2 < MOL_ILO(V, 2) || i2 > MOL_IHI(V, 2))) ? 1 : 0;
#endif
    return out >= 2;
}

// periodic dimensions wrap first (a mixed tap can be off in two dimensions at once; a ghost rule must see valid indices in
// the other dimensions)
template <int V>
__device__ __forceinline__ void mol_wrap_periodic(int& i0, int& i1, int& i2) {
    if (MOL_PER(V, 0) && (i0 < MOL_ILO(V, 0) || i0 > MOL_IHI(V, 0))) i0 += (i0 <= 1) ? (MOL_N0 - 1) : -(MOL_N0 - 1);
#if MOL_NDIM >= 2
    if (MOL_PER(V, 1) && (i1 < MOL_ILO(V, 1) || i1 > MOL_IHI(V, 1))) i1 += (i1 <= 1) ? (MOL_N1 - 1) : -(MOL_N1 - 1);
#endif
#if MOL_NDIM >= 3
    if (MOL_PER(V, 2) && (i2 < MOL_ILO(V, 2) || i2 > MOL_IHI(V, 2))) i2 += (i2 <= 1) ? (MOL_N2 - 1) : -(MOL_N2 - 1);
#endif
}

template <int V, int DX, int DY>
__device__ __forceinline__ double mol_mixed_g(const MolIn& in, const MolCtx& c, int wxo, int sxo, int Lx, int rowx,
                                              int wyo, int syo, int Ly, int rowy, int i0, int i1, int i2) {
    const int* srx = c.tabs + sxo + 2 * rowx;
    const int* sry = c.tabs + syo + 2 * rowy;
    const int startx = __ldg(srx), ntx = __ldg(srx + 1), starty = __ldg(sry), nty = __ldg(sry + 1);
    const double* wx = c.tabw + wxo + (mol_i64)rowx * Lx;
    const double* wy = c.tabw + wyo + (mol_i64)rowy * Ly;
    double acc = 0.0;
    for (int kx = 0; kx < ntx; ++kx) {
        double inner = 0.0;
        for (int ky = 0; ky < nty; ++ky) {
            int j[3] = {i0, i1, i2};
            j[DX] = startx + kx;
            j[DY] = starty + ky;
            mol_wrap_periodic<V>(j[0], j[1], j[2]);
            const double val = mol_is_corner<V>(j[0], j[1], j[2]) ? 0.0 : mol_node<V>(in, c, j[0], j[1], j[2]);
            inner = fma(__ldg(wy + ky), val, inner);
        }
        acc = fma(__ldg(wx + kx), inner, acc);
    }
    return acc;
}

// same row applied to the node-coordinate vector of dimension DIM (interpolated coordinates of
// the nonlinear Laplacian, nonlinear_laplacian.jl:74-84); taps wrap like the field taps do.
template <int V, int DIM>
__device__ __forceinline__ double mol_lin_coord(const MolCtx& c, int woff, int soff, int L, int row) {
    const int* sr = c.tabs + soff + 2 * row;
    const int start = __ldg(sr), nt = __ldg(sr + 1);
    const double* w = c.tabw + woff + (mol_i64)row * L;
    const int n = (DIM == 0) ? MOL_N0 : (DIM == 1 ? MOL_N1 : MOL_N2);
    double acc = 0.0;
    for (int k = 0; k < nt; ++k) {
        int j = start + k;
        if (MOL_PER(V, DIM)) { if (j <= 1) j += n - 1; else if (j > n) j -= n - 1; }
        acc = fma(__ldg(w + k), __ldg(c.grid[DIM] + j - 1), acc);
    }
    return acc;
}

// ---- FP64-pipe economy for the WENO kernels -----------------------------------------------------------------------
// These kernels are bound by ISSUE slots, not by memory: an FP64 instruction occupies its scheduler for two cycles and
// every other instruction for one (measured: the uniform 1-D kernel ran exactly at (2 x 87 FP64 + 135 other) cycles per
// node).  fmax / comparisons on doubles cost an FP64-pipe DSETP plus ~8 selects and moves each, and an IEEE division drags
// a slow-path call with it; the helpers below do the same jobs on the integer pipe where the operands allow it.

// 2^-(exponent of the largest of three POSITIVE doubles), exact: for positive doubles the order of the values is the
// order of their high words, so the maximum is taken on integers (no FP64 compare, no NaN fix-up code)
__device__ __forceinline__ double mol_pow2_inv3(double a, double b, double c) {
    const int h = max(__double2hiint(a), max(__double2hiint(b), __double2hiint(c)));
    return __hiloint2double(0x7fe00000 - (h & 0x7ff00000), 0);
}
__device__ __forceinline__ double mol_pow2_inv(double a) {
    return __hiloint2double(0x7fe00000 - (__double2hiint(a) & 0x7ff00000), 0);
}
// max(x, 0) through the sign bit (x is a rounded sum that is non-negative in exact arithmetic)
__device__ __forceinline__ double mol_clamp0(double x) {
    const int hi = __double2hiint(x), m = ~(hi >> 31);
    return __hiloint2double(hi & m, __double2loint(x) & m);
}
// a / x for x > 0 well inside the normal range (the callers scale their denominators to O(1)): reciprocal seed (>= 20
// bits), one Newton step (>= 40 bits), one residual correction of the quotient, whose error is the product of the errors
// of r and of y = a r (<= 1 ulp); no special-case branch, no slow-path call
__device__ __forceinline__ double mol_div_pos(double a, double x) {
#ifdef MOL_HOST_EMU
    return a / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    const double y = a * r;
    return fma(fma(-x, y, a), r, y);
#endif
}

// q_k proportional to 1 / f_k^2 without a division: the squared product of the other two f, after scaling the largest
// product to [1, 2) (an exact power of two: nothing over- or underflows against the reference's formula)
__device__ __forceinline__ void mol_weno_ratios_sq(double f0, double f1, double f2, double& q0, double& q1, double& q2) {
    double p0 = f1 * f2, p1 = f0 * f2, p2 = f0 * f1;
    const double s = mol_pow2_inv3(p0, p1, p2);
    p0 *= s; p1 *= s; p2 *= s;
    q0 = p0 * p0; q1 = p1 * p1; q2 = p2 * p2;
}
__device__ __forceinline__ double mol_weno_quot(double num, double den) { return mol_div_pos(num, den); }

// ---- WENO5, uniform grid: Jiang-Shu weights, WENO.jl:6-57 ---------------------------------------------------------
// Same quantities as the reference, re-derived for the FP64 pipe: B200 issues 64 FP64 operations per clock and SM and a
// correctly rounded FP64 division costs ~20 of them, which made the literal transcription (22 divisions per evaluation)
// FP64-pipe-bound at 6 % (2-D) / 12 % (1-D) of the HBM roofline.  Each change moves a term by a few ulp, far inside the
// 1e-12 parity bar (tests/test_gpu_parity.py).
__device__ __forceinline__ double mol_weno5_uniform(double u_m2, double u_m1, double u_0, double u_p1,
                                                    double u_p2, double eps, double dx) {
    // Everything in first and second differences of the five values (61 FP64 operations instead of the 87 of the form
    // written in the field values; the kernel's time is 2 x FP64 + other instructions, see above):
    //   * neighbouring nodes of a thread share D_j and S_j (identical operands: the compiler's CSE finds them);
    //   * each candidate flux is 6 u_0 + a two-term combination of the D_j, and 6 u_0 cancels between h+ and h- because both
    //     weight sets are normalised -- less arithmetic and less cancellation;
    //   * the three (eps + beta_k) are carried with a common factor 4 (it cancels in the weight ratios), which makes each a
    //     multiply and two fused multiply-adds.
    const double D1 = u_m1 - u_m2, D2 = u_0 - u_m1, D3 = u_p1 - u_0, D4 = u_p2 - u_p1;
    const double S1 = D2 - D1, S2 = D3 - D2, S3 = D4 - D3;
    // beta_1 = 13/12 S3^2 + 1/4 (D4 - 3 D3)^2,  beta_2 = 13/12 S2^2 + 1/4 (D2 + D3)^2,  beta_3 = 13/12 S1^2 + 1/4 (3 D2 - D1)^2
    const double c133 = 13.0 / 3.0, eps4 = 4.0 * eps;
    const double t2 = fma(-3.0, D3, D4), t4 = D2 + D3, t6 = fma(3.0, D2, -D1);
    const double e1 = fma(t2, t2, fma(c133 * S3, S3, eps4));
    const double e2 = fma(t4, t4, fma(c133 * S2, S2, eps4));
    const double e3 = fma(t6, t6, fma(c133 * S1, S1, eps4));
    // 6 (candidate flux) - 6 u_0
    const double gm1 = fma(2.0, D4, -5.0 * D3), gm2 = fma(-2.0, D2, -D3), gm3 = fma(-4.0, D2, D1);
    const double gp1 = fma(4.0, D3, -D4), gp2 = fma(2.0, D3, D2), gp3 = fma(5.0, D2, -2.0 * D1);
#if MOL_WENO_RATIO
    // The nonlinear weights only enter as ratios, so 1/(eps + beta_k)^2 is replaced by the squared product of the other
    // two (eps + beta) -- the common factor cancels between each weighted sum and its normalisation -- and h+ - h- is formed
    // over the common denominator: ONE division per evaluation instead of the reference's 22.  The products are scaled by an
    // exact power of two (largest to [1, 2)) before they are squared, so nothing over- or underflows.
    double q1, q2, q3;
    mol_weno_ratios_sq(e1, e2, e3, q1, q2, q3);
    // ideal weights x 10: (3, 6, 1) for h+, (1, 6, 3) for h-
    const double w1 = 3.0 * q1, w2 = 6.0 * q2, w3 = 3.0 * q3;
    const double Np = fma(w1, gp1, fma(w2, gp2, q3 * gp3)), Dp = w1 + w2 + q3;
    const double Nm = fma(q1, gm1, fma(w2, gm2, w3 * gm3)), Dm = q1 + w2 + w3;
    return mol_weno_quot(fma(Np, Dm, -(Nm * Dp)), Dp * Dm) * (1.0 / (6.0 * dx));
#else
    const double r1 = 1.0 / (e1 * e1), r2 = 1.0 / (e2 * e2), r3 = 1.0 / (e3 * e3);
    const double w1 = 3.0 * r1, w2 = 6.0 * r2, w3 = 3.0 * r3;
    const double hp = fma(w1, gp1, fma(w2, gp2, r3 * gp3)) / (w1 + w2 + r3);
    const double hm = fma(r1, gm1, fma(w2, gm2, w3 * gm3)) / (r1 + w2 + w3);
    return (hp - hm) * (1.0 / (6.0 * dx));
#endif
}

// ---- WENO5, non-uniform grid (nonuniform_weno.jl:5-163) -----------------------------------------------------------
// What the reference evaluates per point is split here into a u-independent part, built once at plan time on the host
// (csrc/mol_plan.cpp, weno_nu_tables), and the u-dependent part below.  Facts used (each sub-stencil interpolant p_k is a
// quadratic, so p_k' is linear and p_k'' constant):
//   * with D_j = u_{j+1} - u_j, both r_k = p_k'(x_i) and c_k = p_k'' are two-term combinations of D_k, D_{k+1};
//   * Simpson's rule is exact for the quadratic (p_k')^2, so the smoothness indicator is the quadratic form
//         beta_k = A r_k^2 + B r_k c_k + C c_k^2,    A = dx^2, B = dx^2 (sL + sR), C = dx^2 (sL^2 + sL sR + sR^2)/3 + dx^4
//     (cell [xL, xR] of width dx, sL = xL - x_i, sR = xR - x_i; uniform grid: A = h^2, B = 0, C = 13/12 h^4, Jiang-Shu);
//   * the nonlinear weights enter only as ratios, so 1/(eps + beta_k)^2 is replaced by the product of the other two
//     (eps + beta)^2 after an exact power-of-two scaling, and sigma+ R+ - sigma- R- is formed with ONE division.
// No Fornberg recurrences and no geometry divisions are left on the device.
//
// Two sources of the u-independent coefficients:
//   records  (explicit rows: wall targets T = 1, 2, 4, 5, irregular charts) MOL_WREC doubles per row:
//            ra[3], rb[3] (r_k = ra_k D_k + rb_k D_{k+1}), ca[3], cb[3] (c_k likewise), A, B, C, d+[3], d-[3], s+, s-
//   compact  (core rows: centre target on five consecutive nodes) three per-interval arrays h_j = x_{j+1} - x_j, 1/h_j,
//            1/(x_{j+2} - x_j); 24 B per node instead of 184 B, the rest is ~40 multiply-adds (a 1-D sweep would
//            otherwise be bound by reading the records).
#define MOL_WREC 23

// u-dependent tail shared by both sources.  d+ / d- / s+ / s- need not be normalised: `den` is the common factor they
// carry.  (mol_weno_ratios_sq / mol_clamp0 resolve to the dual-number overloads of mol_jvp.cuh when S = MolDual.)
template <class S>
__device__ __forceinline__ S mol_weno_nu_tail(const S& r0, const S& r1, const S& r2, const S& c0, const S& c1, const S& c2,
                                              double A, double B, double C, double dp0, double dp1, double dp2, double dm0,
                                              double dm1, double dm2, double sp, double sm, double den, double eps) {
    S b0 = (A * r0 + B * c0) * r0 + (C * c0) * c0;
    S b1 = (A * r1 + B * c1) * r1 + (C * c1) * c1;
    S b2 = (A * r2 + B * c2) * r2 + (C * c2) * c2;
    b0 = mol_clamp0(b0); b1 = mol_clamp0(b1); b2 = mol_clamp0(b2);
    S q0, q1, q2;
    mol_weno_ratios_sq(eps + b0, eps + b1, eps + b2, q0, q1, q2);
    const S wp0 = dp0 * q0, wp1 = dp1 * q1, wp2 = dp2 * q2;
    const S wm0 = dm0 * q0, wm1 = dm1 * q1, wm2 = dm2 * q2;
    const S Np = wp0 * r0 + wp1 * r1 + wp2 * r2, Dp = wp0 + wp1 + wp2;
    const S Nm = wm0 * r0 + wm1 * r1 + wm2 * r2, Dm = wm0 + wm1 + wm2;
    return mol_weno_quot(sp * (Np * Dm) - sm * (Nm * Dp), den * (Dp * Dm));
}

// explicit row: coefficients from its plan-time record
template <class S>
__device__ __forceinline__ S mol_weno5_nu_rec(const S u[5], const double* __restrict__ R, double eps) {
    const S D0 = u[1] - u[0], D1 = u[2] - u[1], D2 = u[3] - u[2], D3 = u[4] - u[3];
    const S r0 = __ldg(R + 0) * D0 + __ldg(R + 3) * D1, r1 = __ldg(R + 1) * D1 + __ldg(R + 4) * D2,
            r2 = __ldg(R + 2) * D2 + __ldg(R + 5) * D3;
    const S c0 = __ldg(R + 6) * D0 + __ldg(R + 9) * D1, c1 = __ldg(R + 7) * D1 + __ldg(R + 10) * D2,
            c2 = __ldg(R + 8) * D2 + __ldg(R + 11) * D3;
    return mol_weno_nu_tail<S>(r0, r1, r2, c0, c1, c2, __ldg(R + 12), __ldg(R + 13), __ldg(R + 14), __ldg(R + 15), __ldg(R + 16),
                               __ldg(R + 17), __ldg(R + 18), __ldg(R + 19), __ldg(R + 20), __ldg(R + 21), __ldg(R + 22), 1.0, eps);
}

// core row (centre target, nodes i-2 .. i+2): g points at the entry of interval i-2 in the first of three arrays of
// length glen: h, 1/h, 1/(two-interval span).  Everything else is formed here from the four spacings around the node.
// POS: the plan found all three ideal weights positive at every core node of this table (any grid whose neighbouring
// spacings differ by less than a factor ~3): the Shi-Hu-Shu splitting is then the identity (d+ = 2 d, d- = d, hence
// omega+ = omega- and 2 R - R = R exactly, also in floating point), and one weight set with one normalisation remains.
// Layout of the three per-interval quantities: element (a, k) = g[k * ks + a * as] -- (ks, as) = (1, glen) for the
// table's own arrays in global memory, (record stride, 1) for the packed records staged in shared memory (SM = true:
// plain loads; ld.global.nc must not see a shared-memory address).
template <bool SM>
__device__ __forceinline__ double mol_gld(const double* p) { return SM ? *p : __ldg(p); }
template <class S, bool POS, bool SM>
__device__ __forceinline__ S mol_weno5_nu_core(const S& um2, const S& um1, const S& u0, const S& up1, const S& up2,
                                               const double* __restrict__ g, int ks, int as, double eps) {
    const double ha = mol_gld<SM>(g), hb = mol_gld<SM>(g + ks), hc = mol_gld<SM>(g + 2 * ks), hd = mol_gld<SM>(g + 3 * ks);
    const double* gi = g + as;
    const double* gs = gi + as;
    // divided differences of the three sub-stencils: first (f) and second (s = p''/2)
    const S fa = (um1 - um2) * mol_gld<SM>(gi), fb = (u0 - um1) * mol_gld<SM>(gi + ks), fc = (up1 - u0) * mol_gld<SM>(gi + 2 * ks),
            fd = (up2 - up1) * mol_gld<SM>(gi + 3 * ks);
    const S s0 = (fb - fa) * mol_gld<SM>(gs), s1 = (fc - fb) * mol_gld<SM>(gs + ks), s2 = (fd - fc) * mol_gld<SM>(gs + 2 * ks);
    // p_k'(x_i) = f[a,b] + s_k ((x_i - a) + (x_i - b))
    const S r0 = fa + s0 * (ha + 2.0 * hb), r1 = fb + s1 * hb, r2 = fc - s2 * hc;
    // cell [x_i - hb/2, x_i + hc/2]; the quadratic form is written for s = c/2:  A r^2 + (2B) r s + (4C) s^2
    const double dx = 0.5 * (hb + hc), dx2 = dx * dx;
    const double A = dx2, B2 = dx2 * (hc - hb), C4 = dx2 * ((hb * hb - hb * hc + hc * hc) * (1.0 / 3.0) + 4.0 * dx2);
    // ideal weights of the centre target (the five-point first derivative at x_i as a combination of the three
    // sub-stencil derivatives), carried with their common denominator `den` instead of being divided out
    const double s3a = ha + hb + hc, s4 = s3a + hd, s3b = hb + hc + hd;
    double den = s3a * s4 * s3b;
    double d0 = hc * (hc + hd) * s3b, d2 = (ha + hb) * hb * s3a;
    if (POS) {
        const double d1 = den - d0 - d2;
        S b0 = (A * r0 + B2 * s0) * r0 + (C4 * s0) * s0;
        S b1 = (A * r1 + B2 * s1) * r1 + (C4 * s1) * s1;
        S b2 = (A * r2 + B2 * s2) * r2 + (C4 * s2) * s2;
        b0 = mol_clamp0(b0); b1 = mol_clamp0(b1); b2 = mol_clamp0(b2);
        S q0, q1, q2;
        mol_weno_ratios_sq(eps + b0, eps + b1, eps + b2, q0, q1, q2);
        const S w0 = d0 * q0, w1 = d1 * q1, w2 = d2 * q2;             // the common factor of the d's cancels in N / D
        // (q is O(1) after its scaling but the d's carry h^3: bring the denominator back to O(1) before dividing)
        const double sc = mol_pow2_inv(den);
        return mol_weno_quot(sc * (w0 * r0 + w1 * r1 + w2 * r2), sc * (w0 + w1 + w2));
    }
    const double sc = mol_pow2_inv(den);
    den *= sc; d0 *= sc; d2 *= sc;
    const double d1 = den - d0 - d2;
    // positive / negative splitting (theta = 3) of weights that may be negative on strongly non-uniform grids
    const double dp0 = 2.0 * d0, dp2 = 2.0 * d2, dp1 = 0.5 * d1 + 1.5 * fabs(d1);
    const double dm0 = d0, dm2 = d2, dm1 = dp1 - d1;
    const double sp = dp0 + dp1 + dp2, sm = dm0 + dm1 + dm2;
    return mol_weno_nu_tail<S>(r0, r1, r2, s0, s1, s2, A, B2, C4, dp0, dp1, dp2, dm0, dm1, dm2, sp, sm, den, eps);
}

// table-driven WENO row (generic path).  Per row two ints: first tap node, and target T | (record + 1) << 3.
// goff/glo/glen locate the compact arrays of the table (non-uniform only), roff its records.
template <int V, int DIM>
__device__ __forceinline__ double mol_weno_g(const MolIn& in, const MolCtx& c, int soff, int row, double eps,
                                             double dx_uniform, int goff, int glo, int glen, int roff, int pos, int i0, int i1,
                                             int i2) {
    const int* sr = c.tabs + soff + 2 * row;
    const int start = __ldg(sr), code = __ldg(sr + 1);
    double u[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        int j0 = i0, j1 = i1, j2 = i2;
        const int raw = start + k;
        if (DIM == 0) j0 = raw; else if (DIM == 1) j1 = raw; else j2 = raw;
        u[k] = mol_node<V>(in, c, j0, j1, j2);
    }
    if (dx_uniform != 0.0) return mol_weno5_uniform(u[0], u[1], u[2], u[3], u[4], eps, dx_uniform);
    const int rec = (code >> 3) - 1;
    if (rec >= 0) return mol_weno5_nu_rec<double>(u, c.tabw + roff + (mol_i64)rec * MOL_WREC, eps);
    if (pos) return mol_weno5_nu_core<double, true, false>(u[0], u[1], u[2], u[3], u[4], c.tabw + goff + (start - glo), 1, glen, eps);
    return mol_weno5_nu_core<double, false, false>(u[0], u[1], u[2], u[3], u[4], c.tabw + goff + (start - glo), 1, glen, eps);
}

// grid coordinate of a (possibly wrapped) node along DIM, as the taps of variable V see it
template <int V, int DIM>
__device__ __forceinline__ double mol_gx(const MolCtx& c, int j) {
    const int n = (DIM == 0) ? MOL_N0 : (DIM == 1 ? MOL_N1 : MOL_N2);
    if (MOL_PER(V, DIM)) { if (j <= 1) j += n - 1; else if (j > n) j -= n - 1; }
    j = min(max(j, 1), n);      // (overhanging cells of an edge tile are evaluated, never stored: keep their reads inside the array)
    return __ldg(c.grid[DIM] + j - 1);
}

// ---- region + fused Runge-Kutta epilogue -----------------------------------------------------------
struct MolBox { int lo[3]; int hi[3]; };      // inclusive node ranges of a region

// Fused Runge-Kutta epilogues of an FSAL embedded pair (Tsit5: stages 6 and 7), MOL_EPI =
//   2 "PRE": the last-but-one stage.  Its k is never stored: while the inputs a[j] are in registers the loader
//            also forms the partial sums of u+ and of the error estimate, and the epilogue completes them,
//              comb[f] = sum_j cb[j] a_j[f] + cbk k[f]   (= u+, the input of the last stage)
//              eout[f] = sum_j ce[j] a_j[f] + cek k[f]   (= dt sum_{j<s} btilde_j k_j)
//   3 "FIN": the last stage, on the single input u+ (TMA path): stores k (next step's k1) and accumulates
//              sum_f ((e[f] + ek k[f]) / (abstol + max(|u0[f]|, |u+[f]|) reltol))^2
#ifndef MOL_EPI
#define MOL_EPI 0
#endif
#define MOL_EPI_PRE (MOL_EPI == 2)
#define MOL_EPI_FIN (MOL_EPI == 3)
#if MOL_EPI_PRE
struct MolEpi {
    double* comb;
    double* eout;
    double cb[MOL_NIN];
    double ce[MOL_NIN];
    double cbk, cek;
};
// the three combinations of the inputs at one unknown: stage input v, partial u+ p, partial error q
__device__ __forceinline__ void mol_load3(const MolIn& in, const MolEpi& e, mol_i64 idx, double& v, double& p, double& q) {
    const double a0 = __ldg(in.a[0] + idx);
    v = in.c[0] * a0;
    p = e.cb[0] * a0;
    q = e.ce[0] * a0;
#pragma unroll
    for (int j = 1; j < MOL_NIN; ++j) {
        const double aj = __ldg(in.a[j] + idx);
        v = fma(in.c[j], aj, v);
        p = fma(e.cb[j], aj, p);
        q = fma(e.ce[j], aj, q);
    }
}
#elif MOL_EPI_FIN
struct MolEpi {
    const double* e;     // partial error estimate written by the PRE stage
    const double* u0;    // state at the start of the step
    double ek;           // dt * btilde_s
    double abstol, reltol;
    double* err;         // err[blockIdx.x] = this CTA's part of sum_f (utilde_f / sk_f)^2 (one slot per CTA of every launch)
};
__device__ __forceinline__ void mol_fin_point(const MolEpi& e, double ef, double u0f, double k, double unew, double& errsum) {
    const double ut = fma(e.ek, k, ef);
    const double sk = e.abstol + fmax(fabs(u0f), fabs(unew)) * e.reltol;
    const double r = ut / sk;
    errsum = fma(r, r, errsum);
}
#else
struct MolEpi { int unused; };
#endif

#if MOL_DEVDT
// First statement of every kernel of a MOL_DEVDT variant.  Returns false when the sweep is to be skipped (the solve has
// finished or is waiting for the host; CTA-uniform, nothing has been touched yet).  The products are rounded exactly like
// the host-driven loop's (csrc/mol_rk.cu tsit5_attempt: dt * a_sj, t + c_s * dt), so both loops take the same steps.
__device__ __forceinline__ bool mol_devdt_apply(MolIn& in, MolCtx& c, MolEpi* e) {
    const double* ctl = in.ctl;
    if (ctl[2] != 0.0) return false;
    const double dt = ctl[1];
#pragma unroll
    for (int j = 1; j < MOL_NIN; ++j) in.c[j] = __dmul_rn(dt, in.c[j]);
    c.t = __dadd_rn(ctl[0], __dmul_rn(c.t, dt));
#if MOL_EPI_PRE
#pragma unroll
    for (int j = 1; j < MOL_NIN; ++j) { e->cb[j] = __dmul_rn(dt, e->cb[j]); e->ce[j] = __dmul_rn(dt, e->ce[j]); }
    e->cbk = __dmul_rn(dt, e->cbk);
    e->cek = __dmul_rn(dt, e->cek);
#elif MOL_EPI_FIN
    e->ek = __dmul_rn(dt, e->ek);
#endif
    return true;
}
#endif

#if MOL_HAVE_TILE
// ---- shared-memory tile geometry (cells = tile + halo, x fastest) -----------------------------
// 3-D programs march along z (MOL_ZMARCH, kernels/mol_tiled.cuh): a "tile" in shared memory is then ONE xy plane
// (+ halo), kept in a ring of MOL_RING plane slots; the slot of the plane at offset dz from the one being evaluated
// is (lz + dz + R2) mod RING, where lz is the ring slot of the lowest plane the evaluation reads.
#ifndef MOL_ZMARCH
#define MOL_ZMARCH 0
#endif
#define MOL_SX (MOL_TX + 2 * MOL_R0P)
#define MOL_SY ((MOL_NDIM >= 2) ? (MOL_TY + 2 * MOL_R1) : 1)
#if MOL_ZMARCH
#define MOL_SZ 1
#else
#define MOL_SZ ((MOL_NDIM >= 3) ? (MOL_TZ + 2 * MOL_R2) : 1)
#endif
#define MOL_TILE_CELLS (MOL_SX * MOL_SY * MOL_SZ)
#define MOL_TILE_BYTES (MOL_TILE_CELLS * 8)
#define MOL_TILE_STRIDE ((MOL_TILE_BYTES + 127) / 128 * 128 / 8)     // doubles, 128 B aligned
#if MOL_ZMARCH
#define MOL_ZSLOT(lz, dz) (((lz) + (dz) + MOL_R2 >= MOL_RING) ? ((lz) + (dz) + MOL_R2 - MOL_RING) : ((lz) + (dz) + MOL_R2))
#define MOL_CELL(V, lx, ly, lz, dz) \
    ((MOL_ZSLOT(lz, dz) * MOL_NVAR + (V)) * MOL_TILE_STRIDE + ((ly) + MOL_R1) * MOL_SX + (lx) + MOL_R0P)
#else
#define MOL_CELL(V, lx, ly, lz, dz)                                                             \
    ((V) * MOL_TILE_STRIDE +                                                                    \
     (((lz) + (dz) + ((MOL_NDIM >= 3) ? MOL_R2 : 0)) * MOL_SY + ((ly) + ((MOL_NDIM >= 2) ? MOL_R1 : 0))) * MOL_SX + (lx) + MOL_R0P)
#endif
// value of variable V at offset (dx,dy,dz) from the thread's node (lx,ly,lz) of the tile
#define MOL_S(V, dx, dy, dz) sm[MOL_CELL(V, lx + (dx), ly + (dy), lz, dz)]
// the same cell of the u tile and of the v tile as one dual number (tiled Jacobian-vector product, kernels/mol_jvp.cuh)
#define MOL_SD(V, dx, dy, dz) MolDual(sm[MOL_CELL(V, lx + (dx), ly + (dy), lz, dz)], smv[MOL_CELL(V, lx + (dx), ly + (dy), lz, dz)])

// ---- per-node records of the non-uniform axes, staged in shared memory with the tile (csrc/mol_parse.cpp
// weight_records): MOL_WRSd doubles per record along dimension d, MOL_WHLd / MOL_WHHd records of halo, the array in
// c.tabw at MOL_WOFFd starting with node MOL_WLOd.  Staged by the cp.async and cooperative flavours of the 1-D / 2-D
// kernel (MOL_WSTAGE); the TMA flavour and the z-marching kernel read the tables through the read-only path instead.
#ifndef MOL_WRS0
#define MOL_WRS0 0
#define MOL_WRS1 0
#define MOL_WHL0 0
#define MOL_WHH0 0
#define MOL_WHL1 0
#define MOL_WHH1 0
#define MOL_WOFF0 0
#define MOL_WOFF1 0
#define MOL_WLO0 1
#define MOL_WLO1 1
#endif
#ifndef MOL_TMA
#define MOL_TMA 0
#endif
#ifndef MOL_WNREC0
#define MOL_WNREC0 0
#endif
#define MOL_WSTAGE ((MOL_WRS0 > 0 || MOL_WRS1 > 0) && !MOL_TMA && !MOL_ZMARCH && MOL_NDIM <= 2)
#define MOL_WN0 ((MOL_TX + MOL_WHL0 + MOL_WHH0 + 1) / 2 * 2)             // x records per tile (even)
#define MOL_WN1 ((MOL_NDIM >= 2) ? (MOL_TY + MOL_WHL1 + MOL_WHH1) : 0)
#define MOL_WSM_STRIDE (MOL_WRS0 * MOL_WN0 + MOL_WRS1 * MOL_WN1)        // doubles per pipeline stage (even)
// Dimension 0 is FIELD-major in shared memory (and in the table): field f of the node at offset k from the thread's
// node is MOL_WX(k, f); consecutive fields are MOL_WFS0 doubles apart, consecutive nodes 1.  Dimension 1 is node-major
// (a warp reads one row record: broadcast): fields 1 apart, nodes MOL_WRS1 apart.
#define MOL_WX(k, pos) (wsm + (pos) * MOL_WN0 + lx + MOL_WHL0 + (k))
#define MOL_WY(k, pos) (wsm + MOL_WRS0 * MOL_WN0 + (ly + MOL_WHL1 + (k)) * MOL_WRS1 + (pos))
#define MOL_WFS0 MOL_WN0
#define MOL_WFS1 1
#endif

// ---- block-wide sum (warp shuffles, then one value per warp through shared memory) ---------------
__device__ __forceinline__ double mol_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
