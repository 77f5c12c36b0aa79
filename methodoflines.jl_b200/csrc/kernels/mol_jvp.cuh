// mol_jvp.cuh — Jacobian-vector products of the semi-discrete RHS by forward-mode differentiation.
//
// jv = d/d(eps) f(u + eps v) at eps = 0, evaluated exactly (no finite-difference step) by running the generated
// equations on dual numbers: the value part repeats the RHS, the tangent part carries v through every stencil row,
// ghost rule, WENO reconstruction and pointwise nonlinearity.  This is what a matrix-free Newton-Krylov solver needs
// from the reference's stiff test problems (TRBDF2 / Rodas / FBDF with a Jacobian: test/Brusselator/brusselator_eq.jl:71);
// the reference gets J from ModelingToolkit's symbolic Jacobian (MOL_discretization.jl:175-191).
// Table-driven kernel only (one thread per node); compiled when MOL_KERNEL_JVP is set.
#pragma once
#if MOL_KERNEL_JVP

struct MolDual {
    double v, d;
    __device__ __forceinline__ MolDual() : v(0.0), d(0.0) {}
    __device__ __forceinline__ MolDual(double a) : v(a), d(0.0) {}
    __device__ __forceinline__ MolDual(double a, double b) : v(a), d(b) {}
};
#define MOL_DD __device__ __forceinline__
MOL_DD MolDual operator-(const MolDual& a) { return MolDual(-a.v, -a.d); }
MOL_DD MolDual operator+(const MolDual& a, const MolDual& b) { return MolDual(a.v + b.v, a.d + b.d); }
MOL_DD MolDual operator-(const MolDual& a, const MolDual& b) { return MolDual(a.v - b.v, a.d - b.d); }
MOL_DD MolDual operator*(const MolDual& a, const MolDual& b) { return MolDual(a.v * b.v, fma(a.v, b.d, a.d * b.v)); }
MOL_DD MolDual operator/(const MolDual& a, const MolDual& b) {
    const double q = a.v / b.v;
    return MolDual(q, (a.d - q * b.d) / b.v);
}
MOL_DD MolDual operator+(const MolDual& a, double b) { return MolDual(a.v + b, a.d); }
MOL_DD MolDual operator+(double a, const MolDual& b) { return MolDual(a + b.v, b.d); }
MOL_DD MolDual operator-(const MolDual& a, double b) { return MolDual(a.v - b, a.d); }
MOL_DD MolDual operator-(double a, const MolDual& b) { return MolDual(a - b.v, -b.d); }
MOL_DD MolDual operator*(const MolDual& a, double b) { return MolDual(a.v * b, a.d * b); }
MOL_DD MolDual operator*(double a, const MolDual& b) { return MolDual(a * b.v, a * b.d); }
MOL_DD MolDual operator/(const MolDual& a, double b) { return MolDual(a.v / b, a.d / b); }
MOL_DD MolDual operator/(double a, const MolDual& b) {
    const double q = a / b.v;
    return MolDual(q, -q * b.d / b.v);
}
#define MOL_DCMP(op)                                                                         \
    MOL_DD bool operator op(const MolDual& a, const MolDual& b) { return a.v op b.v; }      \
    MOL_DD bool operator op(const MolDual& a, double b) { return a.v op b; }                \
    MOL_DD bool operator op(double a, const MolDual& b) { return a op b.v; }
MOL_DCMP(>) MOL_DCMP(>=) MOL_DCMP(<) MOL_DCMP(<=) MOL_DCMP(==) MOL_DCMP(!=)
#undef MOL_DCMP
MOL_DD MolDual fma(double a, const MolDual& b, const MolDual& c) { return MolDual(fma(a, b.v, c.v), fma(a, b.d, c.d)); }
MOL_DD MolDual sqrt(const MolDual& a) { const double s = sqrt(a.v); return MolDual(s, a.d / (2.0 * s)); }
MOL_DD MolDual exp(const MolDual& a) { const double e = exp(a.v); return MolDual(e, e * a.d); }
MOL_DD MolDual log(const MolDual& a) { return MolDual(log(a.v), a.d / a.v); }
MOL_DD MolDual sin(const MolDual& a) { return MolDual(sin(a.v), cos(a.v) * a.d); }
MOL_DD MolDual cos(const MolDual& a) { return MolDual(cos(a.v), -sin(a.v) * a.d); }
MOL_DD MolDual tan(const MolDual& a) { const double t = tan(a.v); return MolDual(t, (1.0 + t * t) * a.d); }
MOL_DD MolDual sinh(const MolDual& a) { return MolDual(sinh(a.v), cosh(a.v) * a.d); }
MOL_DD MolDual cosh(const MolDual& a) { return MolDual(cosh(a.v), sinh(a.v) * a.d); }
MOL_DD MolDual tanh(const MolDual& a) { const double t = tanh(a.v); return MolDual(t, (1.0 - t * t) * a.d); }
MOL_DD MolDual fabs(const MolDual& a) { return MolDual(fabs(a.v), (a.v < 0.0) ? -a.d : a.d); }
MOL_DD MolDual asin(const MolDual& a) { return MolDual(asin(a.v), a.d / sqrt(1.0 - a.v * a.v)); }
MOL_DD MolDual acos(const MolDual& a) { return MolDual(acos(a.v), -a.d / sqrt(1.0 - a.v * a.v)); }
MOL_DD MolDual atan(const MolDual& a) { return MolDual(atan(a.v), a.d / (1.0 + a.v * a.v)); }
MOL_DD MolDual erf(const MolDual& a) { return MolDual(erf(a.v), 1.1283791670955126 * exp(-a.v * a.v) * a.d); }
MOL_DD MolDual fmin(const MolDual& a, const MolDual& b) { return (a.v <= b.v) ? a : b; }
MOL_DD MolDual fmax(const MolDual& a, const MolDual& b) { return (a.v >= b.v) ? a : b; }
MOL_DD MolDual pow(const MolDual& a, const MolDual& b) {
    const double p = pow(a.v, b.v);
    double d = (b.v != 0.0) ? b.v * pow(a.v, b.v - 1.0) * a.d : 0.0;
    if (b.d != 0.0) d = fma(p * log(a.v), b.d, d);
    return MolDual(p, d);
}

// ---- dual state: value from u, tangent from v -------------------------------------------------------------------------
struct MolJv { const double* v; };

template <int V, int D>
__device__ MolDual mol_ghost_d(const MolIn& in, const MolJv& jv, const MolCtx& c, int i0, int i1, int i2);

// mol_node on dual numbers: same resolution order (periodic wrap or ghost rule, one dimension at a time)
template <int V>
__device__ __forceinline__ MolDual mol_node_d(const MolIn& in, const MolJv& jv, const MolCtx& c, int i0, int i1, int i2) {
    if (i0 < MOL_ILO(V, 0) || i0 > MOL_IHI(V, 0)) {
        if (MOL_PER(V, 0)) i0 += (i0 <= 1) ? (MOL_N0 - 1) : -(MOL_N0 - 1);
        else return mol_ghost_d<V, 0>(in, jv, c, i0, i1, i2);
    }
#if MOL_NDIM >= 2
    if (i1 < MOL_ILO(V, 1) || i1 > MOL_IHI(V, 1)) {
        if (MOL_PER(V, 1)) i1 += (i1 <= 1) ? (MOL_N1 - 1) : -(MOL_N1 - 1);
        else return mol_ghost_d<V, 1>(in, jv, c, i0, i1, i2);
    }
#endif
#if MOL_NDIM >= 3
    if (i2 < MOL_ILO(V, 2) || i2 > MOL_IHI(V, 2)) {
        if (MOL_PER(V, 2)) i2 += (i2 <= 1) ? (MOL_N2 - 1) : -(MOL_N2 - 1);
        else return mol_ghost_d<V, 2>(in, jv, c, i0, i1, i2);
    }
#endif
    const mol_i64 f = mol_flat<V>(c, i0, i1, i2);
    return MolDual(__ldg(in.a[0] + f), __ldg(jv.v + f));
}

template <int V, int DIM>
__device__ __forceinline__ MolDual mol_lin_d(const MolIn& in, const MolJv& jv, const MolCtx& c, int woff, int soff, int L, int row,
                                             int i0, int i1, int i2) {
    const int* sr = c.tabs + soff + 2 * row;
    const int start = __ldg(sr), nt = __ldg(sr + 1);
    const double* w = c.tabw + woff + (mol_i64)row * L;
    MolDual acc;
    for (int k = 0; k < nt; ++k) {
        int j0 = i0, j1 = i1, j2 = i2;
        if (DIM == 0) j0 = start + k; else if (DIM == 1) j1 = start + k; else j2 = start + k;
        acc = fma(__ldg(w + k), mol_node_d<V>(in, jv, c, j0, j1, j2), acc);
    }
    return acc;
}

template <int V, int DX, int DY>
__device__ __forceinline__ MolDual mol_mixed_d(const MolIn& in, const MolJv& jv, const MolCtx& c, int wxo, int sxo, int Lx, int rowx,
                                               int wyo, int syo, int Ly, int rowy, int i0, int i1, int i2) {
    const int* srx = c.tabs + sxo + 2 * rowx;
    const int* sry = c.tabs + syo + 2 * rowy;
    const int startx = __ldg(srx), ntx = __ldg(srx + 1), starty = __ldg(sry), nty = __ldg(sry + 1);
    const double* wx = c.tabw + wxo + (mol_i64)rowx * Lx;
    const double* wy = c.tabw + wyo + (mol_i64)rowy * Ly;
    MolDual acc;
    for (int kx = 0; kx < ntx; ++kx) {
        MolDual inner;
        for (int ky = 0; ky < nty; ++ky) {
            int j[3] = {i0, i1, i2};
            j[DX] = startx + kx;
            j[DY] = starty + ky;
            mol_wrap_periodic<V>(j[0], j[1], j[2]);
            if (!mol_is_corner<V>(j[0], j[1], j[2])) inner = fma(__ldg(wy + ky), mol_node_d<V>(in, jv, c, j[0], j[1], j[2]), inner);
        }
        acc = fma(__ldg(wx + kx), inner, acc);
    }
    return acc;
}

// uniform WENO5 on dual numbers: the form of mol_weno5_uniform (mol_device.cuh) -- first / second differences, candidate
// fluxes relative to 6 u_0, (eps + beta_k) carried with a common factor 4, weights x 10 -- with dual field values; the
// weight ratios and the final quotient are formed directly (mol_weno_ratios_sq / mol_weno_quot overloads below)
__device__ __forceinline__ void mol_weno_ratios_sq(const MolDual& f0, const MolDual& f1, const MolDual& f2, MolDual& q0, MolDual& q1,
                                                   MolDual& q2);
__device__ __forceinline__ MolDual mol_weno_quot(const MolDual& num, const MolDual& den);
__device__ __forceinline__ MolDual mol_weno5_uniform_d(const MolDual& u_m2, const MolDual& u_m1, const MolDual& u_0,
                                                       const MolDual& u_p1, const MolDual& u_p2, double eps, double dx) {
    const MolDual D1 = u_m1 - u_m2, D2 = u_0 - u_m1, D3 = u_p1 - u_0, D4 = u_p2 - u_p1;
    const MolDual S1 = D2 - D1, S2 = D3 - D2, S3 = D4 - D3;
    const double c133 = 13.0 / 3.0, eps4 = 4.0 * eps;
    const MolDual t2 = D4 - 3.0 * D3, t4 = D2 + D3, t6 = 3.0 * D2 - D1;
    const MolDual e1 = t2 * t2 + (c133 * (S3 * S3) + eps4);
    const MolDual e2 = t4 * t4 + (c133 * (S2 * S2) + eps4);
    const MolDual e3 = t6 * t6 + (c133 * (S1 * S1) + eps4);
    const MolDual gm1 = 2.0 * D4 - 5.0 * D3, gm2 = -(2.0 * D2 + D3), gm3 = D1 - 4.0 * D2;
    const MolDual gp1 = 4.0 * D3 - D4, gp2 = 2.0 * D3 + D2, gp3 = 5.0 * D2 - 2.0 * D1;
    MolDual q1, q2, q3;
    mol_weno_ratios_sq(e1, e2, e3, q1, q2, q3);
    const MolDual w1 = 3.0 * q1, w2 = 6.0 * q2, w3 = 3.0 * q3;
    const MolDual Np = w1 * gp1 + w2 * gp2 + q3 * gp3, Dp = w1 + w2 + q3;
    const MolDual Nm = q1 * gm1 + w2 * gm2 + w3 * gm3, Dm = q1 + w2 + w3;
    return mol_weno_quot(Np * Dm - Nm * Dp, Dp * Dm) * (1.0 / (6.0 * dx));
}

// non-uniform WENO5 on dual numbers: the templates of mol_device.cuh (mol_weno5_nu_rec / mol_weno5_nu_core) with
// S = MolDual; the plan-time geometry stays plain FP64.
MOL_DD MolDual mol_clamp0(const MolDual& x) { return (x.v >= 0.0) ? x : MolDual(0.0); }
MOL_DD MolDual mol_weno_quot(const MolDual& num, const MolDual& den) { return num / den; }
// q_k proportional to 1 / f_k^2, as in the FP64 path: the squared product of the other two (the common factor cancels in the
// weight ratios), scaled by an exact power of two taken from the value parts -- no division
MOL_DD void mol_weno_ratios_sq(const MolDual& f0, const MolDual& f1, const MolDual& f2, MolDual& q0, MolDual& q1, MolDual& q2) {
    MolDual p0 = f1 * f2, p1 = f0 * f2, p2 = f0 * f1;
    const double s = mol_pow2_inv3(p0.v, p1.v, p2.v);
    p0 = p0 * s; p1 = p1 * s; p2 = p2 * s;
    q0 = p0 * p0; q1 = p1 * p1; q2 = p2 * p2;
}

template <int V, int DIM>
__device__ __forceinline__ MolDual mol_weno_d(const MolIn& in, const MolJv& jv, const MolCtx& c, int soff, int row, double eps,
                                              double dx_uniform, int goff, int glo, int glen, int roff, int pos, int i0, int i1,
                                              int i2) {
    const int* sr = c.tabs + soff + 2 * row;
    const int start = __ldg(sr), code = __ldg(sr + 1);
    MolDual u[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        int j0 = i0, j1 = i1, j2 = i2;
        const int raw = start + k;
        if (DIM == 0) j0 = raw; else if (DIM == 1) j1 = raw; else j2 = raw;
        u[k] = mol_node_d<V>(in, jv, c, j0, j1, j2);
    }
    if (dx_uniform != 0.0) return mol_weno5_uniform_d(u[0], u[1], u[2], u[3], u[4], eps, dx_uniform);
    const int rec = (code >> 3) - 1;
    if (rec >= 0) return mol_weno5_nu_rec<MolDual>(u, c.tabw + roff + (mol_i64)rec * MOL_WREC, eps);
    (void)pos;      // the general form covers both cases; this kernel is not issue-bound
    return mol_weno5_nu_core<MolDual, false, false>(u[0], u[1], u[2], u[3], u[4], c.tabw + goff + (start - glo), 1, glen, eps);
}
#undef MOL_DD
#endif  // MOL_KERNEL_JVP
