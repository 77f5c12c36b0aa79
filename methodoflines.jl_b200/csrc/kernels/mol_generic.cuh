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
        if (threadIdx.x == 0 && epi.err) atomicAdd(epi.err, v);
    }
#endif
}

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
