// mol_tiled.cuh — the hot kernel: persistent, shared-memory tiled RHS evaluation.
//
// One CTA loops over tiles of the *core box* (the part of the interior where every stencil of every equation has
// the same taps relative to its node — the reference's "core box" of array_discretization.jl:368-420; literal weights
// on uniform axes, per-node table weights on non-uniform ones).  A tile plus its halo is staged in shared memory and
// all variables of the system are evaluated in one pass from it (u and v of the Brusselator are read once, du and dv
// written once => 32 B per grid point).  Each thread owns VX=2 consecutive x nodes (128-bit stores) times PY
// consecutive rows.  Tiles are handed out by a dynamic ticket queue.
//
// Staging flavours (chosen per kernel variant by the library, csrc/mol_plan.cpp):
//   MOL_TMA      one cp.async.bulk.tensor load per tile and variable, mbarrier completion, MOL_STAGES tiles deep
//                (cells outside the stored state are zero-filled by the TMA unit, then patched: periodic images /
//                ghost rules, only in tiles that touch the domain edge);
//   MOL_CPASYNC  the same pipeline with per-thread cp.async copies, for layouts the TMA unit cannot address
//                (row pitch not a multiple of 16 B, 1-D programs);
//   neither      cooperative loader: with MOL_NIN > 1 the input is the Runge-Kutta stage combination
//                u + dt*sum_j a_sj k_j, formed in registers on load; MOL_EPI = 2/3 add the fused epilogues of an FSAL
//                embedded pair (see mol_device.cuh).
// MOL_ZMARCH (3-D programs): xy tiles marching along z through a ring of planes (second half of this file).
#pragma once

#define MOL_NTX (MOL_TX / MOL_VX)
#define MOL_NTXT ((MOL_NTX < MOL_NTHREADS) ? MOL_NTX : MOL_NTHREADS)
#define MOL_PX (MOL_NTX / MOL_NTXT)
#define MOL_NTY (MOL_NTHREADS / MOL_NTXT)
#define MOL_PY ((MOL_TY + MOL_NTY - 1) / MOL_NTY)

struct alignas(64) MolTensorMap { unsigned char bytes[128]; };

#ifdef MOL_HOST_EMU
// ---- tests/cuda_emu only: synchronous host stand-ins for the asynchronous copy machinery, so that the staging logic
// around it (stage / ring bookkeeping, ticket order, which cells are copied and which are patched) runs on a CPU.
// Never defined by the library: the product path below is inline PTX.
#include <atomic>
struct MolEmuMap { const double* base; long long dim[3]; long long stride[3]; int box[3]; };   // stride in elements
__device__ __forceinline__ void mol_mbar_init(mol_u64* bar, unsigned) { *bar = 0; }               // {completed phases << 32 | pending bytes}
__device__ __forceinline__ void mol_mbar_expect_tx(mol_u64* bar, unsigned bytes) {
    reinterpret_cast<std::atomic<mol_u64>*>(bar)->fetch_add(bytes);
}
__device__ __forceinline__ void mol_mbar_wait(mol_u64* bar, unsigned parity) {
    while (((reinterpret_cast<std::atomic<mol_u64>*>(bar)->load() >> 32) & 1u) == parity) std::this_thread::yield();
}
__device__ __forceinline__ void mol_fence_proxy_async() {}
__device__ __forceinline__ void mol_fence_mbar_init() {}
__device__ __forceinline__ void mol_tma_load(void* dst, const MolTensorMap* map, mol_u64* bar, int c0, int c1, int c2) {
    const MolEmuMap* m = reinterpret_cast<const MolEmuMap*>(map->bytes);
    double* d = reinterpret_cast<double*>(dst);
    const int c[3] = {c0, (MOL_NDIM >= 2) ? c1 : 0, (MOL_NDIM >= 3) ? c2 : 0};
    for (int z = 0; z < m->box[2]; ++z)
        for (int y = 0; y < m->box[1]; ++y)
            for (int x = 0; x < m->box[0]; ++x) {
                const long long i0 = c[0] + x, i1 = c[1] + y, i2 = c[2] + z;
                const bool in_ = i0 >= 0 && i0 < m->dim[0] && i1 >= 0 && i1 < m->dim[1] && i2 >= 0 && i2 < m->dim[2];
                *d++ = in_ ? m->base[i0 * m->stride[0] + i1 * m->stride[1] + i2 * m->stride[2]] : 0.0;     // OOB cells are zero-filled
            }
    const mol_u64 bytes = (mol_u64)m->box[0] * m->box[1] * m->box[2] * 8;
    std::atomic<mol_u64>* b = reinterpret_cast<std::atomic<mol_u64>*>(bar);
    if (((b->fetch_sub(bytes) - bytes) & 0xffffffffull) == 0) b->fetch_add(1ull << 32);                       // phase complete
}
__device__ __forceinline__ void mol_tma_prefetch_l2(const MolTensorMap*, int, int, int) {}
#else
// ---- PTX wrappers: mbarrier + TMA ----------------------------------------------------------------
__device__ __forceinline__ unsigned mol_smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mol_mbar_init(mol_u64* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mol_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mol_mbar_expect_tx(mol_u64* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mol_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mol_mbar_wait(mol_u64* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MOL_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MOL_DONE_%=;\n"
        "bra MOL_WAIT_%=;\n"
        "MOL_DONE_%=:\n"
        "}\n" ::"r"(mol_smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mol_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mol_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mol_tma_load(void* dst, const MolTensorMap* map, mol_u64* bar, int c0, int c1, int c2) {
#if MOL_NDIM == 1
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
                 ::"r"(mol_smem_u32(dst)), "l"(map), "r"(mol_smem_u32(bar)), "r"(c0) : "memory");
#elif MOL_NDIM == 2
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(mol_smem_u32(dst)), "l"(map), "r"(mol_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
#else
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(mol_smem_u32(dst)), "l"(map), "r"(mol_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
#endif
}

// L2 prefetch of a tile (no shared memory, no completion tracking): shortens the latency of the bulk-tensor load
// that follows a few steps later
__device__ __forceinline__ void mol_tma_prefetch_l2(const MolTensorMap* map, int c0, int c1, int c2) {
#if MOL_NDIM == 1
    asm volatile("cp.async.bulk.prefetch.tensor.1d.L2.global.tile [%0, {%1}];" ::"l"(map), "r"(c0) : "memory");
#elif MOL_NDIM == 2
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
#else
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
#endif
}
#endif  // MOL_HOST_EMU

struct MolTileMaps { MolTensorMap m[MOL_NVAR]; };

// boxes of nodes [lo, hi] this launch evaluates (the core box, the part of it a slab owns, or the two
// slab-edge parts in one launch) and their decomposition into tiles
struct MolTileBox { int nt0, nt1, nt2, ntiles; int lo[3]; int hi[3]; };
// rot: tiles of the FIRST box are visited in rotated order, tile (k + rot) mod ntiles(box 0) for ticket k: a slab
// launched as one box visits its first row of tiles -- the one that reads the lower ghost planes -- last.
// wflag / wseq (slab mode, fused ghost-plane wait): tiles that read ghost planes first wait until both flags, which the
// neighbouring ranks' copy engines bump after pushing their edge planes into this rank's pool, have reached wseq.
struct MolTiles { MolTileBox b[2]; int ntiles; int rot; int* counter; const unsigned long long* wflag[2]; unsigned long long wseq; };

// Field-by-field selects (constant-bank loads): indexing T.b[] with a run-time value would make the
// compiler copy the kernel parameter to the local stack and read it back with LDL in the tile loop.
#define MOL_TB(T, second, f) ((second) ? (T).b[1].f : (T).b[0].f)
__device__ __forceinline__ void mol_tile_origin(const MolTiles& T, int tile, int& X0, int& Y0, int& Z0, int& H0, int& H1,
                                                int& H2) {
    const bool second = tile >= T.b[0].ntiles;
    if (second) tile -= T.b[0].ntiles;
    else { tile += T.rot; if (tile >= T.b[0].ntiles) tile -= T.b[0].ntiles; }
    const int nt0 = MOL_TB(T, second, nt0), nt1 = MOL_TB(T, second, nt1);
    const int b0 = tile % nt0;
    const int b1 = (tile / nt0) % nt1;
    const int b2 = tile / (nt0 * nt1);
    X0 = MOL_TB(T, second, lo[0]) + b0 * MOL_TX;
    Y0 = (MOL_NDIM >= 2) ? MOL_TB(T, second, lo[1]) + b1 * MOL_TY : 1;
    Z0 = (MOL_NDIM >= 3) ? MOL_TB(T, second, lo[2]) + b2 * MOL_TZ : 1;
    H0 = MOL_TB(T, second, hi[0]);
    H1 = MOL_TB(T, second, hi[1]);
    H2 = MOL_TB(T, second, hi[2]);
}
__device__ __forceinline__ void mol_tile_origin(const MolTiles& T, int tile, int& X0, int& Y0, int& Z0) {
    int h0, h1, h2;
    mol_tile_origin(T, tile, X0, Y0, Z0, h0, h1, h2);
}

// does the tile (with halo) reach outside the interior box of any variable?
__device__ __forceinline__ bool mol_tile_touches_edge(const MolCtx& c, int X0, int Y0, int Z0) {
    bool e = (X0 - MOL_R0 < MOL_ILO_MAX0) || (X0 + MOL_TX - 1 + MOL_R0 > MOL_IHI_MIN0);
#if MOL_DIST
    {   // ... or outside this rank's slab (ghost planes come from the neighbouring rank)
        const int l0 = (MOL_NDIM == 2) ? Y0 - MOL_R1 : Z0 - MOL_R2;
        const int l1 = (MOL_NDIM == 2) ? Y0 + MOL_TY - 1 + MOL_R1 : Z0 + MOL_TZ - 1 + MOL_R2;
        e = e || (l0 < c.loc_lo) || (l1 > c.loc_hi);
    }
#endif
#if MOL_NDIM >= 2
    e = e || (Y0 - MOL_R1 < MOL_ILO_MAX1) || (Y0 + MOL_TY - 1 + MOL_R1 > MOL_IHI_MIN1);
#endif
#if MOL_NDIM >= 3
    e = e || (Z0 - MOL_R2 < MOL_ILO_MAX2) || (Z0 + MOL_TZ - 1 + MOL_R2 > MOL_IHI_MIN2);
#endif
    return e;
}

// Fused ghost-plane wait (slab mode): does this tile read planes of a neighbouring rank?  If so the CTA waits, once per such
// tile, for the two sequence flags of this RHS evaluation (system-scope acquire loads by one thread, then a barrier).
// The planes were pushed by the neighbours' copy engines BEFORE their flag writes (stream order on their side), and this
// kernel has not touched those addresses before, so the loads that follow see them.
__device__ __forceinline__ void mol_wait_ghost_planes(const MolTiles& T, const MolCtx& c, int Y0, int Z0) {
#if MOL_DIST
    if (T.wflag[0] == nullptr && T.wflag[1] == nullptr) return;
    const int l0 = (MOL_NDIM == 2) ? Y0 - MOL_R1 : Z0 - MOL_R2;
    const int l1 = (MOL_NDIM == 2) ? Y0 + MOL_TY - 1 + MOL_R1 : Z0 + MOL_TZ - 1 + MOL_R2;
    if (l0 >= c.loc_lo && l1 <= c.loc_hi) return;            // CTA-uniform
#ifndef MOL_HOST_EMU
    if (threadIdx.x == 0) {
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const unsigned long long* f = T.wflag[side];
            if (f == nullptr) continue;
            unsigned long long v;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            } while (v < T.wseq);
        }
    }
#endif
    __syncthreads();
#endif
}

// fill (or patch) the cells of one variable's tile that the TMA unit could not supply
// PRE epilogue: two more tiles per variable behind the stage-input tiles hold the partial u+ / error sums
// (not in z-march mode, where the ring of planes takes the shared memory: the epilogue re-reads the inputs there)
// The aux tiles cover the tile's own nodes only (no halo): index ((lz * TY) + ly) * TX + lx, variable V's P tile at
// MOL_AUX_P(V), its Q tile at MOL_AUX_Q(V), both relative to the start of the stage-input tiles.
#define MOL_AUX_STRIDE (MOL_TX * MOL_TY * ((MOL_NDIM >= 3) ? MOL_TZ : 1))
#define MOL_AUX_P(V) (MOL_NVAR * MOL_TILE_STRIDE + (V) * MOL_AUX_STRIDE)
#define MOL_AUX_Q(V) (MOL_NVAR * MOL_TILE_STRIDE + (MOL_NVAR + (V)) * MOL_AUX_STRIDE)
#define MOL_PRE_AUX (MOL_EPI_PRE && !MOL_ZMARCH)

template <int V, bool ALL>
__device__ __forceinline__ void mol_tile_fill(double* sm, const MolIn& in, const MolCtx& c, const MolEpi* epi, int X0, int Y0,
                                              int Z0) {
    for (int cell = threadIdx.x; cell < MOL_TILE_CELLS; cell += MOL_NTHREADS) {
        const int sx = cell % MOL_SX;
        const int sy = (cell / MOL_SX) % MOL_SY;
        const int sz = cell / (MOL_SX * MOL_SY);
        const int n0 = X0 - MOL_R0P + sx;
        const int n1 = (MOL_NDIM >= 2) ? Y0 - MOL_R1 + sy : 1;
        const int n2 = (MOL_NDIM >= 3) ? Z0 - MOL_R2 + sz : 1;
        bool inside = (n0 >= MOL_ILO(V, 0)) && (n0 <= MOL_IHI(V, 0));
        bool near_ = (n0 >= MOL_ILO(V, 0) - MOL_R0) && (n0 <= MOL_IHI(V, 0) + MOL_R0);
#if MOL_NDIM >= 2
        inside = inside && (n1 >= MOL_ILO(V, 1)) && (n1 <= MOL_IHI(V, 1));
        near_ = near_ && (n1 >= MOL_ILO(V, 1) - MOL_R1) && (n1 <= MOL_IHI(V, 1) + MOL_R1);
#endif
#if MOL_NDIM >= 3
        inside = inside && (n2 >= MOL_ILO(V, 2)) && (n2 <= MOL_IHI(V, 2));
        near_ = near_ && (n2 >= MOL_ILO(V, 2) - MOL_R2) && (n2 <= MOL_IHI(V, 2) + MOL_R2);
#endif
#if MOL_DIST
        {   // planes outside this rank's slab are ghost planes (mol_node reads them), never state
            const int nl = (MOL_NDIM == 2) ? n1 : n2;
            const int rl = (MOL_NDIM == 2) ? MOL_R1 : MOL_R2;
            near_ = near_ && (nl >= c.loc_lo - rl) && (nl <= c.loc_hi + rl);
            if (nl < c.loc_lo || nl > c.loc_hi) inside = false;
        }
#endif
        if (inside) {
#if MOL_PRE_AUX
            if (ALL) {
                double v, p, q;
                mol_load3(in, *epi, mol_flat<V>(c, n0, n1, n2), v, p, q);
                sm[cell] = v;
                const int ax = sx - MOL_R0P, ay = sy - ((MOL_NDIM >= 2) ? MOL_R1 : 0), az = sz - ((MOL_NDIM >= 3) ? MOL_R2 : 0);
                if (ax >= 0 && ax < MOL_TX && ay >= 0 && ay < MOL_TY && az >= 0 && az < ((MOL_NDIM >= 3) ? MOL_TZ : 1)) {
                    double* tile0 = sm - V * MOL_TILE_STRIDE;
                    tile0[MOL_AUX_P(V) + (az * MOL_TY + ay) * MOL_TX + ax] = p;
                    tile0[MOL_AUX_Q(V) + (az * MOL_TY + ay) * MOL_TX + ax] = q;
                }
            }
#else
            if (ALL) sm[cell] = mol_load(in, mol_flat<V>(c, n0, n1, n2));
#endif
        } else {
            sm[cell] = near_ ? mol_node<V>(in, c, n0, n1, n2) : 0.0;
        }
    }
}

#if MOL_KERNEL_JVP
// Tiled Jacobian-vector product: a second set of tiles holds the direction v.  Cells that are stored state are loaded
// from u and v; every other cell (periodic images, ghost nodes) gets the value AND the tangent of its resolution rule
// (mol_node_d: the tangent of a Dirichlet ghost is 0, of a Neumann / Robin / extrapolated one its linear part applied to v).
#define MOL_JVP_VBASE (MOL_NVAR * MOL_TILE_STRIDE + (MOL_WSTAGE ? MOL_WSM_STRIDE : 0))      // (cooperative flavour: one stage)
template <int V>
__device__ __forceinline__ void mol_tile_fill_d(double* sm, double* smv, const MolIn& in, const MolJv& jv, const MolCtx& c, int X0,
                                                int Y0, int Z0) {
    for (int cell = threadIdx.x; cell < MOL_TILE_CELLS; cell += MOL_NTHREADS) {
        const int sx = cell % MOL_SX;
        const int sy = (cell / MOL_SX) % MOL_SY;
        const int sz = cell / (MOL_SX * MOL_SY);
        const int n0 = X0 - MOL_R0P + sx;
        const int n1 = (MOL_NDIM >= 2) ? Y0 - MOL_R1 + sy : 1;
        const int n2 = (MOL_NDIM >= 3) ? Z0 - MOL_R2 + sz : 1;
        bool inside = (n0 >= MOL_ILO(V, 0)) && (n0 <= MOL_IHI(V, 0));
        bool near_ = (n0 >= MOL_ILO(V, 0) - MOL_R0) && (n0 <= MOL_IHI(V, 0) + MOL_R0);
#if MOL_NDIM >= 2
        inside = inside && (n1 >= MOL_ILO(V, 1)) && (n1 <= MOL_IHI(V, 1));
        near_ = near_ && (n1 >= MOL_ILO(V, 1) - MOL_R1) && (n1 <= MOL_IHI(V, 1) + MOL_R1);
#endif
#if MOL_NDIM >= 3
        inside = inside && (n2 >= MOL_ILO(V, 2)) && (n2 <= MOL_IHI(V, 2));
        near_ = near_ && (n2 >= MOL_ILO(V, 2) - MOL_R2) && (n2 <= MOL_IHI(V, 2) + MOL_R2);
#endif
        if (inside) {
            const mol_i64 f = mol_flat<V>(c, n0, n1, n2);
            sm[cell] = __ldg(in.a[0] + f);
            smv[cell] = __ldg(jv.v + f);
        } else if (near_) {
            const MolDual d = mol_node_d<V>(in, jv, c, n0, n1, n2);
            sm[cell] = d.v;
            smv[cell] = d.d;
        } else {
            sm[cell] = 0.0;
            smv[cell] = 0.0;
        }
    }
}
template <int V>
struct MolFillVarsD {
    static __device__ __forceinline__ void run(double* sm, double* smv, const MolIn& in, const MolJv& jv, const MolCtx& c, int X0, int Y0,
                                               int Z0) {
        mol_tile_fill_d<V>(sm + V * MOL_TILE_STRIDE, smv + V * MOL_TILE_STRIDE, in, jv, c, X0, Y0, Z0);
        MolFillVarsD<V + 1>::run(sm, smv, in, jv, c, X0, Y0, Z0);
    }
};
template <>
struct MolFillVarsD<MOL_NVAR> {
    static __device__ __forceinline__ void run(double*, double*, const MolIn&, const MolJv&, const MolCtx&, int, int, int) {}
};
#endif

// is the tile, including its (even-padded) halo, entirely made of stored state of every variable?
__device__ __forceinline__ bool mol_tile_fully_inside(const MolCtx& c, int X0, int Y0, int Z0) {
    bool in_ = (X0 - MOL_R0P >= MOL_ILO_MAX0) && (X0 + MOL_TX - 1 + MOL_R0P <= MOL_IHI_MIN0);
#if MOL_NDIM >= 2
    in_ = in_ && (Y0 - MOL_R1 >= MOL_ILO_MAX1) && (Y0 + MOL_TY - 1 + MOL_R1 <= MOL_IHI_MIN1);
#endif
#if MOL_NDIM >= 3
    in_ = in_ && (Z0 - MOL_R2 >= MOL_ILO_MAX2) && (Z0 + MOL_TZ - 1 + MOL_R2 <= MOL_IHI_MIN2);
#endif
#if MOL_DIST
    {
        const int l0 = (MOL_NDIM == 2) ? Y0 - MOL_R1 : Z0 - MOL_R2;
        const int l1 = (MOL_NDIM == 2) ? Y0 + MOL_TY - 1 + MOL_R1 : Z0 + MOL_TZ - 1 + MOL_R2;
        in_ = in_ && (l0 >= c.loc_lo) && (l1 <= c.loc_hi);
    }
#endif
    return in_;
}

// Cooperative loader for such tiles: value = sum_j c[j] * a[j][..] formed in registers (the fused Runge-Kutta stage
// input); rows of the tile are contiguous in the state arrays.  MOL_FW = 2 x nodes per 128-bit load when the layout
// guarantees 16 B alignment of every row (even row pitch), else one node per load (odd pitches, e.g. the n - 2
// unknowns per row of a Dirichlet/Neumann problem on n = 2^k + 1 nodes).  MOL_FILL_UNROLL iterations are issued
// together so that about 8-12 independent loads per thread are in flight.
#if MOL_VEC_ST
#define MOL_FW 2
struct MolFv { double x, y; };
__device__ __forceinline__ MolFv mol_fv_ld(const double* p) { const double2 t = __ldg(reinterpret_cast<const double2*>(p)); return {t.x, t.y}; }
__device__ __forceinline__ void mol_fv_st(double* p, const MolFv& v) { *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y); }
#else
#define MOL_FW 1
struct MolFv { double x, y; };
__device__ __forceinline__ MolFv mol_fv_ld(const double* p) { return {__ldg(p), 0.0}; }
__device__ __forceinline__ void mol_fv_st(double* p, const MolFv& v) { *p = v.x; }
#endif
#define MOL_FILL_UNROLL (MOL_PRE_AUX ? 1 : ((MOL_NIN <= 2) ? 5 : ((MOL_NIN == 3) ? 4 : ((MOL_NIN == 4) ? 3 : 2))))
template <int V>
__device__ __forceinline__ void mol_tile_fill_vec(double* sm, const MolIn& in, const MolCtx& c, const MolEpi* epi, int X0,
                                                  int Y0, int Z0) {
    constexpr int SXW = MOL_SX / MOL_FW;
    constexpr int NVW = SXW * MOL_SY * MOL_SZ;
    const mol_i64 base = mol_flat<V>(c, X0 - MOL_R0P, (MOL_NDIM >= 2) ? Y0 - MOL_R1 : 1, (MOL_NDIM >= 3) ? Z0 - MOL_R2 : 1);
    const mol_i64 s1 = MOL_EXT(V, 0);
    const mol_i64 s2 = (mol_i64)MOL_EXT(V, 0) * MOL_EXT(V, 1);
#pragma unroll MOL_FILL_UNROLL
    for (int idx = threadIdx.x; idx < NVW; idx += MOL_NTHREADS) {
        const int sx = (idx % SXW) * MOL_FW;
        const int row = idx / SXW;
        const int sy = row % MOL_SY, sz = row / MOL_SY;
        const mol_i64 f = base + sx + (MOL_NDIM >= 2 ? sy * s1 : 0) + (MOL_NDIM >= 3 ? sz * s2 : 0);
        MolFv v = mol_fv_ld(in.a[0] + f);
#if MOL_PRE_AUX
        MolFv p = {epi->cb[0] * v.x, epi->cb[0] * v.y};
        MolFv q = {epi->ce[0] * v.x, epi->ce[0] * v.y};
#endif
#if MOL_NIN > 1
        v.x *= in.c[0];
        v.y *= in.c[0];
#pragma unroll
        for (int j = 1; j < MOL_NIN; ++j) {
            const MolFv w = mol_fv_ld(in.a[j] + f);
            v.x = fma(in.c[j], w.x, v.x);
            v.y = fma(in.c[j], w.y, v.y);
#if MOL_PRE_AUX
            p.x = fma(epi->cb[j], w.x, p.x);
            p.y = fma(epi->cb[j], w.y, p.y);
            q.x = fma(epi->ce[j], w.x, q.x);
            q.y = fma(epi->ce[j], w.y, q.y);
#endif
        }
#endif
        mol_fv_st(sm + (size_t)row * MOL_SX + sx, v);
#if MOL_PRE_AUX
        {   // R0P is even, so a pair of x nodes is either entirely inside the tile's own nodes or entirely halo
            // (the aux tiles' row pitch MOL_TX is even too, so the 128-bit store is aligned)
            const int ax = sx - MOL_R0P, ay = sy - ((MOL_NDIM >= 2) ? MOL_R1 : 0), az = sz - ((MOL_NDIM >= 3) ? MOL_R2 : 0);
            if (ax >= 0 && ax < MOL_TX && ay >= 0 && ay < MOL_TY && az >= 0 && az < ((MOL_NDIM >= 3) ? MOL_TZ : 1)) {
                double* tile0 = sm - V * MOL_TILE_STRIDE;
                mol_fv_st(tile0 + MOL_AUX_P(V) + (az * MOL_TY + ay) * MOL_TX + ax, p);
                mol_fv_st(tile0 + MOL_AUX_Q(V) + (az * MOL_TY + ay) * MOL_TX + ax, q);
            }
        }
#endif
    }
}

template <int V>
struct MolFillVarsVec {
    static __device__ __forceinline__ void run(double* sm, const MolIn& in, const MolCtx& c, const MolEpi* epi, int X0,
                                               int Y0, int Z0) {
        mol_tile_fill_vec<V>(sm + V * MOL_TILE_STRIDE, in, c, epi, X0, Y0, Z0);
        MolFillVarsVec<V + 1>::run(sm, in, c, epi, X0, Y0, Z0);
    }
};
template <>
struct MolFillVarsVec<MOL_NVAR> {
    static __device__ __forceinline__ void run(double*, const MolIn&, const MolCtx&, const MolEpi*, int, int, int) {}
};

template <int V, bool ALL>
struct MolFillVars {
    static __device__ __forceinline__ void run(double* sm, const MolIn& in, const MolCtx& c, const MolEpi* epi, int X0,
                                               int Y0, int Z0) {
        mol_tile_fill<V, ALL>(sm + V * MOL_TILE_STRIDE, in, c, epi, X0, Y0, Z0);
        MolFillVars<V + 1, ALL>::run(sm, in, c, epi, X0, Y0, Z0);
    }
};
template <bool ALL>
struct MolFillVars<MOL_NVAR, ALL> {
    static __device__ __forceinline__ void run(double*, const MolIn&, const MolCtx&, const MolEpi*, int, int, int) {}
};

// VX consecutive values at flat index f (128-bit access when the layout guarantees alignment); hi0 = last x node stored
__device__ __forceinline__ void mol_tile_store(double* __restrict__ arr, mol_i64 f, const double* v, int i0, int hi0) {
#if MOL_VEC_ST && MOL_VX == 2
    if (i0 + 1 <= hi0) *reinterpret_cast<double2*>(arr + f) = make_double2(v[0], v[1]);
    else arr[f] = v[0];
#else
#pragma unroll
    for (int vx = 0; vx < MOL_VX; ++vx)
        if (i0 + vx <= hi0) arr[f + vx] = v[vx];
#endif
}
__device__ __forceinline__ void mol_tile_load(const double* __restrict__ arr, mol_i64 f, double* v, int i0, int hi0) {
#if MOL_VEC_ST && MOL_VX == 2
    if (i0 + 1 <= hi0) {
        const double2 w = __ldg(reinterpret_cast<const double2*>(arr + f));
        v[0] = w.x;
        v[1] = w.y;
    } else {
        v[0] = __ldg(arr + f);
        v[1] = 0.0;
    }
#else
#pragma unroll
    for (int vx = 0; vx < MOL_VX; ++vx) v[vx] = (i0 + vx <= hi0) ? __ldg(arr + f + vx) : 0.0;
#endif
}

// evaluate + store every equation at VX consecutive x nodes starting at node (i0,i1,i2).
// `ok` only guards the stores: the arithmetic is straight-line so that the fully unrolled row loop
// lets the compiler keep the up/centre/down values of consecutive rows in registers.
template <int V>
struct MolTileVars {
    static __device__ __forceinline__ void run(const double* __restrict__ sm, const double* __restrict__ wsm, const MolIn& in,
                                               const MolCtx& c, int lx, int ly, int lz, int i0, int i1, int i2, bool ok, int hi0,
                                               const double* xc, double yc, double zc,
                                               double* __restrict__ out, const MolEpi* epi, double& errsum) {
        double du[MOL_VX];
#if MOL_KERNEL_JVP
        const double* const smv = sm + MOL_JVP_VBASE;       // the direction's tiles (cooperative flavour: sm is the CTA's base)
#pragma unroll
        for (int vx = 0; vx < MOL_VX; ++vx)
            du[vx] = mol_eq_tile_d<V>(sm, smv, wsm, c, lx + vx, ly, lz, i0 + vx, i1, i2, xc[vx], yc, zc).d;
#else
#pragma unroll
        for (int vx = 0; vx < MOL_VX; ++vx)
            du[vx] = mol_eq_tile<V>(sm, wsm, c, lx + vx, ly, lz, i0 + vx, i1, i2, xc[vx], yc, zc);
#endif
        const mol_i64 f = mol_flat<V>(c, i0, i1, i2);
        if (ok) {
#if MOL_EPI
            // the thread's own cell(s) of the tile: the combined input at this node (u+ of the step in FIN mode)
            const int cidx = MOL_CELL(V, lx, ly, lz, 0);
#endif
#if MOL_EPI_PRE
            double up[MOL_VX], eq[MOL_VX];
#pragma unroll
            for (int vx = 0; vx < MOL_VX; ++vx) {
#if MOL_PRE_AUX
                const double p = sm[MOL_AUX_P(V) + (lz * MOL_TY + ly) * MOL_TX + lx + vx];
                const double q = sm[MOL_AUX_Q(V) + (lz * MOL_TY + ly) * MOL_TX + lx + vx];
#else
                double v_, p = 0.0, q = 0.0;
                if (i0 + vx <= hi0) mol_load3(in, *epi, f + vx, v_, p, q);
#endif
                up[vx] = fma(epi->cbk, du[vx], p);
                eq[vx] = fma(epi->cek, du[vx], q);
            }
            mol_tile_store(epi->comb, f, up, i0, hi0);
            mol_tile_store(epi->eout, f, eq, i0, hi0);
#else
            mol_tile_store(out, f, du, i0, hi0);
#endif
#if MOL_EPI_FIN
            double ef[MOL_VX], uf[MOL_VX];
            mol_tile_load(epi->e, f, ef, i0, hi0);
            mol_tile_load(epi->u0, f, uf, i0, hi0);
#pragma unroll
            for (int vx = 0; vx < MOL_VX; ++vx)
                if (i0 + vx <= hi0) mol_fin_point(*epi, ef[vx], uf[vx], du[vx], sm[cidx + vx], errsum);
#endif
        }
        MolTileVars<V + 1>::run(sm, wsm, in, c, lx, ly, lz, i0, i1, i2, ok, hi0, xc, yc, zc, out, epi, errsum);
    }
};
template <>
struct MolTileVars<MOL_NVAR> {
    static __device__ __forceinline__ void run(const double*, const double*, const MolIn&, const MolCtx&, int, int, int, int, int,
                                               int, bool, int, const double*, double, double, double*, const MolEpi*, double&) {}
};

// node coordinate along dimension J (clamped: overhanging tile cells are computed but never stored)
template <int J>
__device__ __forceinline__ double mol_tile_coord(const MolCtx& c, int node) {
    const int n = (J == 0) ? MOL_N0 : (J == 1 ? MOL_N1 : MOL_N2);
    return __ldg(c.grid[J] + ((node < n) ? node : n) - 1);
}

#if MOL_TMA
// tensor-map coordinates are relative to the first stored node of each dimension: the interior
// lower bound, or the slab's first plane along the split (last) dimension
__device__ __forceinline__ void mol_tma_issue(double* smem, int stage, const MolTileMaps& maps, mol_u64* bar,
                                              const MolCtx& c, int X0, int Y0, int Z0) {
#pragma unroll
    for (int v = 0; v < MOL_NVAR; ++v) {
        const int o0 = mol_ilo_[v][0];
        const int o1 = (MOL_NDIM == 2) ? MOL_LLO(v, c) : mol_ilo_[v][1];
        const int o2 = (MOL_NDIM == 3) ? MOL_LLO(v, c) : mol_ilo_[v][2];
        mol_tma_load(smem + ((size_t)stage * MOL_NVAR + v) * MOL_TILE_STRIDE, &maps.m[v], bar,
                     X0 - MOL_R0P - o0, Y0 - MOL_R1 - o1, Z0 - MOL_R2 - o2);
    }
}
#endif

// ---- cp.async staging (MOL_CPASYNC): the multi-stage pipeline of the TMA path for layouts the TMA unit cannot address
// (row pitch not a multiple of 16 B: the n - 2 unknowns per row of a Dirichlet/Neumann problem on n = 2^k + 1 nodes;
// 1-D programs).  Every thread copies its cells of the NEXT tiles global -> shared asynchronously (no registers
// held), 8 B or 16 B per copy as alignment allows, one commit group per tile; cells that are not stored state are
// skipped here and patched after arrival, exactly as after a TMA load.
#ifndef MOL_CPASYNC
#define MOL_CPASYNC 0
#endif
#if MOL_CPASYNC
#ifdef MOL_HOST_EMU
__device__ __forceinline__ void mol_cp_async(double* dst, const double* src) {       // tests/cuda_emu: synchronous copy
    dst[0] = src[0];
    if (MOL_FW == 2) dst[1] = src[1];
}
__device__ __forceinline__ void mol_cp_commit() {}
template <int N>
__device__ __forceinline__ void mol_cp_wait() {}
#else
__device__ __forceinline__ void mol_cp_async(double* dst, const double* src) {
#if MOL_FW == 2
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(mol_smem_u32(dst)), "l"(src) : "memory");
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(mol_smem_u32(dst)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void mol_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void mol_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

template <int V>
__device__ __forceinline__ void mol_tile_issue_cp(double* sm, const MolIn& in, const MolCtx& c, int X0, int Y0, int Z0,
                                                  bool all_inside) {
    constexpr int SXW = MOL_SX / MOL_FW;
    constexpr int NVW = SXW * MOL_SY * MOL_SZ;
    const mol_i64 base = mol_flat<V>(c, X0 - MOL_R0P, (MOL_NDIM >= 2) ? Y0 - MOL_R1 : 1, (MOL_NDIM >= 3) ? Z0 - MOL_R2 : 1);
    const mol_i64 s1 = MOL_EXT(V, 0);
    const mol_i64 s2 = (mol_i64)MOL_EXT(V, 0) * MOL_EXT(V, 1);
    for (int idx = threadIdx.x; idx < NVW; idx += MOL_NTHREADS) {
        const int sx = (idx % SXW) * MOL_FW;
        const int row = idx / SXW;
        const int sy = row % MOL_SY, sz = row / MOL_SY;
        bool ok = all_inside;
        if (!all_inside) {      // both nodes of a 16 B pair are stored state (pairs never straddle: even bounds when MOL_FW == 2)
            const int n0 = X0 - MOL_R0P + sx;
            const int n1 = (MOL_NDIM >= 2) ? Y0 - MOL_R1 + sy : 1;
            const int n2 = (MOL_NDIM >= 3) ? Z0 - MOL_R2 + sz : 1;
            ok = (n0 >= MOL_ILO(V, 0)) && (n0 + MOL_FW - 1 <= MOL_IHI(V, 0));
            if (MOL_NDIM >= 2) ok = ok && (n1 >= MOL_ILO(V, 1)) && (n1 <= MOL_IHI(V, 1));
            if (MOL_NDIM >= 3) ok = ok && (n2 >= MOL_ILO(V, 2)) && (n2 <= MOL_IHI(V, 2));
#if MOL_DIST
            {
                const int nl = (MOL_NDIM == 2) ? n1 : n2;
                ok = ok && (nl >= c.loc_lo) && (nl <= c.loc_hi);
            }
#endif
        }
        if (ok) mol_cp_async(sm + (size_t)row * MOL_SX + sx,
                             in.a[0] + base + sx + (MOL_NDIM >= 2 ? sy * s1 : 0) + (MOL_NDIM >= 3 ? sz * s2 : 0));
    }
}
template <int V>
struct MolIssueVars {
    static __device__ __forceinline__ void run(double* sm, const MolIn& in, const MolCtx& c, int X0, int Y0, int Z0, bool all_inside) {
        mol_tile_issue_cp<V>(sm + V * MOL_TILE_STRIDE, in, c, X0, Y0, Z0, all_inside);
        MolIssueVars<V + 1>::run(sm, in, c, X0, Y0, Z0, all_inside);
    }
};
template <>
struct MolIssueVars<MOL_NVAR> {
    static __device__ __forceinline__ void run(double*, const MolIn&, const MolCtx&, int, int, int, bool) {}
};
#endif

// ---- staged per-node records of the non-uniform axes (MOL_WSTAGE, see mol_device.cuh) ----------------------------------
// One block of MOL_WSM_STRIDE doubles per pipeline stage behind all tiles: the records of the tile's columns (x) and rows
// (y), with their halo.  ASYNC: 16-byte cp.async copies that join the tile's commit group; otherwise plain loads (the
// cooperative flavour, whose barrier after the fill publishes them).  Source and destination are 16-byte aligned (even
// record strides, even offsets).
#define MOL_NSTAGE_EFF ((MOL_TMA || MOL_CPASYNC) ? MOL_STAGES : 1)
#define MOL_WSM_BASE (MOL_PRE_AUX ? MOL_NVAR * (MOL_TILE_STRIDE + 2 * MOL_AUX_STRIDE) : MOL_NSTAGE_EFF * MOL_NVAR * MOL_TILE_STRIDE)
#if MOL_WSTAGE
// x: one segment of MOL_WN0 nodes per field (8-byte copies: a tile may start at an odd node);  y: the tile's row
// records are one contiguous block (16-byte copies: even record stride, even offsets)
template <bool ASYNC, int BYTES>
__device__ __forceinline__ void mol_wrec_copy(double* dst, const double* src, int ndoubles) {
    constexpr int W = BYTES / 8;
    for (int k = W * (int)threadIdx.x; k < ndoubles; k += W * MOL_NTHREADS) {
#if MOL_CPASYNC && !defined(MOL_HOST_EMU)
        if (ASYNC) {
            if (BYTES == 16) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(mol_smem_u32(dst + k)), "l"(src + k) : "memory");
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(mol_smem_u32(dst + k)), "l"(src + k) : "memory");
            continue;
        }
#endif
        dst[k] = __ldg(src + k);
        if (W == 2) dst[k + 1] = __ldg(src + k + 1);
    }
}
template <bool ASYNC>
__device__ __forceinline__ void mol_wrec_issue(double* wsm, const MolCtx& c, int X0, int Y0) {
    if (MOL_WRS0 > 0) {
        const double* src = c.tabw + MOL_WOFF0 + (X0 - MOL_WHL0 - MOL_WLO0);
#pragma unroll
        for (int f = 0; f < MOL_WRS0; ++f) mol_wrec_copy<ASYNC, 8>(wsm + f * MOL_WN0, src + (mol_i64)f * MOL_WNREC0, MOL_WN0);
    }
    if (MOL_NDIM >= 2 && MOL_WRS1 > 0)
        mol_wrec_copy<ASYNC, 16>(wsm + MOL_WRS0 * MOL_WN0, c.tabw + MOL_WOFF1 + (mol_i64)(Y0 - MOL_WHL1 - MOL_WLO1) * MOL_WRS1,
                                 MOL_WRS1 * MOL_WN1);
}
#endif

#if !MOL_ZMARCH
extern "C" __global__ void __launch_bounds__(MOL_NTHREADS, MOL_MIN_CTAS)
mol_rhs_tiled(MolIn in, MolCtx c, MolTiles T, double* __restrict__ out
#if MOL_TMA
              , const __grid_constant__ MolTileMaps maps
#endif
#if MOL_EPI
              , MolEpi epi
#endif
#if MOL_KERNEL_JVP
              , MolJv jv          // tiled J*v (cooperative flavour, one input, no epilogue): out = (df/du)(u) v
#endif
) {
#if MOL_DEVDT
#if MOL_EPI
    if (!mol_devdt_apply(in, c, &epi)) return;
#else
    if (!mol_devdt_apply(in, c, nullptr)) return;
#endif
#endif
    extern __shared__ __align__(128) unsigned char mol_smem_raw[];
    double* smem = reinterpret_cast<double*>(mol_smem_raw);
    const int tid = threadIdx.x;
    const int tx = tid % MOL_NTXT, ty = tid / MOL_NTXT;
#if MOL_EPI
    double errsum = 0.0;
    const MolEpi* const epip = &epi;
#else
    const MolEpi* const epip = nullptr;
#endif

    // ---- dynamic tile queue ------------------------------------------------------------------------
    // The first tile of a CTA is its block index; every further one is drawn from a global ticket
    // counter, so a CTA that starts late (e.g. behind a communication kernel holding its SM) simply
    // takes fewer tiles instead of stretching the sweep.  Each CTA stops at its first out-of-range
    // ticket, so exactly `ntiles` tickets are drawn per launch and the last draw re-arms the counter.
    __shared__ int tile_q[MOL_STAGES];
    bool drained = false;                           // thread 0 only
#if MOL_EPI_FIN
    // FIN epilogue: STATIC assignment, tiles b, b + G, b + 2G, ... for CTA b of G.  Each thread then adds its share of the
    // scaled error norm in a fixed order, the CTA's partial sum goes to its own slot (no atomics), and the norm -- hence
    // the step-size sequence of an adaptive solve -- is reproducible bit for bit from run to run.
    int static_tile = (int)blockIdx.x;
    auto next_ticket = [&]() -> int {
        static_tile += (int)gridDim.x;
        return static_tile < T.ntiles ? static_tile : T.ntiles;
    };
#else
    auto next_ticket = [&]() -> int {
        if (drained) return T.ntiles;
        const int raw = atomicAdd(T.counter, 1);
        if (raw == T.ntiles - 1) *T.counter = 0;    // the very last draw of this launch
        const int tk = raw + (int)gridDim.x;
        if (tk >= T.ntiles) drained = true;
        return tk < T.ntiles ? tk : T.ntiles;
    };
#endif
    (void)drained;
#if MOL_TMA
    __shared__ __align__(8) mol_u64 full_bar[MOL_STAGES];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < MOL_STAGES; ++s) mol_mbar_init(&full_bar[s], 1);
        mol_fence_mbar_init();
        // prologue: fill all but one stage
        for (int s = 0; s < MOL_STAGES - 1; ++s) {
            const int tq = (s == 0) ? (int)blockIdx.x : next_ticket();
            tile_q[s] = tq;
            if (tq < T.ntiles) {
                int X0, Y0, Z0;
                mol_tile_origin(T, tq, X0, Y0, Z0);
                mol_mbar_expect_tx(&full_bar[s], MOL_NVAR * MOL_TILE_BYTES);
                mol_tma_issue(smem, s, maps, &full_bar[s], c, X0, Y0, Z0);
            }
        }
    }
    __syncthreads();
#elif MOL_CPASYNC
    // tickets are drawn one iteration before their tile is issued (every thread issues, so the index must be published)
    if (tid == 0) {
        tile_q[0] = blockIdx.x;
        for (int s = 1; s < MOL_STAGES; ++s) tile_q[s] = next_ticket();
    }
    __syncthreads();
    auto issue_cp = [&](int s) {
        const int tq = tile_q[s];
        if (tq < T.ntiles) {
            int X1, Y1, Z1;
            mol_tile_origin(T, tq, X1, Y1, Z1);
            MolIssueVars<0>::run(smem + (size_t)s * MOL_NVAR * MOL_TILE_STRIDE, in, c, X1, Y1, Z1,
                                 mol_tile_fully_inside(c, X1, Y1, Z1));
#if MOL_WSTAGE
            mol_wrec_issue<true>(smem + MOL_WSM_BASE + (size_t)s * MOL_WSM_STRIDE, c, X1, Y1);
#endif
        }
        mol_cp_commit();            // one group per stage, empty or not, so that the group count stays uniform
    };
    for (int s = 0; s < MOL_STAGES - 1; ++s) issue_cp(s);
#else
    if (tid == 0) tile_q[0] = blockIdx.x;
    __syncthreads();
#endif

    for (int it = 0;; ++it) {
#if MOL_TMA || MOL_CPASYNC
        const int stage = it % MOL_STAGES;
#else
        const int stage = 0;
#endif
        const int tile = tile_q[stage];
        if (tile >= T.ntiles) break;
        int X0, Y0, Z0, H0, H1, H2;
        mol_tile_origin(T, tile, X0, Y0, Z0, H0, H1, H2);
#if MOL_TMA
        double* sm = smem + (size_t)stage * MOL_NVAR * MOL_TILE_STRIDE;
        if (tid == 0) {   // keep STAGES-1 tiles in flight: refill the stage that iteration it-1 released
            const int s1 = (it + MOL_STAGES - 1) % MOL_STAGES;
            const int tq = next_ticket();
            tile_q[s1] = tq;
            if (tq < T.ntiles) {
                int X1, Y1, Z1;
                mol_tile_origin(T, tq, X1, Y1, Z1);
                mol_mbar_expect_tx(&full_bar[s1], MOL_NVAR * MOL_TILE_BYTES);
                mol_tma_issue(smem, s1, maps, &full_bar[s1], c, X1, Y1, Z1);
            }
        }
        mol_mbar_wait(&full_bar[stage], (it / MOL_STAGES) & 1);
        if (mol_tile_touches_edge(c, X0, Y0, Z0)) {       // CTA-uniform
            mol_wait_ghost_planes(T, c, Y0, Z0);
            MolFillVars<0, false>::run(sm, in, c, epip, X0, Y0, Z0);
            mol_fence_proxy_async();
            __syncthreads();
        }
#elif MOL_CPASYNC
        double* sm = smem + (size_t)stage * MOL_NVAR * MOL_TILE_STRIDE;
        issue_cp((it + MOL_STAGES - 1) % MOL_STAGES);     // refill the stage that iteration it-1 released
        mol_cp_wait<MOL_STAGES - 1>();                    // this thread's copies of the current tile have landed
        __syncthreads();                                  // ... and everybody else's
        if (mol_tile_touches_edge(c, X0, Y0, Z0)) {       // CTA-uniform
            mol_wait_ghost_planes(T, c, Y0, Z0);
            MolFillVars<0, false>::run(sm, in, c, epip, X0, Y0, Z0);
            __syncthreads();
        }
        if (tid == 0) tile_q[stage] = next_ticket();      // issued at the top of the next iteration (published below)
#else
        double* sm = smem;
#if MOL_KERNEL_JVP
        if (mol_tile_fully_inside(c, X0, Y0, Z0)) {                                                         // CTA-uniform
            MolIn inv = in;
            inv.a[0] = jv.v;
            MolFillVarsVec<0>::run(sm, in, c, epip, X0, Y0, Z0);
            MolFillVarsVec<0>::run(smem + MOL_JVP_VBASE, inv, c, epip, X0, Y0, Z0);
        } else {
            MolFillVarsD<0>::run(sm, smem + MOL_JVP_VBASE, in, jv, c, X0, Y0, Z0);
        }
#else
        if (mol_tile_fully_inside(c, X0, Y0, Z0)) MolFillVarsVec<0>::run(sm, in, c, epip, X0, Y0, Z0);      // CTA-uniform
        else { mol_wait_ghost_planes(T, c, Y0, Z0); MolFillVars<0, true>::run(sm, in, c, epip, X0, Y0, Z0); }
#endif
#if MOL_WSTAGE
        mol_wrec_issue<false>(smem + MOL_WSM_BASE, c, X0, Y0);
#endif
        __syncthreads();
        if (tid == 0) tile_q[0] = next_ticket();       // read by everyone after the barrier below
#endif

#if MOL_WSTAGE
        const double* const wsm = smem + MOL_WSM_BASE + (size_t)stage * MOL_WSM_STRIDE;
#else
        const double* const wsm = nullptr;
#endif
        // ---- pointwise evaluation: VX consecutive x nodes x PY consecutive rows per thread ---------
        // No branches around the arithmetic (only the stores are predicated): rows march in registers.
#pragma unroll
        for (int kx = 0; kx < MOL_PX; ++kx) {
            const int lx = (kx * MOL_NTXT + tx) * MOL_VX;
            const int n0 = X0 + lx;
            double xc[MOL_VX];
#pragma unroll
            for (int vx = 0; vx < MOL_VX; ++vx) xc[vx] = MOL_USE_X0 ? mol_tile_coord<0>(c, n0 + vx) : 0.0;
#pragma unroll
            for (int kz = 0; kz < ((MOL_NDIM >= 3) ? MOL_TZ : 1); ++kz) {
                const int n2 = Z0 + kz;
                const double zc = (MOL_NDIM >= 3 && MOL_USE_X2) ? mol_tile_coord<2>(c, n2) : 0.0;
#pragma unroll
                for (int ky = 0; ky < ((MOL_NDIM >= 2) ? MOL_PY : 1); ++ky) {
                    const int ly = (MOL_NDIM >= 2) ? ty * MOL_PY + ky : 0;
                    const int n1 = Y0 + ly;
                    const double yc = (MOL_NDIM >= 2 && MOL_USE_X1) ? mol_tile_coord<1>(c, n1) : 0.0;
                    bool ok = (n0 <= H0);
                    if (MOL_NDIM >= 2) ok = ok && (n1 <= H1);
                    if (MOL_NDIM >= 3) ok = ok && (n2 <= H2);
#if MOL_EPI
                    MolTileVars<0>::run(sm, wsm, in, c, lx, ly, kz, n0, n1, n2, ok, H0, xc, yc, zc, out, &epi, errsum);
#else
                    double dummy = 0.0;
                    MolTileVars<0>::run(sm, wsm, in, c, lx, ly, kz, n0, n1, n2, ok, H0, xc, yc, zc, out, nullptr, dummy);
#endif
                }
            }
        }
        __syncthreads();          // everyone is done with this stage (and tile_q is visible) before the refill
    }

#if MOL_EPI_FIN
    __shared__ double red[MOL_NTHREADS / 32];
    errsum = mol_warp_sum(errsum);
    if ((tid & 31) == 0) red[tid >> 5] = errsum;
    __syncthreads();
    if (tid < 32) {
        double v = (tid < MOL_NTHREADS / 32) ? red[tid] : 0.0;
        v = mol_warp_sum(v);
        if (tid == 0 && epi.err) epi.err[blockIdx.x] = v;      // this CTA's slot (the host offsets epi.err per launch)
    }
#endif
}
#else   // MOL_ZMARCH
// ---- 3-D: xy tiles marching along z --------------------------------------------------------------------------------
// A work item is an xy tile (MOL_TX x MOL_TY nodes) times a chunk of up to MOL_TZ consecutive z planes.  The CTA
// walks the chunk plane by plane; shared memory holds a ring of MOL_RING xy planes (+ halo) per variable, so every
// plane of the state is fetched once per chunk (plus 2*R2 planes of overlap between chunks) instead of once per
// (TZ + 2*R2)/TZ-thick brick.  With TMA, one bulk-tensor load per plane and variable runs MOL_RING - 2*R2 - 1 planes
// ahead of the arithmetic (one mbarrier per ring slot); the fused-stage variants (MOL_NIN > 1) fill one plane per
// step with the cooperative 128-bit loader.

// does plane z of an xy tile need patching (cells the TMA unit could not supply / ghost cells)?
__device__ __forceinline__ bool mol_plane_touches_edge(const MolCtx& c, int X0, int Y0, int z) {
    bool e = (X0 - MOL_R0 < MOL_ILO_MAX0) || (X0 + MOL_TX - 1 + MOL_R0 > MOL_IHI_MIN0);
    e = e || (Y0 - MOL_R1 < MOL_ILO_MAX1) || (Y0 + MOL_TY - 1 + MOL_R1 > MOL_IHI_MIN1);
    e = e || (z < MOL_ILO_MAX2) || (z > MOL_IHI_MIN2);
#if MOL_DIST
    e = e || (z < c.loc_lo) || (z > c.loc_hi);
#endif
    return e;
}
__device__ __forceinline__ bool mol_plane_fully_inside(const MolCtx& c, int X0, int Y0, int z) {
    bool in_ = (X0 - MOL_R0P >= MOL_ILO_MAX0) && (X0 + MOL_TX - 1 + MOL_R0P <= MOL_IHI_MIN0);
    in_ = in_ && (Y0 - MOL_R1 >= MOL_ILO_MAX1) && (Y0 + MOL_TY - 1 + MOL_R1 <= MOL_IHI_MIN1);
    in_ = in_ && (z >= MOL_ILO_MAX2) && (z <= MOL_IHI_MIN2);
#if MOL_DIST
    in_ = in_ && (z >= c.loc_lo) && (z <= c.loc_hi);
#endif
    return in_;
}

// Patch list of an xy tile that touches the domain edge in x or y: the cells of a plane that are not stored state
// but are read by some stencil (periodic images, ghost nodes).  The list is the same for every z-interior plane of
// the work item; each thread owns up to MOL_PATCH_PER_THREAD entries and fetches their values for the NEXT plane
// into registers while the current plane is evaluated, so the wrap/ghost loads are off the critical path.
#define MOL_PATCH_PER_THREAD 2
#define MOL_MAXPATCH (MOL_PATCH_PER_THREAD * MOL_NTHREADS)
__device__ __forceinline__ bool mol_z_outside(const MolCtx& c, int z) {
    bool e = (z < MOL_ILO_MAX2) || (z > MOL_IHI_MIN2);
#if MOL_DIST
    e = e || (z < c.loc_lo) || (z > c.loc_hi);
#endif
    return e;
}
template <int V>
struct MolPatch {
    static __device__ __forceinline__ void load(double (*pv)[MOL_PATCH_PER_THREAD], const MolIn& in, const MolCtx& c,
                                                const unsigned short* list, int npatch, int X0, int Y0, int z) {
#pragma unroll
        for (int q = 0; q < MOL_PATCH_PER_THREAD; ++q) {
            const int k = threadIdx.x + q * MOL_NTHREADS;
            if (k < npatch) {
                const int cell = list[k];
                pv[V][q] = mol_node<V>(in, c, X0 - MOL_R0P + cell % MOL_SX, Y0 - MOL_R1 + cell / MOL_SX, z);
            }
        }
        MolPatch<V + 1>::load(pv, in, c, list, npatch, X0, Y0, z);
    }
    static __device__ __forceinline__ void store(const double (*pv)[MOL_PATCH_PER_THREAD], double* slot,
                                                 const unsigned short* list, int npatch) {
#pragma unroll
        for (int q = 0; q < MOL_PATCH_PER_THREAD; ++q) {
            const int k = threadIdx.x + q * MOL_NTHREADS;
            if (k < npatch) slot[V * MOL_TILE_STRIDE + list[k]] = pv[V][q];
        }
        MolPatch<V + 1>::store(pv, slot, list, npatch);
    }
};
template <>
struct MolPatch<MOL_NVAR> {
    static __device__ __forceinline__ void load(double (*)[MOL_PATCH_PER_THREAD], const MolIn&, const MolCtx&,
                                                const unsigned short*, int, int, int, int) {}
    static __device__ __forceinline__ void store(const double (*)[MOL_PATCH_PER_THREAD], double*, const unsigned short*, int) {}
};

extern "C" __global__ void __launch_bounds__(MOL_NTHREADS, MOL_MIN_CTAS)
mol_rhs_tiled(MolIn in, MolCtx c, MolTiles T, double* __restrict__ out
#if MOL_TMA
              , const __grid_constant__ MolTileMaps maps
#endif
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
    extern __shared__ __align__(128) unsigned char mol_smem_raw[];
    double* sm = reinterpret_cast<double*>(mol_smem_raw);
    const int tid = threadIdx.x;
    const int tx = tid % MOL_NTXT, ty = tid / MOL_NTXT;
#if MOL_EPI
    double errsum = 0.0;
    const MolEpi* const epip = &epi;
#else
    const MolEpi* const epip = nullptr;
#endif
    __shared__ int tile_q[2];
    bool drained = false;                           // thread 0 only
#if MOL_EPI_FIN
    int static_tile = (int)blockIdx.x;              // static assignment: reproducible error norm (see the 2-D kernel)
    auto next_ticket = [&]() -> int {
        static_tile += (int)gridDim.x;
        return static_tile < T.ntiles ? static_tile : T.ntiles;
    };
#else
    auto next_ticket = [&]() -> int {
        if (drained) return T.ntiles;
        const int raw = atomicAdd(T.counter, 1);
        if (raw == T.ntiles - 1) *T.counter = 0;    // the very last draw of this launch re-arms the counter
        const int tk = raw + (int)gridDim.x;
        if (tk >= T.ntiles) drained = true;
        return tk < T.ntiles ? tk : T.ntiles;
    };
#endif
    (void)drained;
#if MOL_TMA
    __shared__ __align__(8) mol_u64 full_bar[MOL_RING];
    __shared__ unsigned short patch_list[MOL_MAXPATCH];
    __shared__ int patch_count;
    unsigned phase = 0;                             // bit s: parity of the next completion of ring slot s
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < MOL_RING; ++s) mol_mbar_init(&full_bar[s], 1);
        mol_fence_mbar_init();
    }
#endif
    if (tid == 0) tile_q[0] = blockIdx.x;
    __syncthreads();

    for (int it = 0;; ++it) {
        const int tile = tile_q[it & 1];
        if (tile >= T.ntiles) break;
        int X0, Y0, Z0, H0, H1, H2;
        mol_tile_origin(T, tile, X0, Y0, Z0, H0, H1, H2);
        const int nz = min(MOL_TZ, H2 - Z0 + 1);            // planes evaluated by this work item
        const int np = nz + 2 * MOL_R2;                     // planes read: r = 0..np-1 <-> z = Z0 - R2 + r
#if MOL_TMA
        auto issue = [&](int r) {                           // thread 0: bulk-tensor load of plane r into its ring slot
            const int s = r % MOL_RING;
            mol_mbar_expect_tx(&full_bar[s], MOL_NVAR * MOL_TILE_BYTES);
#pragma unroll
            for (int v = 0; v < MOL_NVAR; ++v)
                mol_tma_load(sm + ((size_t)s * MOL_NVAR + v) * MOL_TILE_STRIDE, &maps.m[v], &full_bar[s],
                             X0 - MOL_R0P - mol_ilo_[v][0], Y0 - MOL_R1 - mol_ilo_[v][1], Z0 - MOL_R2 + r - MOL_LLO(v, c));
        };
        auto prefetch = [&](int r) {                        // thread 0: pull plane r into L2 ahead of its load
#pragma unroll
            for (int v = 0; v < MOL_NVAR; ++v)
                mol_tma_prefetch_l2(&maps.m[v], X0 - MOL_R0P - mol_ilo_[v][0], Y0 - MOL_R1 - mol_ilo_[v][1],
                                    Z0 - MOL_R2 + r - MOL_LLO(v, c));
        };
        if (tid == 0) {
            tile_q[(it + 1) & 1] = next_ticket();
            for (int r = 0; r < min(np, MOL_RING); ++r) issue(r);
            for (int r = MOL_RING; r < min(np, MOL_RING + MOL_L2_AHEAD); ++r) prefetch(r);
            patch_count = 0;
        }
        const bool xy_edge = (X0 - MOL_R0 < MOL_ILO_MAX0) || (X0 + MOL_TX - 1 + MOL_R0 > MOL_IHI_MIN0) ||
                             (Y0 - MOL_R1 < MOL_ILO_MAX1) || (Y0 + MOL_TY - 1 + MOL_R1 > MOL_IHI_MIN1);     // CTA-uniform
        int npatch = 0;
        if (xy_edge) {      // (all variables share the interior box: a requirement of the tiled path)
            __syncthreads();
            for (int cell = tid; cell < MOL_TILE_CELLS; cell += MOL_NTHREADS) {
                const int n0 = X0 - MOL_R0P + cell % MOL_SX, n1 = Y0 - MOL_R1 + cell / MOL_SX;
                const bool inside = (n0 >= MOL_ILO(0, 0)) && (n0 <= MOL_IHI(0, 0)) && (n1 >= MOL_ILO(0, 1)) && (n1 <= MOL_IHI(0, 1));
                const bool near_ = (n0 >= MOL_ILO(0, 0) - MOL_R0) && (n0 <= MOL_IHI(0, 0) + MOL_R0) &&
                                   (n1 >= MOL_ILO(0, 1) - MOL_R1) && (n1 <= MOL_IHI(0, 1) + MOL_R1);
                if (!inside && near_) {
                    const int k = atomicAdd(&patch_count, 1);
                    if (k < MOL_MAXPATCH) patch_list[k] = (unsigned short)cell;
                }
            }
            __syncthreads();
            npatch = patch_count;
        }
        const bool fast_patch = xy_edge && npatch <= MOL_MAXPATCH;
        double pv[MOL_NVAR][MOL_PATCH_PER_THREAD];
        int pv_plane = -1;                                  // plane whose patch values the registers hold
#else
        if (tid == 0) tile_q[(it + 1) & 1] = next_ticket();
#endif
        // slab mode, fused ghost-plane wait: a chunk that reads planes of a neighbouring rank waits for their arrival here,
        // with its own first planes already in flight
        mol_wait_ghost_planes(T, c, Y0, Z0);
        int arrived = 0;                                    // planes r < arrived are complete in shared memory
        // per-thread x coordinates (hoisted)
        double xcs[MOL_PX][MOL_VX];
#pragma unroll
        for (int kx = 0; kx < MOL_PX; ++kx)
#pragma unroll
            for (int vx = 0; vx < MOL_VX; ++vx)
                xcs[kx][vx] = MOL_USE_X0 ? mol_tile_coord<0>(c, X0 + (kx * MOL_NTXT + tx) * MOL_VX + vx) : 0.0;

        for (int i = 0; i < nz; ++i) {
            // ---- make planes r <= i + 2*R2 available
#if MOL_TMA
            if (tid == 0 && i >= 1) {
                if (i + MOL_RING - 1 < np) issue(i + MOL_RING - 1);                 // refill the slot step i-1 released
                if (MOL_L2_AHEAD > 0 && i + MOL_RING - 1 + MOL_L2_AHEAD < np) prefetch(i + MOL_RING - 1 + MOL_L2_AHEAD);
            }
            bool patched = false;
            while (arrived <= i + 2 * MOL_R2) {
                const int s = arrived % MOL_RING;
                mol_mbar_wait(&full_bar[s], (phase >> s) & 1u);
                phase ^= 1u << s;
                const int z = Z0 - MOL_R2 + arrived;
                double* slot = sm + (size_t)s * MOL_NVAR * MOL_TILE_STRIDE;
                if (mol_z_outside(c, z) || (xy_edge && !fast_patch)) {                // CTA-uniform
                    MolFillVars<0, false>::run(slot, in, c, epip, X0, Y0, z + MOL_R2);
                    patched = true;
                } else if (xy_edge) {
                    if (pv_plane != arrived) MolPatch<0>::load(pv, in, c, patch_list, npatch, X0, Y0, z);
                    MolPatch<0>::store(pv, slot, patch_list, npatch);
                    patched = true;
                }
                ++arrived;
            }
            if (patched) {
                mol_fence_proxy_async();
                __syncthreads();
            }
            if (fast_patch && arrived < np && !mol_z_outside(c, Z0 - MOL_R2 + arrived)) {
                // fetch the wrap/ghost values of the next plane to arrive; they are stored at the next step
                MolPatch<0>::load(pv, in, c, patch_list, npatch, X0, Y0, Z0 - MOL_R2 + arrived);
                pv_plane = arrived;
            }
#else
            while (arrived <= i + 2 * MOL_R2) {
                const int s = arrived % MOL_RING;
                const int z = Z0 - MOL_R2 + arrived;
                double* sp = sm + (size_t)s * MOL_NVAR * MOL_TILE_STRIDE;
                if (mol_plane_fully_inside(c, X0, Y0, z)) MolFillVarsVec<0>::run(sp, in, c, epip, X0, Y0, z + MOL_R2);
                else MolFillVars<0, true>::run(sp, in, c, epip, X0, Y0, z + MOL_R2);
                ++arrived;
            }
            __syncthreads();
#endif
            // ---- evaluate plane z = Z0 + i: VX consecutive x nodes x PY consecutive rows per thread
            const int lz = i % MOL_RING;
            const int n2 = Z0 + i;
            const double zc = MOL_USE_X2 ? mol_tile_coord<2>(c, n2) : 0.0;
#pragma unroll
            for (int kx = 0; kx < MOL_PX; ++kx) {
                const int lx = (kx * MOL_NTXT + tx) * MOL_VX;
                const int n0 = X0 + lx;
#pragma unroll
                for (int ky = 0; ky < MOL_PY; ++ky) {
                    const int ly = ty * MOL_PY + ky;
                    const int n1 = Y0 + ly;
                    const double yc = MOL_USE_X1 ? mol_tile_coord<1>(c, n1) : 0.0;
                    const bool ok = (n0 <= H0) && (n1 <= H1);
#if MOL_EPI
                    MolTileVars<0>::run(sm, nullptr, in, c, lx, ly, lz, n0, n1, n2, ok, H0, xcs[kx], yc, zc, out, &epi, errsum);
#else
                    double dummy = 0.0;
                    MolTileVars<0>::run(sm, nullptr, in, c, lx, ly, lz, n0, n1, n2, ok, H0, xcs[kx], yc, zc, out, nullptr, dummy);
#endif
                }
            }
#if MOL_TMA
            __syncthreads();      // plane r = i is released: its slot is refilled at the top of the next step
#endif
        }
#if !MOL_TMA
        __syncthreads();          // the next work item's first fills reuse the ring
#endif
    }

#if MOL_EPI_FIN
    __shared__ double red[MOL_NTHREADS / 32];
    errsum = mol_warp_sum(errsum);
    if ((tid & 31) == 0) red[tid >> 5] = errsum;
    __syncthreads();
    if (tid < 32) {
        double v = (tid < MOL_NTHREADS / 32) ? red[tid] : 0.0;
        v = mol_warp_sum(v);
        if (tid == 0 && epi.err) epi.err[blockIdx.x] = v;      // this CTA's slot (the host offsets epi.err per launch)
    }
#endif
}
#endif  // MOL_ZMARCH
