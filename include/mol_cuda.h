/* libmol_cuda.so — C ABI of the B200-native backend for MethodOfLines.jl's hot path.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *
 *   mol_fd_weights      calculate_weights            src/discretization/schemes/fornberg_calculate_weights.jl:20-67
 *   mol_plan_create     the lowering + codegen that today is
 *                         discretize_equation!        src/scalar_discretization.jl:1-42  (one symbolic eq / point)
 *                         mtkcompile + ODEProblem     src/discretization/staggered_discretize.jl:6-23,
 *                                                     src/MOL_discretization.jl:175-191 (RuntimeGeneratedFunction f!)
 *                       here: a serialized *stencil program*  is compiled to sm_100a kernels
 *   mol_rhs             the generated f!(du,u,p,t)    docs/src/generated/bruss_code.md:46-118 (artifact),
 *                                                     called as benchmark/weno/suite.jl:42-48 does
 *   mol_rk_*            OrdinaryDiffEq.solve(prob, Euler()/SSPRK33()/Tsit5(); abstol, reltol, dt, adaptive, saveat)
 *                                                     call sites test/Diffusion/MOL_1D_Linear_Diffusion.jl:73,
 *                                                     benchmark/weno/suite.jl:50-54 (third-party, restated)
 *   mol_unpack          PDETimeSeriesSolution unpacking  src/interface/solution/timedep.jl:30-72 (reshape + observed
 *                                                     boundary nodes + zero corners)
 *   mol_dist_*          (no reference equivalent: slab decomposition + halo exchange)
 *
 * Conventions: every function returns 0 on success, <0 on error (MOL_E_*); the message is
 * available from mol_last_error().  Device pointers are raw CUDA device addresses owned by the caller
 * (a Julia CuArray, a torch tensor, ...); the library never frees or retains caller memory.  `stream`
 * is a cudaStream_t passed as void* (NULL = default stream).  State layout = the reference's flat
 * unknown vector: variable-major, first spatial index fastest, interior nodes only.
 * Handles are not thread-safe; use one host thread per device, and ONE stream at a time per plan: the persistent
 * kernels of a plan share one tile-ticket counter, so two mol_rhs calls of the same plan must not run concurrently
 * on different streams (create a second plan for that).
 */
#ifndef MOL_CUDA_H
#define MOL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOL_OK              0
#define MOL_E_PARSE        -1   /* malformed stencil program */
#define MOL_E_UNSUPPORTED  -2   /* pattern outside the kernel set (cf. ArrayDiscretizationError) */
#define MOL_E_COMPILE      -3   /* NVRTC failure (log in mol_last_error) */
#define MOL_E_CUDA         -4   /* CUDA runtime/driver error */
#define MOL_E_ARG          -5   /* bad argument */
#define MOL_E_NOCUDA       -6   /* no CUDA driver / device: there is NO CPU fallback */

typedef struct mol_plan mol_plan;
typedef struct mol_rk   mol_rk;

/* algorithms for mol_rk_init */
#define MOL_ALG_EULER    1
#define MOL_ALG_SSPRK33  2
#define MOL_ALG_RK4      3
#define MOL_ALG_TSIT5    4

/* kernel selection for mol_plan_set_option("kernel", ...) */
#define MOL_KERNEL_AUTO    0   /* tiled (TMA) kernel on the uniform core, generic kernel on the frame */
#define MOL_KERNEL_GENERIC 1   /* table-driven generic kernel everywhere */

typedef struct mol_step_stats {
    double t;          /* time after the step (unchanged if rejected) */
    double dt_next;    /* proposed next step */
    double eest;       /* scaled error estimate (adaptive only) */
    int    accepted;   /* 1 accepted, 0 rejected */
    int    nf;         /* RHS evaluations spent */
} mol_step_stats;

typedef struct mol_solve_stats {
    double   t_final;
    double   dt_last;
    int64_t  nf, naccept, nreject;
    int      retcode;  /* 0 = Success, 1 = MaxIters, 2 = DtLessThanMin/NaN */
} mol_solve_stats;

/* -- a1: finite-difference weights (host, no GPU needed) ------------------------------------------- */
int mol_fd_weights(int order, double x0, const double* x, int n, double* w_out);
/* nrows rows at once: row r = weights at x0[r] on the n nodes x[r*n .. r*n+n-1] (the per-node loops that build the
 * non-uniform tables: centered_diff_weights.jl:94-103, upwind_diff_weights.jl:107-133, half_offset_weights.jl:94-120) */
int mol_fd_weights_rows(int order, int64_t nrows, int n, const double* x0, const double* x, double* w_out);

/* -- plan ---------------------------------------------------------------------------------------------- */
/* device >= 0: compile and load on that CUDA device.  device == -1: compile only (NVRTC to cubin,
 * works without a GPU; mol_rhs etc. then fail with MOL_E_NOCUDA). */
int         mol_plan_create (const char* program, size_t nbytes, int device, mol_plan** out);
int         mol_plan_destroy(mol_plan*);
size_t      mol_plan_state_len(const mol_plan*);              /* number of FP64 unknowns */
int         mol_plan_nvar(const mol_plan*);
int         mol_plan_var_info(const mol_plan*, int var, int64_t* offset, int64_t* extents /*[ndim]*/);
int         mol_plan_set_option(mol_plan*, const char* key, int64_t value);
const char* mol_plan_generated_source(const mol_plan*);        /* CUDA source fed to NVRTC */
int         mol_plan_cubin(mol_plan*, const char* kernel_variant, const void** data, size_t* nbytes);
/* Compile every kernel variant the integrator `alg` (MOL_ALG_*) launches, on several host threads at once (the reference
 * pays this once per problem too: mtkcompile + RuntimeGeneratedFunction, src/MOL_discretization.jl:175-191).
 * mol_rk_init calls it; MOL_COMPILE_THREADS overrides the thread count. */
int         mol_plan_precompile(mol_plan*, int alg);
int64_t     mol_plan_launch_count(const mol_plan*);            /* kernels launched so far through this plan */
/* Introspection: the flattened stencil tables exactly as they are uploaded to the device (row weights, L doubles per
 * row; per row {first tap node, number of taps} / for WENO rows {first tap node, target}).  Host pointers owned by the
 * plan; either output may be NULL. */
int         mol_plan_tables(const mol_plan*, const double** tabw, size_t* ntabw, const int** tabs, size_t* ntabs);

/* Sparsity pattern of the Jacobian d(du)/d(u) of the semi-discrete RHS, read off the stencil program (which unknowns
 * each equation taps, through periodic images and ghost rules): what an implicit solver needs as `jac_prototype`
 * (the reference gets it from ModelingToolkit, `ODEProblem(...; jac = true, sparse = true)`, MOL_discretization.jl:175-191).
 * Compressed sparse column, 0-based, rows sorted within a column: colptr[state_len + 1], rowval[nnz].  Call with
 * rowval == NULL to obtain nnz first.  Both branches of an upwind ifelse are part of the pattern.  Host only. */
int mol_plan_jac_sparsity(const mol_plan*, int64_t* colptr, int64_t* rowval, int64_t* nnz_out);

/* -- a19: du = f(u, p, t) ---------------------------------------------------------------------------------- */
int mol_rhs(mol_plan*, double* du_dev, const double* u_dev, const double* p_host, double t, void* stream);

/* Same call with HOST buffers (state_len doubles each; pin them — cudaHostRegister / CUDA.pin / torch pin_memory —
 * or the copies cannot overlap): the grid is cut into `nchunks` (<= 0: default 16) chunks of planes along the last
 * dimension and the H2D copy, the sweep and the D2H copy of successive chunks overlap on three streams.  The
 * caller's stream completes when du_host is complete.  Single-device plans only. */
int mol_rhs_host(mol_plan*, double* du_host, const double* u_host, const double* p_host, double t, int nchunks, void* stream);

/* -- solution unpacking: what indexing a reference solution with a dependent variable does,
 * sol[u(t,x)]  (src/interface/solution/timedep.jl:30-72): the flat unknown vector becomes the variable on the WHOLE
 * grid -- unknowns from the state, boundary-face nodes from the eliminated boundary equations (`observed`), nodes
 * outside the interior in two or more dimensions 0 (generate_corner_eqs!, generate_bc_eqs.jl:396-416).
 * mol_plan_grid_len returns prod_j n_j (nodes of one variable) and fills nodes[ndim].  mol_unpack converts `nstates`
 * consecutive state vectors (state_len doubles each, e.g. the save_dev block of mol_rk_solve) taken at times
 * t_host[0..nstates) into full_dev: nstates x nvar x (n_0 x n_1 x ..) doubles, first spatial index fastest. */
int64_t mol_plan_grid_len(const mol_plan*, int64_t* nodes /*[ndim]*/);
int     mol_unpack(mol_plan*, double* full_dev, const double* u_dev, int nstates, const double* t_host,
                   const double* p_host, void* stream);

/* -- Jacobian-vector product: jv = J(u, p, t) v with J = d f / d u, evaluated exactly by running the
 * generated equations on dual numbers (no finite-difference step): what a matrix-free Newton-Krylov solver asks of the
 * stiff problems the reference integrates with TRBDF2 / Rodas / FBDF and ModelingToolkit's symbolic Jacobian
 * (test/Brusselator/brusselator_eq.jl:71, MOL_discretization.jl:175-191).  1-D / 2-D programs with a tiled core run the
 * tiled kernel on dual numbers (u tiles and v tiles in shared memory), the rest the table-driven kernel; single-device
 * plans; u_dev, v_dev and jv_dev 16-byte aligned. */
int mol_jvp(mol_plan*, double* jv_dev, const double* u_dev, const double* v_dev, const double* p_host, double t, void* stream);

/* -- a20: explicit Runge-Kutta -------------------------------------------------------------------------- */
int mol_rk_init   (mol_plan*, int alg, double abstol, double reltol, mol_rk** out);
int mol_rk_destroy(mol_rk*);
int mol_rk_set_params(mol_rk*, const double* p_host);
/* one step from (*t, u) with step *dt; adaptive != 0 uses the embedded error estimate + PI controller.  In place: u is
 * overwritten on acceptance (one state copy; mol_rk_step_to avoids it).  The FSAL stage of the previous step is reused
 * only when this call continues it (same array, the time it reached); call mol_rk_reinit after rewriting u in place. */
int mol_rk_step   (mol_rk*, double* u_dev, double* t_inout, double* dt_inout, int adaptive,
                   mol_step_stats* out, void* stream);
/* the same step from u_in into a DIFFERENT array u_out (the caller ping-pongs; no state copy).  After a rejected
 * adaptive step u_out is undefined, *t_inout unchanged and *dt_inout the reduced step. */
int mol_rk_step_to(mol_rk*, const double* u_in_dev, double* u_out_dev, double* t_inout, double* dt_inout, int adaptive,
                   mol_step_stats* out, void* stream);
int mol_rk_reinit (mol_rk*);   /* forget the FSAL stage and the controller history (the caller changed u, p or t) */
/* integrate t0 -> t1 (t1 is a stop time: the last step is shortened to land on it); if nsave > 0 the states at
 * saveat[0..nsave) (non-decreasing, inside [t0, t1]) are stored into save_dev (nsave * state_len doubles, device).
 * Save points never clip a step: they are produced by dense output inside the step that covers them (Tsit5: its
 * 4th-order interpolant; Euler / SSPRK33 / RK4: cubic Hermite), as OrdinaryDiffEq's saveat does.
 * dt0 <= 0 selects the automatic initial step (adaptive Tsit5); fixed-step integration needs dt0 > 0.
 * Who controls the step depends on the problem size (same controller arithmetic, same step sequence in all three):
 * <= 1024 unknowns the whole solve is one launch of a persistent single-CTA kernel; up to 2^21 unknowns (adaptive Tsit5,
 * single device) the controller is a kernel behind the last sweep and whole attempts are queued as a captured CUDA graph
 * on a private stream ordered behind `stream`; beyond that, and in slab mode, the host reads the controller's block back
 * once per attempt.  The call returns after the solve has finished (it synchronises). */
int mol_rk_solve  (mol_rk*, double* u_dev, double t0, double t1, double dt0, int adaptive,
                   const double* saveat, int nsave, double* save_dev, int64_t maxiters,
                   mol_solve_stats* out, void* stream);

/* -- e: slab decomposition over ranks (one process per GPU; split along the slowest spatial axis) --------
 * No reference equivalent (the reference is single-process); SURVEY §8e.  Every rank creates the plan of the
 * GLOBAL problem, then mol_dist_init() restricts it to the rank's slab: the state vectors passed to mol_rhs /
 * mol_rk_* then hold only the local planes (variable-major, state_len_local doubles).  Per RHS evaluation the
 * first/last `halo_planes` planes of every variable go to the neighbouring ranks (ring across a periodic
 * seam; at a non-periodic domain edge the owning rank applies the boundary rule instead).
 *   transport A (built in): mol_dist_unique_id() on rank 0, broadcast the 128 bytes by any means, then
 *     mol_dist_comm_init() on every rank: the library exchanges on a private stream, overlapped with the
 *     interior part of the sweep, inside mol_rhs / mol_rk_*; error norms are all-reduced (ncclAllReduce).
 *     Same-node ranks use CUDA-IPC-mapped ghost pools: copy-engine pushes over NVLink + stream memory-op
 *     flags (no SM is taken from the stencil kernel); otherwise, or with MOL_DIST_TRANSPORT=nccl,
 *     ncclSend/ncclRecv.  Registered arrays (mol_dist_register, mol_rk_init) must be registered in the same
 *     order on every rank.
 *   transport B (caller moves the planes, any fabric): mol_dist_set_halo() + mol_rhs_part(INTERIOR),
 *     move planes, mol_rhs_part(BOUNDARY).
 */
typedef struct mol_dist_info_t {
    int     rank, nranks;
    int     halo_planes;        /* ghost planes per side */
    int     periodic;           /* split axis is periodic (ring) */
    int     prev_rank, next_rank;   /* -1 at a non-periodic domain edge */
    int64_t plane_len;          /* doubles per variable per plane */
    int64_t first_plane;        /* index of this rank's first plane among the global interior planes */
    int64_t n_planes;           /* planes owned by this rank */
    int64_t state_len_local;    /* nvar * n_planes * plane_len */
    int64_t state_len_global;
    int64_t halo_len;           /* doubles per ghost buffer: nvar * halo_planes * plane_len */
} mol_dist_info_t;

#define MOL_PART_INTERIOR 1     /* everything that needs no ghost planes */
#define MOL_PART_BOUNDARY 2     /* the planes next to a neighbouring rank */

int mol_dist_partition(int64_t n_planes, int nranks, int rank, int64_t* first, int64_t* count);  /* host only */
int mol_dist_init (mol_plan*, int rank, int nranks);
int mol_dist_info (const mol_plan*, mol_dist_info_t* out);
int mol_dist_unique_id(void* id_out, size_t nbytes /* >= 128 */);
int mol_dist_comm_init(mol_plan*, const void* unique_id, size_t nbytes);
const char* mol_dist_transport(const mol_plan*);   /* which transport mol_dist_comm_init selected */
int mol_dist_set_halo(mol_plan*, double* lo_recv_dev, double* hi_recv_dev);   /* halo_len doubles each */
int mol_rhs_part(mol_plan*, double* du_dev, const double* u_dev, const double* p_host, double t, int part, void* stream);
int mol_dist_register  (mol_plan*, const double* arr_dev);   /* give a resident array its own ghost planes */
int mol_dist_unregister(mol_plan*, const double* arr_dev);
int mol_dist_invalidate(mol_plan*, const double* arr_dev);   /* the caller rewrote a registered array */
int mol_dist_allreduce_sum(mol_plan*, double* dev, int n, void* stream);

const char* mol_last_error(void);
const char* mol_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MOL_CUDA_H */
