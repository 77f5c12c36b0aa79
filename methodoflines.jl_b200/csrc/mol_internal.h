// Internal structures of libmol_cuda.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/mol_cuda.h"

namespace mol {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

struct Grid {
    int n = 0;
    bool uniform = true;
    double dx = 0.0;
    std::vector<double> x;
};

struct Var {
    std::string name;
    int ilo[3] = {1, 1, 1}, ihi[3] = {1, 1, 1};
    int per[3] = {0, 0, 0};
    int ext(int j) const { return ihi[j] - ilo[j] + 1; }
};

struct Row { int start = 0; std::vector<double> w; };

// generic per-index row table: index r = idx - first
struct Tab {
    int id = 0, L = 0, nrows = 0, first = 0;
    bool has_core = false;
    int core_lo = 0, core_hi = -1, core_off = 0;
    std::vector<double> core_w;
    // "shape core": rows [score_lo, score_hi] tap the same score_n nodes relative to their own node (first tap =
    // node + score_off) with per-node weights (non-uniform grids): tiled kernels read the weights from the table
    bool has_score = false;
    int score_lo = 0, score_hi = -1, score_off = 0, score_n = 0;
    std::vector<Row> rows;          // nrows entries (core rows expanded too)
    std::vector<char> have;
    int woff = 0, soff = 0;         // offsets in the flattened device tables
};

struct WTab {                       // WENO: per node {first tap node, target}
    int id = 0, nrows = 0, first = 0;
    bool has_core = false;
    int core_lo = 0, core_hi = -1;
    std::vector<int> start, target;
    std::vector<char> have;
    int soff = 0;
    // non-uniform grids (set by weno_nu_tables, csrc/mol_parse.cpp): the u-independent part of the reconstruction
    bool nu = false;                // the W token that uses this table carries dx == 0
    int var = -1, dim = -1;         // variable / dimension of that token (periodic wrap of the chart coordinates)
    int goff = 0, glo = 0, glen = 0;    // compact per-interval arrays of the core rows: tabw[goff + a*glen + (j - glo)], a = 0..2
    int roff = 0, nrec = 0;         // records (MOL_WREC doubles each) of the explicit rows
    bool allpos = false;            // all three ideal weights positive at every core node: the +/- splitting is the identity
};
constexpr int kWenoRec = 23;        // = MOL_WREC in kernels/mol_device.cuh

typedef std::vector<std::string> Rpn;

struct GhostTap { int var, node; double coef; };
// tapexpr: per-tap coefficient expressions (parameters, t, coordinates along the boundary); empty = the constant coefs
struct Ghost { int var, dim, node; std::vector<GhostTap> taps; Rpn expr; std::vector<Rpn> tapexpr; };

struct Program {
    int ndim = 0, nvar = 0, nparam = 0;
    std::vector<std::string> pname;
    std::vector<double> pdefault;
    Grid grid[3];
    std::vector<Var> vars;
    std::map<int, Tab> tabs;
    std::map<int, WTab> wtabs;
    std::map<int, Rpn> fns;
    std::vector<Ghost> ghosts;
    std::vector<Rpn> eqs;           // per var
    bool has_core = false;
    int clo[3] = {1, 1, 1}, chi[3] = {0, 0, 0};
    // packed weight records of the tiled kernel on non-uniform axes (weight_records, csrc/mol_parse.cpp): per dimension
    // one record of wrec_stride doubles per core node, holding the row weights of every "shape core" table of that
    // dimension; staged into shared memory with the tile.  wrec_stride == 0: none.
    int wrec_off[3] = {0, 0, 0}, wrec_stride[3] = {0, 0, 0}, wrec_lo[3] = {1, 1, 1}, wrec_n[3] = {0, 0, 0};
    int wrec_hl[3] = {0, 0, 0}, wrec_hh[3] = {0, 0, 0};      // records a tile needs before / after its own nodes (WENO: 2 / 1)
    std::map<int, int> wrec_pos;    // table id -> offset of its weights inside the record
    std::map<int, int> wrec_wpos;   // WENO table id -> offset of {h, 1/h, 1/(two-interval span)} of interval j in record j
    // flattened tables
    std::vector<double> tabw;
    std::vector<int> tabs_flat;
    int64_t voff[8] = {0};
    int64_t nstate = 0;
};

int parse_program(const char* text, size_t nbytes, Program& P);

struct TileCfg {
    bool enabled = false;
    int tx = 0, ty = 1, tz = 1, vx = 2, nthreads = 256, stages = 2;
    int min_ctas = 0;       // __launch_bounds__ minimum CTAs/SM (0: derive from shared-memory footprint)
    int r[3] = {0, 0, 0};
    int r0p = 0;
    bool zmarch = false;    // 3-D: xy tiles marching along z through a ring of `ring` planes; tz = planes per work item
    int ring = 0;
    int l2_ahead = 0;       // planes pulled into L2 ahead of the ring's own loads (measured: evicted before use, off)
    bool tma = false;       // geometry allows TMA (alignment), used when NIN == 1
    bool vec_store = false;
    size_t tile_stride_doubles = 0;
    size_t wstage_doubles = 0;   // per pipeline stage: the tile's per-node records of the non-uniform axes (0: none)
    bool jvp = false;       // the tiled equations also exist on dual numbers (tiled Jacobian-vector product; 1-D / 2-D)
};

struct GenSource {
    std::string prelude;    // constants shared by all variants
    std::string body;       // ghost rules, fns, equations (generic + tile)
    TileCfg tile;
};

int generate_source(const Program& P, GenSource& out);

// driver API entry points resolved at run time (no link-time dependency on libcuda)
struct Driver {
    bool ok = false;
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             CUstream, void**, void**) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
};
int load_driver(Driver& D);

int nvrtc_compile(const std::string& src, const std::vector<std::string>& defines, std::string& cubin,
                  std::string& log);

}  // namespace mol

struct MolVariant {
    std::string key;
    std::string cubin;
    CUmodule module = nullptr;
    CUfunction fn = nullptr;
    int nin = 1;
    int epi = 0;
    bool tiled = false, tma = false;
    size_t smem = 0;
    int min_ctas = 0;            // __launch_bounds__ minimum CTAs/SM the variant was compiled for (after spill back-off)
    int grid_ctas = 0;
};

// ---- slab decomposition state (mol_dist.cpp) ------------------------------------------------------
struct MolHalo {                 // ghost planes of one state-sized array
    double* lo = nullptr;        // nvar * H * plane_max doubles: planes below the slab
    double* hi = nullptr;        // planes above the slab
    bool owned = false;          // allocated by the library (mol_dist_register)
    bool fresh = false;          // planes match the array's current contents
    int slot = -1;               // peer-to-peer transport: slot of the ghost-plane pool
};

// Peer-to-peer transport (same node, NVLink): every rank owns one pool of ghost-plane slots that its
// two neighbours map through CUDA IPC.  A rank PUSHES its edge planes into the neighbour's pool with
// copy-engine transfers and then bumps a sequence flag there (stream memory operation); the
// receiver's stream waits on its own flag.  No SM is used, so the exchange overlaps a persistent
// stencil kernel that fills the whole GPU.  Slots are double buffered by sequence parity.
#define MOL_P2P_SLOTS 16
struct MolP2P {
    bool on = false;
    size_t halo_bytes = 0;       // one side, one parity
    char* pool = nullptr;        // [slot][parity][side][halo_bytes] then flags [slot][side] (uint64)
    size_t flags_off = 0;
    char* prev_pool = nullptr;   // neighbours' pools (IPC mappings); equal when prev == next
    char* next_pool = nullptr;
    unsigned long long seq[MOL_P2P_SLOTS] = {0};
    bool used[MOL_P2P_SLOTS] = {false};
    CUresult (*WriteValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int) = nullptr;
    CUresult (*WaitValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int) = nullptr;
};

struct MolDist {
    bool on = false;
    int rank = 0, nranks = 1;
    int split = 0;               // split dimension (the last one)
    bool periodic = false;       // ring across the seam
    int H = 0;                   // ghost planes per side
    int64_t plane = 0;           // doubles per variable per plane (all variables share the interior box)
    int64_t plane_max = 0;
    int glo = 0, ghi = 0;        // global interior range along the split dimension
    int loc_lo = 0, loc_hi = 0;  // this rank's planes
    int64_t rows = 0;            // loc_hi - loc_lo + 1
    int64_t vstride = 0;         // rows * plane
    int64_t nstate_local = 0;
    int64_t nstate_global = 0;
    int prev = -1, next = -1;    // neighbour ranks (-1: domain edge)
    // built-in transport: NCCL send/recv on a private stream
    void* comm = nullptr;        // ncclComm_t
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
    MolP2P p2p;
    std::map<const double*, MolHalo> halos;     // registered arrays -> ghost planes
    MolHalo scratch;             // ghost planes for unregistered (caller-owned) arrays
    // boxes (global node numbers, inclusive): tiled core part, frame parts that need no ghost planes,
    // and the slab-edge parts that do
    std::vector<int> tile_box;                  // empty: no tiled part
    std::vector<std::vector<int>> inner_frame, edge_frame;
    std::vector<std::vector<int>> edge_tiles;   // slab-edge parts of the core box (tiled kernel, after the exchange)
    bool whole_slab_tiled = false;              // the core box covers every node of this rank's slab (no frame kernels at all)
};

#define MOL_PART_ALL      0
#define MOL_PART_INTERIOR 1      /* everything that needs no ghost planes */
#define MOL_PART_BOUNDARY 2      /* the H planes next to each slab edge */

struct mol_plan {
    mol::Program P;
    mol::GenSource G;
    std::string full_source;
    int device = -1;
    int kernel_mode = MOL_KERNEL_AUTO;
    int sm_count = 148;
    mol::Driver drv;
    std::map<std::string, MolVariant> variants;
    double* d_tabw = nullptr;
    int* d_tabs = nullptr;
    int* d_counter = nullptr;    // tile ticket counter of the persistent tiled kernel (self re-arming)
    double* d_grid[3] = {nullptr, nullptr, nullptr};
    std::vector<double> params;
    int64_t launches = 0;
    // FIN epilogue: every CTA of every launch of one RHS evaluation writes its part of the error sum to its own slot of
    // MolRhsEpi::err; this cursor counts the slots handed out (reset by mol_rhs_launch, read by the caller afterwards)
    int fin_slot = 0;
    // frame boxes (interior minus core box) for the generic kernel
    std::vector<std::vector<int>> frame;     // each {lo0,lo1,lo2,hi0,hi1,hi2}
    MolDist dist;
    bool dist_tiled_edges = false;
    // sub-box override used by the pipelined host-buffer path (mol_rhs_host): evaluate only these boxes
    bool ov_on = false;
    std::vector<int> ov_tile;                        // empty: no tiled part
    std::vector<std::vector<int>> ov_frame;
    struct HostPipe {                                // staging for mol_rhs_host
        double* d_u = nullptr;
        double* d_du = nullptr;
        cudaStream_t s_in = nullptr, s_out = nullptr;
        std::vector<cudaEvent_t> ev_in, ev_cmp;
        cudaEvent_t ev_free_u = nullptr, ev_free_du = nullptr, ev_out = nullptr, ev_start = nullptr;
        bool used = false;
    } hp;
    // tensor maps are cached per input pointer (encoding costs a few microseconds per call); the RK drivers
    // alternate between a handful of arrays, so a few entries are kept (round-robin replacement)
    struct MapSet {
        const double* ptr = nullptr;
        bool dist = false;
        alignas(64) unsigned char maps[8 * 128];
    };
    MapSet mapsets[4];
    int map_next = 0;
};

struct MolRhsIn {
    int nin = 1;
    const double* a[8] = {nullptr};
    double c[8] = {0};
    // device-side step control (MOL_DEVDT kernel variants): address of {t, dt, skip}; the coefficients above, the time
    // argument and the epilogue coefficients are then passed WITHOUT their factor dt (c[j] = a_sj, t = c_s)
    const double* ctl = nullptr;
};
#define MOL_ALG_DEVDT 0x100      // flag on mol_plan_precompile's algorithm code: the MOL_DEVDT variants of that integrator
namespace mol {
// fused ghost-plane wait: see dist_prepare_halos (csrc/mol_dist.cpp) and mol_wait_ghost_planes (kernels/mol_tiled.cuh)
struct MolFuse { bool want = false, on = false; const unsigned long long* flag[2] = {nullptr, nullptr}; unsigned long long seq = 0; };
int dist_prepare_halos(mol_plan* plan, const MolRhsIn& in, const double** hlo, const double** hhi, cudaStream_t st,
                       bool* exchanging, MolFuse* fuse = nullptr);
void dist_mark_stale(mol_plan* plan, const double* arr);
int dist_allreduce_sum(mol_plan* plan, double* dev, int n, cudaStream_t st);
void dist_destroy(mol_plan* plan);
void compute_frame(mol_plan* plan);
// the persistent single-CTA solver kernel (kernels/mol_generic.cuh, MOL_KERNEL_SOLVE); args = its MolSolveArgs block
int mol_plan_solve_small(mol_plan* plan, const void* solve_args, size_t nbytes, double t0, cudaStream_t st);
}
// fused Runge-Kutta epilogues (MolEpi in kernels/mol_device.cuh)
#define MOL_EPI_NONE 0
#define MOL_EPI_PRE  2   /* last-but-one stage: writes u+ (comb) and the partial error estimate (eout) instead of k */
#define MOL_EPI_FIN  3   /* last stage on the single input u+: writes k, accumulates the scaled error norm */
struct MolRhsEpi {
    int mode = MOL_EPI_NONE;
    // PRE
    double* comb = nullptr;
    double* eout = nullptr;
    double cb[8] = {0}, ce[8] = {0};
    double cbk = 0, cek = 0;
    // FIN
    const double* e = nullptr;
    const double* u0 = nullptr;
    double ek = 0, abstol = 0, reltol = 0;
    double* err = nullptr;           // MOL_FIN_SLOTS doubles: one partial sum per CTA (mol_plan::fin_slot of them are written)
};
#define MOL_FIN_SLOTS 65536
int mol_rhs_launch(mol_plan* plan, const MolRhsIn& in, double* out, double t, const MolRhsEpi& epi, cudaStream_t st,
                   int part = MOL_PART_ALL);
