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
};

typedef std::vector<std::string> Rpn;

struct GhostTap { int var, node; double coef; };
struct Ghost { int var, dim, node; std::vector<GhostTap> taps; Rpn expr; };

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
    int r[3] = {0, 0, 0};
    int r0p = 0;
    bool tma = false;       // geometry allows TMA (alignment), used when NIN == 1
    bool vec_store = false;
    size_t tile_stride_doubles = 0;
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
    bool epi = false, tiled = false, tma = false;
    size_t smem = 0;
    int grid_ctas = 0;
};

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
    double* d_grid[3] = {nullptr, nullptr, nullptr};
    std::vector<double> params;
    int64_t launches = 0;
    // frame boxes (interior minus core box) for the generic kernel
    std::vector<std::vector<int>> frame;     // each {lo0,lo1,lo2,hi0,hi1,hi2}
    // dist
    int rank = 0, nranks = 1;
};

struct MolRhsIn {
    int nin = 1;
    const double* a[8] = {nullptr};
    double c[8] = {0};
};
struct MolRhsEpi {
    bool on = false;
    double* comb = nullptr;
    double ec[8] = {0};
    double ek = 0, abstol = 0, reltol = 0;
    double* err = nullptr;
};
int mol_rhs_launch(mol_plan* plan, const MolRhsIn& in, double* out, double t, const MolRhsEpi& epi, cudaStream_t st);
