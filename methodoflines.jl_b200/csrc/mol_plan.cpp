// libmol_cuda.so — plan creation (NVRTC), kernel variants, RHS launches.  See include/mol_cuda.h.
//
// There is deliberately NO CPU fallback: without a CUDA device every compute entry point returns
// MOL_E_NOCUDA.  device == -1 (compile only) exists so the build can be checked on a GPU-less host.
#include <dlfcn.h>
#include <unistd.h>
#include <nvrtc.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <atomic>
#include <thread>

#include "mol_internal.h"
#include <string>
#include "mol_kernels_embed.inc"   // generated at build time from kernels/*.cuh

namespace mol {

const char* last_error_cstr();

// ------------------------------------------------------------------------------------------ NVRTC
struct Nvrtc {
    void* h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    const char* (*GetErrorString)(nvrtcResult);
    nvrtcResult (*Version)(int*, int*);          // optional
};

static Nvrtc* get_nvrtc(std::string& err) {
    static Nvrtc N;
    static bool tried = false;
    if (N.h) return &N;
    if (tried) { err = "libnvrtc not available"; return nullptr; }
    tried = true;
    // MOL_NVRTC_PATH (a full path) wins: a process that already holds another libnvrtc.so.12 (PyTorch bundles its own)
    // would otherwise get that one back from dlopen by soname, and register allocation differs between NVRTC releases
    const char* env = getenv("MOL_NVRTC_PATH");
    if (env && *env) N.h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    const char* cands[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* c : cands) {
        if (N.h) break;
        N.h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
    }
    if (!N.h) { err = "cannot dlopen libnvrtc.so.12 (set MOL_NVRTC_PATH)"; return nullptr; }
#define SYM(field, name)                                                   \
    *(void**)(&N.field) = dlsym(N.h, name);                                \
    if (!N.field) { err = std::string("libnvrtc lacks ") + name; N.h = nullptr; return nullptr; }
    SYM(CreateProgram, "nvrtcCreateProgram")
    SYM(CompileProgram, "nvrtcCompileProgram")
    SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    SYM(GetCUBIN, "nvrtcGetCUBIN")
    SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    SYM(GetProgramLog, "nvrtcGetProgramLog")
    SYM(DestroyProgram, "nvrtcDestroyProgram")
    SYM(GetErrorString, "nvrtcGetErrorString")
#undef SYM
    *(void**)(&N.Version) = dlsym(N.h, "nvrtcVersion");
    return &N;
}

// ---- optional on-disk cache of compiled variants (opt-in: MOL_CUBIN_CACHE=<directory>) ---------------------------------
// NVRTC needs 0.3-2 s per kernel variant and an adaptive Tsit5 solve touches about a dozen of them; with the cache a
// program that was compiled before (same generated source, same defines, same NVRTC) loads its cubins from disk.
// File = "MOLCUBIN1" | u64 log bytes | log | cubin; written to a temporary name and renamed, so readers never see a
// partial file.  The ptxas report is kept because the register-cap back-off reads it.
static uint64_t fnv1a(const std::string& s, uint64_t h) {
    for (unsigned char ch : s) { h ^= ch; h *= 1099511628211ull; }
    return h;
}
static std::string cache_path(const std::string& src, const std::vector<std::string>& opts) {
    const char* dir = getenv("MOL_CUBIN_CACHE");
    if (!dir || !*dir) return "";
    uint64_t a = 14695981039346656037ull, b = 0x9e3779b97f4a7c15ull;
    for (const auto& o : opts) { a = fnv1a(o, a); a = fnv1a("|", a); b = fnv1a(o, b); b = fnv1a("#", b); }
    a = fnv1a(src, a);
    b = fnv1a(src, b);
    char name[64];
    snprintf(name, sizeof name, "/%016llx%016llx.molcubin", (unsigned long long)a, (unsigned long long)b);
    return std::string(dir) + name;
}
static bool cache_read(const std::string& path, std::string& cubin, std::string& log) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[9];
    uint64_t nlog = 0;
    bool ok = fread(magic, 1, 9, f) == 9 && !memcmp(magic, "MOLCUBIN1", 9) && fread(&nlog, 8, 1, f) == 1 && nlog < (1u << 26);
    if (ok) {
        log.assign((size_t)nlog, 0);
        ok = nlog == 0 || fread(&log[0], 1, (size_t)nlog, f) == nlog;
    }
    if (ok) {
        cubin.clear();
        char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) cubin.append(buf, n);
        ok = cubin.size() > 4 && !memcmp(cubin.data(), "\x7f" "ELF", 4);
    }
    fclose(f);
    return ok;
}
static void cache_write(const std::string& path, const std::string& cubin, const std::string& log) {
    const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;                                   // (cache directory missing or read-only: compile every time)
    const uint64_t nlog = log.size();
    bool ok = fwrite("MOLCUBIN1", 1, 9, f) == 9 && fwrite(&nlog, 8, 1, f) == 1 &&
              (nlog == 0 || fwrite(log.data(), 1, log.size(), f) == log.size()) &&
              fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

int nvrtc_compile(const std::string& src, const std::vector<std::string>& defines, std::string& cubin, std::string& log) {
    std::string err;
    Nvrtc* N = get_nvrtc(err);
    if (!N) return fail(MOL_E_COMPILE, err);
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--ptxas-options=-v", "--fmad=true",
                                     "--prec-div=true", "--prec-sqrt=true", "--ftz=false"};
    for (auto& d : defines) opts.push_back("-D" + d);
    std::vector<std::string> keyed = opts;
    if (N->Version) {
        int major = 0, minor = 0;
        N->Version(&major, &minor);
        keyed.push_back("nvrtc " + std::to_string(major) + "." + std::to_string(minor));
    }
    const std::string cpath = cache_path(src, keyed);
    if (!cpath.empty() && cache_read(cpath, cubin, log)) return MOL_OK;
    nvrtcProgram prog;
    nvrtcResult r = N->CreateProgram(&prog, src.c_str(), "mol_program.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return fail(MOL_E_COMPILE, std::string("nvrtcCreateProgram: ") + N->GetErrorString(r));
    std::vector<const char*> o;
    for (auto& s : opts) o.push_back(s.c_str());
    r = N->CompileProgram(prog, (int)o.size(), o.data());
    size_t ls = 0;
    N->GetProgramLogSize(prog, &ls);
    log.assign(ls, 0);
    if (ls) N->GetProgramLog(prog, &log[0]);
    if (r != NVRTC_SUCCESS) {
        N->DestroyProgram(&prog);
        return fail(MOL_E_COMPILE, std::string("NVRTC compile failed: ") + N->GetErrorString(r) + "\n" + log);
    }
    size_t cs = 0;
    N->GetCUBINSize(prog, &cs);
    cubin.assign(cs, 0);
    N->GetCUBIN(prog, &cubin[0]);
    N->DestroyProgram(&prog);
    if (!cpath.empty()) cache_write(cpath, cubin, log);
    return MOL_OK;
}

// ------------------------------------------------------------------------------------------ driver
int load_driver(Driver& D) {
    if (D.ok) return MOL_OK;
    if (cudaFree(0) != cudaSuccess) {
        cudaGetLastError();
        return fail(MOL_E_NOCUDA, "no usable CUDA device/driver (there is no CPU fallback)");
    }
#define ENTRY(field, name)                                                                               \
    {                                                                                                    \
        void* fp = nullptr;                                                                              \
        cudaDriverEntryPointQueryResult qr;                                                              \
        if (cudaGetDriverEntryPoint(name, &fp, cudaEnableDefault, &qr) != cudaSuccess || !fp)            \
            return fail(MOL_E_CUDA, std::string("driver entry point not found: ") + name);               \
        *(void**)(&D.field) = fp;                                                                        \
    }
    ENTRY(ModuleLoadData, "cuModuleLoadData")
    ENTRY(ModuleUnload, "cuModuleUnload")
    ENTRY(ModuleGetFunction, "cuModuleGetFunction")
    ENTRY(LaunchKernel, "cuLaunchKernel")
    ENTRY(FuncSetAttribute, "cuFuncSetAttribute")
    ENTRY(TensorMapEncodeTiled, "cuTensorMapEncodeTiled")
    ENTRY(GetErrorString, "cuGetErrorString")
    ENTRY(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
#undef ENTRY
    D.ok = true;
    return MOL_OK;
}

static std::string cu_err(const Driver& D, CUresult r) {
    const char* s = nullptr;
    if (D.GetErrorString) D.GetErrorString(r, &s);
    return s ? s : "unknown CUDA driver error";
}

}  // namespace mol

using namespace mol;

// ------------------------------------------------------------------------------------------ plan
static std::string build_source(const mol_plan* plan) {
    std::string s;
    s += plan->G.prelude;
    s += MOL_SRC_DEVICE;
    s += MOL_SRC_JVP;           // dual-number runtime of the Jacobian-vector product (compiled when MOL_KERNEL_JVP is set)
    s += plan->G.body;
    s += "#if MOL_KERNEL_TILED\n";
    s += MOL_SRC_TILED;
    s += "#else\n";
    s += MOL_SRC_GENERIC;
    s += "#endif\n";
    return s;
}

static int get_tiled_jvp_variant(mol_plan* plan, MolVariant** out);
static size_t tile_smem_bytes(const mol_plan* plan, bool tma, int epi, bool use_tma_flavour = false) {
    const TileCfg& T = plan->G.tile;
    // z-march (3-D): a ring of xy planes per variable
    if (T.zmarch) return (size_t)T.ring * plan->P.nvar * T.tile_stride_doubles * 8;
    // PRE epilogue: two halo-free aux tiles per variable (partial u+ and error sums, kernels/mol_tiled.cuh)
    // (+ the staged per-node records of the non-uniform axes behind the tiles, one block per stage; never with TMA)
    if (epi == MOL_EPI_PRE)
        return ((size_t)plan->P.nvar * (T.tile_stride_doubles + 2 * (size_t)T.tx * T.ty * (plan->P.ndim >= 3 ? T.tz : 1)) +
                T.wstage_doubles) * 8;
    const size_t stages = tma ? T.stages : 1;      // `tma` here: any multi-stage flavour (TMA or cp.async)
    return stages * (plan->P.nvar * T.tile_stride_doubles + (use_tma_flavour ? 0 : T.wstage_doubles)) * 8;
}

// Compile one kernel variant (NVRTC -> cubin).  Reads the plan, never writes it: callable from several threads at once
// (mol_plan_precompile); the caller inserts the result into plan->variants.
static std::string variant_key(const mol_plan* plan, bool tiled, int nin, int epi, bool& tma, bool& cpasync, bool devdt = false) {
    const TileCfg& T = plan->G.tile;
    tma = tiled && T.tma && nin == 1 && epi != MOL_EPI_PRE;
    // single-input tiles the TMA unit cannot address (odd row pitch, 1-D): the same multi-stage pipeline with cp.async
    cpasync = tiled && !tma && nin == 1 && epi != MOL_EPI_PRE && !T.zmarch && !getenv("MOL_TILE_NO_CPASYNC");
    std::ostringstream k;
    k << (tiled ? "tiled" : "generic") << "_nin" << nin << (epi == MOL_EPI_PRE ? "_pre" : (epi == MOL_EPI_FIN ? "_fin" : ""))
      << (tma ? "_tma" : (cpasync ? "_cpa" : ""))
      << (plan->dist.on ? "_dist" : "") << (devdt ? "_dd" : "");
    return k.str();
}

static int compile_variant(const mol_plan* plan, bool tiled, int nin, int epi, MolVariant& v, bool devdt = false) {
    const TileCfg& T = plan->G.tile;
    bool tma, cpasync;
    v.key = variant_key(plan, tiled, nin, epi, tma, cpasync, devdt);
    v.nin = nin;
    v.epi = epi;
    v.tiled = tiled;
    v.tma = tma;
    std::vector<std::string> defs = {"MOL_NIN=" + std::to_string(nin), "MOL_EPI=" + std::to_string(epi),
                                     "MOL_KERNEL_TILED=" + std::to_string(tiled ? 1 : 0),
                                     "MOL_TMA=" + std::to_string(tma ? 1 : 0),
                                     "MOL_CPASYNC=" + std::to_string(cpasync ? 1 : 0)};
    if (plan->dist.on) {
        defs.push_back("MOL_DIST=1");
        defs.push_back("MOL_HALO=" + std::to_string(plan->dist.H));
    }
    if (devdt) defs.push_back("MOL_DEVDT=1");      // step size / time / skip flag from a device-resident block (mol_rk.cu)
    if (const char* wr = getenv("MOL_WENO_RATIO"))      // A/B switch (kernels/mol_device.cuh, mol_weno5_uniform); default 1
        defs.push_back(std::string("MOL_WENO_RATIO=") + ((*wr && *wr != '0') ? "1" : "0"));
    if (tiled) {
        v.smem = tile_smem_bytes(plan, tma || cpasync, epi, tma);
        int ctas = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(v.smem, 1)));
        if (T.min_ctas > 0) ctas = T.min_ctas;
        // PRE epilogue: three tiles per variable in shared memory and three accumulators per load in the loader
        if (epi == MOL_EPI_PRE) {
            // three accumulators per load in the loader: at the 64-register cap of 4 CTAs/SM it spills (measured
            // 1758 vs 1534 us per 4096^2 Tsit5 step), so at most 3 CTAs/SM
            ctas = std::max(1, std::min(std::min(ctas, 3), (int)((220 * 1024) / std::max<size_t>(v.smem, 1))));
            const char* e = getenv("MOL_TILE_PRE_MINCTAS");      // tuning override (experiments only)
            if (e && *e) ctas = std::max(1, atoi(e));
        }
        // Register-cap back-off: `ctas` resident CTAs/SM cap the kernel at 65536 / (ctas * threads) registers.
        // Heavy stencil programs (many terms, table-driven weights) spill under the cap of the default
        // occupancy; ptxas reports the spill traffic (--resource-usage), and a variant that spills more than a
        // few registers is recompiled for one CTA less per SM (measured on the PRE epilogue: 4 CTAs/SM with
        // spills 1758 us, 3 CTAs/SM without 1534 us per Tsit5 step; non-uniform 2-D Burgers 4096^2: 457 / 427 / 476 us
        // at 4 / 3 / 2 CTAs/SM, so the back-off stops at 3).  An explicit MOL_TILE_MINCTAS is honoured.
        const char* forced = getenv("MOL_TILE_MINCTAS");
        std::string log;
        for (;;) {
            std::vector<std::string> d2 = defs;
            d2.push_back("MOL_MIN_CTAS=" + std::to_string(ctas));
            int rc = nvrtc_compile(plan->full_source, d2, v.cubin, log);
            if (rc != MOL_OK) return rc;
            long spill = 0;      // largest "N bytes spill stores" of any function in the ptxas report
            for (size_t pos = log.find("bytes spill stores"); pos != std::string::npos;
                 pos = log.find("bytes spill stores", pos + 1)) {
                const size_t b = log.rfind(',', pos);
                if (b != std::string::npos) spill = std::max(spill, atol(log.c_str() + b + 1));
            }
            // (one more step, to 2 CTAs/SM, for kernels that still spill at 3 -- measured on a B200: non-uniform WENO5 2-D,
            // 288 B of spills at 80 registers, 84.5 us at 3 CTAs/SM vs 63.1 us at 2 (2048^2); non-uniform 2-D Burgers with
            // staged weight records, 120 B, 219.6 vs 201.0 us (4097^2))
            if (getenv("MOL_DEBUG_SPILL")) fprintf(stderr, "[mol] %s at %d CTAs/SM: %ld bytes of spill stores\n", v.key.c_str(), ctas, spill);
            if (spill <= 48 || ctas <= 2 || (ctas == 3 && spill <= 64) || (forced && *forced)) break;
            --ctas;
        }
        v.min_ctas = ctas;
    } else {
        std::string log;
        int rc = nvrtc_compile(plan->full_source, defs, v.cubin, log);
        if (rc != MOL_OK) return rc;
    }
    return MOL_OK;
}

static int get_variant(mol_plan* plan, bool tiled, int nin, int epi, MolVariant** out, bool devdt = false) {
    bool tma, cpasync;
    const std::string key = variant_key(plan, tiled, nin, epi, tma, cpasync, devdt);
    auto it = plan->variants.find(key);
    if (it == plan->variants.end()) {
        MolVariant v;
        int rc = compile_variant(plan, tiled, nin, epi, v, devdt);
        if (rc != MOL_OK) return rc;
        plan->variants[v.key] = v;
        it = plan->variants.find(v.key);
    }
    MolVariant& v = it->second;
    if (plan->device >= 0 && !v.fn) {
        CUresult r = plan->drv.ModuleLoadData(&v.module, v.cubin.data());
        if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "cuModuleLoadData: " + cu_err(plan->drv, r));
        r = plan->drv.ModuleGetFunction(&v.fn, v.module, tiled ? "mol_rhs_tiled" : "mol_rhs_generic");
        if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "cuModuleGetFunction: " + cu_err(plan->drv, r));
        if (tiled) {
            r = plan->drv.FuncSetAttribute(v.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)v.smem);
            if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "cuFuncSetAttribute(smem): " + cu_err(plan->drv, r));
            int nb = 1;
            plan->drv.OccupancyMaxActiveBlocksPerMultiprocessor(&nb, v.fn, plan->G.tile.nthreads, v.smem);
            // PRE epilogue: the residency the variant was compiled for is also the residency it runs at.  The register
            // allocator may stay under the next occupancy step by itself (NVRTC 12.8 gives this kernel 64-70 registers under
            // a cap of 85, depending on unrelated source details), and a fourth resident CTA of this kernel -- 8 input and 2
            // output streams per CTA -- costs 230 us per 4096^2 sweep (605 vs 371 us, profiles/r02_kernels.md): pad the
            // dynamic shared memory request until only min_ctas fit.
            const char* cap = getenv("MOL_TILE_CAP_RESIDENCY");      // "1": all tiled variants (experiments), "0": none
            if (((epi == MOL_EPI_PRE && !(cap && *cap == '0')) || (cap && *cap == '1')) && v.min_ctas > 0) {
                size_t smem = v.smem;
                while (nb > v.min_ctas && smem + 2048 <= 227 * 1024) {
                    smem += 2048;
                    if (plan->drv.FuncSetAttribute(v.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem) != CUDA_SUCCESS) break;
                    plan->drv.OccupancyMaxActiveBlocksPerMultiprocessor(&nb, v.fn, plan->G.tile.nthreads, smem);
                    v.smem = smem;
                }
            }
            if (getenv("MOL_DEBUG_SPILL"))
                fprintf(stderr, "[mol] %s: compiled for %d CTAs/SM, %d resident, %zu bytes of shared memory\n", v.key.c_str(), v.min_ctas, nb, v.smem);
            v.grid_ctas = std::max(1, nb) * plan->sm_count;
        } else {
            int nb = 1;
            plan->drv.OccupancyMaxActiveBlocksPerMultiprocessor(&nb, v.fn, 256, 0);
            v.grid_ctas = std::max(1, nb) * plan->sm_count;
        }
    }
    *out = &v;
    return MOL_OK;
}

// B \ T for nested boxes {lo0,lo1,lo2,hi0,hi1,hi2}: per dimension a lower and an upper slab, with the
// dimensions already peeled restricted to T (the reference's frame around the core box,
// array_discretization.jl:368-420)
static void peel(const std::vector<int>& B, const std::vector<int>& T, int ndim, std::vector<std::vector<int>>& out) {
    std::vector<int> cur = B;
    for (int j = 0; j < ndim; ++j) {
        if (T[j] > cur[j]) {
            std::vector<int> b = cur;
            b[3 + j] = T[j] - 1;
            out.push_back(b);
        }
        if (T[3 + j] < cur[3 + j]) {
            std::vector<int> b = cur;
            b[j] = T[3 + j] + 1;
            out.push_back(b);
        }
        cur[j] = T[j];
        cur[3 + j] = T[3 + j];
    }
}

static bool box_empty(const std::vector<int>& b, int ndim) {
    for (int j = 0; j < ndim; ++j)
        if (b[3 + j] < b[j]) return true;
    return false;
}

namespace mol {
void compute_frame(mol_plan* plan) {
    const Program& P = plan->P;
    plan->frame.clear();
    std::vector<int> B = {1, 1, 1, 1, 1, 1};
    for (int j = 0; j < P.ndim; ++j) {
        B[j] = P.vars[0].ilo[j];
        B[3 + j] = P.vars[0].ihi[j];
        for (int v = 1; v < P.nvar; ++v) {
            B[j] = std::min(B[j], P.vars[v].ilo[j]);
            B[3 + j] = std::max(B[3 + j], P.vars[v].ihi[j]);
        }
    }
    const bool tiled = plan->G.tile.enabled && plan->kernel_mode == MOL_KERNEL_AUTO;
    std::vector<int> core = {P.clo[0], P.clo[1], P.clo[2], P.chi[0], P.chi[1], P.chi[2]};
    MolDist& D = plan->dist;
    if (!D.on) {
        if (!tiled) plan->frame.push_back(B);
        else peel(B, core, P.ndim, plan->frame);
        return;
    }
    // slab: planes [loc_lo, loc_hi] of the split dimension; the E planes next to a neighbouring rank need
    // ghost planes (boundary part, after the exchange), everything else does not (interior part).  E is one
    // tile thick when the tiled kernel is used, so the boundary part is made of whole tiles.
    const int s = D.split;
    D.tile_box.clear();
    D.inner_frame.clear();
    D.edge_frame.clear();
    D.edge_tiles.clear();
    B[s] = D.loc_lo;
    B[3 + s] = D.loc_hi;
    const int tdim[3] = {plan->G.tile.tx, plan->G.tile.ty, plan->G.tile.tz};
    // Measured on 2 x B200 (profiles/r01_scaling.md): handing the slab edges to the tiled kernel serialises them
    // behind the persistent interior sweep (Brusselator 4096^2: 113 us vs 92 us; 3-D 1024^2 x 128: 562 vs 468 us),
    // whereas the table-driven kernel co-runs with it on the communication stream.  Default: table-driven edges.
    const bool tiled_edges = tiled && plan->dist_tiled_edges;
    const int E = tiled_edges ? std::max(D.H, tdim[s]) : D.H;
    std::vector<int> inner = B;
    std::vector<std::vector<int>> edges;
    if (D.prev >= 0) {
        std::vector<int> e = B;
        e[3 + s] = std::min(D.loc_hi, D.loc_lo + E - 1);
        edges.push_back(e);
        inner[s] = e[3 + s] + 1;
    }
    if (D.next >= 0 && inner[s] <= D.loc_hi) {
        std::vector<int> e = B;
        e[s] = std::max(inner[s], D.loc_hi - E + 1);
        edges.push_back(e);
        inner[3 + s] = e[s] - 1;
    }
    auto split_box = [&](const std::vector<int>& box, std::vector<int>* tile, std::vector<std::vector<int>>& frame) {
        if (box_empty(box, P.ndim)) return;
        if (!tiled) { frame.push_back(box); return; }
        std::vector<int> T = core;
        T[s] = std::max(core[s], box[s]);
        T[3 + s] = std::min(core[3 + s], box[3 + s]);
        if (box_empty(T, P.ndim)) { frame.push_back(box); return; }
        *tile = T;
        peel(box, T, P.ndim, frame);
    };
    {   // does the core box cover the whole slab?  (then one tiled launch can do everything: fused ghost-plane wait)
        bool whole = tiled;
        for (int j = 0; j < P.ndim && whole; ++j) {
            const int lo = j == s ? std::max(core[j], D.loc_lo) : core[j], hi = j == s ? std::min(core[3 + j], D.loc_hi) : core[3 + j];
            if (lo != B[j] || hi != B[3 + j]) whole = false;
        }
        D.whole_slab_tiled = whole;
    }
    split_box(inner, &D.tile_box, D.inner_frame);
    for (auto& e : edges) {
        std::vector<int> T;
        if (tiled_edges) split_box(e, &T, D.edge_frame);
        else D.edge_frame.push_back(e);
        if (!T.empty()) D.edge_tiles.push_back(T);
    }
}
}  // namespace mol

extern "C" int mol_plan_create(const char* program, size_t nbytes, int device, mol_plan** out) {
    if (!program || !out) return fail(MOL_E_ARG, "null argument");
    mol_plan* plan = new mol_plan();
    int rc = parse_program(program, nbytes, plan->P);
    if (rc == MOL_OK) rc = generate_source(plan->P, plan->G);
    if (rc != MOL_OK) { delete plan; return rc; }
    plan->full_source = build_source(plan);
    plan->params = plan->P.pdefault;
    plan->device = device;
    compute_frame(plan);
    if (device >= 0) {
        if (cudaSetDevice(device) != cudaSuccess) {
            cudaGetLastError();
            delete plan;
            return fail(MOL_E_NOCUDA, "cudaSetDevice failed: no usable CUDA device (there is no CPU fallback)");
        }
        rc = load_driver(plan->drv);
        if (rc != MOL_OK) { delete plan; return rc; }
        cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, device);
        const Program& P = plan->P;
        cudaError_t e = cudaMalloc(&plan->d_tabw, P.tabw.size() * 8);
        if (e == cudaSuccess) e = cudaMemcpy(plan->d_tabw, P.tabw.data(), P.tabw.size() * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&plan->d_counter, 64);
        if (e == cudaSuccess) e = cudaMemset(plan->d_counter, 0, 64);
        if (e == cudaSuccess) e = cudaMalloc(&plan->d_tabs, P.tabs_flat.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(plan->d_tabs, P.tabs_flat.data(), P.tabs_flat.size() * 4, cudaMemcpyHostToDevice);
        for (int j = 0; j < P.ndim && e == cudaSuccess; ++j) {
            e = cudaMalloc(&plan->d_grid[j], P.grid[j].n * 8);
            if (e == cudaSuccess) e = cudaMemcpy(plan->d_grid[j], P.grid[j].x.data(), P.grid[j].n * 8, cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) {
            std::string m = cudaGetErrorString(e);
            mol_plan_destroy(plan);
            return fail(MOL_E_CUDA, "table upload: " + m);
        }
    }
    // compile the plain RHS variants eagerly so errors surface at discretize time
    MolVariant* v = nullptr;
    bool tiled = plan->G.tile.enabled;
    if (tiled) rc = get_variant(plan, true, 1, false, &v);
    if (rc == MOL_OK && (!tiled || !plan->frame.empty() || true)) rc = get_variant(plan, false, 1, false, &v);
    if (rc != MOL_OK) { mol_plan_destroy(plan); return rc; }
    *out = plan;
    return MOL_OK;
}

static void hostpipe_destroy(mol_plan* plan);

extern "C" int mol_plan_destroy(mol_plan* plan) {
    if (!plan) return MOL_OK;
    if (plan->device >= 0) {
        hostpipe_destroy(plan);
        dist_destroy(plan);
        for (auto& kv : plan->variants)
            if (kv.second.module && plan->drv.ModuleUnload) plan->drv.ModuleUnload(kv.second.module);
        if (plan->d_tabw) cudaFree(plan->d_tabw);
        if (plan->d_tabs) cudaFree(plan->d_tabs);
        if (plan->d_counter) cudaFree(plan->d_counter);
        for (int j = 0; j < 3; ++j)
            if (plan->d_grid[j]) cudaFree(plan->d_grid[j]);
    }
    delete plan;
    return MOL_OK;
}

// Compile, on several host threads at once, every kernel variant one time integrator will launch (NVRTC is thread-safe;
// 0.3-2 s per variant, a dozen variants for adaptive Tsit5): the wait before the first step of a solve drops from the sum
// to about the longest single compile.  Modules are still loaded lazily by the launching thread.
extern "C" int mol_plan_precompile(mol_plan* plan, int alg) {
    if (!plan) return fail(MOL_E_ARG, "null plan");
    struct Want { bool tiled; int nin, epi; };
    std::vector<Want> want;
    std::vector<std::pair<int, int>> stages;        // (nin, epilogue) of every sweep of a step
    const bool devdt = (alg & MOL_ALG_DEVDT) != 0;  // the variants of the queued adaptive solve (csrc/mol_rk.cu)
    alg &= ~MOL_ALG_DEVDT;
    if (devdt && alg != MOL_ALG_TSIT5) return fail(MOL_E_ARG, "device-side step control exists for Tsit5 only");
    switch (alg) {
        case MOL_ALG_EULER: stages = {{1, 0}}; break;
        case MOL_ALG_SSPRK33: stages = {{1, 0}, {2, 0}, {3, 0}}; break;
        case MOL_ALG_RK4: stages = {{1, 0}, {2, 0}}; break;
        case MOL_ALG_TSIT5: stages = {{1, 0}, {2, 0}, {3, 0}, {4, 0}, {5, 0}, {6, MOL_EPI_PRE}, {1, MOL_EPI_FIN}}; break;
        default: return fail(MOL_E_ARG, "unknown algorithm");
    }
    if (devdt) stages.erase(stages.begin());        // (k1 of the first step is evaluated by the host-driven path)
    const bool tiling = plan->G.tile.enabled && plan->kernel_mode == MOL_KERNEL_AUTO;
    for (auto& sg : stages) {
        if (tiling) want.push_back({true, sg.first, sg.second});
        if (!tiling || !plan->frame.empty() || plan->dist.on) want.push_back({false, sg.first, sg.second});
    }
    std::vector<Want> todo;
    for (auto& w : want) {
        bool tma, cpa;
        if (!plan->variants.count(variant_key(plan, w.tiled, w.nin, w.epi, tma, cpa, devdt))) todo.push_back(w);
    }
    if (todo.empty()) return MOL_OK;
    unsigned nthreads = std::thread::hardware_concurrency();
    if (const char* e = getenv("MOL_COMPILE_THREADS")) nthreads = (unsigned)std::max(1, atoi(e));
    nthreads = std::max(1u, std::min<unsigned>(nthreads, (unsigned)todo.size()));
    std::vector<MolVariant> done(todo.size());
    std::vector<int> rcs(todo.size(), MOL_OK);
    std::vector<std::string> errs(todo.size());
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (size_t i = next++; i < todo.size(); i = next++) {
            rcs[i] = compile_variant(plan, todo[i].tiled, todo[i].nin, todo[i].epi, done[i], devdt);
            if (rcs[i] != MOL_OK) errs[i] = last_error_cstr();            // (the error text is per thread)
        }
    };
    std::vector<std::thread> pool;
    for (unsigned k = 1; k < nthreads; ++k) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    for (size_t i = 0; i < todo.size(); ++i) {
        if (rcs[i] != MOL_OK) return fail(rcs[i], errs[i]);
        plan->variants[done[i].key] = done[i];
    }
    return MOL_OK;
}

extern "C" size_t mol_plan_state_len(const mol_plan* plan) {
    if (!plan) return 0;
    return (size_t)(plan->dist.on ? plan->dist.nstate_local : plan->P.nstate);      // local planes only in slab mode
}
extern "C" int mol_plan_nvar(const mol_plan* plan) { return plan ? plan->P.nvar : 0; }
extern "C" int mol_plan_var_info(const mol_plan* plan, int var, int64_t* offset, int64_t* extents) {
    if (!plan || var < 0 || var >= plan->P.nvar) return fail(MOL_E_ARG, "bad variable index");
    if (offset) *offset = plan->P.voff[var];
    if (extents)
        for (int j = 0; j < plan->P.ndim; ++j) extents[j] = plan->P.vars[var].ext(j);
    return MOL_OK;
}
extern "C" int mol_plan_set_option(mol_plan* plan, const char* key, int64_t value) {
    if (!plan || !key) return fail(MOL_E_ARG, "null argument");
    if (!strcmp(key, "kernel")) {
        plan->kernel_mode = (int)value;
        compute_frame(plan);
        return MOL_OK;
    }
    if (!strcmp(key, "dist_tiled_edges")) {      // slab edges through the tiled kernel (after the interior sweep)
        plan->dist_tiled_edges = value != 0;
        compute_frame(plan);
        return MOL_OK;
    }
    return fail(MOL_E_ARG, std::string("unknown option ") + key);
}
extern "C" const char* mol_plan_generated_source(const mol_plan* plan) { return plan ? plan->full_source.c_str() : ""; }
static int get_special_variant(mol_plan* plan, const char* key, const char* define, const char* entry, MolVariant** out,
                               int nin = 1);

extern "C" int mol_plan_cubin(mol_plan* plan, const char* key, const void** data, size_t* nbytes) {
    if (!plan || !key) return fail(MOL_E_ARG, "null argument");
    auto it = plan->variants.find(key);
    if (it == plan->variants.end() && !strcmp(key, "tiled_jvp")) {
        if (!plan->G.tile.enabled || !plan->G.tile.jvp || plan->G.tile.zmarch) return fail(MOL_E_ARG, "this program has no tiled J*v");
        MolVariant* sv = nullptr;
        int rc = get_tiled_jvp_variant(plan, &sv);
        if (rc != MOL_OK) return rc;
        it = plan->variants.find(key);
    }
    if (it == plan->variants.end() && (!strcmp(key, "jvp") || !strcmp(key, "unpack") || !strcmp(key, "solve"))) {
        MolVariant* sv = nullptr;
        int rc = !strcmp(key, "jvp")      ? get_special_variant(plan, "jvp", "MOL_KERNEL_JVP=1", "mol_jvp_generic", &sv)
                 : !strcmp(key, "solve") ? get_special_variant(plan, "solve", "MOL_KERNEL_SOLVE=1", "mol_solve_small", &sv, 7)
                                         : get_special_variant(plan, "unpack", "MOL_KERNEL_UNPACK=1", "mol_unpack_full", &sv);
        if (rc != MOL_OK) return rc;
        it = plan->variants.find(key);
    }
    if (it == plan->variants.end()) {
        // compile on demand: "<tiled|generic>_nin<K>[_pre|_fin][_tma][_dist][_dd]"
        std::string k(key);
        int nin = 0;
        const bool tiled = k.compare(0, 5, "tiled") == 0;
        const size_t pos = k.find("_nin");
        if ((tiled || k.compare(0, 7, "generic") == 0) && pos != std::string::npos) nin = atoi(k.c_str() + pos + 4);
        if (nin >= 1 && nin <= 8 && (!tiled || plan->G.tile.enabled)) {
            MolVariant* v = nullptr;
            const int epi = k.find("_pre") != std::string::npos ? MOL_EPI_PRE : (k.find("_fin") != std::string::npos ? MOL_EPI_FIN : MOL_EPI_NONE);
            int rc = get_variant(plan, tiled, nin, epi, &v, k.size() > 3 && k.compare(k.size() - 3, 3, "_dd") == 0);
            if (rc != MOL_OK) return rc;
            it = plan->variants.find(v->key);       // the library may pick a staging flavour (_tma / _cpa) by itself
        }
    }
    if (it == plan->variants.end()) {
        std::string have;
        for (auto& kv : plan->variants) have += kv.first + " ";
        return fail(MOL_E_ARG, "no such kernel variant; have: " + have);
    }
    if (data) *data = it->second.cubin.data();
    if (nbytes) *nbytes = it->second.cubin.size();
    return MOL_OK;
}
extern "C" int mol_plan_tables(const mol_plan* plan, const double** tabw, size_t* ntabw, const int** tabs, size_t* ntabs) {
    if (!plan) return fail(MOL_E_ARG, "null plan");
    if (tabw) *tabw = plan->P.tabw.data();
    if (ntabw) *ntabw = plan->P.tabw.size();
    if (tabs) *tabs = plan->P.tabs_flat.data();
    if (ntabs) *ntabs = plan->P.tabs_flat.size();
    return MOL_OK;
}
// ---- Jacobian sparsity from the stencil program (first half) --------------------------------------------
namespace {
struct Sparsity {
    const Program& P;
    explicit Sparsity(const Program& p) : P(p) {}
    int64_t flat(int v, const int* idx) const {
        int64_t f = P.voff[v], stride = 1;
        for (int d = 0; d < P.ndim; ++d) { f += (int64_t)(idx[d] - P.vars[v].ilo[d]) * stride; stride *= P.vars[v].ext(d); }
        return f;
    }
    // unknowns the value of variable v at node idx depends on: itself, its periodic image, or the taps of its ghost rule
    void node(int v, const int* idx_in, std::vector<int64_t>& deps, int depth = 0) const {
        int idx[3] = {idx_in[0], idx_in[1], idx_in[2]};
        for (int d = 0; d < P.ndim; ++d) {
            if (idx[d] >= P.vars[v].ilo[d] && idx[d] <= P.vars[v].ihi[d]) continue;
            if (P.vars[v].per[d]) { idx[d] += (idx[d] <= 1) ? (P.grid[d].n - 1) : -(P.grid[d].n - 1); continue; }
            if (depth > 8) return;
            for (const Ghost& g : P.ghosts) {
                if (g.var != v || g.dim != d || g.node != idx[d]) continue;
                for (const GhostTap& tp : g.taps) {
                    int j[3] = {idx[0], idx[1], idx[2]};
                    j[d] = tp.node;
                    node(tp.var, j, deps, depth + 1);
                }
            }
            return;                     // no rule: the node's value is 0
        }
        for (int d = 0; d < P.ndim; ++d)
            if (idx[d] < P.vars[v].ilo[d] || idx[d] > P.vars[v].ihi[d]) return;     // wrapped onto a non-unknown
        deps.push_back(flat(v, idx));
    }
    void row_taps(const Tab& T, int var, int dim, int row_idx, const int* idx, std::vector<int64_t>& deps) const {
        const int r = row_idx - T.first;
        if (r < 0 || r >= T.nrows) return;
        for (size_t k = 0; k < T.rows[r].w.size(); ++k) {
            if (T.rows[r].w[k] == 0.0) continue;
            int j[3] = {idx[0], idx[1], idx[2]};
            j[dim] = T.rows[r].start + (int)k;
            node(var, j, deps);
        }
    }
    int equation(int v, const int* idx, std::vector<int64_t>& deps) const {
        for (const std::string& tk : P.eqs[v]) {
            int a = 0, b = 0, c = 0, d = 0, e = 0, f = 0;
            if (tk[0] == 'u' && sscanf(tk.c_str(), "u:%d", &a) == 1) {
                node(a, idx, deps);
            } else if (tk[0] == 's' && tk.size() > 1 && tk[1] == ':' && sscanf(tk.c_str(), "s:%d", &a) == 1) {
                if (a < 0 || a >= P.nvar) return MOL_E_PARSE;
                int j[3] = {1, 1, 1};
                for (int q = 0; q < P.ndim; ++q) j[q] = P.vars[a].ilo[q];
                node(a, j, deps);
            } else if (tk[0] == 'L' && sscanf(tk.c_str(), "L:%d:%d:%d", &a, &b, &c) == 3) {
                auto it = P.tabs.find(a);
                if (it == P.tabs.end()) return MOL_E_PARSE;
                row_taps(it->second, b, c, idx[c], idx, deps);
            } else if (tk[0] == 'W' && sscanf(tk.c_str(), "W:%d:%d:%d", &a, &b, &c) == 3) {
                auto it = P.wtabs.find(a);
                if (it == P.wtabs.end()) return MOL_E_PARSE;
                const int r = idx[c] - it->second.first;
                if (r < 0 || r >= it->second.nrows) continue;
                for (int k = 0; k < 5; ++k) {
                    int j[3] = {idx[0], idx[1], idx[2]};
                    j[c] = it->second.start[r] + k;
                    node(b, j, deps);
                }
            } else if (tk[0] == 'M' && sscanf(tk.c_str(), "M:%d:%d:%d:%d:%d", &a, &b, &c, &d, &e) == 5) {
                // product of two rows: every (x tap, y tap) pair that is not a corner node
                if (!P.tabs.count(a) || !P.tabs.count(b)) return MOL_E_PARSE;
                const Tab &TX = P.tabs.at(a), &TY = P.tabs.at(b);
                const int rx = idx[d] - TX.first, ry = idx[e] - TY.first;
                if (rx < 0 || rx >= TX.nrows || ry < 0 || ry >= TY.nrows) continue;
                for (size_t kx = 0; kx < TX.rows[rx].w.size(); ++kx)
                    for (size_t ky = 0; ky < TY.rows[ry].w.size(); ++ky) {
                        if (TX.rows[rx].w[kx] == 0.0 || TY.rows[ry].w[ky] == 0.0) continue;
                        int j[3] = {idx[0], idx[1], idx[2]};
                        j[d] = TX.rows[rx].start + (int)kx;
                        j[e] = TY.rows[ry].start + (int)ky;
                        int outside = 0;
                        for (int q = 0; q < P.ndim; ++q)
                            outside += (!P.vars[c].per[q] && (j[q] < P.vars[c].ilo[q] || j[q] > P.vars[c].ihi[q])) ? 1 : 0;
                        if (outside < 2) node(c, j, deps);
                    }
            } else if (tk[0] == 'N' && sscanf(tk.c_str(), "N:%d:%d:%d:%d:%d:%d", &a, &b, &c, &d, &e, &f) == 6) {
                // sum over half points m of wo_m * a(u~_m) * (D u)_m: u~ interpolates every variable the coefficient reads
                if (!P.tabs.count(d) || !P.tabs.count(e) || !P.tabs.count(f) || !P.fns.count(c)) return MOL_E_PARSE;
                const Tab& TO = P.tabs.at(f);
                const int r = idx[b] - TO.first;
                if (r < 0 || r >= TO.nrows) continue;
                for (size_t k = 0; k < TO.rows[r].w.size(); ++k) {
                    if (TO.rows[r].w[k] == 0.0) continue;
                    const int m = TO.rows[r].start + (int)k;
                    for (const std::string& ft : P.fns.at(c)) {
                        int w = 0;
                        if (ft[0] == 'u' && sscanf(ft.c_str(), "u:%d", &w) == 1) row_taps(P.tabs.at(d), w, b, m, idx, deps);
                    }
                    row_taps(P.tabs.at(e), a, b, m, idx, deps);
                }
            }
        }
        return MOL_OK;
    }
};
}  // namespace

extern "C" int mol_plan_jac_sparsity(const mol_plan* plan, int64_t* colptr, int64_t* rowval, int64_t* nnz_out) {
    if (!plan) return fail(MOL_E_ARG, "null plan");
    if (plan->dist.on) return fail(MOL_E_UNSUPPORTED, "the Jacobian pattern is that of the global problem: ask a plan without slabs");
    const Program& P = plan->P;
    Sparsity S(P);
    std::vector<std::vector<int64_t>> cols((size_t)P.nstate);      // per column: rows that depend on it
    std::vector<int64_t> deps;
    for (int v = 0; v < P.nvar; ++v) {
        int idx[3] = {1, 1, 1};
        const Var& V = P.vars[v];
        for (int i2 = (P.ndim >= 3 ? V.ilo[2] : 1); i2 <= (P.ndim >= 3 ? V.ihi[2] : 1); ++i2)
            for (int i1 = (P.ndim >= 2 ? V.ilo[1] : 1); i1 <= (P.ndim >= 2 ? V.ihi[1] : 1); ++i1)
                for (int i0 = V.ilo[0]; i0 <= V.ihi[0]; ++i0) {
                    idx[0] = i0; idx[1] = i1; idx[2] = i2;
                    deps.clear();
                    int rc = S.equation(v, idx, deps);
                    if (rc != MOL_OK) return fail(rc, "malformed equation while reading off the Jacobian pattern");
                    const int64_t row = S.flat(v, idx);
                    std::sort(deps.begin(), deps.end());
                    deps.erase(std::unique(deps.begin(), deps.end()), deps.end());
                    for (int64_t c : deps) cols[(size_t)c].push_back(row);
                }
    }
    int64_t nnz = 0;
    for (auto& c : cols) { std::sort(c.begin(), c.end()); nnz += (int64_t)c.size(); }
    if (nnz_out) *nnz_out = nnz;
    if (colptr) {
        int64_t p = 0;
        for (int64_t j = 0; j < P.nstate; ++j) { colptr[j] = p; p += (int64_t)cols[(size_t)j].size(); }
        colptr[P.nstate] = p;
    }
    if (rowval) {
        int64_t p = 0;
        for (auto& c : cols)
            for (int64_t r : c) rowval[p++] = r;
    }
    return MOL_OK;
}

extern "C" int64_t mol_plan_launch_count(const mol_plan* plan) { return plan ? plan->launches : 0; }
extern "C" const char* mol_last_error(void) { return mol::last_error_cstr(); }
extern "C" const char* mol_version(void) { return "mol_cuda 0.1 (sm_100a, NVRTC-specialised stencil programs)"; }

// ------------------------------------------------------------------------------------------ launch
namespace {
struct ArgBuf {
    std::vector<unsigned char> b;
    template <class T>
    void put(const T& v) {
        size_t a = alignof(T) > 8 ? 8 : alignof(T);
        while (b.size() % a) b.push_back(0);
        const unsigned char* p = reinterpret_cast<const unsigned char*>(&v);
        b.insert(b.end(), p, p + sizeof(T));
    }
    void pad(size_t a) { while (b.size() % a) b.push_back(0); }
};
}  // namespace

// all boxes of one part in a single launch (MolBoxes in mol_generic.cuh: {n, pad, b[8] = {lo[3], hi[3]}, start[9]})
// FIN epilogue: point MolEpi::err (byte offset 40 of the argument block: e, u0, ek, abstol, reltol, err) at the next `grid`
// free slots of the caller's array: one slot per CTA of this launch
static int fin_take_slots(mol_plan* plan, ArgBuf& aepi, double* err_base, int grid) {
    if (!err_base) return MOL_OK;
    if (plan->fin_slot + grid > MOL_FIN_SLOTS) return fail(MOL_E_ARG, "internal: error-norm slots exhausted");
    double* p = err_base + plan->fin_slot;
    memcpy(aepi.b.data() + 40, &p, sizeof p);
    plan->fin_slot += grid;
    return MOL_OK;
}

static int launch_generic_boxes(mol_plan* plan, MolVariant* v, const std::vector<std::vector<int>>& boxes, ArgBuf& ain,
                                ArgBuf& actx, ArgBuf& aepi, bool epi_on, double* out, cudaStream_t st, double* fin_err = nullptr) {
    const Program& P = plan->P;
    const int MAXB = 8;
    for (size_t first = 0; first < boxes.size(); first += MAXB) {
        const int nb = (int)std::min<size_t>(MAXB, boxes.size() - first);
        ArgBuf ab;
        ab.put(nb);
        ab.put((int)0);
        long long start[MAXB + 1] = {0};
        for (int k = 0; k < MAXB; ++k) {
            int64_t total = 0;
            if (k < nb) {
                const std::vector<int>& b = boxes[first + k];
                total = 1;
                for (int j = 0; j < P.ndim; ++j) total *= std::max(0, b[3 + j] - b[j] + 1);
                for (int q = 0; q < 6; ++q) ab.put(b[q]);
            } else {
                for (int q = 0; q < 6; ++q) ab.put((int)(q < 3 ? 1 : 0));
            }
            start[k + 1] = start[k] + total;
        }
        for (int k = 0; k <= MAXB; ++k) ab.put(start[k]);
        const int64_t total = start[nb];
        if (total <= 0) continue;
        int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)v->grid_ctas * 4);
        if (int rcs = fin_take_slots(plan, aepi, fin_err, grid)) return rcs;
        void* args[6];
        int na = 0;
        args[na++] = ain.b.data();
        args[na++] = actx.b.data();
        args[na++] = ab.b.data();
        args[na++] = &out;
        if (epi_on) args[na++] = aepi.b.data();
        CUresult r = plan->drv.LaunchKernel(v->fn, grid, 1, 1, 256, 1, 1, 0, (CUstream)st, args, nullptr);
        if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "launch mol_rhs_generic: " + cu_err(plan->drv, r));
        plan->launches++;
    }
    return MOL_OK;
}

int mol_rhs_launch(mol_plan* plan, const MolRhsIn& in, double* out, double t, const MolRhsEpi& epi, cudaStream_t st,
                   int part) {
    if (!plan) return fail(MOL_E_ARG, "null plan");
    if (plan->device < 0) return fail(MOL_E_NOCUDA, "plan was created compile-only (device = -1); there is no CPU fallback");
    const Program& P = plan->P;
    const TileCfg& T = plan->G.tile;
    MolDist& D = plan->dist;
    const int nin = in.nin;
    if (nin < 1 || nin > 8) return fail(MOL_E_ARG, "nin out of range");
    if (!D.on) part = MOL_PART_ALL;
    // ---- ghost planes of every input array (slab decomposition)
    const double* hlo[8] = {nullptr};
    const double* hhi[8] = {nullptr};
    bool exchanging = false;
    // Fused ghost-plane wait (peer-to-peer transport, tiled programs whose core box covers the whole slab, 2-D or
    // z-marching 3-D): ONE tiled launch over the whole slab; the tiles that read ghost planes -- visited last -- wait for the
    // neighbours' sequence flags inside the kernel, instead of a second launch behind a stream-side wait.
    MolFuse fuse;
    if (D.on) {
        const char* fe = getenv("MOL_DIST_FUSED");
        fuse.want = part == MOL_PART_ALL && T.enabled && plan->kernel_mode == MOL_KERNEL_AUTO && D.whole_slab_tiled &&
                    (P.ndim == 2 || T.zmarch) && !(fe && *fe == '0');
        // part == ALL: the library moves the planes itself (NCCL on its private stream, overlapped with
        // the interior part below); otherwise the caller moved them into the registered buffers
        int rc = dist_prepare_halos(plan, in, hlo, hhi, st, part == MOL_PART_ALL ? &exchanging : nullptr, &fuse);
        if (rc != MOL_OK) return rc;
    }
    // ---- argument blocks shared by both kernels (layouts of MolIn / MolCtx / MolEpi in mol_device.cuh)
    ArgBuf ain;
    for (int j = 0; j < nin; ++j) ain.put(in.a[j]);
    for (int j = 0; j < nin; ++j) ain.put(in.c[j]);
    if (D.on) {
        for (int j = 0; j < nin; ++j) ain.put(hlo[j]);
        for (int j = 0; j < nin; ++j) ain.put(hhi[j]);
    }
    const bool devdt = in.ctl != nullptr;           // MOL_DEVDT variants: MolIn ends with the control block's address
    if (devdt) ain.put(in.ctl);
    ArgBuf actx;
    actx.put(t);
    for (int k = 0; k < std::max(1, P.nparam); ++k) actx.put(k < P.nparam ? plan->params[k] : 0.0);
    for (int j = 0; j < 3; ++j) actx.put((const double*)plan->d_grid[j]);
    actx.put((const double*)plan->d_tabw);
    actx.put((const int*)plan->d_tabs);
    const int last = P.ndim - 1;
    actx.put((int)(D.on ? D.loc_lo : P.vars[0].ilo[last]));
    actx.put((int)(D.on ? D.loc_hi : P.vars[0].ihi[last]));
    actx.put((long long)(D.on ? D.vstride : 0));
    ArgBuf aepi;
    const bool epi_on = epi.mode != MOL_EPI_NONE;
    if (epi.mode == MOL_EPI_PRE) {
        if (!epi.comb || !epi.eout) return fail(MOL_E_ARG, "PRE epilogue needs comb and eout arrays");
        aepi.put(epi.comb);
        aepi.put(epi.eout);
        for (int j = 0; j < nin; ++j) aepi.put(epi.cb[j]);
        for (int j = 0; j < nin; ++j) aepi.put(epi.ce[j]);
        aepi.put(epi.cbk);
        aepi.put(epi.cek);
    } else if (epi.mode == MOL_EPI_FIN) {
        if (!epi.e || !epi.u0 || !out) return fail(MOL_E_ARG, "FIN epilogue needs e, u0 and an output array");
        aepi.put(epi.e);
        aepi.put(epi.u0);
        aepi.put(epi.ek);
        aepi.put(epi.abstol);
        aepi.put(epi.reltol);
        aepi.put(epi.err);               // (re-pointed per launch: fin_take_slots)
        plan->fin_slot = 0;
    } else if (!out) return fail(MOL_E_ARG, "null output array");
    double* const fin_err = epi.mode == MOL_EPI_FIN ? epi.err : nullptr;
    const bool tiling = T.enabled && plan->kernel_mode == MOL_KERNEL_AUTO;
    // the tiled kernel on one or two boxes of nodes (MolTiles in mol_tiled.cuh)
    int fuse_rot = 0;
    auto launch_tiled = [&](const std::vector<std::vector<int>>& boxes) -> int {
        MolVariant* v = nullptr;
        int rc = get_variant(plan, true, nin, epi.mode, &v, devdt);
        if (rc != MOL_OK) return rc;
        const bool use_tma = v->tma;
        if ((use_tma || T.vec_store) && (reinterpret_cast<uintptr_t>(in.a[0]) % 16 != 0))
            return fail(MOL_E_ARG, "state pointer must be 16-byte aligned (128-bit loads / TMA)");
        const int tdim[3] = {T.tx, T.ty, T.tz};
        ArgBuf at;
        int total = 0;
        for (int k = 0; k < 2; ++k) {
            int nt[3] = {1, 1, 1}, lo[3] = {1, 1, 1}, hi[3] = {0, 0, 0}, n = 0;
            if (k < (int)boxes.size()) {
                n = 1;
                for (int j = 0; j < P.ndim; ++j) {
                    lo[j] = boxes[k][j];
                    hi[j] = boxes[k][3 + j];
                    nt[j] = (hi[j] - lo[j] + 1 + tdim[j] - 1) / tdim[j];
                    n *= nt[j];
                }
                for (int j = P.ndim; j < 3; ++j) hi[j] = 1;
            }
            for (int j = 0; j < 3; ++j) at.put(nt[j]);
            at.put(n);
            for (int j = 0; j < 3; ++j) at.put(lo[j]);
            for (int j = 0; j < 3; ++j) at.put(hi[j]);
            total += n;
        }
        if (total <= 0) return MOL_OK;
        at.put(total);
        at.put((int)(fuse.on ? fuse_rot : 0));
        at.put((int*)plan->d_counter);
        at.put(fuse.on ? fuse.flag[0] : (const unsigned long long*)nullptr);
        at.put(fuse.on ? fuse.flag[1] : (const unsigned long long*)nullptr);
        at.put((unsigned long long)(fuse.on ? fuse.seq : 0));
        mol_plan::MapSet* ms = nullptr;
        if (use_tma) {
            for (auto& m : plan->mapsets)
                if (m.ptr == in.a[0] && m.dist == D.on) ms = &m;
        }
        if (use_tma && !ms) {
            ms = &plan->mapsets[plan->map_next];
            plan->map_next = (plan->map_next + 1) % 4;
            ms->ptr = nullptr;
            const int sx = T.tx + 2 * T.r0p, sy = T.ty + 2 * T.r[1], sz = T.zmarch ? 1 : T.tz + 2 * T.r[2];
            for (int var = 0; var < P.nvar; ++var) {
                cuuint64_t gdim[3] = {(cuuint64_t)P.vars[var].ext(0), (cuuint64_t)(P.ndim >= 2 ? P.vars[var].ext(1) : 1),
                                      (cuuint64_t)(P.ndim >= 3 ? P.vars[var].ext(2) : 1)};
                if (D.on) gdim[last] = (cuuint64_t)D.rows;
                cuuint64_t gstr[2] = {gdim[0] * 8, gdim[0] * gdim[1] * 8};
                cuuint32_t box[3] = {(cuuint32_t)sx, (cuuint32_t)(P.ndim >= 2 ? sy : 1), (cuuint32_t)(P.ndim >= 3 ? sz : 1)};
                cuuint32_t estr[3] = {1, 1, 1};
                const double* base = in.a[0] + (D.on ? (int64_t)var * D.vstride : P.voff[var]);
                CUresult r = plan->drv.TensorMapEncodeTiled(
                    reinterpret_cast<CUtensorMap*>(ms->maps + 128 * var), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)P.ndim,
                    (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "cuTensorMapEncodeTiled: " + cu_err(plan->drv, r));
            }
            ms->ptr = in.a[0];
            ms->dist = D.on;
        }
        void* args[8];
        int na = 0;
        args[na++] = ain.b.data();
        args[na++] = actx.b.data();
        args[na++] = at.b.data();
        args[na++] = &out;
        if (use_tma) args[na++] = ms->maps;
        if (epi_on) args[na++] = aepi.b.data();
        int grid = std::min(total, v->grid_ctas);
        if (int rcs = fin_take_slots(plan, aepi, fin_err, grid)) return rcs;
        CUresult r = plan->drv.LaunchKernel(v->fn, grid, 1, 1, T.nthreads, 1, 1, (unsigned)v->smem, (CUstream)st, args, nullptr);
        if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "launch mol_rhs_tiled: " + cu_err(plan->drv, r));
        plan->launches++;
        return MOL_OK;
    };
    auto launch_generic = [&](const std::vector<std::vector<int>>& boxes, cudaStream_t s2) -> int {
        if (boxes.empty()) return MOL_OK;
        MolVariant* v = nullptr;
        int rc = get_variant(plan, false, nin, epi.mode, &v, devdt);
        if (rc != MOL_OK) return rc;
        return launch_generic_boxes(plan, v, boxes, ain, actx, aepi, epi_on, out, s2, fin_err);
    };
    int rc = MOL_OK;
    if (fuse.on) {
        // the whole slab in one launch: first row of tiles (lower ghost planes) rotated to the end of the ticket order
        const int s = D.split;
        std::vector<int> box = {P.clo[0], P.clo[1], P.clo[2], P.chi[0], P.chi[1], P.chi[2]};
        box[s] = std::max(P.clo[s], D.loc_lo);
        box[3 + s] = std::min(P.chi[s], D.loc_hi);
        fuse_rot = (box[3] - box[0] + 1 + T.tx - 1) / T.tx;            // work items per layer along the split dimension
        if (P.ndim == 3) fuse_rot *= (box[4] - box[1] + 1 + T.ty - 1) / T.ty;
        if ((rc = launch_tiled({box}))) return rc;
        // the copy engines may still be reading this rank's edge planes: order later work on `st` behind the pushes
        cudaError_t e = cudaEventRecord(D.ev_done, D.comm_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, D.ev_done, 0);
        if (e != cudaSuccess) return fail(MOL_E_CUDA, std::string("ghost-plane exchange (join): ") + cudaGetErrorString(e));
        if (out) dist_mark_stale(plan, out);
        if (epi.mode == MOL_EPI_PRE) { dist_mark_stale(plan, epi.comb); dist_mark_stale(plan, epi.eout); }
        return MOL_OK;
    }
    // ---- interior part: tiled core + frame boxes that need no ghost planes
    if (part != MOL_PART_BOUNDARY) {
        std::vector<std::vector<int>> tb;
        if (tiling) {
            if (plan->ov_on) { if (!plan->ov_tile.empty()) tb.push_back(plan->ov_tile); }
            else if (D.on) { if (!D.tile_box.empty()) tb.push_back(D.tile_box); }
            else tb.push_back({P.clo[0], P.clo[1], P.clo[2], P.chi[0], P.chi[1], P.chi[2]});
        }
        if (!tb.empty() && (rc = launch_tiled(tb))) return rc;
        if ((rc = launch_generic(plan->ov_on ? plan->ov_frame : (D.on ? D.inner_frame : plan->frame), st))) return rc;
    }
    // ---- boundary part: the planes next to a neighbouring rank, after the ghost planes have landed.
    // Tiled programs: one more tiled launch over both slab edges (whole tiles) on the caller's stream.
    // Table-driven programs: the generic kernel is queued on the communication stream right behind the
    // exchange (it is small enough to co-run with the interior sweep).
    if (D.on && part != MOL_PART_INTERIOR) {
        const bool tiled_edges = tiling && !D.edge_tiles.empty();
        if (exchanging && tiled_edges) {
            cudaError_t e = cudaEventRecord(D.ev_done, D.comm_stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(st, D.ev_done, 0);
            if (e != cudaSuccess) return fail(MOL_E_CUDA, std::string("ghost-plane exchange (join): ") + cudaGetErrorString(e));
        }
        if (tiled_edges) {
            if ((rc = launch_tiled(D.edge_tiles))) return rc;
            if ((rc = launch_generic(D.edge_frame, st))) return rc;
        } else {
            if ((rc = launch_generic(D.edge_frame, exchanging ? D.comm_stream : st))) return rc;
            if (exchanging) {
                cudaError_t e = cudaEventRecord(D.ev_done, D.comm_stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(st, D.ev_done, 0);
                if (e != cudaSuccess) return fail(MOL_E_CUDA, std::string("ghost-plane exchange (join): ") + cudaGetErrorString(e));
            }
        }
    }
    if (D.on) {
        if (out) dist_mark_stale(plan, out);
        if (epi.mode == MOL_EPI_PRE) { dist_mark_stale(plan, epi.comb); dist_mark_stale(plan, epi.eout); }
    }
    return MOL_OK;
}

extern "C" int mol_rhs(mol_plan* plan, double* du_dev, const double* u_dev, const double* p_host, double t, void* stream) {
    if (!plan || !du_dev || !u_dev) return fail(MOL_E_ARG, "null argument");
    if (p_host)
        for (int k = 0; k < plan->P.nparam; ++k) plan->params[k] = p_host[k];
    MolRhsIn in;
    in.nin = 1;
    in.a[0] = u_dev;
    in.c[0] = 1.0;
    MolRhsEpi epi;
    return mol_rhs_launch(plan, in, du_dev, t, epi, (cudaStream_t)stream);
}


// single-purpose table-driven kernels (solution unpacking, Jacobian-vector product): one extra define, one entry point
static int get_special_variant(mol_plan* plan, const char* key, const char* define, const char* entry, MolVariant** out,
                               int nin) {
    auto it = plan->variants.find(key);
    if (it == plan->variants.end()) {
        MolVariant v;
        v.key = key;
        std::string log;
        int rc = nvrtc_compile(plan->full_source, {"MOL_NIN=" + std::to_string(nin), "MOL_EPI=0", "MOL_KERNEL_TILED=0", "MOL_TMA=0",
                                                   "MOL_CPASYNC=0", define},
                               v.cubin, log);
        if (rc != MOL_OK) return rc;
        plan->variants[v.key] = v;
        it = plan->variants.find(key);
    }
    MolVariant& v = it->second;
    if (plan->device >= 0 && !v.fn) {
        CUresult r = plan->drv.ModuleLoadData(&v.module, v.cubin.data());
        if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "cuModuleLoadData: " + cu_err(plan->drv, r));
        r = plan->drv.ModuleGetFunction(&v.fn, v.module, entry);
        if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, std::string("cuModuleGetFunction(") + entry + "): " + cu_err(plan->drv, r));
        int nb = 1;
        plan->drv.OccupancyMaxActiveBlocksPerMultiprocessor(&nb, v.fn, 256, 0);
        v.grid_ctas = std::max(1, nb) * plan->sm_count;
    }
    *out = &v;
    return MOL_OK;
}

// ---- persistent solver kernel (kernels/mol_generic.cuh, MOL_KERNEL_SOLVE): one launch per solve ----------------------------
// args: the MolSolveArgs block of the kernel (pointers, times, tolerances ...), marshalled by csrc/mol_rk.cu
namespace mol {
int mol_plan_solve_small(mol_plan* plan, const void* solve_args, size_t nbytes, double t0, cudaStream_t st) {
    if (!plan || plan->device < 0) return fail(MOL_E_NOCUDA, "plan was created compile-only; there is no CPU fallback");
    const Program& P = plan->P;
    MolVariant* v = nullptr;
    int rc = get_special_variant(plan, "solve", "MOL_KERNEL_SOLVE=1", "mol_solve_small", &v, 7);
    if (rc != MOL_OK) return rc;
    ArgBuf actx;
    actx.put(t0);
    for (int q = 0; q < std::max(1, P.nparam); ++q) actx.put(q < P.nparam ? plan->params[q] : 0.0);
    for (int j = 0; j < 3; ++j) actx.put((const double*)plan->d_grid[j]);
    actx.put((const double*)plan->d_tabw);
    actx.put((const int*)plan->d_tabs);
    const int last = P.ndim - 1;
    actx.put((int)P.vars[0].ilo[last]);
    actx.put((int)P.vars[0].ihi[last]);
    actx.put((long long)0);
    // one bounding box of every variable's interior (MolGenericVars tests each variable's own box)
    ArgBuf ab;
    ab.put((int)1);
    ab.put((int)0);
    long long start[9] = {0};
    long long total = 1;
    for (int k = 0; k < 8; ++k) {
        for (int q = 0; q < 6; ++q) {
            int val = q < 3 ? 1 : 0;
            if (k == 0) {
                const int j = q % 3;
                val = 1;
                if (j < P.ndim) {
                    val = q < 3 ? P.vars[0].ilo[j] : P.vars[0].ihi[j];
                    for (int w = 1; w < P.nvar; ++w) val = q < 3 ? std::min(val, P.vars[w].ilo[j]) : std::max(val, P.vars[w].ihi[j]);
                }
            }
            ab.put(val);
        }
    }
    for (int j = 0; j < P.ndim; ++j) {
        int lo = P.vars[0].ilo[j], hi = P.vars[0].ihi[j];
        for (int w = 1; w < P.nvar; ++w) { lo = std::min(lo, P.vars[w].ilo[j]); hi = std::max(hi, P.vars[w].ihi[j]); }
        total *= (hi - lo + 1);
    }
    for (int k = 1; k <= 8; ++k) start[k] = total;
    for (int k = 0; k <= 8; ++k) ab.put(start[k]);
    std::vector<unsigned char> sa((const unsigned char*)solve_args, (const unsigned char*)solve_args + nbytes);
    void* args[3] = {actx.b.data(), ab.b.data(), sa.data()};
    const int threads = (int)std::min<long long>(256, std::max<long long>(64, (total + 31) / 32 * 32));
    CUresult r = plan->drv.LaunchKernel(v->fn, 1, 1, 1, threads, 1, 1, 0, (CUstream)st, args, nullptr);
    if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "launch mol_solve_small: " + cu_err(plan->drv, r));
    plan->launches++;
    return MOL_OK;
}
}  // namespace mol

// ---- solution unpacking on the device (interface/solution/timedep.jl:30-72) ------------------------------
extern "C" int64_t mol_plan_grid_len(const mol_plan* plan, int64_t* nodes /*[ndim]*/) {
    if (!plan) return 0;
    int64_t n = 1;
    for (int j = 0; j < plan->P.ndim; ++j) {
        n *= plan->P.grid[j].n;
        if (nodes) nodes[j] = plan->P.grid[j].n;
    }
    return n;
}

extern "C" int mol_unpack(mol_plan* plan, double* full_dev, const double* u_dev, int nstates, const double* t_host,
                          const double* p_host, void* stream) {
    if (!plan || !full_dev || !u_dev || nstates < 1 || !t_host) return fail(MOL_E_ARG, "bad argument");
    if (plan->device < 0) return fail(MOL_E_NOCUDA, "plan was created compile-only (device = -1); there is no CPU fallback");
    if (plan->dist.on) return fail(MOL_E_UNSUPPORTED, "mol_unpack works on the whole grid: gather the slabs first");
    const Program& P = plan->P;
    if (p_host)
        for (int k = 0; k < P.nparam; ++k) plan->params[k] = p_host[k];
    MolVariant* vp = nullptr;
    int rcv = get_special_variant(plan, "unpack", "MOL_KERNEL_UNPACK=1", "mol_unpack_full", &vp);
    if (rcv != MOL_OK) return rcv;
    MolVariant& v = *vp;
    const int64_t nodes = mol_plan_grid_len(plan, nullptr);
    const int last = P.ndim - 1;
    for (int k = 0; k < nstates; ++k) {
        ArgBuf ain;
        ain.put(u_dev + (int64_t)k * P.nstate);
        ain.put(1.0);
        ArgBuf actx;
        actx.put(t_host[k]);
        for (int q = 0; q < std::max(1, P.nparam); ++q) actx.put(q < P.nparam ? plan->params[q] : 0.0);
        for (int j = 0; j < 3; ++j) actx.put((const double*)plan->d_grid[j]);
        actx.put((const double*)plan->d_tabw);
        actx.put((const int*)plan->d_tabs);
        actx.put((int)P.vars[0].ilo[last]);
        actx.put((int)P.vars[0].ihi[last]);
        actx.put((long long)0);
        double* out = full_dev + (int64_t)k * nodes * P.nvar;
        void* args[3] = {ain.b.data(), actx.b.data(), &out};
        const int grid = (int)std::min<int64_t>((nodes + 255) / 256, (int64_t)plan->sm_count * 16);
        CUresult r = plan->drv.LaunchKernel(v.fn, grid, 1, 1, 256, 1, 1, 0, (CUstream)stream, args, nullptr);
        if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "launch mol_unpack_full: " + cu_err(plan->drv, r));
        plan->launches++;
    }
    return MOL_OK;
}

// ---- Jacobian-vector product: jv = d/d(eps) f(u + eps v, p, t) at eps = 0 ----------------------------------
// The tiled kernel compiled on dual numbers (cooperative flavour, one input): u tiles, staged records, v tiles.
static int get_tiled_jvp_variant(mol_plan* plan, MolVariant** out) {
    const TileCfg& T = plan->G.tile;
    const char* key = "tiled_jvp";
    auto it = plan->variants.find(key);
    if (it == plan->variants.end()) {
        MolVariant v;
        v.key = key;
        v.tiled = true;
        v.smem = tile_smem_bytes(plan, false, MOL_EPI_NONE, false) + (size_t)plan->P.nvar * T.tile_stride_doubles * 8;
        // dual arithmetic roughly doubles the live values: two CTAs/SM (128 registers), fewer when the tiles do not fit
        v.min_ctas = (int)std::max<size_t>(1, std::min<size_t>(2, (200 * 1024) / std::max<size_t>(v.smem, 1)));
        std::string log;
        int rc = nvrtc_compile(plan->full_source, {"MOL_NIN=1", "MOL_EPI=0", "MOL_KERNEL_TILED=1", "MOL_TMA=0", "MOL_CPASYNC=0",
                                                   "MOL_KERNEL_JVP=1", "MOL_MIN_CTAS=" + std::to_string(v.min_ctas)},
                               v.cubin, log);
        if (rc != MOL_OK) return rc;
        plan->variants[v.key] = v;
        it = plan->variants.find(key);
    }
    MolVariant& v = it->second;
    if (plan->device >= 0 && !v.fn) {
        CUresult r = plan->drv.ModuleLoadData(&v.module, v.cubin.data());
        if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "cuModuleLoadData: " + cu_err(plan->drv, r));
        r = plan->drv.ModuleGetFunction(&v.fn, v.module, "mol_rhs_tiled");
        if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "cuModuleGetFunction(mol_rhs_tiled, J*v): " + cu_err(plan->drv, r));
        r = plan->drv.FuncSetAttribute(v.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)v.smem);
        if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "cuFuncSetAttribute(smem): " + cu_err(plan->drv, r));
        int nb = 1;
        plan->drv.OccupancyMaxActiveBlocksPerMultiprocessor(&nb, v.fn, T.nthreads, v.smem);
        v.grid_ctas = std::max(1, nb) * plan->sm_count;
    }
    *out = &v;
    return MOL_OK;
}

extern "C" int mol_jvp(mol_plan* plan, double* jv_dev, const double* u_dev, const double* v_dev, const double* p_host, double t,
                       void* stream) {
    if (!plan || !jv_dev || !u_dev || !v_dev) return fail(MOL_E_ARG, "null argument");
    if (plan->device < 0) return fail(MOL_E_NOCUDA, "plan was created compile-only (device = -1); there is no CPU fallback");
    if (plan->dist.on) return fail(MOL_E_UNSUPPORTED, "mol_jvp is not available in slab mode yet");
    const Program& P = plan->P;
    if (p_host)
        for (int k = 0; k < P.nparam; ++k) plan->params[k] = p_host[k];
    ArgBuf ain;
    ain.put(u_dev);
    ain.put(1.0);
    ArgBuf ajv;
    ajv.put(v_dev);
    ArgBuf actx;
    actx.put(t);
    for (int k = 0; k < std::max(1, P.nparam); ++k) actx.put(k < P.nparam ? plan->params[k] : 0.0);
    for (int j = 0; j < 3; ++j) actx.put((const double*)plan->d_grid[j]);
    actx.put((const double*)plan->d_tabw);
    actx.put((const int*)plan->d_tabs);
    const int last = P.ndim - 1;
    actx.put((int)P.vars[0].ilo[last]);
    actx.put((int)P.vars[0].ihi[last]);
    actx.put((long long)0);
    // the table-driven kernel on a list of boxes (MolBoxes in mol_generic.cuh)
    auto launch_generic_jvp = [&](const std::vector<std::vector<int>>& boxes) -> int {
        if (boxes.empty()) return MOL_OK;
        MolVariant* v = nullptr;
        int rc = get_special_variant(plan, "jvp", "MOL_KERNEL_JVP=1", "mol_jvp_generic", &v);
        if (rc != MOL_OK) return rc;
        const int MAXB = 8;
        for (size_t first = 0; first < boxes.size(); first += MAXB) {
            const int nb = (int)std::min<size_t>(MAXB, boxes.size() - first);
            ArgBuf ab;
            ab.put(nb);
            ab.put((int)0);
            long long start[MAXB + 1] = {0};
            for (int k = 0; k < MAXB; ++k) {
                int64_t tot = 0;
                if (k < nb) {
                    const std::vector<int>& b = boxes[first + k];
                    tot = 1;
                    for (int j = 0; j < P.ndim; ++j) tot *= std::max(0, b[3 + j] - b[j] + 1);
                    for (int q = 0; q < 6; ++q) ab.put(b[q]);
                } else {
                    for (int q = 0; q < 6; ++q) ab.put((int)(q < 3 ? 1 : 0));
                }
                start[k + 1] = start[k] + tot;
            }
            for (int k = 0; k <= MAXB; ++k) ab.put(start[k]);
            const int64_t total = start[nb];
            if (total <= 0) continue;
            void* args[5] = {ain.b.data(), ajv.b.data(), actx.b.data(), ab.b.data(), &jv_dev};
            const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)v->grid_ctas * 4);
            CUresult r = plan->drv.LaunchKernel(v->fn, grid, 1, 1, 256, 1, 1, 0, (CUstream)stream, args, nullptr);
            if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "launch mol_jvp_generic: " + cu_err(plan->drv, r));
            plan->launches++;
        }
        return MOL_OK;
    };
    const TileCfg& T = plan->G.tile;
    const char* ge = getenv("MOL_JVP_GENERIC");
    if (T.enabled && T.jvp && !T.zmarch && plan->kernel_mode == MOL_KERNEL_AUTO && !(ge && *ge && *ge != '0')) {
        // Tiled J*v: the core box through mol_rhs_tiled compiled on dual numbers (u tiles + v tiles in shared memory,
        // cooperative loader), the frame around it through the table-driven kernel
        MolVariant* v = nullptr;
        int rc = get_tiled_jvp_variant(plan, &v);
        if (rc != MOL_OK) return rc;
        if ((T.vec_store) && (reinterpret_cast<uintptr_t>(u_dev) % 16 != 0 || reinterpret_cast<uintptr_t>(v_dev) % 16 != 0 ||
                              reinterpret_cast<uintptr_t>(jv_dev) % 16 != 0))
            return fail(MOL_E_ARG, "state pointers must be 16-byte aligned (128-bit loads)");
        const int tdim[3] = {T.tx, T.ty, T.tz};
        ArgBuf at;
        int total = 0;
        for (int k = 0; k < 2; ++k) {
            int nt[3] = {1, 1, 1}, lo[3] = {1, 1, 1}, hi[3] = {0, 0, 0}, n = 0;
            if (k == 0) {
                n = 1;
                for (int j = 0; j < P.ndim; ++j) {
                    lo[j] = P.clo[j];
                    hi[j] = P.chi[j];
                    nt[j] = (hi[j] - lo[j] + 1 + tdim[j] - 1) / tdim[j];
                    n *= nt[j];
                }
                for (int j = P.ndim; j < 3; ++j) hi[j] = 1;
            }
            for (int j = 0; j < 3; ++j) at.put(nt[j]);
            at.put(n);
            for (int j = 0; j < 3; ++j) at.put(lo[j]);
            for (int j = 0; j < 3; ++j) at.put(hi[j]);
            total += n;
        }
        at.put(total);
        at.put((int)0);
        at.put((int*)plan->d_counter);
        at.put((const unsigned long long*)nullptr);
        at.put((const unsigned long long*)nullptr);
        at.put((unsigned long long)0);
        if (total > 0) {
            void* args[5] = {ain.b.data(), actx.b.data(), at.b.data(), &jv_dev, ajv.b.data()};
            const int grid = std::min(total, v->grid_ctas);
            CUresult r = plan->drv.LaunchKernel(v->fn, grid, 1, 1, T.nthreads, 1, 1, (unsigned)v->smem, (CUstream)stream, args, nullptr);
            if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "launch mol_rhs_tiled (J*v): " + cu_err(plan->drv, r));
            plan->launches++;
        }
        return launch_generic_jvp(plan->frame);
    }
    // one box: the union of the interior boxes of all variables
    std::vector<int> box(6, 1);
    for (int j = 0; j < P.ndim; ++j) {
        box[j] = P.vars[0].ilo[j];
        box[3 + j] = P.vars[0].ihi[j];
        for (int w = 1; w < P.nvar; ++w) { box[j] = std::min(box[j], P.vars[w].ilo[j]); box[3 + j] = std::max(box[3 + j], P.vars[w].ihi[j]); }
    }
    return launch_generic_jvp({box});
}

// ---- reference-facing call with HOST buffers -------------------------------------------------------------
// f!(du, u, p, t) on host arrays (what a CPU caller of the reference's generated function holds): the grid is cut into
// chunks of planes along the last dimension and H2D copy / stencil sweep / D2H copy of successive chunks overlap on
// three streams, so a call costs about max(H2D, D2H) over PCIe instead of H2D + sweep + D2H.
static void hostpipe_destroy(mol_plan* plan) {
    auto& H = plan->hp;
    if (!H.used) return;
    for (auto e : H.ev_in) cudaEventDestroy(e);
    for (auto e : H.ev_cmp) cudaEventDestroy(e);
    if (H.ev_free_u) cudaEventDestroy(H.ev_free_u);
    if (H.ev_free_du) cudaEventDestroy(H.ev_free_du);
    if (H.ev_out) cudaEventDestroy(H.ev_out);
    if (H.ev_start) cudaEventDestroy(H.ev_start);
    if (H.s_in) cudaStreamDestroy(H.s_in);
    if (H.s_out) cudaStreamDestroy(H.s_out);
    if (H.d_u) cudaFree(H.d_u);
    if (H.d_du) cudaFree(H.d_du);
    H = mol_plan::HostPipe();
}

extern "C" int mol_rhs_host(mol_plan* plan, double* du_host, const double* u_host, const double* p_host, double t,
                            int nchunks, void* stream) {
    if (!plan || !du_host || !u_host) return fail(MOL_E_ARG, "null argument");
    if (plan->device < 0) return fail(MOL_E_NOCUDA, "plan was created compile-only (device = -1); there is no CPU fallback");
    if (plan->dist.on) return fail(MOL_E_UNSUPPORTED, "mol_rhs_host is a single-device call; in slab mode keep the state resident");
    const Program& P = plan->P;
    cudaStream_t st = (cudaStream_t)stream;
    auto& H = plan->hp;
    const int last = P.ndim - 1;
    int lo = P.vars[0].ilo[last], hi = P.vars[0].ihi[last];
    bool same = true;
    for (int v = 1; v < P.nvar; ++v)
        if (P.vars[v].ilo[last] != lo || P.vars[v].ihi[last] != hi) same = false;
    const int64_t rows = (int64_t)hi - lo + 1;
    if (nchunks <= 0) {
        const char* e = getenv("MOL_HOST_CHUNKS");      // tuning override (experiments only)
        nchunks = (e && *e) ? std::max(1, atoi(e)) : 16;
    }
    if (!same || P.ndim == 1) nchunks = 1;                       // 1-D: one chunk (alignment of 128-bit stores)
    nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, rows / 16));
    cudaError_t e = cudaSuccess;
    if (!H.used) {
        e = cudaMalloc(&H.d_u, P.nstate * 8);
        if (e == cudaSuccess) e = cudaMalloc(&H.d_du, P.nstate * 8);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&H.s_in, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&H.s_out, cudaStreamNonBlocking);
        for (cudaEvent_t* ev : {&H.ev_free_u, &H.ev_free_du, &H.ev_out, &H.ev_start})
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        H.used = true;
        if (e != cudaSuccess) { hostpipe_destroy(plan); return fail(MOL_E_CUDA, std::string("mol_rhs_host setup: ") + cudaGetErrorString(e)); }
        // the staging buffers start out free
        cudaEventRecord(H.ev_free_u, st);
        cudaEventRecord(H.ev_free_du, st);
    }
    while ((int)H.ev_in.size() < nchunks) {
        cudaEvent_t a, b;
        if (cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess)
            return fail(MOL_E_CUDA, "mol_rhs_host: event creation failed");
        H.ev_in.push_back(a);
        H.ev_cmp.push_back(b);
    }
    if (p_host)
        for (int k = 0; k < P.nparam; ++k) plan->params[k] = p_host[k];
    // planes [r0, r1) of chunk c; plane length per variable
    auto r0 = [&](int c) { return rows * c / nchunks; };
    std::vector<int64_t> plane(P.nvar, 1);
    for (int v = 0; v < P.nvar; ++v)
        for (int j = 0; j < last; ++j) plane[v] *= P.vars[v].ext(j);
    auto copy_rows = [&](bool in, int64_t a, int64_t b, cudaStream_t s) {
        for (int v = 0; v < P.nvar && e == cudaSuccess; ++v) {
            // variables with their own plane range (nchunks == 1): copy the whole variable
            const int64_t off = P.voff[v] + (same ? a * plane[v] : 0);
            const int64_t len = same ? (b - a) * plane[v] : (int64_t)plane[v] * P.vars[v].ext(last);
            if (in) e = cudaMemcpyAsync(H.d_u + off, u_host + off, len * 8, cudaMemcpyHostToDevice, s);
            else e = cudaMemcpyAsync(du_host + off, H.d_du + off, len * 8, cudaMemcpyDeviceToHost, s);
        }
    };
    // order after everything already queued on the caller's stream, and after the previous call released the buffers
    cudaEventRecord(H.ev_start, st);
    cudaStreamWaitEvent(H.s_in, H.ev_start, 0);
    cudaStreamWaitEvent(H.s_in, H.ev_free_u, 0);
    cudaStreamWaitEvent(st, H.ev_free_du, 0);
    const int reach = 8;        // planes of the neighbouring chunks a chunk may read (>= any stencil / one-sided row reach)
    for (int c = 0; c < nchunks && e == cudaSuccess; ++c) {
        copy_rows(true, r0(c), r0(c + 1), H.s_in);
        if (c == 0 && nchunks > 1) copy_rows(true, rows - reach, rows, H.s_in);      // periodic wrap / far-edge rows
        if (e == cudaSuccess) e = cudaEventRecord(H.ev_in[c], H.s_in);
    }
    int rc = MOL_OK;
    MolRhsIn in;
    in.nin = 1;
    in.a[0] = H.d_u;
    in.c[0] = 1.0;
    MolRhsEpi epi;
    std::vector<int> B = {1, 1, 1, 1, 1, 1};
    for (int j = 0; j < P.ndim; ++j) {
        B[j] = P.vars[0].ilo[j];
        B[3 + j] = P.vars[0].ihi[j];
        for (int v = 1; v < P.nvar; ++v) { B[j] = std::min(B[j], P.vars[v].ilo[j]); B[3 + j] = std::max(B[3 + j], P.vars[v].ihi[j]); }
    }
    const bool tiled = plan->G.tile.enabled && plan->kernel_mode == MOL_KERNEL_AUTO;
    for (int c = 0; c < nchunks && e == cudaSuccess && rc == MOL_OK; ++c) {
        cudaStreamWaitEvent(st, H.ev_in[std::min(c + 1, nchunks - 1)], 0);
        if (nchunks > 1) {
            std::vector<int> Bc = B;
            Bc[last] = lo + (int)r0(c);
            Bc[3 + last] = lo + (int)r0(c + 1) - 1;
            plan->ov_on = true;
            plan->ov_tile.clear();
            plan->ov_frame.clear();
            if (tiled) {
                std::vector<int> Tc = {P.clo[0], P.clo[1], P.clo[2], P.chi[0], P.chi[1], P.chi[2]};
                Tc[last] = std::max(Tc[last], Bc[last]);
                Tc[3 + last] = std::min(Tc[3 + last], Bc[3 + last]);
                if (!box_empty(Tc, P.ndim)) { plan->ov_tile = Tc; peel(Bc, Tc, P.ndim, plan->ov_frame); }
                else plan->ov_frame.push_back(Bc);
            } else plan->ov_frame.push_back(Bc);
        }
        rc = mol_rhs_launch(plan, in, H.d_du, t, epi, st);
        plan->ov_on = false;
        if (rc != MOL_OK) break;
        e = cudaEventRecord(H.ev_cmp[c], st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(H.s_out, H.ev_cmp[c], 0);
        if (e == cudaSuccess) copy_rows(false, r0(c), r0(c + 1), H.s_out);
    }
    if (rc != MOL_OK) return rc;
    if (e == cudaSuccess) e = cudaEventRecord(H.ev_free_u, st);          // last sweep done: d_u may be overwritten
    if (e == cudaSuccess) e = cudaEventRecord(H.ev_out, H.s_out);
    if (e == cudaSuccess) e = cudaEventRecord(H.ev_free_du, H.s_out);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, H.ev_out, 0);      // the caller's stream completes when du_host is complete
    if (e != cudaSuccess) return fail(MOL_E_CUDA, std::string("mol_rhs_host: ") + cudaGetErrorString(e));
    return MOL_OK;
}

// ---- a1: Fornberg weights, operation order of fornberg_calculate_weights.jl:20-67 ---------------
extern "C" int mol_fd_weights(int order, double x0, const double* x, int n, double* w_out) {
    if (!x || !w_out || n < 1 || order < 0) return fail(MOL_E_ARG, "bad argument");
    if (order >= n) return fail(MOL_E_ARG, "Not enough points for the requested order.");
    const int M = order;
    std::vector<double> C((size_t)n * (M + 1), 0.0);
    auto at = [&](int i, int s) -> double& { return C[(size_t)i * (M + 1) + s]; };
    double c1 = 1.0, c4 = x[0] - x0;
    at(0, 0) = 1.0;
    for (int i = 1; i < n; ++i) {
        const int mn = std::min(i, M);
        double c2 = 1.0;
        const double c5 = c4;
        c4 = x[i] - x0;
        for (int j = 0; j < i; ++j) {
            const double c3 = x[i] - x[j];
            c2 *= c3;
            if (j == i - 1) {
                for (int s = mn; s >= 1; --s) at(i, s) = c1 * (s * at(i - 1, s - 1) - c5 * at(i - 1, s)) / c2;
                at(i, 0) = -c1 * c5 * at(i - 1, 0) / c2;
            }
            for (int s = mn; s >= 1; --s) at(j, s) = (c4 * at(j, s) - s * at(j, s - 1)) / c3;
            at(j, 0) = c4 * at(j, 0) / c3;
        }
        c1 = c2;
    }
    double sum = 0.0;
    for (int i = 0; i < n; ++i) { w_out[i] = at(i, M); sum += w_out[i]; }
    if (order != 0) w_out[n / 2] -= sum;     // the reference's sum-to-zero fix (:62-65)
    return MOL_OK;
}

// one row of weights per node: the per-node loops that build non-uniform tables (centered_diff_weights.jl:94-103,
// upwind_diff_weights.jl:107-133, half_offset_weights.jl:94-120) in one call
extern "C" int mol_fd_weights_rows(int order, int64_t nrows, int n, const double* x0, const double* x, double* w_out) {
    if (!x0 || !x || !w_out || nrows < 0) return fail(MOL_E_ARG, "bad argument");
    for (int64_t r = 0; r < nrows; ++r) {
        const int rc = mol_fd_weights(order, x0[r], x + r * n, n, w_out + r * n);
        if (rc != MOL_OK) return rc;
    }
    return MOL_OK;
}

extern "C" int mol_rhs_part(mol_plan* plan, double* du_dev, const double* u_dev, const double* p_host, double t, int part,
                            void* stream) {
    if (!plan || !du_dev || !u_dev) return fail(MOL_E_ARG, "null argument");
    if (part != MOL_PART_INTERIOR && part != MOL_PART_BOUNDARY) return fail(MOL_E_ARG, "part must be MOL_PART_INTERIOR or MOL_PART_BOUNDARY");
    if (!plan->dist.on) return fail(MOL_E_ARG, "mol_rhs_part needs mol_dist_init first");
    if (p_host)
        for (int k = 0; k < plan->P.nparam; ++k) plan->params[k] = p_host[k];
    MolRhsIn in;
    in.nin = 1;
    in.a[0] = u_dev;
    in.c[0] = 1.0;
    MolRhsEpi epi;
    return mol_rhs_launch(plan, in, du_dev, t, epi, (cudaStream_t)stream, part);
}
