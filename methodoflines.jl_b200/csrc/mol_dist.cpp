// Slab decomposition across ranks: one process per GPU, the grid split along the
// slowest-varying (last) spatial dimension, H ghost planes per side exchanged with the two
// neighbouring ranks once per RHS evaluation — ncclSend/ncclRecv inside one group on a private
// stream, overlapped with the interior part of the stencil sweep (mol_rhs_launch).  A periodic
// split dimension closes the ring (rank 0 <-> rank P-1); at a non-periodic domain edge the owning
// rank applies the boundary rule instead of receiving.  The reference has no multi-process path;
// this is the B200-side extension the north star specifies.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy the host process already loaded,
// e.g. the one bundled with torch or NCCL_jll) so libmol_cuda.so has no link-time dependency on it.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "mol_internal.h"

namespace mol {

// ---- minimal NCCL surface (ABI-stable since 2.x) ---------------------------------------------------
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { kNcclFloat64 = 8, kNcclSum = 0 };

struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

static Nccl* get_nccl(std::string& err) {
    static Nccl N;
    if (N.h) return &N;
    const char* env = getenv("MOL_NCCL_PATH");
    const char* cands[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* c : cands) {
        if (!c) continue;
        N.h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (N.h) break;
    }
    if (!N.h) { err = "cannot dlopen libnccl.so.2 (set MOL_NCCL_PATH)"; return nullptr; }
#define SYM(field, name)                                                \
    *(void**)(&N.field) = dlsym(N.h, name);                             \
    if (!N.field) { err = std::string("libnccl lacks ") + name; N.h = nullptr; return nullptr; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return &N;
}

static int nccl_fail(Nccl* N, int rc, const char* what) {
    return fail(MOL_E_CUDA, std::string(what) + ": " + (N && N->GetErrorString ? N->GetErrorString(rc) : "NCCL error"));
}

static int cuda_fail(cudaError_t e, const char* what) {
    return fail(MOL_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// planes of the split dimension owned by `rank`: as even as possible, the first `rem` ranks one more
void slab_range(int glo, int ghi, int rank, int nranks, int* lo, int* hi) {
    const int64_t total = (int64_t)ghi - glo + 1;
    const int64_t base = total / nranks, rem = total % nranks;
    const int64_t start = (int64_t)rank * base + std::min<int64_t>(rank, rem);
    const int64_t len = base + (rank < rem ? 1 : 0);
    *lo = glo + (int)start;
    *hi = glo + (int)(start + len) - 1;
}

// largest distance (in planes) between a node and the taps of its stencil rows along `dim`,
// ignoring one-sided rows anchored at a domain edge (those only reach inward on the edge rank)
static int row_reach(const Program& P, int dim) {
    int reach = 0;
    auto tab_reach = [&](const Tab& T, int n, int shift_lo, int shift_hi) {
        int r = 0;
        for (int k = 0; k < T.nrows; ++k) {
            const int idx = T.first + k;
            const std::vector<double>& w = T.rows[k].w;
            int a = 0, b = (int)w.size() - 1;                             // first / last non-zero weight
            while (a <= b && w[a] == 0.0) ++a;
            while (b >= a && w[b] == 0.0) --b;
            if (a > b) continue;
            const int st = T.rows[k].start + a, en = T.rows[k].start + b;
            if (!T.has_core || idx < T.core_lo || idx > T.core_hi) {
                if (T.rows[k].start <= 1 || T.rows[k].start + (int)w.size() - 1 >= n) continue;   // anchored one-sided row
            }
            r = std::max(r, std::max(idx + shift_lo - st, en - idx - shift_hi));
        }
        return r;
    };
    const int n = P.grid[dim].n;
    for (const Rpn& eq : P.eqs) {
        for (const std::string& tk : eq) {
            int a = 0, b = 0, c = 0, d = 0, e = 0, f = 0;
            if (sscanf(tk.c_str(), "L:%d:%d:%d", &a, &b, &c) == 3 && tk[0] == 'L') {
                if (c == dim && P.tabs.count(a)) reach = std::max(reach, tab_reach(P.tabs.at(a), n, 0, 0));
            } else if (tk[0] == 'W' && sscanf(tk.c_str(), "W:%d:%d:%d", &a, &b, &c) == 3) {
                if (c == dim) reach = std::max(reach, 2);
            } else if (tk[0] == 'M' && sscanf(tk.c_str(), "M:%d:%d:%d:%d:%d", &a, &b, &c, &d, &e) == 5) {
                // mixed derivative: the first-derivative row that lies along the split dimension sets the reach
                if (d == dim && P.tabs.count(a)) reach = std::max(reach, tab_reach(P.tabs.at(a), n, 0, 0));
                if (e == dim && P.tabs.count(b)) reach = std::max(reach, tab_reach(P.tabs.at(b), n, 0, 0));
            } else if (tk[0] == 'N' && sscanf(tk.c_str(), "N:%d:%d:%d:%d:%d:%d", &a, &b, &c, &d, &e, &f) == 6) {
                if (b == dim && P.tabs.count(d) && P.tabs.count(e) && P.tabs.count(f)) {
                    // node -> half points (outer row) -> nodes (interpolation / derivative rows)
                    const int ro = tab_reach(P.tabs.at(f), n, 0, 0);
                    const int ri = std::max(tab_reach(P.tabs.at(d), n, 0, 0), tab_reach(P.tabs.at(e), n, 0, 0));
                    reach = std::max(reach, ro + ri + 1);
                }
            }
        }
    }
    return reach;
}

void dist_destroy(mol_plan* plan) {
    MolDist& D = plan->dist;
    if (!D.on) return;
    for (auto& kv : D.halos)
        if (kv.second.owned) { cudaFree(kv.second.lo); cudaFree(kv.second.hi); }
    D.halos.clear();
    if (D.scratch.owned) { cudaFree(D.scratch.lo); cudaFree(D.scratch.hi); }
    if (D.p2p.on) {
        if (D.p2p.prev_pool) cudaIpcCloseMemHandle(D.p2p.prev_pool);
        if (D.p2p.next_pool && D.p2p.next_pool != D.p2p.prev_pool) cudaIpcCloseMemHandle(D.p2p.next_pool);
        D.p2p.on = false;
    }
    if (D.p2p.pool) { cudaFree(D.p2p.pool); D.p2p.pool = nullptr; }
    if (D.comm) {
        std::string err;
        Nccl* N = get_nccl(err);
        if (N) N->CommDestroy(D.comm);
        D.comm = nullptr;
    }
    if (D.ev_ready) cudaEventDestroy(D.ev_ready);
    if (D.ev_done) cudaEventDestroy(D.ev_done);
    if (D.comm_stream) cudaStreamDestroy(D.comm_stream);
    D.on = false;
}

static size_t halo_doubles(const mol_plan* plan) {
    return (size_t)plan->P.nvar * plan->dist.H * plan->dist.plane_max;
}

// one grouped exchange of the first / last H planes of every variable of `arr`
static int exchange(mol_plan* plan, Nccl* N, const double* arr, MolHalo& h) {
    MolDist& D = plan->dist;
    const size_t cnt = (size_t)D.H * D.plane;
    const int nv = plan->P.nvar;
    int rc;
    // Posting order matters when prev == next (two ranks, periodic): the k-th send to a peer pairs
    // with that peer's k-th receive, so "top planes -> next" must meet "lower ghosts <- prev" first.
    for (int v = 0; v < nv; ++v)
        if (D.next >= 0 && (rc = N->Send(arr + (int64_t)v * D.vstride + (D.rows - D.H) * D.plane, cnt, kNcclFloat64, D.next, D.comm, D.comm_stream)))
            return nccl_fail(N, rc, "ncclSend");
    for (int v = 0; v < nv; ++v)
        if (D.prev >= 0 && (rc = N->Send(arr + (int64_t)v * D.vstride, cnt, kNcclFloat64, D.prev, D.comm, D.comm_stream)))
            return nccl_fail(N, rc, "ncclSend");
    for (int v = 0; v < nv; ++v)
        if (D.prev >= 0 && (rc = N->Recv(h.lo + (int64_t)v * D.H * D.plane_max, cnt, kNcclFloat64, D.prev, D.comm, D.comm_stream)))
            return nccl_fail(N, rc, "ncclRecv");
    for (int v = 0; v < nv; ++v)
        if (D.next >= 0 && (rc = N->Recv(h.hi + (int64_t)v * D.H * D.plane_max, cnt, kNcclFloat64, D.next, D.comm, D.comm_stream)))
            return nccl_fail(N, rc, "ncclRecv");
    return MOL_OK;
}

// Resolves the ghost-plane buffers of every input array.  With `exchanging` non-null the library is
// the transport: stale planes are exchanged on the private stream (after everything already queued
// on `st`), and *exchanging tells the caller to wait on ev_done before the boundary part.
static inline size_t p2p_off(const MolP2P& X, int slot, unsigned long long q, int side) {
    return (((size_t)slot * 2 + (size_t)(q & 1)) * 2 + side) * X.halo_bytes;
}

// push the edge planes of `arr` into the neighbours' pools (copy engines), bump their flags, wait for ours
// wait_on_stream = false: the caller's tiled kernel waits for the flags itself (fused ghost-plane wait)
static int p2p_exchange(mol_plan* plan, const double* arr, int slot, bool wait_on_stream) {
    MolDist& D = plan->dist;
    MolP2P& X = D.p2p;
    const unsigned long long q = ++X.seq[slot];
    const size_t width = (size_t)D.H * D.plane * 8, spitch = (size_t)D.vstride * 8, dpitch = (size_t)D.H * D.plane_max * 8;
    const int nv = plan->P.nvar;
    cudaError_t e = cudaSuccess;
    CUresult r = CUDA_SUCCESS;
    if (D.next >= 0) {      // my top planes are the neighbour's lower ghosts
        e = cudaMemcpy2DAsync(X.next_pool + p2p_off(X, slot, q, 0), dpitch, arr + (D.rows - D.H) * D.plane, spitch, width, nv,
                              cudaMemcpyDeviceToDevice, D.comm_stream);
        if (e == cudaSuccess)
            r = X.WriteValue64((CUstream)D.comm_stream, (CUdeviceptr)(X.next_pool + X.flags_off + ((size_t)slot * 2 + 0) * 8), q, 0);
    }
    if (e == cudaSuccess && r == CUDA_SUCCESS && D.prev >= 0) {      // my bottom planes are the neighbour's upper ghosts
        e = cudaMemcpy2DAsync(X.prev_pool + p2p_off(X, slot, q, 1), dpitch, arr, spitch, width, nv, cudaMemcpyDeviceToDevice,
                              D.comm_stream);
        if (e == cudaSuccess)
            r = X.WriteValue64((CUstream)D.comm_stream, (CUdeviceptr)(X.prev_pool + X.flags_off + ((size_t)slot * 2 + 1) * 8), q, 0);
    }
    if (wait_on_stream && e == cudaSuccess && r == CUDA_SUCCESS && D.prev >= 0)
        r = X.WaitValue64((CUstream)D.comm_stream, (CUdeviceptr)(X.pool + X.flags_off + ((size_t)slot * 2 + 0) * 8), q,
                          CU_STREAM_WAIT_VALUE_GEQ);
    if (wait_on_stream && e == cudaSuccess && r == CUDA_SUCCESS && D.next >= 0)
        r = X.WaitValue64((CUstream)D.comm_stream, (CUdeviceptr)(X.pool + X.flags_off + ((size_t)slot * 2 + 1) * 8), q,
                          CU_STREAM_WAIT_VALUE_GEQ);
    if (e != cudaSuccess) return cuda_fail(e, "peer-to-peer ghost-plane push");
    if (r != CUDA_SUCCESS) return fail(MOL_E_CUDA, "stream memory operation failed in the ghost-plane exchange");
    return MOL_OK;
}

// Resolves the ghost-plane buffers of every input array.  With `exchanging` non-null the library is
// the transport: stale planes are exchanged on the private stream (after everything already queued
// on `st`), and *exchanging tells the caller to wait on ev_done before the boundary part.
// `fuse` (in/out, may be null): on entry *fuse->want says the caller can wait for the flags inside its kernel; if exactly
// one array is exchanged over the peer-to-peer transport, the stream-side waits are left out and its flag addresses and
// sequence number are returned (fuse->on = true).
int dist_prepare_halos(mol_plan* plan, const MolRhsIn& in, const double** hlo, const double** hhi, cudaStream_t st,
                       bool* exchanging, MolFuse* fuse) {
    MolDist& D = plan->dist;
    MolP2P& X = D.p2p;
    MolHalo* hs[8];
    bool any_stale = false;
    int nscratch = 0;
    const bool lib_transport = exchanging != nullptr;
    for (int j = 0; j < in.nin; ++j) {
        auto it = D.halos.find(in.a[j]);
        if (it != D.halos.end()) hs[j] = &it->second;
        else {
            hs[j] = &D.scratch;
            D.scratch.fresh = false;
            if (++nscratch > 1) return fail(MOL_E_ARG, "only one unregistered array per RHS evaluation in slab mode");
        }
        if (!(lib_transport && X.on) && (!hs[j]->lo || !hs[j]->hi))
            return fail(MOL_E_ARG, "ghost-plane buffers missing (mol_dist_set_halo / mol_dist_comm_init)");
        if (!hs[j]->fresh) any_stale = true;
    }
    const bool have_peers = D.prev >= 0 || D.next >= 0;
    if (lib_transport) {
        *exchanging = false;
        if (any_stale && have_peers) {
            if (!D.comm) return fail(MOL_E_ARG, "no transport: call mol_dist_comm_init, or move the ghost planes yourself and use mol_rhs_part");
            cudaError_t e = cudaEventRecord(D.ev_ready, st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(D.comm_stream, D.ev_ready, 0);
            if (e != cudaSuccess) return cuda_fail(e, "ghost-plane exchange (stream order)");
            int rc;
            if (X.on) {
                int nstale = 0;
                for (int j = 0; j < in.nin; ++j) nstale += hs[j]->fresh ? 0 : 1;
                const bool fused = fuse && fuse->want && nstale == 1;
                for (int j = 0; j < in.nin; ++j) {
                    if (hs[j]->fresh) continue;
                    if ((rc = p2p_exchange(plan, in.a[j], hs[j]->slot, !fused)) != MOL_OK) return rc;
                    if (fused) {
                        const int slot = hs[j]->slot;
                        fuse->on = true;
                        fuse->seq = X.seq[slot];
                        fuse->flag[0] = D.prev >= 0 ? reinterpret_cast<const unsigned long long*>(X.pool + X.flags_off + ((size_t)slot * 2 + 0) * 8) : nullptr;
                        fuse->flag[1] = D.next >= 0 ? reinterpret_cast<const unsigned long long*>(X.pool + X.flags_off + ((size_t)slot * 2 + 1) * 8) : nullptr;
                    }
                    hs[j]->fresh = (hs[j] != &D.scratch);
                }
            } else {
                std::string err;
                Nccl* N = get_nccl(err);
                if (!N) return fail(MOL_E_CUDA, err);
                if ((rc = N->GroupStart())) return nccl_fail(N, rc, "ncclGroupStart");
                for (int j = 0; j < in.nin; ++j) {
                    if (hs[j]->fresh) continue;
                    if ((rc = exchange(plan, N, in.a[j], *hs[j])) != MOL_OK) { N->GroupEnd(); return rc; }
                    hs[j]->fresh = (hs[j] != &D.scratch);
                }
                if ((rc = N->GroupEnd())) return nccl_fail(N, rc, "ncclGroupEnd");
            }
            *exchanging = true;      // mol_rhs_launch queues the boundary part behind it and records ev_done
        }
    }
    for (int j = 0; j < in.nin; ++j) {
        if (lib_transport && X.on) {       // current parity of the slot
            hlo[j] = reinterpret_cast<const double*>(X.pool + p2p_off(X, hs[j]->slot, X.seq[hs[j]->slot], 0));
            hhi[j] = reinterpret_cast<const double*>(X.pool + p2p_off(X, hs[j]->slot, X.seq[hs[j]->slot], 1));
        } else {
            hlo[j] = hs[j]->lo;
            hhi[j] = hs[j]->hi;
        }
    }
    return MOL_OK;
}

void dist_mark_stale(mol_plan* plan, const double* arr) {
    auto it = plan->dist.halos.find(arr);
    if (it != plan->dist.halos.end()) it->second.fresh = false;
}

int dist_allreduce_sum(mol_plan* plan, double* dev, int n, cudaStream_t st) {
    MolDist& D = plan->dist;
    if (!D.on || D.nranks == 1) return MOL_OK;
    if (!D.comm) return fail(MOL_E_ARG, "all-reduce needs mol_dist_comm_init");
    std::string err;
    Nccl* N = get_nccl(err);
    if (!N) return fail(MOL_E_CUDA, err);
    int rc = N->AllReduce(dev, dev, (size_t)n, kNcclFloat64, kNcclSum, D.comm, st);
    return rc ? nccl_fail(N, rc, "ncclAllReduce") : MOL_OK;
}

}  // namespace mol

using namespace mol;

namespace mol {
// One pool per rank, mapped by both neighbours.  Bootstrap: IPC handles are all-gathered over NCCL.
static int p2p_setup(mol_plan* plan, Nccl* N) {
    MolDist& D = plan->dist;
    MolP2P& X = D.p2p;
    if (D.prev < 0 && D.next < 0) return MOL_OK;
    // Every rank takes part in BOTH collectives below whatever happens locally (a rank that returned early would leave
    // the others blocked in them): local failures only clear `ok`, and the all-reduced flag decides for everybody.
    int ok = 1;
    std::string why;
    void* f1 = nullptr;
    void* f2 = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &f1, cudaEnableDefault, &qr) != cudaSuccess || !f1 ||
        cudaGetDriverEntryPoint("cuStreamWaitValue64", &f2, cudaEnableDefault, &qr) != cudaSuccess || !f2) {
        cudaGetLastError();
        ok = 0;
        why = "stream memory operations are not available";
    }
    *(void**)(&X.WriteValue64) = f1;
    *(void**)(&X.WaitValue64) = f2;
    X.halo_bytes = (halo_doubles(plan) * 8 + 255) / 256 * 256;
    X.flags_off = (size_t)MOL_P2P_SLOTS * 4 * X.halo_bytes;
    const size_t total = X.flags_off + (size_t)MOL_P2P_SLOTS * 2 * 8;
    cudaError_t e = ok ? cudaMalloc(&X.pool, total) : cudaSuccess;
    if (ok && e == cudaSuccess) e = cudaMemset(X.pool, 0, total);
    if (e != cudaSuccess) { cudaGetLastError(); ok = 0; why = std::string("ghost-plane pool: ") + cudaGetErrorString(e); X.pool = nullptr; }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok) {
        e = cudaIpcGetMemHandle(&mine, X.pool);
        if (e != cudaSuccess) { cudaGetLastError(); ok = 0; why = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); }
    }
    // all-gather the handles (device staging buffer, NCCL as the bootstrap channel)
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    char* d_all = nullptr;
    std::vector<char> all(hb * D.nranks);
    e = cudaMalloc(&d_all, hb * D.nranks);
    if (e == cudaSuccess) e = cudaMemcpy(d_all + hb * D.rank, &mine, hb, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "IPC handle staging");        // out of memory for 64 bytes per rank: nothing sensible left
    int rc = N->AllGather(d_all + hb * D.rank, d_all, hb, /*ncclInt8*/ 0, D.comm, D.comm_stream);
    if (rc) { cudaFree(d_all); return nccl_fail(N, rc, "ncclAllGather (IPC handles)"); }
    e = cudaStreamSynchronize(D.comm_stream);
    if (e == cudaSuccess) e = cudaMemcpy(all.data(), d_all, hb * D.nranks, cudaMemcpyDeviceToHost);
    cudaFree(d_all);
    if (e != cudaSuccess) return cuda_fail(e, "IPC handle exchange");
    auto open_peer = [&](int peer, char** out) -> int {
        cudaIpcMemHandle_t h;
        memcpy(&h, all.data() + hb * peer, hb);
        void* p = nullptr;
        cudaError_t ee = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (ee != cudaSuccess) { cudaGetLastError(); return fail(MOL_E_UNSUPPORTED, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(ee)); }
        *out = (char*)p;
        return MOL_OK;
    };
    // every rank must take the same decision: all-reduce a success flag before switching transports
    if (ok && D.prev >= 0 && open_peer(D.prev, &X.prev_pool) != MOL_OK) ok = 0;
    if (ok && D.next >= 0) {
        if (D.next == D.prev) X.next_pool = X.prev_pool;
        else if (open_peer(D.next, &X.next_pool) != MOL_OK) ok = 0;
    }
    double* d_ok = nullptr;
    double h_ok = ok ? 0.0 : 1.0;
    e = cudaMalloc(&d_ok, 8);
    if (e == cudaSuccess) e = cudaMemcpy(d_ok, &h_ok, 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "p2p agreement");
    rc = N->AllReduce(d_ok, d_ok, 1, kNcclFloat64, kNcclSum, D.comm, D.comm_stream);
    if (rc) { cudaFree(d_ok); return nccl_fail(N, rc, "ncclAllReduce (p2p agreement)"); }
    cudaStreamSynchronize(D.comm_stream);
    cudaMemcpy(&h_ok, d_ok, 8, cudaMemcpyDeviceToHost);
    cudaFree(d_ok);
    if (h_ok != 0.0) {      // some rank could not map: nobody switches; release what this rank had opened
        if (X.prev_pool) cudaIpcCloseMemHandle(X.prev_pool);
        if (X.next_pool && X.next_pool != X.prev_pool) cudaIpcCloseMemHandle(X.next_pool);
        X.prev_pool = X.next_pool = nullptr;
        if (X.pool) { cudaFree(X.pool); X.pool = nullptr; }
        return fail(MOL_E_UNSUPPORTED, "peer-to-peer mapping failed on at least one rank; using NCCL send/recv" +
                                           (why.empty() ? std::string() : " (" + why + ")"));
    }
    X.on = true;
    X.used[0] = true;                 // slot 0: unregistered (caller-owned) arrays
    D.scratch.slot = 0;
    for (auto& kv : D.halos) {        // arrays registered before the transport came up
        for (int sidx = 1; sidx < MOL_P2P_SLOTS; ++sidx)
            if (!X.used[sidx]) { X.used[sidx] = true; kv.second.slot = sidx; break; }
    }
    return MOL_OK;
}
}  // namespace mol

extern "C" int mol_dist_partition(int64_t n_planes, int nranks, int rank, int64_t* first, int64_t* count) {
    if (n_planes < 1 || nranks < 1 || rank < 0 || rank >= nranks || !first || !count) return fail(MOL_E_ARG, "bad argument");
    int lo, hi;
    slab_range(0, (int)n_planes - 1, rank, nranks, &lo, &hi);
    *first = lo;
    *count = (int64_t)hi - lo + 1;
    return MOL_OK;
}

extern "C" int mol_dist_init(mol_plan* plan, int rank, int nranks) {
    if (!plan) return fail(MOL_E_ARG, "null plan");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(MOL_E_ARG, "bad rank / nranks");
    if (plan->dist.on) return fail(MOL_E_ARG, "mol_dist_init was already called on this plan");
    if (nranks == 1) return MOL_OK;
    const Program& P = plan->P;
    if (P.ndim < 2) return fail(MOL_E_UNSUPPORTED, "slab decomposition needs at least 2 spatial dimensions");
    for (int v = 1; v < P.nvar; ++v)
        for (int j = 0; j < P.ndim; ++j)
            if (P.vars[v].ilo[j] != P.vars[0].ilo[j] || P.vars[v].ihi[j] != P.vars[0].ihi[j] || P.vars[v].per[j] != P.vars[0].per[j])
                return fail(MOL_E_UNSUPPORTED, "slab decomposition needs all variables on the same interior box");
    MolDist D;
    D.rank = rank;
    D.nranks = nranks;
    D.split = P.ndim - 1;
    D.periodic = P.vars[0].per[D.split] != 0;
    D.H = std::max(1, row_reach(P, D.split));
    if (plan->G.tile.enabled) D.H = std::max(D.H, plan->G.tile.r[D.split]);
    D.glo = P.vars[0].ilo[D.split];
    D.ghi = P.vars[0].ihi[D.split];
    slab_range(D.glo, D.ghi, rank, nranks, &D.loc_lo, &D.loc_hi);
    D.rows = (int64_t)D.loc_hi - D.loc_lo + 1;
    int minrows = (int)((int64_t)(D.ghi - D.glo + 1) / nranks);
    if (minrows < std::max(2 * D.H, 8))
        return fail(MOL_E_ARG, "too few planes per rank along the split dimension for this stencil");
    D.plane = 1;
    for (int j = 0; j < D.split; ++j) D.plane *= P.vars[0].ext(j);
    D.plane_max = D.plane;
    D.vstride = D.rows * D.plane;
    D.nstate_local = D.vstride * P.nvar;
    D.nstate_global = P.nstate;
    D.prev = rank > 0 ? rank - 1 : (D.periodic ? nranks - 1 : -1);
    D.next = rank < nranks - 1 ? rank + 1 : (D.periodic ? 0 : -1);
    D.on = true;
    plan->dist = D;
    for (auto& m : plan->mapsets) m.ptr = nullptr;
    compute_frame(plan);
    return MOL_OK;
}

extern "C" int mol_dist_info(const mol_plan* plan, mol_dist_info_t* out) {
    if (!plan || !out) return fail(MOL_E_ARG, "null argument");
    const MolDist& D = plan->dist;
    memset(out, 0, sizeof *out);
    out->rank = D.rank;
    out->nranks = D.on ? D.nranks : 1;
    out->halo_planes = D.H;
    out->plane_len = D.on ? D.plane : 0;
    out->first_plane = D.on ? D.loc_lo - D.glo : 0;
    out->n_planes = D.on ? D.rows : (plan->P.vars[0].ext(plan->P.ndim - 1));
    out->state_len_local = D.on ? D.nstate_local : plan->P.nstate;
    out->state_len_global = plan->P.nstate;
    out->halo_len = D.on ? (int64_t)halo_doubles(plan) : 0;
    out->prev_rank = D.on ? D.prev : -1;
    out->next_rank = D.on ? D.next : -1;
    out->periodic = D.periodic ? 1 : 0;
    return MOL_OK;
}

extern "C" int mol_dist_set_halo(mol_plan* plan, double* lo_dev, double* hi_dev) {
    if (!plan || !lo_dev || !hi_dev) return fail(MOL_E_ARG, "null argument");
    MolDist& D = plan->dist;
    if (!D.on) return fail(MOL_E_ARG, "mol_dist_init first");
    if (D.scratch.owned) { cudaFree(D.scratch.lo); cudaFree(D.scratch.hi); }
    D.scratch.lo = lo_dev;
    D.scratch.hi = hi_dev;
    D.scratch.owned = false;
    D.scratch.fresh = false;
    return MOL_OK;
}

extern "C" int mol_dist_unique_id(void* id_out, size_t nbytes) {
    if (!id_out || nbytes < sizeof(NcclUniqueId)) return fail(MOL_E_ARG, "unique-id buffer must hold 128 bytes");
    std::string err;
    Nccl* N = get_nccl(err);
    if (!N) return fail(MOL_E_CUDA, err);
    NcclUniqueId id;
    int rc = N->GetUniqueId(&id);
    if (rc) return nccl_fail(N, rc, "ncclGetUniqueId");
    memcpy(id_out, &id, sizeof id);
    return MOL_OK;
}

extern "C" int mol_dist_comm_init(mol_plan* plan, const void* unique_id, size_t nbytes) {
    if (!plan || !unique_id || nbytes < sizeof(NcclUniqueId)) return fail(MOL_E_ARG, "bad argument");
    MolDist& D = plan->dist;
    if (!D.on) return fail(MOL_E_ARG, "mol_dist_init first");
    if (plan->device < 0) return fail(MOL_E_NOCUDA, "plan was created compile-only; there is no CPU fallback");
    if (D.comm) return fail(MOL_E_ARG, "communicator already initialised");
    std::string err;
    Nccl* N = get_nccl(err);
    if (!N) return fail(MOL_E_CUDA, err);
    cudaError_t e = cudaSetDevice(plan->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    NcclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    int rc = N->CommInitRank(&D.comm, D.nranks, id, D.rank);
    if (rc) { D.comm = nullptr; return nccl_fail(N, rc, "ncclCommInitRank"); }
    int lo_pri = 0, hi_pri = 0;
    cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
    e = cudaStreamCreateWithPriority(&D.comm_stream, cudaStreamNonBlocking, hi_pri);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&D.ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&D.ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess && !D.scratch.lo) {
        e = cudaMalloc(&D.scratch.lo, halo_doubles(plan) * 8);
        if (e == cudaSuccess) e = cudaMalloc(&D.scratch.hi, halo_doubles(plan) * 8);
        if (e == cudaSuccess) e = cudaMemset(D.scratch.lo, 0, halo_doubles(plan) * 8);
        if (e == cudaSuccess) e = cudaMemset(D.scratch.hi, 0, halo_doubles(plan) * 8);
        D.scratch.owned = true;
    }
    if (e != cudaSuccess) return cuda_fail(e, "mol_dist_comm_init");
    const char* tr = getenv("MOL_DIST_TRANSPORT");
    if (!(tr && !strcmp(tr, "nccl"))) {
        int rc2 = p2p_setup(plan, N);
        if (rc2 != MOL_OK && tr && !strcmp(tr, "p2p")) return rc2;      // explicitly requested: report why it failed
    }
    return MOL_OK;
}

extern "C" const char* mol_dist_transport(const mol_plan* plan) {
    if (!plan || !plan->dist.on) return "none";
    if (plan->dist.p2p.on) return "p2p (CUDA IPC pools, copy-engine push + stream memory-op flags over NVLink)";
    if (plan->dist.comm) return "nccl (ncclSend/ncclRecv)";
    return "external (caller moves the ghost planes)";
}

// Registers a library-side array (RK stage vector) so that it gets its own ghost planes, exchanged
// lazily: once after each time the array is rewritten.
extern "C" int mol_dist_register(mol_plan* plan, const double* arr_dev) {
    if (!plan || !arr_dev) return fail(MOL_E_ARG, "null argument");
    MolDist& D = plan->dist;
    if (!D.on) return MOL_OK;
    if (D.halos.count(arr_dev)) return MOL_OK;
    MolHalo h;
    cudaError_t e = cudaMalloc(&h.lo, halo_doubles(plan) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&h.hi, halo_doubles(plan) * 8);
    if (e == cudaSuccess) e = cudaMemset(h.lo, 0, halo_doubles(plan) * 8);
    if (e == cudaSuccess) e = cudaMemset(h.hi, 0, halo_doubles(plan) * 8);
    if (e != cudaSuccess) return cuda_fail(e, "mol_dist_register");
    h.owned = true;
    h.fresh = false;
    if (D.p2p.on) {      // slots are assigned in registration order: every rank must register in the same order
        for (int sidx = 1; sidx < MOL_P2P_SLOTS && h.slot < 0; ++sidx)
            if (!D.p2p.used[sidx]) { D.p2p.used[sidx] = true; h.slot = sidx; }
        if (h.slot < 0) { cudaFree(h.lo); cudaFree(h.hi); return fail(MOL_E_ARG, "out of ghost-plane slots"); }
    }
    D.halos[arr_dev] = h;
    return MOL_OK;
}

extern "C" int mol_dist_unregister(mol_plan* plan, const double* arr_dev) {
    if (!plan) return fail(MOL_E_ARG, "null argument");
    auto it = plan->dist.halos.find(arr_dev);
    if (it == plan->dist.halos.end()) return MOL_OK;
    if (it->second.owned) { cudaFree(it->second.lo); cudaFree(it->second.hi); }
    if (it->second.slot > 0) plan->dist.p2p.used[it->second.slot] = false;
    plan->dist.halos.erase(it);
    return MOL_OK;
}

extern "C" int mol_dist_invalidate(mol_plan* plan, const double* arr_dev) {
    if (!plan) return fail(MOL_E_ARG, "null argument");
    dist_mark_stale(plan, arr_dev);
    return MOL_OK;
}

extern "C" int mol_dist_allreduce_sum(mol_plan* plan, double* dev, int n, void* stream) {
    if (!plan || !dev || n < 1) return fail(MOL_E_ARG, "bad argument");
    return dist_allreduce_sum(plan, dev, n, (cudaStream_t)stream);
}
