// Host shim that lets g++ compile the NVRTC source of a stencil program (prelude + kernels/mol_device.cuh + generated
// ghost rules / equations + kernels/mol_generic.cuh) as plain C++ and run the TABLE-DRIVEN kernel as one emulated
// CTA of one warp: 32 host threads that meet at __syncthreads and exchange values in warp shuffles.  Test infrastructure only: it checks
// what the code generator emits and what the device runtime computes on a machine without a GPU.  It is not a
// product path (the product has no CPU fallback), and the tiled kernel (TMA / cp.async / mbarrier PTX) is out of its reach.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static          /* (tests/cuda_emu rewrites `extern __shared__` to a plain extern array) */
#define __align__(x)
#define __grid_constant__
#define __constant__ static const

struct EmuDim3 { unsigned x, y, z; };
// One emulated CTA: EMU_THREADS host threads that meet at __syncthreads and exchange values in warp shuffles
// (one barrier per warp of 32).  CTAs are run one after the other (gridDim = 1).
#include <barrier>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#ifndef EMU_THREADS
#define EMU_THREADS 32
#endif
static thread_local EmuDim3 threadIdx = {0, 0, 0};
static EmuDim3 blockIdx = {0, 0, 0}, blockDim = {EMU_THREADS, 1, 1}, gridDim = {1, 1, 1};
static std::barrier<> emu_bar(EMU_THREADS);
static std::unique_ptr<std::barrier<>> emu_warp_bar[EMU_THREADS / 32];
static std::mutex emu_mutex;
static double emu_lane_val[EMU_THREADS];

struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return {x, y}; }
inline unsigned long long __cvta_generic_to_shared(const void* p) { return (unsigned long long)p; }

template <class T>
inline T __ldg(const T* p) { return *p; }
inline int __double2hiint(double x) { unsigned long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
inline int __double2loint(double x) { unsigned long long b; std::memcpy(&b, &x, 8); return (int)(unsigned)(b & 0xffffffffull); }
inline double __hiloint2double(int hi, int lo) {
    unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double x; std::memcpy(&x, &b, 8); return x;
}
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }      // (volatile: no fma contraction)
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline void __syncthreads() { emu_bar.arrive_and_wait(); }
inline double __shfl_xor_sync(unsigned, double v, int mask) {
    std::barrier<>& wb = *emu_warp_bar[threadIdx.x / 32];
    emu_lane_val[threadIdx.x] = v;
    wb.arrive_and_wait();
    const double r = emu_lane_val[threadIdx.x ^ (unsigned)mask];
    wb.arrive_and_wait();
    return r;
}
inline double atomicAdd(double* p, double v) { std::lock_guard<std::mutex> g(emu_mutex); const double o = *p; *p += v; return o; }
inline int atomicAdd(int* p, int v) { std::lock_guard<std::mutex> g(emu_mutex); const int o = *p; *p += v; return o; }
using std::max;
using std::min;
// run `kernel` once per thread of the CTA
inline void emu_launch(const std::function<void()>& kernel) {
    for (auto& b : emu_warp_bar) b = std::make_unique<std::barrier<>>(32);
    std::vector<std::thread> lanes;
    for (unsigned tid = 0; tid < EMU_THREADS; ++tid)
        lanes.emplace_back([tid, &kernel]() { threadIdx.x = tid; kernel(); });
    for (auto& t : lanes) t.join();
}
