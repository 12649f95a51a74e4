// tests/emu/cuda_runtime.h -- stand-in for <cuda_runtime.h> when the kernels of mosdepth_b200/csrc are compiled by g++ for the CPU
// emulation test (tests/emu/bedgz_emu.cpp): CUDA's execution model for the few features K10's kernels use, on host threads.
//   * a launch runs its CTAs one after the other; the threads of a CTA are real host threads;
//   * __syncthreads is a barrier over the CTA, warp collectives (full mask only) go through a per-warp exchange buffer and barrier;
//   * __shared__ variables are function-local statics (one CTA at a time, so one copy is the CTA's copy);
//   * atomicOr is a real atomic, so the concurrent ORs of neighbouring lines are exercised as on the device.
// TEST INFRASTRUCTURE ONLY.  It checks indexing, barriers and launch geometry -- not performance, not the memory model.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __CUDA_ARCH__ 1000            // take the device branches of the MSD_HD functions (atomicOr, __clz)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
struct uint4 { uint32_t x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
inline int4 make_int4(int x, int y, int z, int w) { int4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

// ---- the few runtime calls the host sequence of K10 makes (bedgz_host.cuh): everything is synchronous host memory -----------------
typedef int cudaError_t;
typedef int cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emulated allocation failure" : "no error"; }

namespace emu {
struct Dim { unsigned x = 1, y = 1, z = 1; };
inline thread_local Dim t_thread;
inline Dim g_block, g_grid, g_bdim;
struct Warp { uint64_t slot[32]; std::unique_ptr<std::barrier<>> bar; };
inline std::vector<Warp> g_warps;
inline std::unique_ptr<std::barrier<>> g_cta_bar;
inline Warp& warp() { return g_warps[t_thread.x >> 5]; }
inline unsigned lane() { return t_thread.x & 31u; }

template <class F> void launch(unsigned grid, unsigned block, F body) {
    g_grid.x = grid; g_bdim.x = block;
    const unsigned nw = (block + 31) / 32;
    for (unsigned b = 0; b < grid; b++) {
        g_block.x = b;
        g_warps.clear(); g_warps.resize(nw);
        for (unsigned w = 0; w < nw; w++) g_warps[w].bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)std::min(32u, block - 32 * w));
        g_cta_bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)block);
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < block; t++) ts.emplace_back([&, t]() { t_thread.x = t; body(); g_cta_bar->arrive_and_drop(); warp().bar->arrive_and_drop(); });   // exited threads count as arrived
        for (auto& th : ts) th.join();
    }
}
}  // namespace emu

#define threadIdx (emu::t_thread)
#define blockIdx (emu::g_block)
#define blockDim (emu::g_bdim)
#define gridDim (emu::g_grid)

inline void __syncthreads() { emu::g_cta_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp().bar->arrive_and_wait(); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned o) {
    emu::Warp& w = emu::warp(); const unsigned l = emu::lane();
    uint64_t bits = 0; memcpy(&bits, &v, sizeof v); w.slot[l] = bits;
    w.bar->arrive_and_wait();
    T r = v; if (l >= o) memcpy(&r, &w.slot[l - o], sizeof r);
    w.bar->arrive_and_wait();
    return r;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    emu::Warp& w = emu::warp(); w.slot[emu::lane()] = v;
    w.bar->arrive_and_wait();
    unsigned s = 0; for (int i = 0; i < 32; i++) s += (unsigned)w.slot[i];
    w.bar->arrive_and_wait();
    return s;
}
inline unsigned atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) {
    emu::Warp& w = emu::warp();
    uint64_t bits = 0; memcpy(&bits, &v, sizeof v); w.slot[emu::lane()] = bits;
    w.bar->arrive_and_wait();
    T r; memcpy(&r, &w.slot[src & 31], sizeof r);
    w.bar->arrive_and_wait();
    return r;
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned o) {
    emu::Warp& w = emu::warp(); const unsigned l = emu::lane();
    uint64_t bits = 0; memcpy(&bits, &v, sizeof v); w.slot[l] = bits;
    w.bar->arrive_and_wait();
    T r = v; if (l + o < 32) memcpy(&r, &w.slot[l + o], sizeof r);
    w.bar->arrive_and_wait();
    return r;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) {
    emu::Warp& w = emu::warp(); const unsigned l = emu::lane();
    uint64_t bits = 0; memcpy(&bits, &v, sizeof v); w.slot[l] = bits;
    w.bar->arrive_and_wait();
    T r; memcpy(&r, &w.slot[(l ^ (unsigned)m) & 31], sizeof r);
    w.bar->arrive_and_wait();
    return r;
}
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Warp& w = emu::warp(); w.slot[emu::lane()] = pred ? 1 : 0;
    w.bar->arrive_and_wait();
    unsigned m = 0; for (int i = 0; i < 32; i++) m |= (unsigned)(w.slot[i] & 1) << i;
    w.bar->arrive_and_wait();
    return m;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
// atomics: every CUDA overload the emulated kernels use, on the GCC builtins
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T emu_atomic_minmax(T* p, T v, bool is_min) { T o = __atomic_load_n(p, __ATOMIC_RELAXED); while ((is_min ? v < o : v > o) && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline unsigned atomicMin(unsigned* p, unsigned v) { return emu_atomic_minmax(p, v, true); }
inline unsigned atomicMax(unsigned* p, unsigned v) { return emu_atomic_minmax(p, v, false); }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { return emu_atomic_minmax(p, v, false); }
inline long long atomicMin(long long* p, long long v) { return emu_atomic_minmax(p, v, true); }
inline long long atomicMax(long long* p, long long v) { return emu_atomic_minmax(p, v, false); }
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    emu::Warp& w = emu::warp(); w.slot[emu::lane()] = v;
    w.bar->arrive_and_wait();
    unsigned m = 0xffffffffu; for (int i = 0; i < 32; i++) m = std::min(m, (unsigned)w.slot[i]);
    w.bar->arrive_and_wait();
    return m;
}
// Votes under divergence: the stand-in's threads are independent, so "the lanes that are converged here" is the caller alone -- a
// legal answer for __activemask() on any Volta-or-later part, and what code written for independent thread scheduling must accept.
// A single-lane mask needs no exchange; the aggregated atomics of records.cuh then degenerate to one atomic per thread.
inline unsigned __activemask() { return 1u << emu::lane(); }
// full mask: a real exchange (the K6 histogram aggregation); a single-lane mask (the aggregated atomics of records.cuh under divergence): the caller alone
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    if (!(mask & (mask - 1))) return mask;
    if (mask != 0xffffffffu) { fprintf(stderr, "emu: __match_any_sync over a partial mask is not provided\n"); abort(); }
    emu::Warp& w = emu::warp(); const unsigned l = emu::lane();
    uint64_t bits = 0; memcpy(&bits, &v, sizeof v); w.slot[l] = bits;
    w.bar->arrive_and_wait();
    const unsigned n = std::min(32u, (unsigned)emu::g_bdim.x - 32u * (emu::t_thread.x >> 5));
    unsigned m = 0; for (unsigned i = 0; i < n; i++) if (w.slot[i] == bits) m |= 1u << i;
    w.bar->arrive_and_wait();
    return m;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { if (mask & (mask - 1)) { fprintf(stderr, "emu: __reduce_max_sync over several lanes is not provided\n"); abort(); } return v; }
inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }      // (x86-64 SSE2: IEEE double, round to nearest)
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
