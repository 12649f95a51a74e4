// runs.cuh -- K8 rle_runs (gen_depths, mosdepth.nim:58-96) and K9 quantize_runs (gen_quantized :124-141 with
// linear_search :98-109): run-length segmentation of f(depth[i]) over [0, n), f = identity or quantize bin.
// Two passes over the depth array (count boundaries per tile, then write them at their scanned rank); the
// boundary positions and values are what the host formats into BED lines.
#pragma once
#include "common.cuh"

namespace msd {

constexpr int kRunThreads = 256;
constexpr int kRunItems = 16;
constexpr int kRunTile = kRunThreads * kRunItems;
constexpr int kMaxQuants = 64;

struct QuantArgs { long long q[kMaxQuants]; int n; };

struct IdentityF { MSD_D int32_t operator()(int32_t v) const { return v; } };
struct QuantF {
    QuantArgs qa;
    // linear_search (mosdepth.nim:98-109): -1 outside [q0, q_last]; index of the last value <= v otherwise
    MSD_D int32_t operator()(int32_t v) const {
        const long long x = v;
        if (x < qa.q[0] || x > qa.q[qa.n - 1]) return -1;
        for (int i = 0; i < qa.n; i++) { if (qa.q[i] > x) return i - 1; if (qa.q[i] == x) return i; }
        return qa.n - 1;
    }
};

template <class F>
__global__ void __launch_bounds__(kRunThreads)
runs_count_kernel(const int32_t* __restrict__ arr, uint32_t n, F f, uint32_t* __restrict__ tile_count) {
    __shared__ uint32_t s_w[kRunThreads / 32];
    const uint32_t base = blockIdx.x * kRunTile + threadIdx.x * kRunItems;
    uint32_t c = 0;
    if (base < n) {
        int32_t prev = base ? f(arr[base - 1]) : 0;
#pragma unroll 4
        for (int k = 0; k < kRunItems; k++) {
            const uint32_t i = base + k;
            if (i >= n) break;
            const int32_t cur = f(arr[i]);
            c += (i == 0) || (cur != prev);
            prev = cur;
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t s = 0; for (int w = 0; w < kRunThreads / 32; w++) s += s_w[w]; tile_count[blockIdx.x] = s; }
}

// single CTA: exclusive scan of tile_count (64-bit totals) -> tile_base, total
__global__ void __launch_bounds__(1024, 1)
runs_scan_kernel(const uint32_t* __restrict__ tile_count, uint32_t n_tiles, unsigned long long* __restrict__ tile_base,
                 unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_scan[32];
    __shared__ unsigned long long s_run;
    const int t = threadIdx.x, nt = blockDim.x;
    if (t == 0) s_run = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += nt) {
        const uint32_t i = base + t;
        unsigned long long v = i < n_tiles ? tile_count[i] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xffffffffu, x, o); if ((t & 31) >= o) x += y; }
        if ((t & 31) == 31) s_scan[t >> 5] = x;
        __syncthreads();
        if (t < 32) {
            unsigned long long w = t < (nt >> 5) ? s_scan[t] : 0, z = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xffffffffu, z, o); if (t >= o) z += y; }
            s_scan[t] = z - w;
        }
        __syncthreads();
        const unsigned long long excl = s_run + s_scan[t >> 5] + x - v;
        if (i < n_tiles) tile_base[i] = excl;
        __syncthreads();
        if (t == nt - 1) s_run = excl + v;
        __syncthreads();
    }
    if (t == 0) *total = s_run;
}

template <class F>
__global__ void __launch_bounds__(kRunThreads)
runs_write_kernel(const int32_t* __restrict__ arr, uint32_t n, F f, const unsigned long long* __restrict__ tile_base,
                  int32_t* __restrict__ out_start, int32_t* __restrict__ out_val) {
    __shared__ uint32_t s_w[kRunThreads / 32];
    const uint32_t base = blockIdx.x * kRunTile + threadIdx.x * kRunItems;
    uint32_t flags = 0; int32_t vals[kRunItems];
    uint32_t c = 0;
    if (base < n) {
        int32_t prev = base ? f(arr[base - 1]) : 0;
#pragma unroll
        for (int k = 0; k < kRunItems; k++) {
            const uint32_t i = base + k;
            int32_t cur = prev;
            if (i < n) { cur = f(arr[i]); if (i == 0 || cur != prev) { flags |= 1u << k; c++; } }
            vals[k] = cur; prev = cur;
        }
    }
    // block exclusive scan of c
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    uint32_t wex = 0;
    for (int w = 0; w < warp; w++) wex += s_w[w];
    unsigned long long o = tile_base[blockIdx.x] + wex + x - c;
#pragma unroll
    for (int k = 0; k < kRunItems; k++) if (flags & (1u << k)) { out_start[o] = (int32_t)(base + k); out_val[o] = vals[k]; o++; }
}

// stop[i] = start[i+1], last = end
__global__ void runs_stops_kernel(const int32_t* __restrict__ start, unsigned long long n_runs, int32_t end, int32_t* __restrict__ stop) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_runs; i += (unsigned long long)gridDim.x * blockDim.x)
        stop[i] = (i + 1 < n_runs) ? start[i + 1] : end;
}


// K9's filter: gen_quantized only yields the segments whose bin has a label (bin != -1 and bin < n_labels, mosdepth.nim:136,:140).  The
// stops are fixed before the filter (a dropped segment leaves a gap), so this is a second, order-preserving compaction over the runs
// themselves: count the kept runs per tile, runs_scan_kernel, then move (start, stop, bin) to their final rank.  r1 did this in a
// host loop over the pinned arrays.
MSD_D bool quant_keep(int32_t bin, int32_t n_labels) { return bin != -1 && bin < n_labels; }

__global__ void __launch_bounds__(kRunThreads)
keep_count_kernel(const int32_t* __restrict__ val, unsigned long long n_runs, int32_t n_labels, uint32_t* __restrict__ tile_count) {
    __shared__ uint32_t s_w[kRunThreads / 32];
    const unsigned long long base = (unsigned long long)blockIdx.x * kRunTile;
    uint32_t c = 0;
    for (int k = 0; k < kRunItems; k++) { const unsigned long long i = base + (unsigned long long)k * kRunThreads + threadIdx.x; if (i < n_runs) c += quant_keep(val[i], n_labels); }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < kRunThreads / 32; w++) t += s_w[w]; tile_count[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(kRunThreads)
keep_write_kernel(const int32_t* __restrict__ start, const int32_t* __restrict__ stop, const int32_t* __restrict__ val, unsigned long long n_runs, int32_t n_labels,
                  const unsigned long long* __restrict__ tile_base, int32_t* __restrict__ o_start, int32_t* __restrict__ o_stop, int32_t* __restrict__ o_val) {
    __shared__ uint32_t s_w[kRunThreads / 32];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = tile_base[blockIdx.x];
    __syncthreads();
    const unsigned long long base = (unsigned long long)blockIdx.x * kRunTile;
    for (int k = 0; k < kRunItems; k++) {      // rows of 256 consecutive runs: coalesced, and the order is kept row by row
        const unsigned long long i = base + (unsigned long long)k * kRunThreads + threadIdx.x;
        const bool keep = i < n_runs && quant_keep(val[i], n_labels);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        uint32_t wex = 0, row = 0;
        for (int w = 0; w < kRunThreads / 32; w++) { const uint32_t x = s_w[w]; if (w < warp) wex += x; row += x; }
        if (keep) { const unsigned long long o = s_base + wex + __popc(m & ((1u << lane) - 1u)); o_start[o] = start[i]; o_stop[o] = stop[i]; o_val[o] = val[i]; }
        __syncthreads();
        if (threadIdx.x == 0) s_base += row;
        __syncthreads();
    }
}

}  // namespace msd
