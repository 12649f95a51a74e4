// scan.cuh -- K6 scan_fused: per-contig int32 inclusive prefix sum (to_coverage -> hts-nim cumsum,
// mosdepth.nim:54-56) fused with the two whole-contig consumers that are order independent:
// inc(chrom_global_distribution, arr, 0, len-1) (:417-437,:746) and newDepthStat(arr[0..<len-1]) (:747,
// depthstat.nim:39-45).  Single pass, decoupled look-back between tiles, 128-bit coalesced loads and stores.
// Persistent CTAs take tiles from an atomic ticket, so the look-back never waits on a tile that has not started.
//
// HBM traffic per base: 4 B read + 4 B written (the depth array is kept for the window / region / run kernels).
#pragma once
#include "common.cuh"

namespace msd {

constexpr int kScanThreads = 256;
constexpr int kScanVecPerThread = 4;                                   // int4 per thread
constexpr int kScanTile = kScanThreads * kScanVecPerThread * 4;        // 4096 cells = 16 KB
constexpr int kHistSmemBins = 1024;

struct ContigAccum {            // device-side accumulators of one contig (also reused per query)
    unsigned long long cum_depth;
    unsigned int min_depth, max_depth;
    unsigned int neg_seen, pad;
};

MSD_D void hist_add(uint32_t* __restrict__ sh, unsigned long long* __restrict__ gh, int32_t v, uint32_t n) {
    if (v < 0) return;                                   // :436
    if (v > (int32_t)kMaxCoverage) v = kMaxCoverage - 10;   // :429-430
    if (v < kHistSmemBins) atomicAdd(&sh[v], n);
    else atomicAdd(&gh[v], (unsigned long long)n);
}

// tile_state[i]: (flag << 32) | value, flag 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix
__global__ void __launch_bounds__(kScanThreads)
scan_fused_kernel(int32_t* __restrict__ arr, uint32_t n_scan, uint32_t n_stat, unsigned long long* __restrict__ tile_state,
                  uint32_t* __restrict__ ticket, unsigned long long* __restrict__ hist, ContigAccum* __restrict__ acc, int do_stats) {
    __shared__ uint32_t s_hist[kHistSmemBins];
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ uint32_t s_tile, s_prefix;
    __shared__ unsigned long long s_sum[kScanThreads / 32];
    __shared__ uint32_t s_mn[kScanThreads / 32], s_mx[kScanThreads / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t n_tiles = (n_scan + kScanTile - 1) / kScanTile;
    if (do_stats) for (int i = t; i < kHistSmemBins; i += kScanThreads) s_hist[i] = 0;
    unsigned long long my_sum = 0; uint32_t my_mn = 0xffffffffu, my_mx = 0;
    __syncthreads();
    for (;;) {
        if (t == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;
        const uint32_t tile_base = tile * kScanTile;
        // warp w owns cells [tile_base + w*512, +512): 4 rounds of 32 lanes x int4
        int4 v[kScanVecPerThread];
        uint32_t run = 0;
#pragma unroll
        for (int j = 0; j < kScanVecPerThread; j++) {
            const uint32_t idx = tile_base + warp * 512 + j * 128 + lane * 4;
            int4 x = make_int4(0, 0, 0, 0);
            if (idx + 3 < n_scan) x = *reinterpret_cast<const int4*>(arr + idx);
            else {
                if (idx < n_scan) x.x = arr[idx];
                if (idx + 1 < n_scan) x.y = arr[idx + 1];
                if (idx + 2 < n_scan) x.z = arr[idx + 2];
            }
            // local inclusive scan of the 4 cells (uint32 wrap-around like the reference's int32 cumsum)
            uint32_t a = (uint32_t)x.x, b = a + (uint32_t)x.y, c = b + (uint32_t)x.z, d = c + (uint32_t)x.w;
            uint32_t incl = d;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            const uint32_t excl = incl - d + run;
            v[j] = make_int4((int)(a + excl), (int)(b + excl), (int)(c + excl), (int)(d + excl));
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_warp[warp] = run;
        __syncthreads();
        uint32_t warp_excl = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; w++) { uint32_t x = s_warp[w]; if (w < warp) warp_excl += x; tile_total += x; }
        // decoupled look-back (thread 0)
        if (t == 0) {
            uint32_t prefix = 0;
            if (tile == 0) {
                atomicExch(&tile_state[0], (2ull << 32) | tile_total);
            } else {
                atomicExch(&tile_state[tile], (1ull << 32) | tile_total);
                int64_t k = (int64_t)tile - 1;
                for (;;) {
                    unsigned long long s = *((volatile unsigned long long*)&tile_state[k]);
                    uint32_t flag = (uint32_t)(s >> 32);
                    if (flag == 0) continue;
                    prefix += (uint32_t)s;
                    if (flag == 2) break;
                    k--;
                }
                atomicExch(&tile_state[tile], (2ull << 32) | (uint32_t)(prefix + tile_total));
            }
            s_prefix = prefix;
        }
        __syncthreads();
        const uint32_t add = s_prefix + warp_excl;
#pragma unroll
        for (int j = 0; j < kScanVecPerThread; j++) {
            const uint32_t idx = tile_base + warp * 512 + j * 128 + lane * 4;
            int4 x = make_int4((int)((uint32_t)v[j].x + add), (int)((uint32_t)v[j].y + add), (int)((uint32_t)v[j].z + add), (int)((uint32_t)v[j].w + add));
            if (idx + 3 < n_scan) *reinterpret_cast<int4*>(arr + idx) = x;
            else {
                if (idx < n_scan) arr[idx] = x.x;
                if (idx + 1 < n_scan) arr[idx + 1] = x.y;
                if (idx + 2 < n_scan) arr[idx + 2] = x.z;
            }
            if (do_stats) {
                const int32_t e[4] = {x.x, x.y, x.z, x.w};
                int32_t cur = 0; uint32_t cnt = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (idx + q >= n_stat) break;
                    const int32_t d = e[q];
                    my_sum += (unsigned long long)(long long)d;          // dp.uint64 (depthstat.nim:43)
                    const uint32_t u = (uint32_t)d;
                    my_mn = u < my_mn ? u : my_mn; my_mx = u > my_mx ? u : my_mx;
                    if (cnt && d == cur) cnt++;
                    else { if (cnt) hist_add(s_hist, hist, cur, cnt); cur = d; cnt = 1; }
                }
                if (cnt) hist_add(s_hist, hist, cur, cnt);
            }
        }
        __syncthreads();   // s_tile / s_warp reuse
    }
    if (!do_stats) return;
    // block reduction of the stats, flush the shared histogram
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_sum += __shfl_down_sync(0xffffffffu, my_sum, o);
        uint32_t a = __shfl_down_sync(0xffffffffu, my_mn, o), b = __shfl_down_sync(0xffffffffu, my_mx, o);
        my_mn = a < my_mn ? a : my_mn; my_mx = b > my_mx ? b : my_mx;
    }
    if (lane == 0) { s_sum[warp] = my_sum; s_mn[warp] = my_mn; s_mx[warp] = my_mx; }
    __syncthreads();
    if (t == 0) {
        unsigned long long S = 0; uint32_t mn = 0xffffffffu, mx = 0;
        for (int w = 0; w < kScanThreads / 32; w++) { S += s_sum[w]; mn = s_mn[w] < mn ? s_mn[w] : mn; mx = s_mx[w] > mx ? s_mx[w] : mx; }
        atomicAdd(&acc->cum_depth, S); atomicMin(&acc->min_depth, mn); atomicMax(&acc->max_depth, mx);
    }
    for (int i = t; i < kHistSmemBins; i += kScanThreads) { uint32_t c = s_hist[i]; if (c) atomicAdd(&hist[i], (unsigned long long)c); }
}

// The two whole-contig consumers (inc :417-437, newDepthStat depthstat.nim:39-45) over cells [0, n) of an array that is
// already a depth array: the shard of a contig after msd_depth_add() (msd_contig_finalize).  Same arithmetic as the
// do_stats branch of scan_fused_kernel.
__global__ void __launch_bounds__(256)
range_stats_kernel(const int32_t* __restrict__ arr, uint32_t n, unsigned long long* __restrict__ hist, ContigAccum* __restrict__ acc) {
    __shared__ uint32_t s_hist[kHistSmemBins];
    for (int i = threadIdx.x; i < kHistSmemBins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    unsigned long long my_sum = 0; uint32_t my_mn = 0xffffffffu, my_mx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t d = __ldg(arr + i);
        my_sum += (unsigned long long)(long long)d;
        const uint32_t u = (uint32_t)d;
        my_mn = u < my_mn ? u : my_mn; my_mx = u > my_mx ? u : my_mx;
        hist_add(s_hist, hist, d, 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_sum += __shfl_down_sync(0xffffffffu, my_sum, o);
        uint32_t a = __shfl_down_sync(0xffffffffu, my_mn, o), b = __shfl_down_sync(0xffffffffu, my_mx, o);
        my_mn = a < my_mn ? a : my_mn; my_mx = b > my_mx ? b : my_mx;
    }
    if ((threadIdx.x & 31) == 0) { if (my_sum) atomicAdd(&acc->cum_depth, my_sum); atomicMin(&acc->min_depth, my_mn); atomicMax(&acc->max_depth, my_mx); }
    __syncthreads();
    for (int i = threadIdx.x; i < kHistSmemBins; i += blockDim.x) { const uint32_t c = s_hist[i]; if (c) atomicAdd(&hist[i], (unsigned long long)c); }
}

__global__ void add_cells_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, uint32_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

}  // namespace msd
