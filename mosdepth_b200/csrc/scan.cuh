// scan.cuh -- K6 scan_segments: int32 inclusive prefix sum of EVERY contig's array in one launch (to_coverage -> hts-nim
// cumsum, mosdepth.nim:54-56, which the reference runs once per contig inside `for target in sub_targets`, :691,:713),
// fused with the two whole-contig consumers that are order independent: inc(chrom_global_distribution, arr, 0, len-1)
// (:417-437,:746) and newDepthStat(arr[0..<len-1]) (:747, depthstat.nim:39-45).
//
// A *segment* is a contiguous piece of the depth arena that gets its own prefix sum: a whole contig (tlen + 1 cells, the
// first tlen of which feed the stats) or the owned range of a contig that is sharded across GPUs.  All segments are cut
// into tiles (8192 cells = 32 KB in the scan launch) that are numbered through (segment s owns tiles [first_tile, first_tile +
// ceil(n_scan / tile))); persistent CTAs take tile numbers from an atomic ticket AT THE MOMENT they are ready for one, so the
// decoupled look-back never waits for a tile that has not started, and it simply stops at the first tile of its own segment.
// Single pass, 128-bit coalesced loads and stores, warp-wide look-back (32 predecessors per probe).
//
// What round 2 measured (profiles/README.md, r2j-r2n), in the order it mattered:
//   * every load of a tile must be issued before anything waits for one (the bounds checks of a partial tile sat between the loads:
//     four DRAM latencies per tile) -- 38 % -> 54 % of the HBM peak together with
//   * lean stats: a cell below 1024 only touches its shared-memory bin, one atomic per run of equal neighbours, no branches; sum / min /
//     max are read off the bins when a CTA leaves a segment (43 -> 26 thread instructions per cell);
//   * 32-KB tiles on three CTAs per SM instead of 16 KB on four, and an L2 prefetch of the tile one grid further on (tickets ascend,
//     so that is about the tile somebody takes next): 54 % -> 66-68 %;
//   * tickets must NOT be taken ahead of time: a kernel that fed a shared-memory ring with bulk asynchronous copies three tiles ahead
//     (profiles/r2k_scan_pipe_kernel.cuh.txt) ran at 27 % -- a tile that is assigned early waits for its CTA while every later tile
//     waits for its aggregate; a look-back that reads 128 states per probe changed nothing (the walk is short; the wait is for aggregates).
//
// HBM traffic per base: 4 B read + 4 B written (the depth array is kept for the window / region / run kernels).
// r1 ran one launch + one 3.2 MB histogram memset + one 3.2 MB histogram add per contig (25 x 3 launches for a human genome);
// this is one launch for the genome, and the per-contig results stay on the device until one asynchronous copy.
#pragma once
#include "common.cuh"

namespace msd {

constexpr int kScanThreads = 256;
constexpr int kScanVecPerThread = 4;                                   // int4 per thread (the stats-only pass and the default scan)
constexpr int kScanTile = kScanThreads * kScanVecPerThread * 4;        // 4096 cells = 16 KB
constexpr int scan_tile_cells(int vec) { return kScanThreads * vec * 4; }
constexpr int kHistSmemBins = 1024;
constexpr int kHistKeep = 4096;   // dist bins kept per segment on the device (deeper contigs are counted again, one at a time)

struct ContigAccum {            // device-side accumulators of one contig (also reused per query)
    unsigned long long cum_depth;
    unsigned int min_depth, max_depth;
    unsigned int neg_seen, pad;
};

struct ScanSeg {
    unsigned long long arr_off;   // first cell of the segment in the depth arena (a multiple of 32: 128-bit accesses stay aligned)
    uint32_t n_scan;              // cells that get the prefix sum
    uint32_t n_stat;              // cells [0, n_stat) feed depth_stat + dist (0: none -- a shard gets them in msd_contig_finalize)
    uint32_t first_tile;          // number of the segment's first tile among all tiles of the launch
    uint32_t slot;                // row of acc_all / hist_all that receives the segment's stats
};

MSD_D void hist_add(uint32_t* __restrict__ sh, unsigned long long* __restrict__ gh, int32_t v, uint32_t n) {
    if (v < 0) return;                                   // :436
    if (v > (int32_t)kMaxCoverage) v = kMaxCoverage - 10;   // :429-430
    if (v < kHistSmemBins) atomicAdd(&sh[v], n);
    else atomicAdd(&gh[v], (unsigned long long)n);
}
struct SegStats { unsigned long long sum; uint32_t mn, mx; };

MSD_D void prefetch_l2(const void* p) {
#if defined(__CUDACC__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// flush the shared histogram into hist_all[slot] and block-reduce the stats into acc[slot]; all threads call it.  Cells whose depth is below
// kHistSmemBins only ever touch their shared bin (seg_stats4): their part of the sum, the minimum and the maximum is read off the bins here --
// sum = sum of bin x count, min / max = lowest / highest bin in use -- which takes the per-cell 64-bit add and the two compares out of the tile loop.
MSD_D void seg_flush(SegStats& my, uint32_t* __restrict__ s_hist, unsigned long long* __restrict__ s_sum, uint32_t* __restrict__ s_mn, uint32_t* __restrict__ s_mx,
                     ContigAccum* __restrict__ acc, unsigned long long* __restrict__ hist) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    __syncthreads();                                     // every warp's shared-memory atomics of the last tile have landed
    for (int i = t; i < kHistSmemBins; i += kScanThreads) {
        const uint32_t c = s_hist[i];
        if (c) {
            atomicAdd(&hist[i], (unsigned long long)c); s_hist[i] = 0;
            my.sum += (unsigned long long)c * (unsigned long long)i;
            my.mn = (uint32_t)i < my.mn ? (uint32_t)i : my.mn; my.mx = (uint32_t)i > my.mx ? (uint32_t)i : my.mx;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my.sum += __shfl_down_sync(0xffffffffu, my.sum, o);
        const uint32_t a = __shfl_down_sync(0xffffffffu, my.mn, o), b = __shfl_down_sync(0xffffffffu, my.mx, o);
        my.mn = a < my.mn ? a : my.mn; my.mx = b > my.mx ? b : my.mx;
    }
    if (lane == 0) { s_sum[warp] = my.sum; s_mn[warp] = my.mn; s_mx[warp] = my.mx; }
    __syncthreads();
    if (t == 0) {
        unsigned long long S = 0; uint32_t mn = 0xffffffffu, mx = 0;
        for (int w = 0; w < kScanThreads / 32; w++) { S += s_sum[w]; mn = s_mn[w] < mn ? s_mn[w] : mn; mx = s_mx[w] > mx ? s_mx[w] : mx; }
        if (S) atomicAdd(&acc->cum_depth, S);
        if (mn != 0xffffffffu || mx != 0) { atomicMin(&acc->min_depth, mn); atomicMax(&acc->max_depth, mx); }
    }
    my.sum = 0; my.mn = 0xffffffffu; my.mx = 0;
    __syncthreads();
}

// The stats of four consecutive cells starting at segment index idx (cells at or beyond n_stat do not count).
// Common case -- four counted cells, all in [0, kHistSmemBins): only the shared bins are touched (sum / min / max follow in seg_flush), one
// atomic per run of equal neighbours, no branches (r2l: the per-cell version spent ~30 instructions per cell, 14 % of them branches).
// (r2e tried grouping the lanes of a warp by value with match.any, one atomic per group: no gain, dropped.)
MSD_D void seg_stats4(const int4 x, uint32_t idx, uint32_t n_stat, SegStats& my, uint32_t* __restrict__ s_hist, unsigned long long* __restrict__ hist) {
    if (idx + 3 < n_stat && (uint32_t)(x.x | x.y | x.z | x.w) < (uint32_t)kHistSmemBins) {
        const bool p1 = x.y != x.x, p2 = x.z != x.y, p3 = x.w != x.z;
        const uint32_t l2 = p3 ? 1u : 2u, l1 = p2 ? 1u : 1u + l2, l0 = p1 ? 1u : 1u + l1;     // length of the run that starts at cell 2 / 1 / 0
        atomicAdd(&s_hist[x.x], l0);
        if (p1) atomicAdd(&s_hist[x.y], l1);
        if (p2) atomicAdd(&s_hist[x.z], l2);
        if (p3) atomicAdd(&s_hist[x.w], 1u);
        return;
    }
    const int32_t e[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (idx + q >= n_stat) break;
        const int32_t d = e[q];
        if (d >= 0 && d < kHistSmemBins) { atomicAdd(&s_hist[d], 1u); continue; }
        my.sum += (unsigned long long)(long long)d;          // dp.uint64 (depthstat.nim:43)
        const uint32_t u = (uint32_t)d;
        my.mn = u < my.mn ? u : my.mn; my.mx = u > my.mx ? u : my.mx;
        if (d >= 0 && d < kHistKeep) atomicAdd(&hist[d], 1ull);      // deeper values: the segment's max_depth tells the host to count this contig again
    }
}

// Decoupled look-back of one tile, by one warp.  tile_state[i] = (flag << 32) | value, flag 0 = not ready, 1 = tile aggregate, 2 = inclusive
// prefix.  Publishes the tile's aggregate, walks back kGroups x 32 tiles per probe (lane l of group g reads tile k - 32 g - l; tiles in front of
// the segment count as "inclusive prefix 0"), publishes the inclusive prefix and returns the exclusive one (the same value in every lane).
template <int kGroups>
MSD_D uint32_t look_back(unsigned long long* __restrict__ tile_state, uint32_t tile, uint32_t first_tile, uint32_t tile_total, int lane) {
    if (tile == first_tile) {
        if (lane == 0) atomicExch(&tile_state[tile], (2ull << 32) | tile_total);
        return 0u;
    }
    if (lane == 0) atomicExch(&tile_state[tile], (1ull << 32) | tile_total);
    uint32_t prefix = 0;
    int64_t k = (int64_t)tile - 1;
    bool done = false;
    while (!done) {
        unsigned long long s[kGroups];
#pragma unroll
        for (int g = 0; g < kGroups; g++) {
            const int64_t idx = k - 32 * g - lane;
            s[g] = (2ull << 32);
            if (idx >= (int64_t)first_tile) s[g] = *((volatile unsigned long long*)&tile_state[idx]);
        }
        bool stall = false;
#pragma unroll
        for (int g = 0; g < kGroups; g++) {
            if (done || stall) continue;                                           // (warp-uniform)
            const uint32_t flag = (uint32_t)(s[g] >> 32);
            const unsigned ready = __ballot_sync(0xffffffffu, flag != 0), incl = __ballot_sync(0xffffffffu, flag == 2);
            const int f = incl ? __ffs((int)incl) - 1 : 32;                       // nearest tile that knows its inclusive prefix
            const unsigned need = f >= 31 ? 0xffffffffu : ((2u << f) - 1u);       // lanes 0..f (all 32 when none does)
            if ((ready & need) != need) { stall = true; continue; }                // some tile in between has not published yet
            uint32_t val = (lane <= f) ? (uint32_t)s[g] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
            prefix += __shfl_sync(0xffffffffu, val, 0);
            if (f < 32) done = true; else k -= 32;
        }
    }
    if (lane == 0) atomicExch(&tile_state[tile], (2ull << 32) | (uint32_t)(prefix + tile_total));
    return prefix;
}

// tile_state[i]: (flag << 32) | value, flag 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix.
// kScan = false: the arrays already hold depth (a shard after msd_exchange_spills): only the stats pass, tiles taken by stride.
// kVec int4 per thread: a tile is 256 x kVec x 4 cells (4: 16 KB, four CTAs per SM; 8: 32 KB, three CTAs per SM).  kAhead: every CTA asks L2 for the
// lines of the tile gridDim.x tiles further on -- tickets are handed out in order, so that is about the tile some CTA will take next.
template <bool kScan, int kGroups = 1, int kVec = kScanVecPerThread, bool kAhead = false>
__global__ void __launch_bounds__(kScanThreads, kVec <= 4 ? 4 : 3)
scan_segments_kernel(int32_t* __restrict__ arena, const ScanSeg* __restrict__ segs, int n_segs, uint32_t n_tiles,
                     unsigned long long* __restrict__ tile_state, uint32_t* __restrict__ ticket,
                     ContigAccum* __restrict__ acc_all, unsigned long long* __restrict__ hist_all /* [slot][kHistKeep] */) {
    __shared__ uint32_t s_hist[kHistSmemBins];
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ uint32_t s_tile, s_prefix;
    __shared__ unsigned long long s_sum[kScanThreads / 32];
    __shared__ uint32_t s_mn[kScanThreads / 32], s_mx[kScanThreads / 32];
    constexpr int kTile = kScanThreads * kVec * 4;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < kHistSmemBins; i += kScanThreads) s_hist[i] = 0;
    SegStats my{0, 0xffffffffu, 0};
    int cur_seg = -1; bool cur_stats = false; uint32_t cur_slot = 0;
    __syncthreads();
    uint32_t tile = blockIdx.x;
    for (;; tile += gridDim.x) {
        if (kScan) {
            if (t == 0) s_tile = atomicAdd(ticket, 1u);
            __syncthreads();
            tile = s_tile;
        }
        if (tile >= n_tiles) break;
        // the segment of this tile: the last one with first_tile <= tile (uniform across the CTA; a CTA's tiles ascend, so the
        // previous answer is right most of the time)
        int sg = cur_seg;
        if (sg < 0 || tile < segs[sg].first_tile || (sg + 1 < n_segs && tile >= segs[sg + 1].first_tile)) {
            int lo = 0, hi = n_segs - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (segs[mid].first_tile <= tile) lo = mid; else hi = mid - 1; }
            sg = lo;
        }
        if (sg != cur_seg) {
            if (cur_stats) seg_flush(my, s_hist, s_sum, s_mn, s_mx, acc_all + cur_slot, hist_all + (size_t)cur_slot * kHistKeep);
            cur_seg = sg; cur_stats = segs[sg].n_stat != 0; cur_slot = segs[sg].slot;
        }
        const ScanSeg S = segs[sg];
        int32_t* __restrict__ arr = arena + S.arr_off;
        const uint32_t n_scan = S.n_scan, rel = tile - S.first_tile, tile_base = rel * (uint32_t)kTile;
        unsigned long long* __restrict__ hist = hist_all + (size_t)S.slot * kHistKeep;
        // warp w owns cells [tile_base + w * kVec * 128, + kVec * 128): kVec rounds of 32 lanes x int4.  All the loads are issued before anything waits for one
        // (r2l: with the bounds checks of a partial tile between them, each load was waited for in turn -- four DRAM latencies per tile).
        int4 v[kVec];
        const bool whole = tile_base + (uint32_t)kTile <= n_scan && tile_base + (uint32_t)kTile > tile_base;     // (CTA-uniform)
        if (whole) {
#pragma unroll
            for (int j = 0; j < kVec; j++) v[j] = *reinterpret_cast<const int4*>(arr + tile_base + warp * (kVec * 128) + j * 128 + lane * 4);
        } else {
#pragma unroll
            for (int j = 0; j < kVec; j++) {
                const uint32_t idx = tile_base + warp * (kVec * 128) + j * 128 + lane * 4;
                int4 x = make_int4(0, 0, 0, 0);
                if (idx + 3 < n_scan) x = *reinterpret_cast<const int4*>(arr + idx);
                else {
                    if (idx < n_scan) x.x = arr[idx];
                    if (idx + 1 < n_scan) x.y = arr[idx + 1];
                    if (idx + 2 < n_scan) x.z = arr[idx + 2];
                }
                v[j] = x;
            }
        }
        if (kAhead && t < kTile / 32) {          // one 128-byte line per thread
            const uint64_t c = (uint64_t)tile_base + (uint64_t)gridDim.x * kTile + (uint64_t)t * 32u;
            if (c < n_scan) prefetch_l2(arr + c);
        }
        uint32_t run = 0;
        if (kScan) {
#pragma unroll
            for (int j = 0; j < kVec; j++) {
                // local inclusive scan of the 4 cells (uint32 wrap-around like the reference's int32 cumsum)
                const int4 x = v[j];
                uint32_t a = (uint32_t)x.x, b = a + (uint32_t)x.y, c = b + (uint32_t)x.z, d = c + (uint32_t)x.w;
                uint32_t incl = d;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                const uint32_t excl = incl - d + run;
                v[j] = make_int4((int)(a + excl), (int)(b + excl), (int)(c + excl), (int)(d + excl));
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        uint32_t add = 0;
        if (kScan) {
            if (lane == 0) s_warp[warp] = run;
            __syncthreads();
            uint32_t warp_excl = 0, tile_total = 0;
#pragma unroll
            for (int w = 0; w < kScanThreads / 32; w++) { uint32_t x = s_warp[w]; if (w < warp) warp_excl += x; tile_total += x; }
            if (warp == 0) {
                const uint32_t prefix = look_back<kGroups>(tile_state, tile, S.first_tile, tile_total, lane);
                if (lane == 0) s_prefix = prefix;
            }
            __syncthreads();
            add = s_prefix + warp_excl;
        }
#pragma unroll
        for (int j = 0; j < kVec; j++) {
            const uint32_t idx = tile_base + warp * (kVec * 128) + j * 128 + lane * 4;
            const int4 x = make_int4((int)((uint32_t)v[j].x + add), (int)((uint32_t)v[j].y + add), (int)((uint32_t)v[j].z + add), (int)((uint32_t)v[j].w + add));
            if (kScan) {
                if (whole || idx + 3 < n_scan) *reinterpret_cast<int4*>(arr + idx) = x;
                else {
                    if (idx < n_scan) arr[idx] = x.x;
                    if (idx + 1 < n_scan) arr[idx + 1] = x.y;
                    if (idx + 2 < n_scan) arr[idx + 2] = x.z;
                }
            }
            if (cur_stats) seg_stats4(x, idx, S.n_stat, my, s_hist, hist);
        }
        if (kScan) __syncthreads();   // s_tile / s_warp / s_prefix reuse
    }
    if (cur_stats) seg_flush(my, s_hist, s_sum, s_mn, s_mx, acc_all + cur_slot, hist_all + (size_t)cur_slot * kHistKeep);
}

// sum_into (mosdepth.nim:491-499, :761-764) and depth_stat `+` (depthstat.nim:53-57) for every segment with stats, in one launch:
// the genome-wide `total` accumulators get each segment's bins up to its own maximum depth (not 400,001 bins per contig).
// kTotalsParts CTAs per segment.  Segments deeper than kHistKeep are skipped here (the host counts them again and adds them itself).
constexpr int kTotalsParts = 4;
__global__ void totals_add_kernel(const ScanSeg* __restrict__ segs, int n_segs, const ContigAccum* __restrict__ acc_all,
                                  const unsigned long long* __restrict__ hist_all, long long* __restrict__ tot_hist, long long* __restrict__ tot_stats) {
    const int sg = blockIdx.x / kTotalsParts, part = blockIdx.x % kTotalsParts;
    if (sg >= n_segs) return;
    const ScanSeg S = segs[sg];
    if (!S.n_stat) return;
    const ContigAccum a = acc_all[S.slot];
    if (a.max_depth >= (uint32_t)kHistKeep) return;
    const unsigned long long* __restrict__ h = hist_all + (size_t)S.slot * kHistKeep;
    for (uint32_t i = part * blockDim.x + threadIdx.x; i <= a.max_depth; i += kTotalsParts * blockDim.x) { const unsigned long long c = h[i]; if (c) atomicAdd((unsigned long long*)&tot_hist[i], c); }
    if (part == 0 && threadIdx.x == 0) {
        atomicAdd((unsigned long long*)&tot_stats[0], (unsigned long long)S.n_stat); atomicAdd((unsigned long long*)&tot_stats[1], a.cum_depth);
        atomicMin(&tot_stats[2], (long long)a.min_depth); atomicMax(&tot_stats[3], (long long)a.max_depth);
    }
}

// The two whole-contig consumers (inc :417-437, newDepthStat depthstat.nim:39-45) over cells [0, n) of ONE array with the full-width
// dist table: contigs deeper than kHistKeep, counted again.  Same arithmetic as the stats branch of scan_segments_kernel.
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
