// regions.cuh -- K7 region_reduce: the per-region consumers of the depth array.
//
//   imean() mean branch, mosdepth.nim:410-413: result += float64(v[i]) / L, strictly in index order.  The
//   formatted value depends on that order at rounding ties, so each region's chain is run by
//   one thread with IEEE round-to-nearest division and addition (__ddiv_rn / __dadd_rn, no reassociation, no
//   fast-math).  The quotient is only recomputed when the depth value changes (same bits, fewer divisions).
//   imean() median branch :404-408 + CountStat depthstat.nim:16-30: the stop_n-th smallest of min(v, 65535),
//   found by bisection on the value range instead of a 65536-bin table.
//   newDepthStat over the region slice :723-724, window-dist bin :737-739, per-base region dist :741,
//   threshold counts :549-553.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace msd {

struct RegionAccum {
    unsigned long long cum_length, cum_depth;
    unsigned int min_depth, max_depth;
    unsigned int neg_seen, pad;
};

// CountStat.median over arr[a, b) (depthstat.nim:23-30): values clipped to 65535; negative -> error flag
MSD_D double region_median(const int32_t* __restrict__ arr, uint32_t a, uint32_t b, unsigned int* neg_seen) {
    const int64_t n = (int64_t)b - (int64_t)a;
    if (n <= 0) return 0.0;
    const int64_t stop_n = (int64_t)(0.5 + (double)n * 0.5);
    int32_t lo = 65535, hi = 0;
    for (uint32_t i = a; i < b; i++) {
        int32_t v = arr[i];
        if (v < 0) { atomicExch(neg_seen, 1u); v = 0; }
        if (v > 65535) v = 65535;
        lo = v < lo ? v : lo; hi = v > hi ? v : hi;
    }
    if (stop_n <= 0) return 0.0;   // cum(0) >= 0 at i = 0
    // smallest x in [lo, hi] with #{v <= x} >= stop_n
    while (lo < hi) {
        const int32_t mid = lo + ((hi - lo) >> 1);
        int64_t c = 0;
        for (uint32_t i = a; i < b; i++) { int32_t v = arr[i]; if (v < 0) v = 0; if (v > 65535) v = 65535; c += (v <= mid); }
        if (c >= stop_n) hi = mid; else lo = mid + 1;
    }
    return (double)lo;
}

// sequential float64 mean of arr[a, e) with divisor L (mosdepth.nim:411-413) + slice stats
MSD_D double region_mean_seq(const int32_t* __restrict__ arr, uint32_t a, uint32_t e, double L) {
    double acc = 0.0, q = 0.0; int32_t last = 0; bool have = false;
    for (uint32_t i = a; i < e; i++) {
        const int32_t v = __ldg(arr + i);
        if (!have || v != last) { q = __ddiv_rn((double)v, L); last = v; have = true; }
        acc = __dadd_rn(acc, q);
    }
    return acc;
}

MSD_D void slice_stats(const int32_t* __restrict__ arr, uint32_t a, uint32_t b, unsigned long long& sum, uint32_t& mn, uint32_t& mx) {
    for (uint32_t i = a; i < b; i++) { const int32_t d = __ldg(arr + i); sum += (unsigned long long)(long long)d; const uint32_t u = (uint32_t)d; mn = u < mn ? u : mn; mx = u > mx ? u : mx; }
}

MSD_D void block_accumulate(RegionAccum* __restrict__ acc, unsigned long long len, unsigned long long sum, uint32_t mn, uint32_t mx) {
    // warp reduce, then one atomic set per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        len += __shfl_down_sync(0xffffffffu, len, o); sum += __shfl_down_sync(0xffffffffu, sum, o);
        uint32_t a = __shfl_down_sync(0xffffffffu, mn, o), b = __shfl_down_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn; mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        if (len) atomicAdd(&acc->cum_length, len);
        if (sum) atomicAdd(&acc->cum_depth, sum);
        atomicMin(&acc->min_depth, mn); atomicMax(&acc->max_depth, mx);
    }
}

// window mode: window_gen (mosdepth.nim:382-388) evaluated in closed form, one thread per window
__global__ void __launch_bounds__(256)
windows_kernel(const int32_t* __restrict__ arr, uint32_t tlen, uint32_t window, long long n_windows, int use_median,
               double* __restrict__ value_out, RegionAccum* __restrict__ acc, unsigned long long* __restrict__ dist512) {
    __shared__ uint32_t s_hist[kWindowDistBins];
    for (int i = threadIdx.x; i < kWindowDistBins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    unsigned long long len = 0, sum = 0; uint32_t mn = 0xffffffffu, mx = 0;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < n_windows; w += (long long)gridDim.x * blockDim.x) {
        const unsigned long long s64 = (unsigned long long)w * window;
        const uint32_t start = (uint32_t)s64;
        const uint32_t stop = (s64 + window < tlen) ? (uint32_t)(s64 + window) : tlen;
        double me;
        if (use_median) me = region_median(arr, start, stop, &acc->neg_seen);
        else me = region_mean_seq(arr, start, stop, (double)(uint32_t)(stop - start));
        value_out[w] = me;
        slice_stats(arr, start, stop, sum, mn, mx);
        len += stop - start;
        // chrom_region_distribution[min(me.toInt, 511)] += 1 (:737-739); toInt rounds half away from zero
        long long bin = me >= 0 ? (long long)__dadd_rn(me, 0.5) : (long long)__dadd_rn(me, -0.5);
        if (bin > kWindowDistBins - 1) bin = kWindowDistBins - 1;
        if (bin >= 0) atomicAdd(&s_hist[bin], 1u);
    }
    block_accumulate(acc, len, sum, mn, mx);
    __syncthreads();
    for (int i = threadIdx.x; i < kWindowDistBins; i += blockDim.x) { uint32_t c = s_hist[i]; if (c) atomicAdd(&dist512[i], (unsigned long long)c); }
}

// window mode, mean: the same results as windows_kernel(use_median = 0) with coalesced loads.  A warp owns 32
// consecutive windows, one per lane.  The windows are walked in column chunks of 32 cells: the warp loads the chunk
// of every window with one coalesced 128-byte access per window into a padded shared-memory tile, then every lane
// runs its own window's strictly sequential float64 chain (Q1) and slice stats over its row of the tile.
// (windows_kernel reads with one scalar load per cell and lane, 2 KB apart between lanes, twice: 93 GB/s.)
constexpr int kWinWarps = 8;
constexpr int kWinQuot = 1024;
__global__ void __launch_bounds__(kWinWarps * 32)
windows_mean_kernel(const int32_t* __restrict__ arr, uint32_t tlen, uint32_t window, long long n_windows,
                    double* __restrict__ value_out, RegionAccum* __restrict__ acc, unsigned long long* __restrict__ dist512) {
    __shared__ uint32_t s_hist[kWindowDistBins];
    __shared__ int32_t s_tile[kWinWarps][32][33];
    // float64(v) / window for v < 1024, the very doubles the sequential chain adds (every full window divides by the
    // same L).  In lock step a warp would otherwise run the ~300-cycle division at every cell, because some lane's depth
    // changes at almost every cell (r1k launch list: 133 us per launch whatever the grid, 290 cycles per cell).
    __shared__ double s_q[kWinQuot];
    for (int i = threadIdx.x; i < kWindowDistBins; i += blockDim.x) s_hist[i] = 0;
    for (int i = threadIdx.x; i < kWinQuot; i += blockDim.x) s_q[i] = __ddiv_rn((double)i, (double)window);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t (*tile)[33] = s_tile[warp];
    unsigned long long len = 0, sum = 0; uint32_t mn = 0xffffffffu, mx = 0;
    const long long n_groups = (n_windows + 31) / 32;
    for (long long g = (long long)blockIdx.x * kWinWarps + warp; g < n_groups; g += (long long)gridDim.x * kWinWarps) {
        const long long w0 = g * 32, w = w0 + lane;
        const bool valid = w < n_windows;
        const unsigned long long s64 = (unsigned long long)w * window;
        const unsigned long long e64 = (s64 + window < tlen) ? s64 + window : (unsigned long long)tlen;   // window_gen :382-388
        const long long my_len = valid ? (long long)(e64 - s64) : 0;
        const double L = (double)(uint32_t)my_len;
        const int n_rows = (int)((n_windows - w0) < 32 ? (n_windows - w0) : 32);
        double a = 0.0, q = 0.0; int32_t last = 0; bool have = false;
        const bool full_len = my_len == (long long)window;
        const unsigned long long g0 = (unsigned long long)w0 * window;                 // first cell of the group (< tlen)
        const uint32_t max_len = (uint32_t)(((unsigned long long)tlen - g0) < window ? ((unsigned long long)tlen - g0) : window);   // its first window is the longest
        for (uint32_t c0 = 0; c0 < max_len; c0 += 32) {
            unsigned long long idx = g0 + c0 + lane;
            const bool col_ok = c0 + lane < window;
#pragma unroll 8
            for (int r = 0; r < n_rows; r++, idx += window) tile[r][lane] = (col_ok && idx < tlen) ? __ldg(arr + idx) : 0;
            __syncwarp();
            long long n = my_len - (long long)c0; if (n > 32) n = 32;
            for (int j = 0; j < (int)n; j++) {
                const int32_t v = tile[lane][j];
                if (full_len && (uint32_t)v < (uint32_t)kWinQuot) q = s_q[v];
                else if (full_len) q = __ddiv_rn((double)v, L);                                        // depth >= 1024, or negative
                else if (!have || v != last) { q = __ddiv_rn((double)v, L); last = v; have = true; }   // the contig's last, shorter window
                a = __dadd_rn(a, q);
                sum += (unsigned long long)(long long)v;
                const uint32_t u = (uint32_t)v; mn = u < mn ? u : mn; mx = u > mx ? u : mx;
            }
            __syncwarp();
        }
        if (valid) {
            value_out[w] = a;
            len += (unsigned long long)my_len;
            long long bin = a >= 0 ? (long long)__dadd_rn(a, 0.5) : (long long)__dadd_rn(a, -0.5);   // toInt, :737-739
            if (bin > kWindowDistBins - 1) bin = kWindowDistBins - 1;
            if (bin >= 0) atomicAdd(&s_hist[bin], 1u);
        }
    }
    block_accumulate(acc, len, sum, mn, mx);
    __syncthreads();
    for (int i = threadIdx.x; i < kWindowDistBins; i += blockDim.x) { uint32_t c = s_hist[i]; if (c) atomicAdd(&dist512[i], (unsigned long long)c); }
}

// window mode, mean, EVERY contig in one launch (msd_prefetch_windows): the float64 chain makes one window a fixed
// ~100 us of latency whatever the grid, so 25 per-contig launches cost 25 x that; one launch over all contigs costs it once.
// Same per-window arithmetic as windows_mean_kernel.  A warp owns a group of 32 consecutive windows of ONE contig.
struct WinContig {
    unsigned long long arr_off;     // first cell of the contig in the depth arena
    long long first_group;          // index of the contig's first 32-window group among all groups
    long long first_window;         // index of its first window in value_out
    long long n_windows;
    uint32_t tlen, pad;
};
__global__ void __launch_bounds__(kWinWarps * 32)
windows_all_kernel(const int32_t* __restrict__ arena, const WinContig* __restrict__ contigs, int n_contigs, long long n_groups, uint32_t window,
                   double* __restrict__ value_out, RegionAccum* __restrict__ acc /*[n_contigs]*/, unsigned long long* __restrict__ dist512 /*[n_contigs][512]*/) {
    __shared__ int32_t s_tile[kWinWarps][32][33];
    __shared__ double s_q[kWinQuot];
    for (int i = threadIdx.x; i < kWinQuot; i += blockDim.x) s_q[i] = __ddiv_rn((double)i, (double)window);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t (*tile)[33] = s_tile[warp];
    for (long long g = (long long)blockIdx.x * kWinWarps + warp; g < n_groups; g += (long long)gridDim.x * kWinWarps) {
        int lo = 0, hi = n_contigs - 1;                                   // last contig with first_group <= g
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (contigs[mid].first_group <= g) lo = mid; else hi = mid - 1; }
        const WinContig C = contigs[lo];
        const int32_t* __restrict__ arr = arena + C.arr_off;
        const uint32_t tlen = C.tlen;
        const long long w0 = (g - C.first_group) * 32, w = w0 + lane;
        const bool valid = w < C.n_windows;
        const unsigned long long s64 = (unsigned long long)w * window;
        const unsigned long long e64 = (s64 + window < tlen) ? s64 + window : (unsigned long long)tlen;
        const long long my_len = valid ? (long long)(e64 - s64) : 0;
        const double L = (double)(uint32_t)my_len;
        const int n_rows = (int)((C.n_windows - w0) < 32 ? (C.n_windows - w0) : 32);
        double a = 0.0, q = 0.0; int32_t last = 0; bool have = false;
        const bool full_len = my_len == (long long)window;
        unsigned long long sum = 0; uint32_t mn = 0xffffffffu, mx = 0;
        const unsigned long long g0 = (unsigned long long)w0 * window;
        const uint32_t max_len = (uint32_t)(((unsigned long long)tlen - g0) < window ? ((unsigned long long)tlen - g0) : window);
        for (uint32_t c0 = 0; c0 < max_len; c0 += 32) {
            unsigned long long idx = g0 + c0 + lane;
            const bool col_ok = c0 + lane < window;
#pragma unroll 8
            for (int r = 0; r < n_rows; r++, idx += window) tile[r][lane] = (col_ok && idx < tlen) ? __ldg(arr + idx) : 0;
            __syncwarp();
            long long n = my_len - (long long)c0; if (n > 32) n = 32;
            for (int j = 0; j < (int)n; j++) {
                const int32_t v = tile[lane][j];
                if (full_len && (uint32_t)v < (uint32_t)kWinQuot) q = s_q[v];
                else if (full_len) q = __ddiv_rn((double)v, L);
                else if (!have || v != last) { q = __ddiv_rn((double)v, L); last = v; have = true; }
                a = __dadd_rn(a, q);
                sum += (unsigned long long)(long long)v;
                const uint32_t u = (uint32_t)v; mn = u < mn ? u : mn; mx = u > mx ? u : mx;
            }
            __syncwarp();
        }
        if (valid) {
            value_out[C.first_window + w] = a;
            long long bin = a >= 0 ? (long long)__dadd_rn(a, 0.5) : (long long)__dadd_rn(a, -0.5);
            if (bin > kWindowDistBins - 1) bin = kWindowDistBins - 1;
            if (bin >= 0) atomicAdd(&dist512[(size_t)lo * kWindowDistBins + bin], 1ull);
        }
        block_accumulate(acc + lo, (unsigned long long)my_len, sum, mn, mx);
    }
}

// BED mode (r2): one WARP per region (regions sorted by start on the host, mosdepth.nim:376-377).  r1 ran one thread per region:
// 3 + n_thr + up to 16 bisection passes of scalar loads 4 bytes apart between consecutive iterations and kilobases apart between
// lanes, one ~300-cycle division on the chain per depth change -- 1.12 ms for 4.8 k regions (profiles/r2a_regions_full_summary.txt,
// 12 % of the warps active).  Here the warp reads the region in coalesced chunks of 32 cells, ONE pass feeds everything that is
// order independent -- slice stats (:723-724), the per-base region dist (inc, :741), the threshold counts (:549-553: a ballot and
// a popc per threshold), the value range for the median -- and the one thing that is not, the float64 mean (Q1), keeps its exact
// chain: every lane divides its own cell (32 divisions at once instead of one after the other), then all lanes add the 32 quotients
// in index order (shuffle + __dadd_rn), so the chain costs an addition per cell and nothing else.  The median (CountStat,
// depthstat.nim:16-30) is the same bisection on the value range as before, each probe a coalesced pass + one warp reduction.
constexpr int kRegWarps = 8;
constexpr int kRegMaxThr = 16;      // thresholds counted in the main pass; more of them take extra passes of 16

__global__ void __launch_bounds__(kRegWarps * 32)
regions_kernel(const int32_t* __restrict__ arr, uint32_t tlen, long long n, const uint32_t* __restrict__ starts,
               const uint32_t* __restrict__ stops, int use_median, double* __restrict__ value_out, RegionAccum* __restrict__ acc,
               unsigned long long* __restrict__ dist, const int32_t* __restrict__ thresholds, int n_thr,
               unsigned long long* __restrict__ thr_counts) {
    __shared__ uint32_t s_hist[kHistSmemBins];
    for (int i = threadIdx.x; i < kHistSmemBins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t L = tlen + 1;   // len(arr), mosdepth.nim:274,719
    unsigned long long len = 0, sum = 0; uint32_t mn = 0xffffffffu, mx = 0;     // lane 0 keeps the warp's totals
    int32_t thr[kRegMaxThr];
    const int nt0 = n_thr < kRegMaxThr ? n_thr : kRegMaxThr;
#pragma unroll
    for (int j = 0; j < kRegMaxThr; j++) thr[j] = j < nt0 ? thresholds[j] : 0x7fffffff;
    for (long long r = (long long)blockIdx.x * kRegWarps + warp; r < n; r += (long long)gridDim.x * kRegWarps) {
        const uint32_t start = starts[r], stop = stops[r];
        const uint32_t e = stop < L ? stop : L;                              // imean / inc / thresholds stop at the array end (Q5)
        const bool in_arr = start < L;
        const double Ld = (double)(uint32_t)(stop - start);                  // the mean divides by the unclamped length (:411)
        double a = 0.0;
        unsigned long long rsum = 0; uint32_t rmn = 0xffffffffu, rmx = 0; uint32_t cnt[kRegMaxThr];
#pragma unroll
        for (int j = 0; j < kRegMaxThr; j++) cnt[j] = 0;
        int32_t vlo = 65535, vhi = 0; bool neg = false;
        if (in_arr && e > start) {
            for (uint32_t c0 = start; c0 < e; c0 += 32) {
                const uint32_t i = c0 + lane;
                const bool ok = i < e;
                const int32_t v = ok ? __ldg(arr + i) : 0;
                const int nn = (int)((e - c0) < 32u ? (e - c0) : 32u);
                if (!use_median) {                                           // :410-413, strictly in index order
                    const double q = __ddiv_rn((double)v, Ld);
                    for (int j = 0; j < nn; j++) a = __dadd_rn(a, __shfl_sync(0xffffffffu, q, j));
                }
                if (ok) {
                    rsum += (unsigned long long)(long long)v;                // slice stats: the clamped slice equals [start, e) here (:724)
                    const uint32_t u = (uint32_t)v; rmn = u < rmn ? u : rmn; rmx = u > rmx ? u : rmx;
                    hist_add(s_hist, dist, v, 1u);                           // inc(chrom_region_distribution, arr, start, stop) :741
                    int32_t w = v; if (w < 0) { neg = true; w = 0; } if (w > 65535) w = 65535;
                    vlo = w < vlo ? w : vlo; vhi = w > vhi ? w : vhi;
                }
#pragma unroll
                for (int j = 0; j < kRegMaxThr; j++) if (j < nt0) cnt[j] += __popc(__ballot_sync(0xffffffffu, ok && v >= thr[j]));
            }
        }
        // warp totals of the one pass
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rsum += __shfl_down_sync(0xffffffffu, rsum, o);
            const uint32_t x = __shfl_down_sync(0xffffffffu, rmn, o), y = __shfl_down_sync(0xffffffffu, rmx, o);
            rmn = x < rmn ? x : rmn; rmx = y > rmx ? y : rmx;
        }
        double me = a;
        if (use_median) {
            me = 0.0;
            const long long nn = in_arr ? (long long)e - (long long)start : 0;
            if (nn > 0) {
                if (__any_sync(0xffffffffu, neg)) { if (lane == 0) atomicExch(&acc->neg_seen, 1u); }
                int32_t lo = vlo, hi = vhi;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const int32_t x = __shfl_xor_sync(0xffffffffu, lo, o), y = __shfl_xor_sync(0xffffffffu, hi, o); lo = x < lo ? x : lo; hi = y > hi ? y : hi; }
                const long long stop_n = (long long)(0.5 + (double)nn * 0.5);     // depthstat.nim:24
                while (lo < hi) {                                                  // smallest x with #{v <= x} >= stop_n
                    const int32_t mid = lo + ((hi - lo) >> 1);
                    uint32_t c = 0;
                    for (uint32_t i = start + lane; i < e; i += 32) { int32_t v = __ldg(arr + i); if (v < 0) v = 0; if (v > 65535) v = 65535; c += (v <= mid); }
                    long long tot = (long long)__reduce_add_sync(0xffffffffu, c);
                    if (nn > 0xffffffffll / 32) { /* regions longer than 2^27 cells: the 32-bit reduction cannot overflow per lane, but widen the sum */
                        unsigned long long c64 = c;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) c64 += __shfl_xor_sync(0xffffffffu, c64, o);
                        tot = (long long)c64;
                    }
                    if (tot >= stop_n) hi = mid; else lo = mid + 1;
                }
                me = (double)lo;
            }
        }
        if (lane == 0) {
            value_out[r] = me;
            const uint32_t sa = start < L ? start : L, sb = L < stop ? L : stop;   // :724
            if (sb > sa) { len += sb - sa; sum += rsum; mn = rmn < mn ? rmn : mn; mx = rmx > mx ? rmx : mx; }
#pragma unroll
            for (int j = 0; j < kRegMaxThr; j++) if (j < nt0) thr_counts[r * n_thr + j] = cnt[j];
        }
        // thresholds beyond the first kRegMaxThr: one more coalesced pass per group
        for (int j0 = kRegMaxThr; j0 < n_thr; j0 += kRegMaxThr) {
            const int m = (n_thr - j0) < kRegMaxThr ? (n_thr - j0) : kRegMaxThr;
            uint32_t c2[kRegMaxThr];
#pragma unroll
            for (int j = 0; j < kRegMaxThr; j++) c2[j] = 0;
            if (in_arr) for (uint32_t c0 = start; c0 < e; c0 += 32) {
                const uint32_t i = c0 + lane; const bool ok = i < e; const int32_t v = ok ? __ldg(arr + i) : 0;
#pragma unroll
                for (int j = 0; j < kRegMaxThr; j++) if (j < m) c2[j] += __popc(__ballot_sync(0xffffffffu, ok && v >= thresholds[j0 + j]));
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < kRegMaxThr; j++) if (j < m) thr_counts[r * n_thr + j0 + j] = c2[j];
            }
        }
    }
    if (lane != 0) { len = 0; sum = 0; mn = 0xffffffffu; mx = 0; }
    block_accumulate(acc, len, sum, mn, mx);
    __syncthreads();
    for (int i = threadIdx.x; i < kHistSmemBins; i += blockDim.x) { uint32_t c = s_hist[i]; if (c) atomicAdd(&dist[i], (unsigned long long)c); }
}

}  // namespace msd
