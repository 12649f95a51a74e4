// scan.cuh -- K6 scan_segments: int32 inclusive prefix sum of EVERY contig's array in one launch (to_coverage -> hts-nim
// cumsum, mosdepth.nim:54-56, which the reference runs once per contig inside `for target in sub_targets`, :691,:713),
// fused with the two whole-contig consumers that are order independent: inc(chrom_global_distribution, arr, 0, len-1)
// (:417-437,:746) and newDepthStat(arr[0..<len-1]) (:747, depthstat.nim:39-45).
//
// A *segment* is a contiguous piece of the depth arena that gets its own prefix sum: a whole contig (tlen + 1 cells, the
// first tlen of which feed the stats) or the owned range of a contig that is sharded across GPUs.  All segments are cut
// into 4096-cell tiles that are numbered through (segment s owns tiles [first_tile, first_tile + ceil(n_scan / 4096)));
// persistent CTAs take tile numbers from an atomic ticket, so the decoupled look-back never waits for a tile that has not
// started, and it simply stops at the first tile of its own segment.  Single pass, 128-bit coalesced loads and stores,
// warp-wide look-back (32 predecessors per probe).
//
// HBM traffic per base: 4 B read + 4 B written (the depth array is kept for the window / region / run kernels).
// r1 ran one launch + one 3.2 MB histogram memset + one 3.2 MB histogram add per contig (25 x 3 launches for a human genome);
// this is one launch for the genome, and the per-contig results stay on the device until one asynchronous copy.
#pragma once
#include "common.cuh"

namespace msd {

constexpr int kScanThreads = 256;
constexpr int kScanVecPerThread = 4;                                   // int4 per thread
constexpr int kScanTile = kScanThreads * kScanVecPerThread * 4;        // 4096 cells = 16 KB
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
// same rule, for a per-segment table of kHistKeep bins: deeper values are not counted here -- the segment's max_depth tells the
// host that it has to count this one contig again with the full-width table (range_stats_kernel)
MSD_D void hist_add_keep(uint32_t* __restrict__ sh, unsigned long long* __restrict__ gh, int32_t v, uint32_t n) {
    if (v < 0) return;
    if (v < kHistSmemBins) atomicAdd(&sh[v], n);
    else if (v < kHistKeep) atomicAdd(&gh[v], (unsigned long long)n);
}

struct SegStats { unsigned long long sum; uint32_t mn, mx; };

// block-reduce the per-thread stats into acc[slot], flush the shared histogram into hist_all[slot]; all threads call it
MSD_D void seg_flush(SegStats& my, uint32_t* __restrict__ s_hist, unsigned long long* __restrict__ s_sum, uint32_t* __restrict__ s_mn, uint32_t* __restrict__ s_mx,
                     ContigAccum* __restrict__ acc, unsigned long long* __restrict__ hist) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
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
    for (int i = t; i < kHistSmemBins; i += kScanThreads) { const uint32_t c = s_hist[i]; if (c) { atomicAdd(&hist[i], (unsigned long long)c); s_hist[i] = 0; } }
    my.sum = 0; my.mn = 0xffffffffu; my.mx = 0;
    __syncthreads();
}

// the stats of four consecutive cells starting at segment index idx (cells at or beyond n_stat do not count).  Equal neighbouring cells
// of a lane share one histogram atomic.  (r2e tried grouping the lanes of a warp by value with match.any, one atomic per group: 2.13 TB/s at
// f = 0.12 against 2.41 / 2.70 TB/s at f = 0.06 / 0.5 for this version -- no gain, dropped.)
MSD_D void seg_stats4(const int4 x, uint32_t idx, uint32_t n_stat, SegStats& my, uint32_t* __restrict__ s_hist, unsigned long long* __restrict__ hist) {
    const int32_t e[4] = {x.x, x.y, x.z, x.w};
    int32_t cur = 0; uint32_t cnt = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (idx + q >= n_stat) break;
        const int32_t d = e[q];
        my.sum += (unsigned long long)(long long)d;          // dp.uint64 (depthstat.nim:43)
        const uint32_t u = (uint32_t)d;
        my.mn = u < my.mn ? u : my.mn; my.mx = u > my.mx ? u : my.mx;
        if (cnt && d == cur) cnt++;
        else { if (cnt) hist_add_keep(s_hist, hist, cur, cnt); cur = d; cnt = 1; }
    }
    if (cnt) hist_add_keep(s_hist, hist, cur, cnt);
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
template <bool kScan>
__global__ void __launch_bounds__(kScanThreads, 4)
scan_segments_kernel(int32_t* __restrict__ arena, const ScanSeg* __restrict__ segs, int n_segs, uint32_t n_tiles,
                     unsigned long long* __restrict__ tile_state, uint32_t* __restrict__ ticket,
                     ContigAccum* __restrict__ acc_all, unsigned long long* __restrict__ hist_all /* [slot][kHistKeep] */) {
    __shared__ uint32_t s_hist[kHistSmemBins];
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ uint32_t s_tile, s_prefix;
    __shared__ unsigned long long s_sum[kScanThreads / 32];
    __shared__ uint32_t s_mn[kScanThreads / 32], s_mx[kScanThreads / 32];
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
        const uint32_t n_scan = S.n_scan, rel = tile - S.first_tile, tile_base = rel * kScanTile;
        unsigned long long* __restrict__ hist = hist_all + (size_t)S.slot * kHistKeep;
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
            if (!kScan) { v[j] = x; continue; }
            // local inclusive scan of the 4 cells (uint32 wrap-around like the reference's int32 cumsum)
            uint32_t a = (uint32_t)x.x, b = a + (uint32_t)x.y, c = b + (uint32_t)x.z, d = c + (uint32_t)x.w;
            uint32_t incl = d;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            const uint32_t excl = incl - d + run;
            v[j] = make_int4((int)(a + excl), (int)(b + excl), (int)(c + excl), (int)(d + excl));
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        uint32_t add = 0;
        if (kScan) {
            if (lane == 0) s_warp[warp] = run;
            __syncthreads();
            uint32_t warp_excl = 0, tile_total = 0;
#pragma unroll
            for (int w = 0; w < kScanThreads / 32; w++) { uint32_t x = s_warp[w]; if (w < warp) warp_excl += x; tile_total += x; }
            if (warp == 0) {
                const uint32_t prefix = look_back<1>(tile_state, tile, S.first_tile, tile_total, lane);
                if (lane == 0) s_prefix = prefix;
            }
            __syncthreads();
            add = s_prefix + warp_excl;
        }
#pragma unroll
        for (int j = 0; j < kScanVecPerThread; j++) {
            const uint32_t idx = tile_base + warp * 512 + j * 128 + lane * 4;
            const int4 x = make_int4((int)((uint32_t)v[j].x + add), (int)((uint32_t)v[j].y + add), (int)((uint32_t)v[j].z + add), (int)((uint32_t)v[j].w + add));
            if (kScan) {
                if (idx + 3 < n_scan) *reinterpret_cast<int4*>(arr + idx) = x;
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

// ---- K6, pipelined (r2j) ---------------------------------------------------------------------------------------------------------------
// The kernel above pays every latency of a tile in sequence -- ticket, segment table, the tile's loads, the look-back, the stores and
// stats -- so a CTA has loads in flight for ~1 us of the ~7 us a tile takes and four CTAs per SM keep ~9 KB of reads outstanding where the
// HBM rate needs ~35 KB (r2c: 37-41 % of the measured peak).  Here each CTA owns a ring of kScanStages 16-KB tiles in shared memory that one
// thread fills with bulk asynchronous copies (cp.async.bulk global -> shared, completion counted on an mbarrier), kScanStages tiles ahead of
// the tile the CTA is scanning; the ticket of the tile after those is already on its way.  The rest -- int4 per thread, local scan, warp-wide
// look-back, stores and stats from registers -- is the arithmetic of the kernel above, unchanged.
// Deadlock freedom is the ticket argument again: a CTA's tickets ascend and it scans them in order, so the smallest unfinished tile of the
// launch is always the current tile of its CTA and never waits.
constexpr int kScanStages = 3;
constexpr int kScanProducer = 32;      // thread that takes tickets and issues the copies (lane 0 of warp 1: warp 0 walks the look-back)
constexpr size_t kScanPipeSmem = (size_t)kScanStages * kScanTile * 4;

MSD_D void pipe_bar_init(unsigned long long* bar) {
#if defined(__CUDACC__)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)));
#else
    (void)bar;
#endif
}
// the source must be 16-byte aligned and `bytes` a multiple of 16
MSD_D void pipe_issue(void* dst_smem, const void* src, uint32_t bytes, unsigned long long* bar) {
#if defined(__CUDACC__)
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
#else
    (void)bar; memcpy(dst_smem, src, bytes);      // the emulation copies at once; a __syncthreads lies between every issue and the first read
#endif
}
MSD_D void pipe_wait(unsigned long long* bar, uint32_t parity) {
#if defined(__CUDACC__)
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
#else
    (void)bar; (void)parity;
#endif
}
// the segment that owns `tile`: the last one with first_tile <= tile
MSD_D int seg_of_tile(const ScanSeg* __restrict__ segs, int n_segs, uint32_t tile, int hint) {
    if (hint >= 0 && tile >= segs[hint].first_tile && (hint + 1 >= n_segs || tile < segs[hint + 1].first_tile)) return hint;
    int lo = 0, hi = n_segs - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (segs[mid].first_tile <= tile) lo = mid; else hi = mid - 1; }
    return lo;
}

template <int kGroups>
__global__ void __launch_bounds__(kScanThreads, 4)
scan_segments_pipe_kernel(int32_t* __restrict__ arena, const ScanSeg* __restrict__ segs, int n_segs, uint32_t n_tiles,
                          unsigned long long* __restrict__ tile_state, uint32_t* __restrict__ ticket,
                          ContigAccum* __restrict__ acc_all, unsigned long long* __restrict__ hist_all /* [slot][kHistKeep] */) {
#if defined(__CUDACC__)
    extern __shared__ __align__(128) int4 s_ring[];                              // kScanStages tiles of 1024 int4
#else
    __shared__ int4 s_ring[kScanStages * (kScanTile / 4)];
#endif
    __shared__ unsigned long long s_full[kScanStages];
    __shared__ uint32_t s_slot_tile[kScanStages];
    __shared__ uint32_t s_hist[kHistSmemBins];
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ uint32_t s_prefix;
    __shared__ unsigned long long s_sum[kScanThreads / 32];
    __shared__ uint32_t s_mn[kScanThreads / 32], s_mx[kScanThreads / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < kHistSmemBins; i += kScanThreads) s_hist[i] = 0;
    SegStats my{0, 0xffffffffu, 0};
    int cur_seg = -1; bool cur_stats = false; uint32_t cur_slot = 0;
    // producer state (thread kScanProducer only): the ticket that has no copy yet, and the segment of the last tile it looked up
    uint32_t pending = 0; int prod_seg = -1;
    auto fill = [&](int slot, uint32_t tk) {         // producer: tile tk goes to ring slot `slot` (nothing to copy once the tickets run out)
        s_slot_tile[slot] = tk;
        if (tk >= n_tiles) return;
        prod_seg = seg_of_tile(segs, n_segs, tk, prod_seg);
        const ScanSeg P = segs[prod_seg];
        const uint32_t base = (tk - P.first_tile) * kScanTile, cells = P.n_scan - base < (uint32_t)kScanTile ? P.n_scan - base : (uint32_t)kScanTile;
        pipe_issue(s_ring + (size_t)slot * (kScanTile / 4), arena + P.arr_off + base, (cells * 4u + 15u) & ~15u, &s_full[slot]);
    };
    if (t == kScanProducer) {
        for (int k = 0; k < kScanStages; k++) pipe_bar_init(&s_full[k]);
#if defined(__CUDACC__)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
        for (int k = 0; k < kScanStages; k++) fill(k, atomicAdd(ticket, 1u));
        pending = atomicAdd(ticket, 1u);
    }
    __syncthreads();
    for (uint32_t it = 0;; it++) {
        const int slot = (int)(it % kScanStages);
        const uint32_t tile = s_slot_tile[slot];
        if (tile >= n_tiles) break;                     // a CTA's tickets ascend: the later slots hold nothing either
        const int sg = seg_of_tile(segs, n_segs, tile, cur_seg);
        if (sg != cur_seg) {
            if (cur_stats) seg_flush(my, s_hist, s_sum, s_mn, s_mx, acc_all + cur_slot, hist_all + (size_t)cur_slot * kHistKeep);
            cur_seg = sg; cur_stats = segs[sg].n_stat != 0; cur_slot = segs[sg].slot;
        }
        const ScanSeg S = segs[sg];
        int32_t* __restrict__ arr = arena + S.arr_off;
        const uint32_t n_scan = S.n_scan, rel = tile - S.first_tile, tile_base = rel * kScanTile;
        unsigned long long* __restrict__ hist = hist_all + (size_t)S.slot * kHistKeep;
        pipe_wait(&s_full[slot], (it / kScanStages) & 1u);
        const int4* __restrict__ ring = s_ring + (size_t)slot * (kScanTile / 4);
        // warp w owns cells [tile_base + w*512, +512): 4 rounds of 32 lanes x int4 (what lies beyond n_scan in the slot is stale: masked)
        int4 v[kScanVecPerThread];
        uint32_t run = 0;
#pragma unroll
        for (int j = 0; j < kScanVecPerThread; j++) {
            const uint32_t off = warp * 512 + j * 128 + lane * 4, idx = tile_base + off;
            int4 x = ring[off >> 2];
            if (idx + 3 >= n_scan) { if (idx >= n_scan) x.x = 0; if (idx + 1 >= n_scan) x.y = 0; if (idx + 2 >= n_scan) x.z = 0; x.w = 0; }
            uint32_t a = (uint32_t)x.x, b = a + (uint32_t)x.y, c = b + (uint32_t)x.z, d = c + (uint32_t)x.w;
            uint32_t incl = d;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            const uint32_t excl = incl - d + run;
            v[j] = make_int4((int)(a + excl), (int)(b + excl), (int)(c + excl), (int)(d + excl));
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_warp[warp] = run;
        __syncthreads();                                // every thread has read its part of the slot: it may be filled again
        if (t == kScanProducer) { fill(slot, pending); pending = atomicAdd(ticket, 1u); }
        uint32_t warp_excl = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; w++) { uint32_t x = s_warp[w]; if (w < warp) warp_excl += x; tile_total += x; }
        if (warp == 0) {
            const uint32_t prefix = look_back<kGroups>(tile_state, tile, S.first_tile, tile_total, lane);
            if (lane == 0) s_prefix = prefix;
        }
        __syncthreads();
        const uint32_t add = s_prefix + warp_excl;
#pragma unroll
        for (int j = 0; j < kScanVecPerThread; j++) {
            const uint32_t idx = tile_base + warp * 512 + j * 128 + lane * 4;
            const int4 x = make_int4((int)((uint32_t)v[j].x + add), (int)((uint32_t)v[j].y + add), (int)((uint32_t)v[j].z + add), (int)((uint32_t)v[j].w + add));
            if (idx + 3 < n_scan) *reinterpret_cast<int4*>(arr + idx) = x;
            else {
                if (idx < n_scan) arr[idx] = x.x;
                if (idx + 1 < n_scan) arr[idx + 1] = x.y;
                if (idx + 2 < n_scan) arr[idx + 2] = x.z;
            }
            if (cur_stats) seg_stats4(x, idx, S.n_stat, my, s_hist, hist);
        }
        // (no barrier here: s_warp is written again after the next tile's first barrier-free stretch only by threads that have passed the
        // second barrier above, and s_prefix by warp 0 after the next first barrier, which every thread reaches after reading s_prefix)
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
