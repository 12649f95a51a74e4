// reduce_emu.cpp -- K6 (scan.cuh: scan_segments_kernel, totals_add_kernel) and K7 (regions.cuh: windows_mean_kernel, windows_all_kernel, windows_kernel, regions_kernel)
// executed on the CPU under tests/emu/cuda_runtime.h with the launch geometry of run_scan / msd_windows / msd_regions
// (mosdepth_b200/csrc/msd_api.cu), against a plain restatement of the reference's arithmetic:
//   to_coverage mosdepth.nim:54-56 (int32 prefix sum, wraps), inc :417-437, newDepthStat depthstat.nim:39-45, imean :399-413
//   (sequential float64 chain / CountStat median depthstat.nim:16-30), window bins :737-739, thresholds :549-553.
// Float means must be BIT-identical (Q1).  TEST INFRASTRUCTURE ONLY; prints "ok ..." or the first difference and exits non-zero.
//   usage: reduce_emu <tlen> <window> <max_ctas> [seed]
#include <cinttypes>
#include <cstdlib>
#include <string>
#include "../../mosdepth_b200/csrc/scan.cuh"
#include "../../mosdepth_b200/csrc/regions.cuh"
#include "../../mosdepth_b200/csrc/runs.cuh"

using namespace msd;
typedef unsigned long long ull;

static unsigned grid_for(uint64_t n, int threads, int max_ctas) { uint64_t g = (n + threads - 1) / threads; if (g < 1) g = 1; if (g > (uint64_t)max_ctas) g = max_ctas; return (unsigned)g; }
static int fail(const char* what, long long at, double a, double b) { printf("DIFF %s at %lld: kernel %.17g, expected %.17g\n", what, at, a, b); return 1; }
static int64_t nim_to_int(double f) { return f >= 0 ? (int64_t)(f + 0.5) : (int64_t)(f - 0.5); }

struct Stat { ull len = 0, sum = 0; uint32_t mn = 0xffffffffu, mx = 0; };
static void stat_add(Stat& s, const int32_t* v, int64_t n) { s.len += (ull)(n > 0 ? n : 0); for (int64_t i = 0; i < n; i++) { const uint32_t x = (uint32_t)v[i]; s.sum += x; if (x < s.mn) s.mn = x; if (x > s.mx) s.mx = x; } }
static double imean(const int32_t* arr, uint32_t len, uint32_t start, uint32_t stop, bool median) {
    if (start > len) return 0;
    const uint32_t e = stop < len ? stop : len;
    if (median) {
        static std::vector<uint32_t> cnt(65536); std::fill(cnt.begin(), cnt.end(), 0u); int64_t n = 0;
        for (uint32_t i = start; i < e; i++) { const int32_t v = arr[i]; cnt[v > 65535 ? 65535 : v]++; n++; }
        if (n == 0) return 0;
        const int64_t stop_n = (int64_t)(0.5 + (double)n * 0.5); int64_t cum = 0;
        for (int i = 0; i < 65536; i++) { cum += cnt[i]; if (cum >= stop_n) return (double)i; }
        return 0;
    }
    const double L = (double)(uint32_t)(stop - start); volatile double r = 0;
    for (uint32_t i = start; i < e; i++) { volatile double q = (double)arr[i] / L; r = r + q; }
    return r;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: reduce_emu tlen window max_ctas [seed]\n"); return 2; }
    const uint32_t tlen = (uint32_t)strtoul(argv[1], nullptr, 10), window = (uint32_t)strtoul(argv[2], nullptr, 10); const int max_ctas = atoi(argv[3]);
    uint64_t x = 0x9e3779b97f4a7c15ull ^ (argc > 4 ? strtoull(argv[4], nullptr, 10) : 0);
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    // +1/-1 deltas of reads of 30..150 bases (depth stays >= 0), a few very deep spots (bins beyond the shared-memory histogram)
    const uint32_t L = tlen + 1;
    int32_t* arr = nullptr; if (posix_memalign((void**)&arr, 128, ((size_t)L + 64) * 4)) return 2;
    memset(arr, 0, ((size_t)L + 64) * 4);
    const uint64_t n_reads = (uint64_t)tlen * 30 / 100 + 3;
    for (uint64_t r = 0; r < n_reads; r++) { const uint32_t s = (uint32_t)(rnd() % tlen), len = 30 + (uint32_t)(rnd() % 121); const uint32_t e = s + len < tlen ? s + len : tlen; arr[s] += 1; arr[e] -= 1; }
    for (int k = 0; k < 3 && tlen > 1000; k++) { const uint32_t s = (uint32_t)(rnd() % (tlen - 600)); arr[s] += 2000 + 70000 * k; arr[s + 500] -= 2000 + 70000 * k; }
    std::vector<int32_t> want(arr, arr + L);
    { uint32_t s = 0; for (uint32_t i = 0; i < L; i++) { s += (uint32_t)want[i]; want[i] = (int32_t)s; } }
    // ---- K6: run_segments(): three segments in one launch -- this contig (tlen + 1 cells scanned, tlen feed the stats), a copy of its
    //      first third as a second contig, and a shard-like segment without stats; then the stats-only pass (msd_contig_finalize) ----
    {
        const uint32_t n2 = tlen / 3 + 1, n3 = std::min<uint32_t>(tlen, 5000) + 1;
        std::vector<ull> off(3); ull cells = 0; const uint32_t ns[3] = {L, n2 + 1, n3};
        for (int c = 0; c < 3; c++) { off[c] = cells; cells += ((ull)ns[c] + 31) / 32 * 32; }
        int32_t* arena = nullptr; if (posix_memalign((void**)&arena, 128, (cells + 64) * 4)) return 2;
        memset(arena, 0, (cells + 64) * 4);
        for (int c = 0; c < 3; c++) memcpy(arena + off[c], arr, (size_t)ns[c] * 4);          // deltas (a prefix of a delta array is a delta array)
        std::vector<ScanSeg> segs(3); uint32_t tiles = 0;
        for (int c = 0; c < 3; c++) { segs[c].arr_off = off[c]; segs[c].n_scan = ns[c]; segs[c].n_stat = c == 0 ? tlen : c == 1 ? n2 : 0; segs[c].first_tile = tiles; segs[c].slot = (uint32_t)c; tiles += (ns[c] + kScanTile - 1) / kScanTile; }
        std::vector<ull> tile_state(tiles, 0), hist_all(3 * (size_t)kHistKeep, 0); uint32_t ticket = 0;
        std::vector<ContigAccum> acc(3, ContigAccum{0, 0xffffffffu, 0, 0, 0});
        // the other instantiations of the scan first, on copies of the deltas (wide look-back; 32-KB tiles with the L2 prefetch): same arrays, same stats
        {
            std::vector<int32_t> deltas(arena, arena + cells);
            for (int variant : {2, 6}) {
                const uint32_t tc = variant == 6 ? (uint32_t)scan_tile_cells(8) : (uint32_t)kScanTile;
                std::vector<ScanSeg> sv = segs; uint32_t tl = 0;
                for (int c = 0; c < 3; c++) { sv[c].first_tile = tl; tl += (ns[c] + tc - 1) / tc; }
                std::vector<ull> ts2(tl, 0), h2(3 * (size_t)kHistKeep, 0); uint32_t tk2 = 0; std::vector<ContigAccum> a2(3, ContigAccum{0, 0xffffffffu, 0, 0, 0});
                const unsigned g = (unsigned)std::min<uint64_t>(tl, (uint64_t)max_ctas);
                if (variant == 2) emu::launch(g, kScanThreads, [&]() { scan_segments_kernel<true, 4, 4, false>(arena, sv.data(), 3, tl, ts2.data(), &tk2, a2.data(), h2.data()); });
                else emu::launch(g, kScanThreads, [&]() { scan_segments_kernel<true, 1, 8, true>(arena, sv.data(), 3, tl, ts2.data(), &tk2, a2.data(), h2.data()); });
                for (int c = 0; c < 3; c++) for (uint32_t i = 0; i < ns[c]; i++) if (arena[off[c] + i] != want[i]) return fail("prefix sum (scan variant)", (long long)c * 1000000000ll + i, arena[off[c] + i], want[i]);
                for (int c = 0; c < 3; c++) {
                    const uint32_t nst = segs[c].n_stat;
                    std::vector<ull> h(kHistKeep, 0); Stat st;
                    for (uint32_t i = 0; i < nst; i++) { const int32_t v = want[i]; if (v >= 0 && v < kHistKeep) h[v]++; }
                    stat_add(st, want.data(), nst);
                    if (nst == 0) { st.mn = 0xffffffffu; st.mx = 0; }
                    for (int b = 0; b < kHistKeep; b++) if (h[b] != h2[(size_t)c * kHistKeep + b]) return fail("segment dist bin (scan variant)", (long long)c * 1000000 + b, (double)h2[(size_t)c * kHistKeep + b], (double)h[b]);
                    if (a2[c].cum_depth != st.sum || a2[c].min_depth != st.mn || a2[c].max_depth != st.mx) return fail("segment stats (scan variant)", c, (double)a2[c].cum_depth, (double)st.sum);
                }
                memcpy(arena, deltas.data(), cells * 4);
            }
        }
        emu::launch((unsigned)std::min<uint64_t>(tiles, (uint64_t)max_ctas), kScanThreads, [&]() { scan_segments_kernel<true>(arena, segs.data(), 3, tiles, tile_state.data(), &ticket, acc.data(), hist_all.data()); });
        for (int c = 0; c < 3; c++) for (uint32_t i = 0; i < ns[c]; i++) if (arena[off[c] + i] != want[i]) return fail("prefix sum (segment)", (long long)c * 1000000000ll + i, arena[off[c] + i], want[i]);
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) {   // the stats-only pass over the same (now depth) arrays, taken by stride: must give the same numbers
                std::fill(hist_all.begin(), hist_all.end(), 0ull); for (auto& a : acc) a = ContigAccum{0, 0xffffffffu, 0, 0, 0};
                emu::launch((unsigned)std::min<uint64_t>(tiles, (uint64_t)std::max(1, max_ctas / 2)), kScanThreads, [&]() { scan_segments_kernel<false>(arena, segs.data(), 3, tiles, tile_state.data(), &ticket, acc.data(), hist_all.data()); });
                for (int c = 0; c < 3; c++) for (uint32_t i = 0; i < ns[c]; i++) if (arena[off[c] + i] != want[i]) return fail("stats pass changed the array", i, arena[off[c] + i], want[i]);
            }
            for (int c = 0; c < 3; c++) {
                const uint32_t nst = segs[c].n_stat;
                std::vector<ull> h(kHistKeep, 0); Stat st;
                for (uint32_t i = 0; i < nst; i++) { const int32_t v = want[i]; if (v >= 0 && v < kHistKeep) h[v]++; }
                stat_add(st, want.data(), nst);
                if (nst == 0) { st.mn = 0xffffffffu; st.mx = 0; }
                for (int b = 0; b < kHistKeep; b++) if (h[b] != hist_all[(size_t)c * kHistKeep + b]) return fail(pass ? "segment dist bin (stats pass)" : "segment dist bin", (long long)c * 1000000 + b, (double)hist_all[(size_t)c * kHistKeep + b], (double)h[b]);
                if (acc[c].cum_depth != st.sum || acc[c].min_depth != st.mn || acc[c].max_depth != st.mx) return fail("segment stats (sum)", c, (double)acc[c].cum_depth, (double)st.sum);
            }
        }
        // sum_into for the segments: bins up to each segment's maximum (segments deeper than kHistKeep are left to the host)
        std::vector<long long> tot(kDistBins, 0), ts = {0, 0, 0xffffffffll, 0};
        emu::launch(kTotalsParts * 3, 256, [&]() { totals_add_kernel(segs.data(), 3, acc.data(), hist_all.data(), tot.data(), ts.data()); });
        { std::vector<long long> h(kDistBins, 0); long long len = 0, sum = 0, mn = 0xffffffffll, mx = 0;
          for (int c = 0; c < 3; c++) { if (!segs[c].n_stat || acc[c].max_depth >= (uint32_t)kHistKeep) continue; for (uint32_t i = 0; i < segs[c].n_stat; i++) if (want[i] >= 0) h[want[i]]++; len += segs[c].n_stat; sum += (long long)acc[c].cum_depth; mn = std::min<long long>(mn, acc[c].min_depth); mx = std::max<long long>(mx, acc[c].max_depth); }
          for (int b = 0; b < kDistBins; b++) if (h[b] != tot[b]) return fail("total dist bin", b, (double)tot[b], (double)h[b]);
          if (ts[0] != len || ts[1] != sum || ts[2] != mn || ts[3] != mx) return fail("total stats", 0, (double)ts[1], (double)sum); }
        memcpy(arr, arena, (size_t)L * 4);      // the kernels below read the contig's depth array
        free(arena);
    }
    const uint64_t n_tiles = ((uint64_t)L + kScanTile - 1) / kScanTile;
    // ---- K7 windows: msd_windows(mean) and msd_windows(median) ----------------------------------------------------------------------
    const long long nw = tlen ? ((long long)tlen + window - 1) / window : 0;
    for (int median = 0; median < 2; median++) {
        std::vector<double> val((size_t)nw + 1, -1.0); RegionAccum ra{0, 0, 0xffffffffu, 0, 0, 0}; std::vector<ull> d512(kWindowDistBins, 0);
        if (median) emu::launch(grid_for((uint64_t)nw, 256, max_ctas), 256, [&]() { windows_kernel(arr, tlen, window, nw, 1, val.data(), &ra, d512.data()); });
        else emu::launch(grid_for((uint64_t)nw, kWinWarps * 32, std::max(1, max_ctas / 2)), kWinWarps * 32, [&]() { windows_mean_kernel(arr, tlen, window, nw, val.data(), &ra, d512.data()); });
        std::vector<ull> w512(kWindowDistBins, 0); Stat st;
        for (long long k = 0; k < nw; k++) {
            const uint32_t s = (uint32_t)((ull)k * window), e = (uint32_t)std::min<ull>(((ull)k + 1) * window, tlen);
            const double me = imean(want.data(), L, s, e, median != 0);
            if (memcmp(&me, &val[(size_t)k], 8) != 0) return fail(median ? "window median" : "window mean (bit pattern)", k, val[(size_t)k], me);
            const int64_t bin = nim_to_int(me); w512[bin < 511 ? bin : 511]++;
            const uint32_t a = s < L ? s : L, b = L < e ? L : e; stat_add(st, want.data() + a, b > a ? (int64_t)(b - a) : 0);
        }
        for (int b = 0; b < kWindowDistBins; b++) if (w512[b] != d512[b]) return fail("window dist bin", b, (double)d512[b], (double)w512[b]);
        if (ra.cum_length != st.len || ra.cum_depth != st.sum || ra.min_depth != st.mn || ra.max_depth != st.mx) return fail("window region stats (sum)", median, (double)ra.cum_depth, (double)st.sum);
    }
    // ---- K7 regions: msd_regions with thresholds, mean and median; intervals of 1..3000 bases, some reaching past the contig end (Q5) -----
    const long long nr = std::min<long long>(300, (long long)tlen / 150 + 5);      // (a warp-per-region kernel costs ~1,000 barrier exchanges per region under the stand-in)
    std::vector<uint32_t> rs((size_t)nr), re((size_t)nr);
    for (long long k = 0; k < nr; k++) { const uint32_t s = (uint32_t)(rnd() % ((ull)tlen + 40)), len = 1 + (uint32_t)(rnd() % (k % 7 == 0 ? 3000 : 300)); rs[(size_t)k] = s; re[(size_t)k] = s + len; }
    // 4 thresholds (the C4 case), or 18 on odd contig lengths: more than the kernel counts in its main pass (kRegMaxThr)
    const int32_t thr_all[18] = {1, 10, 20, 30, 31, 32, 33, 34, 35, 36, 37, 38, 40, 45, 50, 100, 1000, 50000}; const int32_t* thr = thr_all; const int NT = (tlen & 1) ? 18 : 4;
    for (int median = 0; median < 2; median++) {
        std::vector<double> val((size_t)nr, -1.0); RegionAccum ra{0, 0, 0xffffffffu, 0, 0, 0}; std::vector<ull> dist(kDistBins, 0), cnt((size_t)nr * NT, 0);
        emu::launch(grid_for((uint64_t)nr, kRegWarps, max_ctas), kRegWarps * 32, [&]() { regions_kernel(arr, tlen, nr, rs.data(), re.data(), median, val.data(), &ra, dist.data(), thr, NT, cnt.data()); });
        std::vector<ull> h(kDistBins, 0); Stat st;
        for (long long k = 0; k < nr; k++) {
            const uint32_t s = rs[(size_t)k], e = re[(size_t)k];
            const double me = imean(want.data(), L, s, e, median != 0);
            if (memcmp(&me, &val[(size_t)k], 8) != 0) return fail(median ? "region median" : "region mean (bit pattern)", k, val[(size_t)k], me);
            const uint32_t a = s < L ? s : L, b = L < e ? L : e; stat_add(st, want.data() + a, b > a ? (int64_t)(b - a) : 0);
            if (s < L) for (uint32_t i = s; i < (e < L ? e : L); i++) { int32_t v = want[i]; if (v > (int32_t)kMaxCoverage) v = kMaxCoverage - 10; if (v >= 0) h[v]++; }      // inc, :417-437
            for (uint32_t i = s; i < (e < L ? e : L); i++) for (int j = 0; j < NT; j++) { if (want[i] < thr[j]) break; if (--cnt[(size_t)k * NT + j] == ~0ull) return fail("threshold count", k, -1, 0); }
        }
        for (size_t i = 0; i < cnt.size(); i++) if (cnt[i]) return fail("threshold count", (long long)(i / NT), (double)(long long)cnt[i], 0);
        for (int b = 0; b < kDistBins; b++) if (h[b] != dist[b]) return fail("region dist bin", b, (double)dist[b], (double)h[b]);
        if (ra.cum_length != st.len || ra.cum_depth != st.sum || ra.min_depth != st.mn || ra.max_depth != st.mx) return fail("region stats (length)", median, (double)ra.cum_length, (double)st.len);
    }
    // ---- K7 windows_all_kernel (msd_prefetch_windows: every contig's windows in ONE launch, the kernel of the --by <int> -n -x path) ----------
    {   // an arena of three contigs: this one, its first third, and a 1-base contig; each padded to 128 bytes like the library's arena
        const uint32_t tl[3] = {tlen, tlen / 3 + 1, 1};
        std::vector<ull> off(3); ull cells = 0;
        for (int c = 0; c < 3; c++) { off[c] = cells; cells += (((ull)tl[c] + 1) * 4 + 127) / 128 * 32; }
        int32_t* arena = nullptr; if (posix_memalign((void**)&arena, 128, (cells + 64) * 4)) return 2;
        memset(arena, 0, (cells + 64) * 4);
        for (int c = 0; c < 3; c++) for (uint32_t i = 0; i <= tl[c] && i < L; i++) arena[off[c] + i] = want[i];      // (the sentinel cell of a prefix is a depth like any other)
        std::vector<WinContig> table; long long groups = 0, windows = 0;
        for (int c = 0; c < 3; c++) { const long long n = ((long long)tl[c] + window - 1) / window; WinContig w; w.arr_off = off[c]; w.first_group = groups; w.first_window = windows; w.n_windows = n; w.tlen = tl[c]; w.pad = 0; table.push_back(w); groups += (n + 31) / 32; windows += n; }
        std::vector<double> val((size_t)windows, -1.0); std::vector<RegionAccum> ra(3, RegionAccum{0, 0, 0xffffffffu, 0, 0, 0}); std::vector<ull> d512(3 * kWindowDistBins, 0);
        const unsigned grid = (unsigned)std::min<long long>((groups + kWinWarps - 1) / kWinWarps, (long long)std::max(1, max_ctas / 2));
        emu::launch(grid, kWinWarps * 32, [&]() { windows_all_kernel(arena, table.data(), 3, groups, window, val.data(), ra.data(), d512.data()); });
        for (int c = 0; c < 3; c++) {
            const int32_t* a0 = arena + off[c]; const uint32_t Lc = tl[c] + 1;
            std::vector<ull> w512(kWindowDistBins, 0); Stat st;
            for (long long k = 0; k < table[(size_t)c].n_windows; k++) {
                const uint32_t s = (uint32_t)((ull)k * window), e = (uint32_t)std::min<ull>(((ull)k + 1) * window, tl[c]);
                const double me = imean(a0, Lc, s, e, false), got = val[(size_t)(table[(size_t)c].first_window + k)];
                if (memcmp(&me, &got, 8) != 0) return fail("windows_all mean (bit pattern)", k, got, me);
                const int64_t bin = nim_to_int(me); w512[bin < 511 ? bin : 511]++;
                stat_add(st, a0 + s, (int64_t)e - s);
            }
            for (int b = 0; b < kWindowDistBins; b++) if (w512[b] != d512[(size_t)c * kWindowDistBins + b]) return fail("windows_all dist bin", b, (double)d512[(size_t)c * kWindowDistBins + b], (double)w512[b]);
            if (ra[(size_t)c].cum_length != st.len || ra[(size_t)c].cum_depth != st.sum || ra[(size_t)c].min_depth != st.mn || ra[(size_t)c].max_depth != st.mx) return fail("windows_all region stats (sum)", c, (double)ra[(size_t)c].cum_depth, (double)st.sum);
        }
        free(arena);
    }
    // ---- K9's filter: keep_count_kernel / runs_scan_kernel / keep_write_kernel (order-preserving compaction of the quantized runs) ----
    {
        const ull nr2 = (ull)tlen / 3 + 7; const int32_t n_labels = 3;
        std::vector<int32_t> st(nr2), en(nr2), bn(nr2), o_st(nr2), o_en(nr2), o_bn(nr2);
        for (ull i = 0; i < nr2; i++) { st[i] = (int32_t)(3 * i); en[i] = (int32_t)(3 * i + 3); bn[i] = (int32_t)(rnd() % 6) - 1; }      // bins -1 .. 4: -1, 3 and 4 are dropped
        const unsigned kt = (unsigned)((nr2 + kRunTile - 1) / kRunTile);
        std::vector<uint32_t> tc(kt); std::vector<ull> tb(kt); ull total = 0;
        emu::launch(kt, kRunThreads, [&]() { keep_count_kernel(bn.data(), nr2, n_labels, tc.data()); });
        emu::launch(1, 1024, [&]() { runs_scan_kernel(tc.data(), kt, tb.data(), &total); });
        emu::launch(kt, kRunThreads, [&]() { keep_write_kernel(st.data(), en.data(), bn.data(), nr2, n_labels, tb.data(), o_st.data(), o_en.data(), o_bn.data()); });
        ull w = 0;
        for (ull i = 0; i < nr2; i++) { if (bn[i] == -1 || bn[i] >= n_labels) continue; if (w >= total || o_st[w] != st[i] || o_en[w] != en[i] || o_bn[w] != bn[i]) return fail("kept quantized run", (long long)i, w < total ? o_st[w] : -1, st[i]); w++; }
        if (w != total) return fail("kept quantized runs (count)", 0, (double)total, (double)w);
    }
    printf("ok tlen %u window %u: %llu scan tiles, %lld windows, %lld regions\n", tlen, window, (ull)n_tiles, nw, nr);
    free(arr);
    return 0;
}
