// bedgz_emu.cpp -- the KERNELS of K8 (runs.cuh) and K10 (bedgz.cuh) executed on the CPU under tests/emu/cuda_runtime.h: K8 launched in
// the order and geometry of compute_runs (msd_api.cu), K10 through the library's own host sequence bz_run (bedgz_host.cuh: buffers,
// carving, launches, read-backs -- the code msd_per_base_bgzf runs), from a depth array to BGZF members.  tests/test_kernels_emu.py
// gunzips the result against the printf text of a plain host run-length pass.  K1c is replaced
// by zlib's crc32 (its kernel is inline PTX and has its own GPU tests).  TEST INFRASTRUCTURE ONLY.
//   usage: bedgz_emu <prefix> <chrom> <n_bases> <max_ctas> [seed] [quant]      (quant: K9 + label mode, i.e. msd_quantized_bgzf)
#include <cstdlib>
#include <string>
#include <zlib.h>
#include "../../mosdepth_b200/csrc/bedgz_host.cuh"

using namespace msd;
typedef unsigned long long ull;

static unsigned grid_for(uint64_t n, int threads, int max_ctas) { uint64_t g = (n + threads - 1) / threads; if (g < 1) g = 1; if (g > (uint64_t)max_ctas) g = max_ctas; return (unsigned)g; }

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: bedgz_emu prefix chrom n_bases max_ctas [seed]\n"); return 2; }
    const std::string prefix = argv[1], chrom = argv[2];
    const uint32_t tlen = (uint32_t)strtoul(argv[3], nullptr, 10); const int max_ctas = atoi(argv[4]);
    uint64_t x = 0x9e3779b97f4a7c15ull ^ (argc > 5 ? strtoull(argv[5], nullptr, 10) : 0);
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    std::vector<int32_t> arr(tlen + 1, 0);                 // a depth array like to_coverage() leaves it: plateaus of 1..12 bases
    { int32_t d = 25; uint32_t i = 0; while (i < tlen) { uint32_t len = 1 + (uint32_t)(rnd() % 12); d += (int32_t)(rnd() % 3) - 1; if (d < 0) d = 0; if (rnd() % 400 == 0) d = (int32_t)(rnd() % 3000); while (len-- && i < tlen) arr[i++] = d; } }
    const bool quant = argc > 6 && std::string(argv[6]) == "quant";
    static const char* const labels[4] = {"NO_COVERAGE", "1:25", "CALLABLE", "27:inf"};
    QuantF qf; qf.qa.n = 5; { const long long q[5] = {0, 1, 25, 27, 0x7fffffffffffffffll}; for (int k = 0; k < 5; k++) qf.qa.q[k] = q[k]; }
    // ---- K8 / K9: compute_runs (identity, or the quantize bins over [0, tlen-1) as msd_quantized_runs asks for them) ---------------
    const uint32_t n_scan = quant ? (tlen >= 2 ? tlen - 1 : (tlen ? 1u : 0u)) : tlen;
    const unsigned n_tiles = (n_scan + kRunTile - 1) / kRunTile;
    std::vector<uint32_t> tile_count(n_tiles); std::vector<ull> tile_base(n_tiles); ull total = 0;
    if (quant) emu::launch(n_tiles, kRunThreads, [&]() { runs_count_kernel<QuantF>(arr.data(), n_scan, qf, tile_count.data()); });
    else emu::launch(n_tiles, kRunThreads, [&]() { runs_count_kernel<IdentityF>(arr.data(), n_scan, IdentityF(), tile_count.data()); });
    emu::launch(1, 1024, [&]() { runs_scan_kernel(tile_count.data(), n_tiles, tile_base.data(), &total); });
    std::vector<int32_t> st(total + 1), en(total + 1), dp(total + 1);
    if (quant) emu::launch(n_tiles, kRunThreads, [&]() { runs_write_kernel<QuantF>(arr.data(), n_scan, qf, tile_base.data(), st.data(), dp.data()); });
    else emu::launch(n_tiles, kRunThreads, [&]() { runs_write_kernel<IdentityF>(arr.data(), n_scan, IdentityF(), tile_base.data(), st.data(), dp.data()); });
    emu::launch(grid_for(total, 256, max_ctas), 256, [&]() { runs_stops_kernel(st.data(), total, (int32_t)tlen, en.data()); });
    uint64_t n = total;
    if (quant) {      // the reference's filter, as quantized_runs_impl applies it on the host
        uint64_t w = 0;
        for (uint64_t i = 0; i < total; i++) { const int32_t b = dp[i]; if (b == -1 || b >= qf.qa.n - 1) continue; st[w] = st[i]; en[w] = en[i]; dp[w] = b; w++; }
        n = w;
    }
    // ---- K10: the library's own host sequence (bedgz_host.cuh), CRC step = zlib ---------------------------------------------------
    BzBuffers B;
    const int rc = bz_run(B, 0, chrom.c_str(), st.data(), en.data(), dp.data(), st.data(), en.data(), dp.data(), n, max_ctas,
                          [](cudaStream_t, const uint8_t* text, const BlockDesc* desc, int nb, uint32_t*, const uint32_t*, uint32_t*, uint32_t*, uint32_t* crc_out) {
                              for (int b = 0; b < nb; b++) crc_out[b] = (uint32_t)crc32(crc32(0L, Z_NULL, 0), text + desc[b].dst, desc[b].isize);
                          }, quant ? labels : nullptr, quant ? 4 : 0);
    if (rc) { fprintf(stderr, "bz_run: %d %s\n", rc, B.err); return 5; }
    const uint8_t* out = B.h_out; const ull out_total = B.out_bytes, nb = B.members.size();
    ull text_total = 0; for (const auto& m : B.members) text_total += m.isize;
    // ---- results ----------------------------------------------------------------------------------------------------------
    FILE* f = fopen((prefix + ".gz").c_str(), "wb"); if (!f) return 1;
    fwrite(out, 1, out_total, f);
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, f); fclose(f);
    f = fopen((prefix + ".txt").c_str(), "wb"); if (!f) return 1;                  // plain host pass over the same array
    uint64_t n_host = 0;
    if (!quant) { for (uint32_t a = 0, i = 1; i <= tlen; i++) if (i == tlen || arr[i] != arr[a]) { fprintf(f, "%s\t%u\t%u\t%d\n", chrom.c_str(), a, i, arr[a]); a = i; n_host++; } }
    else if (tlen) {     // gen_quantized (mosdepth.nim:124-141) restated: bins over pos in [0, tlen-1), the last segment closed at tlen
        auto bin = [&](int32_t v) { long long x = v; if (x < qf.qa.q[0] || x > qf.qa.q[4]) return -1; for (int k = 0; k < 5; k++) { if (qf.qa.q[k] > x) return k - 1; if (qf.qa.q[k] == x) return k; } return 4; };
        int last_q = bin(arr[0]); int64_t last_pos = 0;
        for (int64_t pos = 0; pos < (int64_t)tlen - 1; pos++) {
            const int q = bin(arr[pos]);
            if (q == last_q) continue;
            if (last_q != -1 && last_q < 4) { fprintf(f, "%s\t%lld\t%lld\t%s\n", chrom.c_str(), (long long)last_pos, (long long)pos, labels[last_q]); n_host++; }
            last_q = q; last_pos = pos;
        }
        if (last_q != -1 && last_pos < (int64_t)tlen && last_q < 4) { fprintf(f, "%s\t%lld\t%u\t%s\n", chrom.c_str(), (long long)last_pos, tlen, labels[last_q]); n_host++; }
    }
    fclose(f);
    printf("%llu %llu %llu %llu %llu\n", (ull)n, (ull)n_host, text_total, out_total + 28, (ull)nb);
    return n == n_host ? 0 : 4;
}
