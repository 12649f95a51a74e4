// bedgz_emu.cpp -- the KERNELS of K8 (runs.cuh) and K10 (bedgz.cuh) executed on the CPU under tests/emu/cuda_runtime.h, launched in
// the order and with the geometry msd_per_base_bgzf / compute_runs use (mosdepth_b200/csrc/msd_api.cu), from a depth array to BGZF
// members.  tests/test_bedgz_model.py gunzips the result against the printf text of a plain host run-length pass.  K1c is replaced
// by zlib's crc32 (its kernel is inline PTX and has its own GPU tests).  TEST INFRASTRUCTURE ONLY.
//   usage: bedgz_emu <prefix> <chrom> <n_bases> <max_ctas> [seed]
#include <cstdlib>
#include <string>
#include <zlib.h>
#include "../../mosdepth_b200/csrc/runs.cuh"
#include "../../mosdepth_b200/csrc/bedgz.cuh"

using namespace msd;
typedef unsigned long long ull;

static unsigned grid_for(uint64_t n, int threads, int max_ctas) { uint64_t g = (n + threads - 1) / threads; if (g < 1) g = 1; if (g > (uint64_t)max_ctas) g = max_ctas; return (unsigned)g; }
static void bz_scan(const uint32_t* in, uint64_t n, uint32_t* tile_sum, ull* tile_base, ull* total, ull* out) {
    const unsigned nt = (unsigned)((n + kBzScanTile - 1) / kBzScanTile);
    emu::launch(nt, kBzThreads, [&]() { bz_tile_sum_kernel(in, n, tile_sum); });
    emu::launch(1, 1024, [&]() { runs_scan_kernel(tile_sum, nt, tile_base, total); });
    emu::launch(nt, kBzThreads, [&]() { bz_tile_scan_kernel(in, n, tile_base, out); });
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: bedgz_emu prefix chrom n_bases max_ctas [seed]\n"); return 2; }
    const std::string prefix = argv[1], chrom = argv[2];
    const uint32_t tlen = (uint32_t)strtoul(argv[3], nullptr, 10); const int max_ctas = atoi(argv[4]);
    uint64_t x = 0x9e3779b97f4a7c15ull ^ (argc > 5 ? strtoull(argv[5], nullptr, 10) : 0);
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    std::vector<int32_t> arr(tlen + 1, 0);                 // a depth array like to_coverage() leaves it: plateaus of 1..12 bases
    { int32_t d = 25; uint32_t i = 0; while (i < tlen) { uint32_t len = 1 + (uint32_t)(rnd() % 12); d += (int32_t)(rnd() % 3) - 1; if (d < 0) d = 0; if (rnd() % 400 == 0) d = (int32_t)(rnd() % 3000); while (len-- && i < tlen) arr[i++] = d; } }
    // ---- K8: compute_runs ---------------------------------------------------------------------------------------------
    const unsigned n_tiles = (tlen + kRunTile - 1) / kRunTile;
    std::vector<uint32_t> tile_count(n_tiles); std::vector<ull> tile_base(n_tiles); ull total = 0;
    emu::launch(n_tiles, kRunThreads, [&]() { runs_count_kernel<IdentityF>(arr.data(), tlen, IdentityF(), tile_count.data()); });
    emu::launch(1, 1024, [&]() { runs_scan_kernel(tile_count.data(), n_tiles, tile_base.data(), &total); });
    std::vector<int32_t> st(total + 1), en(total + 1), dp(total + 1);
    emu::launch(n_tiles, kRunThreads, [&]() { runs_write_kernel<IdentityF>(arr.data(), tlen, IdentityF(), tile_base.data(), st.data(), dp.data()); });
    emu::launch(grid_for(total, 256, max_ctas), 256, [&]() { runs_stops_kernel(st.data(), total, (int32_t)tlen, en.data()); });
    const uint64_t n = total;
    // ---- K10: msd_per_base_bgzf -----------------------------------------------------------------------------------------
    BzCodes codes;
    if (!bz_build_codes(chrom.c_str(), st.data(), en.data(), dp.data(), n, &codes)) return 3;
    const uint64_t nt = (n + kBzScanTile - 1) / kBzScanTile;
    std::vector<uint32_t> text_len(n), bit_len(n), tsum(nt); std::vector<ull> text_off(n), bit_off(n), tbase(nt); ull totals[4] = {0, 0, 0, 0};
    const unsigned g_lines = grid_for(n, kBzThreads, max_ctas);
    emu::launch(g_lines, kBzThreads, [&]() { bz_len_kernel(&codes, st.data(), en.data(), dp.data(), n, text_len.data()); });
    bz_scan(text_len.data(), n, tsum.data(), tbase.data(), totals + 0, text_off.data());
    const ull text_total = totals[0], last_off = text_off[n - 1];
    const uint64_t nb = (uint64_t)bz_block_of(last_off) + 1;
    const uint64_t nbt = (nb + kBzScanTile - 1) / kBzScanTile;
    std::vector<ull> blk_first(nb, ~0ull), member_off(nb), btbase(nbt); std::vector<BlockDesc> desc(nb); std::vector<uint32_t> payload(nb), member_len(nb), crc(nb), btsum(nbt);
    std::vector<uint8_t> text(text_total + 512, 0);
    emu::launch(g_lines, kBzThreads, [&]() { bz_mark_kernel(&codes, st.data(), en.data(), dp.data(), n, text_off.data(), bit_len.data(), blk_first.data()); });
    bz_scan(bit_len.data(), n, tsum.data(), tbase.data(), totals + 1, bit_off.data());
    const ull bit_total = totals[1];
    emu::launch((unsigned)((nb + kBzThreads - 1) / kBzThreads), kBzThreads, [&]() { bz_blocks_kernel(&codes, (uint32_t)nb, n, blk_first.data(), text_off.data(), text_total, bit_off.data(), bit_total, desc.data(), payload.data(), member_len.data()); });
    bz_scan(member_len.data(), nb, btsum.data(), btbase.data(), totals + 2, member_off.data());
    const ull out_total = totals[2];
    std::vector<uint32_t> out32((out_total + 64) / 4 + 1, 0);
    emu::launch(g_lines, kBzThreads, [&]() { bz_emit_kernel(&codes, st.data(), en.data(), dp.data(), n, text_off.data(), bit_off.data(), blk_first.data(), member_off.data(), text.data(), out32.data()); });
    for (uint64_t b = 0; b < nb; b++) crc[b] = (uint32_t)crc32(crc32(0L, Z_NULL, 0), text.data() + desc[b].dst, desc[b].isize);      // K1c
    uint8_t* out = reinterpret_cast<uint8_t*>(out32.data());
    emu::launch((unsigned)((nb + kBzThreads - 1) / kBzThreads), kBzThreads, [&]() { bz_frame_kernel(&codes, (uint32_t)nb, desc.data(), payload.data(), member_off.data(), crc.data(), out); });
    // ---- results ----------------------------------------------------------------------------------------------------------
    FILE* f = fopen((prefix + ".gz").c_str(), "wb"); if (!f) return 1;
    fwrite(out, 1, out_total, f);
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, f); fclose(f);
    f = fopen((prefix + ".txt").c_str(), "wb"); if (!f) return 1;                  // plain host run-length pass over the same array
    uint64_t n_host = 0;
    for (uint32_t a = 0, i = 1; i <= tlen; i++) if (i == tlen || arr[i] != arr[a]) { fprintf(f, "%s\t%u\t%u\t%d\n", chrom.c_str(), a, i, arr[a]); a = i; n_host++; }
    fclose(f);
    printf("%llu %llu %llu %llu %llu\n", (ull)n, (ull)n_host, text_total, out_total + 28, (ull)nb);
    return n == n_host ? 0 : 4;
}
