// bedgz_model.cpp -- host replay of K10 (mosdepth_b200/csrc/bedgz_core.cuh): the per-line arithmetic the kernels of
// bedgz.cuh run, driven here by plain loops (lines emitted in a scrambled order, as threads would), so that the encoder can
// be checked against zlib / gzip on the CPU (tests/test_bedgz_model.py).  Test infrastructure only.
//   usage: bedgz_model <out_prefix> <chrom> <n_runs> <pattern> [seed]
//          pattern: wgs (short runs, depth walks +-1), sparse (long runs), wide (start near 2^31, negative and 10-digit values),
//                   tiny (n_runs lines from position 0: 1-digit fields), quant (label mode: the 4th field is one of four labels,
//                   runs with gaps between them, as msd_quantized_runs returns them)
//   writes <prefix>.gz (BGZF members + EOF), <prefix>.txt (printf of the same runs), prints "lines text_bytes gz_bytes members"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <zlib.h>
#include "../mosdepth_b200/csrc/bedgz_core.cuh"

using namespace msd;

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: bedgz_model prefix chrom n_runs pattern [seed]\n"); return 2; }
    const std::string prefix = argv[1], chrom = argv[2], pattern = argv[4];
    const uint64_t n = strtoull(argv[3], nullptr, 10);
    uint64_t x = 0x9e3779b97f4a7c15ull ^ (argc > 5 ? strtoull(argv[5], nullptr, 10) : 0);
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    std::vector<int32_t> st(n), en(n), dp(n);
    int64_t pos = pattern == "wide" ? 2147483647ll - (int64_t)n * 3 - 10 : 0; int32_t d = 30;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t r = rnd();
        int64_t len = 1 + (int64_t)(r % 9);
        if (pattern == "sparse") len = 1 + (int64_t)(r % 200000);
        if (pattern == "wide") len = 1 + (int64_t)(r % 3);
        if (pattern == "tiny") len = 1;
        if (pattern == "sparse" && pos + len > 2147483000ll) len = 1;
        if (pattern == "quant") { len = 1 + (int64_t)(r % 3000); if ((r >> 44) % 3 == 0) pos += 1 + (int64_t)((r >> 48) % 500); }   // dropped bins leave gaps
        st[i] = (int32_t)pos; pos += len; en[i] = (int32_t)pos;
        const int32_t prev = d;
        if (pattern == "wide") d = (int32_t)(r >> 17);                                  // any int32, negative included
        else { d += (int32_t)((r >> 20) % 3) - 1; if (d < 0) d = 1; if ((r >> 30) % 50 == 0) d = (int32_t)((r >> 36) % 5000); }
        if (d == prev) d = prev + 1;
        if (pattern == "quant") d = (int32_t)((r >> 33) % 4);
        dp[i] = d;
    }
    static const char* const labels[4] = {"NO_COVERAGE", "0:1", "CALLABLE", "a_label_of_forty-eight_bytes_which_is_the_maximu"};
    const bool quant = pattern == "quant";
    BzCodes C;
    if (!bz_build_codes(chrom.c_str(), st.data(), en.data(), dp.data(), n, &C, quant ? labels : nullptr, quant ? 4 : 0)) { fprintf(stderr, "bz_build_codes refused\n"); return 3; }
    // bz_len_kernel + scan
    std::vector<uint64_t> text_off(n + 1, 0), bit_off(n + 1, 0);
    for (uint64_t i = 0; i < n; i++) text_off[i + 1] = text_off[i] + bz_line_len(C, st[i], en[i], dp[i]);
    const uint64_t text_total = text_off[n];
    const uint32_t nb = n ? bz_block_of(text_off[n - 1]) + 1 : 0;
    // bz_mark_kernel + scan
    std::vector<uint64_t> blk_first(nb, ~0ull);
    for (uint64_t i = 0; i < n; i++) {
        const BzLineCtx lc = bz_line_ctx(C, st.data(), en.data(), dp.data(), i, n, text_off[i]);
        if (lc.first) blk_first[bz_block_of(text_off[i])] = i;
        bit_off[i + 1] = bit_off[i] + bz_line_bits(C, C.litlen, C.dist, st.data(), en.data(), dp.data(), i, lc);
    }
    const uint64_t bit_total = bit_off[n];
    // bz_blocks_kernel + scan
    std::vector<BlockDesc> desc(nb); std::vector<uint32_t> payload(nb); std::vector<uint64_t> member_off(nb + 1, 0);
    for (uint32_t b = 0; b < nb; b++) {
        const uint64_t f = blk_first[b], f1 = b + 1 == nb ? n : blk_first[b + 1];
        if (f == ~0ull || f1 <= f) { fprintf(stderr, "member %u has no first line\n", b); return 4; }
        const uint64_t t1 = b + 1 == nb ? text_total : text_off[f1], b1 = b + 1 == nb ? bit_total : bit_off[f1];
        desc[b].dst = text_off[f]; desc[b].isize = (uint32_t)(t1 - text_off[f]);
        payload[b] = bz_payload_bytes(C.hdr_bits, b1 - bit_off[f]);
        if (desc[b].isize > kBzMaxBlockText || payload[b] + 26 > 65536) { fprintf(stderr, "member %u too large: text %u payload %u\n", b, desc[b].isize, payload[b]); return 5; }
        member_off[b + 1] = member_off[b] + payload[b] + 26;
    }
    // bz_emit_kernel: any order
    std::vector<uint8_t> text(text_total + 64, 0);
    std::vector<uint32_t> out32(member_off[nb] / 4 + 4, 0);
    std::vector<uint64_t> order(n); for (uint64_t i = 0; i < n; i++) order[i] = i;
    for (uint64_t i = n; i > 1; i--) std::swap(order[i - 1], order[rnd() % i]);
    for (uint64_t k = 0; k < n; k++) {
        const uint64_t i = order[k];
        const BzLineCtx lc = bz_line_ctx(C, st.data(), en.data(), dp.data(), i, n, text_off[i]);
        const uint32_t b = bz_block_of(text_off[i]);
        const uint64_t bitpos = (member_off[b] + 18) * 8 + C.hdr_bits + (bit_off[i] - bit_off[blk_first[b]]);
        bz_line_emit(C, C.litlen, C.dist, st.data(), en.data(), dp.data(), i, lc, text.data(), text_off[i], out32.data(), bitpos);
    }
    // K1c (zlib here) + bz_frame_kernel
    uint8_t* out = reinterpret_cast<uint8_t*>(out32.data());
    for (uint32_t b = 0; b < nb; b++) {
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), text.data() + desc[b].dst, desc[b].isize);
        bz_frame(out + member_off[b], C, payload[b], crc, desc[b].isize);
    }
    FILE* f = fopen((prefix + ".gz").c_str(), "wb"); if (!f) return 1;
    fwrite(out, 1, member_off[nb], f);
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, f); fclose(f);
    f = fopen((prefix + ".txt").c_str(), "wb"); if (!f) return 1;
    for (uint64_t i = 0; i < n; i++) { if (quant) fprintf(f, "%s\t%d\t%d\t%s\n", chrom.c_str(), st[i], en[i], labels[dp[i]]); else fprintf(f, "%s\t%d\t%d\t%d\n", chrom.c_str(), st[i], en[i], dp[i]); }
    fclose(f);
    f = fopen((prefix + ".replay.txt").c_str(), "wb"); if (!f) return 1;       // the text bytes the emit step wrote (what K1c sums)
    fwrite(text.data(), 1, text_total, f); fclose(f);
    printf("%llu %llu %llu %u %u\n", (unsigned long long)n, (unsigned long long)text_total, (unsigned long long)member_off[nb] + 28, nb, C.hdr_bits);
    return 0;
}
