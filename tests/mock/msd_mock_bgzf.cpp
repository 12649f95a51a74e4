// msd_mock_bgzf.cpp -- part of the TEST DOUBLE (see msd_mock.c): msd_per_base_bgzf answered by the host replay of K10's per-line
// arithmetic (mosdepth_b200/csrc/bedgz_core.cuh, the same functions tools/bedgz_model.cpp drives), so that the CLI's
// MOSDEPTH_B200_DEVICE_BGZF=1 branch and BgzWriter::write_members can be tested on the CPU.  CRCs come from zlib here (K1c on the GPU).
#include <string>
#include <vector>
#include <zlib.h>
#include "../../include/mosdepth_b200.h"
#include "../../mosdepth_b200/csrc/bedgz_core.cuh"

using namespace msd;

static int replay(const char* chrom, const int32_t* st, const int32_t* en, const int32_t* dp, int64_t nr, const char* const* labels, int n_labels, const uint8_t** bgzf, int64_t* n_bytes,
                  const msd_bgzf_member** members, int64_t* n_members);

extern "C" int msd_mock_per_base_bgzf(msd_ctx* ctx, int tid, const char* chrom, const uint8_t** bgzf, int64_t* n_bytes, const msd_bgzf_member** members, int64_t* n_members,
                                      const int32_t** start, const int32_t** stop, const int32_t** depth, int64_t* n_runs) {
    const int32_t *st, *en, *dp; int64_t nr = 0;
    int rc = msd_per_base_runs(ctx, tid, &st, &en, &dp, &nr); if (rc) return rc;
    *start = st; *stop = en; *depth = dp; *n_runs = nr;
    return replay(chrom, st, en, dp, nr, nullptr, 0, bgzf, n_bytes, members, n_members);
}
extern "C" int msd_quantized_bgzf(msd_ctx* ctx, int tid, const int64_t* quants, int n_q, const char* chrom, const char* const* labels, int n_labels, const uint8_t** bgzf, int64_t* n_bytes,
                                  const msd_bgzf_member** members, int64_t* n_members, const int32_t** start, const int32_t** stop, const int32_t** bin, int64_t* n_runs) {
    const int32_t *st, *en, *bn; int64_t nr = 0;
    int rc = msd_quantized_runs(ctx, tid, quants, n_q, &st, &en, &bn, &nr); if (rc) return rc;
    *start = st; *stop = en; *bin = bn; *n_runs = nr;
    return replay(chrom, st, en, bn, nr, labels, n_labels, bgzf, n_bytes, members, n_members);
}

static int replay(const char* chrom, const int32_t* st, const int32_t* en, const int32_t* dp, int64_t nr, const char* const* labels, int n_labels, const uint8_t** bgzf, int64_t* n_bytes,
                  const msd_bgzf_member** members, int64_t* n_members) {
    static std::vector<uint32_t> out32; static std::vector<msd_bgzf_member> mem;
    *bgzf = nullptr; *n_bytes = 0; *members = nullptr; *n_members = 0;
    if (nr == 0) return MSD_OK;
    const uint64_t n = (uint64_t)nr;
    BzCodes C;
    if (!bz_build_codes(chrom, st, en, dp, n, &C, labels, n_labels)) return MSD_EINVAL;
    std::vector<uint64_t> text_off(n + 1, 0), bit_off(n + 1, 0);
    for (uint64_t i = 0; i < n; i++) text_off[i + 1] = text_off[i] + bz_line_len(C, st[i], en[i], dp[i]);
    const uint32_t nb = bz_block_of(text_off[n - 1]) + 1;
    std::vector<uint64_t> blk_first(nb, ~0ull);
    for (uint64_t i = 0; i < n; i++) {
        const BzLineCtx lc = bz_line_ctx(C, st, en, dp, i, n, text_off[i]);
        if (lc.first) blk_first[bz_block_of(text_off[i])] = i;
        bit_off[i + 1] = bit_off[i] + bz_line_bits(C, C.litlen, C.dist, st, en, dp, i, lc);
    }
    std::vector<BlockDesc> desc(nb); std::vector<uint32_t> payload(nb); std::vector<uint64_t> member_off(nb + 1, 0);
    for (uint32_t b = 0; b < nb; b++) {
        const uint64_t f = blk_first[b], f1 = b + 1 == nb ? n : blk_first[b + 1];
        desc[b].dst = text_off[f]; desc[b].isize = (uint32_t)(text_off[f1] - text_off[f]);
        payload[b] = bz_payload_bytes(C.hdr_bits, bit_off[f1] - bit_off[f]);
        member_off[b + 1] = member_off[b] + payload[b] + 26;
    }
    std::vector<uint8_t> text(text_off[n] + 64, 0);
    out32.assign(member_off[nb] / 4 + 4, 0);
    for (uint64_t k = 0; k < n; k++) {
        const uint64_t i = n - 1 - k;          // any order
        const BzLineCtx lc = bz_line_ctx(C, st, en, dp, i, n, text_off[i]);
        const uint32_t b = bz_block_of(text_off[i]);
        bz_line_emit(C, C.litlen, C.dist, st, en, dp, i, lc, text.data(), text_off[i], out32.data(), (member_off[b] + 18) * 8 + C.hdr_bits + (bit_off[i] - bit_off[blk_first[b]]));
    }
    uint8_t* out = reinterpret_cast<uint8_t*>(out32.data());
    mem.resize(nb);
    for (uint32_t b = 0; b < nb; b++) {
        bz_frame(out + member_off[b], C, payload[b], (uint32_t)crc32(crc32(0L, Z_NULL, 0), text.data() + desc[b].dst, desc[b].isize), desc[b].isize);
        mem[b].offset = member_off[b]; mem[b].first_run = blk_first[b]; mem[b].csize = payload[b] + 26; mem[b].isize = desc[b].isize;
    }
    *bgzf = out; *n_bytes = (int64_t)member_off[nb]; *members = mem.data(); *n_members = nb;
    return MSD_OK;
}
