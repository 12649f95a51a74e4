// inflate_core.cuh -- arithmetic shared by the two inflate kernels (K1a huffman_decode, K1b lz77_resolve):
// table-entry encodings, canonical-Huffman bookkeeping, the per-lane region layout and the token format.
// Everything here is MSD_HD so that a host program (tools/inflate_model.cpp) can replay the same formulas against
// zlib without a GPU; the product path is the CUDA kernels in inflate.cuh only.
//
// Replaces htslib's BGZF inflate [ext] behind `open(bam, threads=)`, mosdepth.nim:888,968.  RFC 1951 throughout.
#pragma once
#include "common.cuh"

namespace msd {
namespace inf {

// ---- per-decoder (per-lane) shared-memory region: 354 halfwords; halfword i of lane L lives at smem16[i * 32 + L]
//      (one shift-add per lookup; at most 2-way bank conflicts, between the two lanes that share a 32-bit word) ------
constexpr int kLitSymHw = 0;         // lit/len symbols in canonical order, as table entries (<= 288: the fixed code has
                                     // 288), + slot 288 = "no such code"
constexpr int kLitBadSlot = 288;
constexpr int kDistSymHw = 289;      // distance symbols in canonical order, as table entries (<= 30), + slot 30 = "no such code"
constexpr int kDistBadSlot = 30;
constexpr int kLitBaseHw = 320;      // [l], l = 1..15: (index of the first length-l symbol - first length-l code), int16;
                                     // [16] = the bad slot (a window that matches no code counts 16 limits)
constexpr int kDistBaseHw = 337;
constexpr int kLaneHw = 354;         // 708 B per decoder
// while a dynamic header is being read the (dead) tables hold, as bytes (byte b = half (b & 1) of halfword b >> 1):
constexpr int kClTabByte = 0;        // 128-entry code-length-code table, one byte per entry: sym << 3 | len
constexpr int kLensByte = 128;       // hlit + hdist code lengths, one byte each (<= 316)

// ---- table entries (16 bit; the code length is not stored: the canonical search yields it) --------------------------
// lit/len: bit 7 = not a length code; [6:4] extra bits of a length code (0 otherwise, so the bits to drop are always
//          code length + [6:4]); [1:0] 1 = end of block, 2 = symbol that must not occur; [15:8] literal byte or
//          length base - 3
// distance: [7:4] extra bits 0..13 (15 = symbol that must not occur), [9:8] mantissa m:
//          distance = 1 + (m << extra) + extra bits
constexpr uint32_t kEntNotLen = 0x80, kEntEob = 0x81, kEntBad = 0x82, kDistNibBad = 15;

MSD_HD uint32_t lit_entry_of(uint32_t sym) {
    if (sym < 256) return (sym << 8) | kEntNotLen;
    if (sym == 256) return kEntEob;
    if (sym > 285) return kEntBad;
    const uint32_t idx = sym - 257;
    uint32_t base, extra;
    if (idx < 8) { base = 3 + idx; extra = 0; }
    else if (idx == 28) { base = 258; extra = 0; }
    else { extra = (idx - 4) >> 2; base = 3 + ((4 + (idx & 3)) << extra); }
    return ((base - 3) << 8) | (extra << 4);
}
MSD_HD uint32_t dist_entry_of(uint32_t sym) {
    if (sym > 29) return kDistNibBad << 4;
    if (sym < 4) return sym << 8;
    const uint32_t extra = (sym >> 1) - 1;
    return ((2 + (sym & 1)) << 8) | (extra << 4);
}

// ---- canonical code bookkeeping for one alphabet, from the per-length symbol counts --------------------------------
// first[l]: first code of length l; offs[l]: number of coded symbols shorter than l (index of the first length-l symbol
// in canonical order); end[l] = first[l] + cnt[l].  A code is over-subscribed iff end[l] > 2^l for some l.
struct Canon {
    uint32_t first[16], offs[16], end[16];
    bool oversubscribed;
};
MSD_HD void canon_from_counts(const uint32_t* cnt /*[16], cnt[0] ignored*/, Canon& c) {
    uint32_t code = 0, idx = 0;
    c.oversubscribed = false;
    c.first[0] = c.offs[0] = c.end[0] = 0;
    for (int l = 1; l < 16; l++) {
        code <<= 1;
        c.first[l] = code; c.offs[l] = idx;
        code += cnt[l]; idx += cnt[l];
        c.end[l] = code;
        if (code > (1u << l)) c.oversubscribed = true;
    }
}
// limit of length l left-aligned to 15 bits: a 15-bit MSB-first window c starts with a code of length <= l iff
// c < limit; the limits are non-decreasing in l, so the code length is 1 + #{l in 1..14 : c >= limit(l)}
MSD_HD uint32_t canon_limit(const Canon& c, int l) { return c.end[l] << (15 - l); }
// index base (mod 2^16) into the canonical-order symbol list: idx = (base + (c >> (15 - l))) & 0xffff
MSD_HD uint32_t canon_base(const Canon& c, int l) { return (c.offs[l] - c.first[l]) & 0xffffu; }

// ---- match tokens handed from K1a to K1b (32 bit) ---------------------------------------------------------------
// [31:23] literals since the previous token (0..510), [22:8] distance - 1, [7:0] length - 3
// [31:23] == 511: no match, [22:0] literals to skip (emitted in front of a match whose literal run is > 510)
constexpr uint32_t kTokSkip = 511;
constexpr int kTokenCap = 22016;     // per BGZF block: 65536 / 3 matches + 65536 / 511 skips, rounded up
MSD_HD uint32_t make_token(uint32_t lit_run, uint32_t len, uint32_t dist) { return (lit_run << 23) | ((dist - 1u) << 8) | (len - 3u); }

enum Status : uint32_t {
    kOk = 0, kBadBlockType = 1, kBadStored = 2, kBadCodeLengths = 3, kBadSymbol = 4, kBadDistance = 5,
    kOverflow = 6, kShort = 7, kOversubscribed = 8, kTruncated = 9, kBadCrc = 10
};

// ---- CRC-32 of the gzip trailer (RFC 1952; reflected polynomial 0xEDB88320), computed in parallel ---------------------
// Polynomials over GF(2) modulo the CRC polynomial in the reflected representation zlib uses: bit 31 is x^0.
constexpr uint32_t kCrcPoly = 0xedb88320u;
MSD_HD uint32_t crc_multmodp(uint32_t a, uint32_t b) {          // a(x) * b(x) mod p(x)
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 31; i >= 0; i--) {
        p ^= (0u - ((a >> i) & 1u)) & b;
        b = (b >> 1) ^ ((0u - (b & 1u)) & kCrcPoly);
    }
    return p;
}
// x2n[k] = x^(2^k) mod p, k = 0..31 (x2n[0] = x^1)
MSD_HD void crc_build_x2n(uint32_t* x2n) {
    uint32_t p = 1u << 30;
    x2n[0] = p;
    for (int n = 1; n < 32; n++) { p = crc_multmodp(p, p); x2n[n] = p; }
}
// x^(8 * n_bytes) mod p: the operator that appends n_bytes zero bytes to a message (crc32_combine: crc(A||B) =
// multmodp(crc_shift_op(|B|), crc(A)) ^ crc(B) for finalised CRCs)
MSD_HD uint32_t crc_shift_op(uint32_t n_bytes, const uint32_t* x2n) {
    uint32_t p = 1u << 31;
    int k = 3;
    while (n_bytes) { if (n_bytes & 1u) p = crc_multmodp(x2n[k & 31], p); n_bytes >>= 1; k++; }
    return p;
}
// slice-by-4 tables: t[0] is the classic byte table, t[k][i] = crc of byte i followed by k zero bytes
MSD_HD uint32_t crc_table0_entry(uint32_t i) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((0u - (c & 1u)) & kCrcPoly); return c; }

}  // namespace inf
}  // namespace msd
