// inflate_core.cuh -- arithmetic shared by the two inflate kernels (K1a huffman_decode, K1b lz77_resolve):
// table-entry encodings, canonical-Huffman bookkeeping, the per-lane region layout and the token format.
// Everything here is MSD_HD so that a host program (tools/inflate_model.cpp) can replay the same formulas against
// zlib without a GPU; the product path is the CUDA kernels in inflate.cuh only.
//
// Replaces htslib's BGZF inflate [ext] behind `open(bam, threads=)`, mosdepth.nim:888,968.  RFC 1951 throughout.
#pragma once
#include <cstring>
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

// code length of the canonical code that the 15-bit MSB-first window c starts with, minus 1 (15 = no such code):
// the number of per-length limits that are <= c.  The 15 limits (each <= 0x8000; the 16th half-word is 0x8000, never
// reached) are packed two per register; (0x8000 + c) - limit is computed for both halves by ONE subtraction (no borrow can
// cross: each half stays >= 0) and its bit 15 says c >= limit.
// r1a-r1n gathered the 16 flags with a shift + mask-or per register and one popc (24 instructions, an 8-deep dependent
// chain).  r1o: PRMT in sign-replicate mode turns the four flags of a register pair into four bytes 0xFF / 0x00 and DP4A
// against -1 counts them: 8 subtractions + 4 PRMT + 4 DP4A in two chains + 1 add = 17 instructions, 4 deep.
#ifndef MSD_CANON_DP4A
#define MSD_CANON_DP4A 1
#endif
MSD_HD uint32_t prmt_sign_bytes_1357(uint32_t a, uint32_t b) {     // bytes 1, 3 of a and 1, 3 of b, each replaced by 8 copies of its msb
#if defined(__CUDA_ARCH__)
    uint32_t r; asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return ((a >> 15) & 1u ? 0xffu : 0u) | ((a >> 31) & 1u ? 0xff00u : 0u) | ((b >> 15) & 1u ? 0xff0000u : 0u) | ((b >> 31) & 1u ? 0xff000000u : 0u);
#endif
}
MSD_HD int dp4a_minus1(uint32_t r, int acc) {                       // acc + sum over the four signed bytes of r of (byte * -1)
#if defined(__CUDA_ARCH__)
    return __dp4a((int)r, -1, acc);
#else
    for (int k = 0; k < 4; k++) acc -= (int)(int8_t)(r >> (8 * k));
    return acc;
#endif
}
MSD_HD uint32_t canon_len_minus1(uint32_t c, const uint32_t (&limp)[8]) {
    const uint32_t cc = c * 0x10001u + 0x80008000u;
#if MSD_CANON_DP4A
    int a0 = dp4a_minus1(prmt_sign_bytes_1357(cc - limp[0], cc - limp[1]), 0);
    int a1 = dp4a_minus1(prmt_sign_bytes_1357(cc - limp[2], cc - limp[3]), 0);
    a0 = dp4a_minus1(prmt_sign_bytes_1357(cc - limp[4], cc - limp[5]), a0);
    a1 = dp4a_minus1(prmt_sign_bytes_1357(cc - limp[6], cc - limp[7]), a1);
    return (uint32_t)(a0 + a1);
#else
    uint32_t t[8];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int j = 0; j < 8; j++) t[j] = ((cc - limp[j]) >> j) & (0x80008000u >> j);
    const uint32_t m = ((t[0] | t[1] | t[2]) | (t[3] | t[4] | t[5])) | (t[6] | t[7]);
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(m);
#else
    return (uint32_t)__builtin_popcount(m);
#endif
#endif
}

// ---- bit reader, second form (MSD_BITREADER2): no bit buffer.  The state is the bit position; a decode step reads its
// 32-bit window straight from the lane's 16-word ring (two loads + a funnel shift) and `drop` is one addition.  The first
// form keeps a 64-bit buffer plus two look-ahead words in registers and merges a word whenever fewer than 33 bits are
// left: ~28 instructions per refill, two refills per symbol iteration (lit/len and distance), executed by the warp in
// nearly every iteration because some lane always needs one.  Here the ring is topped up once per iteration at most
// (maintain(): a compare, and every 16 consumed bytes four stores + the 128-bit load of the chunk three ahead), since an
// iteration consumes at most 48 bits and the ring always holds the chunk of the current bit and the two after it.
// The price: the window loads sit on the decode chain (two shared-memory latencies per symbol).
template <int kStride> struct RingRef {        // word x of the lane at base[(x & 15) * kStride] (kernel: 32, host replay: 1)
    uint32_t* base;
    MSD_HD uint32_t ld(uint32_t w) const { return base[(w & 15u) * kStride]; }
    MSD_HD void st4(uint32_t chunk, uint32_t x, uint32_t y, uint32_t z, uint32_t w) const {
        uint32_t* r = base + (chunk & 3u) * 4u * kStride;
        r[0] = x; r[kStride] = y; r[2 * kStride] = z; r[3 * kStride] = w;
    }
};
template <int kStride> struct BitReader2 {
    uint32_t bitpos;            // next unread bit, counted from the 16-byte boundary at or before the stream's first byte
    uint32_t cur;               // the ring holds chunks cur, cur + 1, cur + 2 (16 bytes each); bitpos >> 7 is cur or cur + 1
    uint32_t p0, p1, p2, p3;    // chunk cur + 3, in flight or landed
    RingRef<kStride> ring;
    MSD_HD void fetch(const uint4* __restrict__ chunks, uint32_t clast, uint32_t chunk, bool pred) {
        const uint4* addr = chunks + (chunk < clast ? chunk : clast);
#if defined(__CUDA_ARCH__)
        // predicated load whose destinations ARE p0..p3 (a conditional C++ load becomes load-into-temporaries + moves, i.e.
        // a wait for the L2 latency in the same iteration; profiles/README.md r1i)
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
                     : "+r"(p0), "+r"(p1), "+r"(p2), "+r"(p3) : "l"(addr), "r"((uint32_t)pred));
#else
        if (pred) { uint32_t v[4]; memcpy(v, addr, 16); p0 = v[0]; p1 = v[1]; p2 = v[2]; p3 = v[3]; }   // (the replay's base is only 8-byte aligned)
#endif
    }
    MSD_HD void seed(const uint4* __restrict__ chunks, uint32_t clast, uint32_t* /*ring (set once)*/, uint32_t byte_pos) {
        bitpos = byte_pos * 8u; cur = byte_pos >> 4;
        for (uint32_t i = 0; i < 3; i++) { fetch(chunks, clast, cur + i, true); ring.st4(cur + i, p0, p1, p2, p3); }
        fetch(chunks, clast, cur + 3, true);
    }
    // call at least once per 128 consumed bits (the kernels: once per symbol iteration, once per header field group)
    MSD_HD void refill(const uint4* __restrict__ chunks, uint32_t clast, uint32_t* /*ring*/) {
        const bool moved = (bitpos >> 7) != cur;
        if (moved) { cur++; ring.st4(cur + 2, p0, p1, p2, p3); }
        fetch(chunks, clast, cur + 3, moved);
    }
    MSD_HD uint32_t lo() const {
        const uint32_t w = bitpos >> 5, sh = bitpos & 31u;
        const uint32_t x = ring.ld(w), y = ring.ld(w + 1u);
#if defined(__CUDA_ARCH__)
        return __funnelshift_r(x, y, sh);
#else
        return sh ? (x >> sh) | (y << (32u - sh)) : x;
#endif
    }
    MSD_HD void drop(int n) { bitpos += (uint32_t)n; }
    MSD_HD uint32_t take(int n) { const uint32_t v = lo() & ((1u << n) - 1u); drop(n); return v; }
    MSD_HD uint32_t bits_to_byte() const { return (0u - bitpos) & 7u; }
    MSD_HD uint32_t byte_pos() const { return bitpos >> 3; }      // byte index in the aligned stream (call on a byte boundary)
    MSD_HD uint32_t word_pos() const { return bitpos >> 5; }
};

// ---- CRC-32 of the gzip trailer (RFC 1952; reflected polynomial 0xEDB88320), computed in parallel ---------------------
// Polynomials over GF(2) modulo the CRC polynomial in the reflected representation zlib uses: bit 31 is x^0.
constexpr uint32_t kCrcPoly = 0xedb88320u;
MSD_HD uint32_t crc_multmodp(uint32_t a, uint32_t b) {          // a(x) * b(x) mod p(x)
    uint32_t p = 0;
#if defined(__CUDACC__)
#pragma unroll 4
#endif
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
// the classic byte table: entry i = byte i pushed through 8 zero bits
MSD_HD uint32_t crc_table0_entry(uint32_t i) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((0u - (c & 1u)) & kCrcPoly); return c; }
// x^e mod p for any 32-bit exponent.  x has order 2^32 - 1 modulo the CRC-32 polynomial, so exponents live modulo
// 2^32 - 1 and a negative power is x^(0xffffffff - k): the combine below divides as well as multiplies.
MSD_HD uint32_t crc_xpow(uint32_t e, const uint32_t* x2n) {
    uint32_t p = 1u << 31;
    for (int k = 0; e; k++, e >>= 1) if (e & 1u) p = crc_multmodp(x2n[k], p);
    return p;
}

// ---- K1c: row-strided CRC-32 (bgzf_crc_kernel and its host replay in tools/inflate_model.cpp) -------------------------
// A warp checks one block.  The message is extended to the left to a 16-byte boundary with `pad` zero bytes (leading zeros
// do not change a CRC computed from state 0; the 0xffffffff start state is added at the end as 0xffffffff * x^(8n)), and
// cut into rows of 512 bytes: in row r lane l loads the 16 bytes at 512 r + 16 l (one coalesced 512-byte request per
// warp) and owns four word STREAMS, one per component of the uint4.  Consecutive words of a stream are 512 bytes apart,
// so the stream state is kept pre-multiplied, d = (state at the end of its last word) * x^(8*508), and one step is
//     d' = (d ^ w) * x^(8*512)  =  U[0][b0] ^ U[1][b1] ^ U[2][b2] ^ U[3][b3],   U[m][i] = table0[i] * x^(8*(511-m)),
// four lookups per word like slice-by-4 (b0 = lowest byte).  Every lane indexes ITS OWN copy of the tables
// (entry (m, i) of lane l at word (m * 256 + i) * 32 + l, 128 KB per CTA): bank = lane, no conflicts whatever the data.
// After R full rows, stream (l, c) contributes d * x^(8*(tail - 16 l - 4 c)), tail = (n + pad) mod 512.  The four
// streams of a lane and its 16-byte piece t of the tail (zero padded: appended zeros are divided out again) fold into
//     G = (((d0 * x^32 ^ d1) * x^32 ^ d2) * x^32 ^ d3) * x^32 ^ t          (16 classic byte-table steps)
// which contributes G * x^(8*tail) * x^(-128 (l + 1)): one product with a per-lane constant, an XOR across the warp,
// one product with x^(8*tail), and the start-state term.
constexpr uint32_t kCrcRow = 512;
constexpr int kCrcXaN = 512;      // XA[t]  = x^(8 t), t < 512
constexpr int kCrcXbN = 129;      // XBI[r] = 0xffffffff * x^(4096 r): start state pushed through 512 r bytes (r <= 128: ISIZE <= 65536)
MSD_HD uint32_t crc_u_entry(int m, uint32_t i, const uint32_t* x2n) { return crc_multmodp(crc_xpow(8u * (511u - (uint32_t)m), x2n), crc_table0_entry(i)); }
MSD_HD uint32_t crc_lane_weight(int lane, const uint32_t* x2n) { return crc_xpow(0xffffffffu - 128u * (uint32_t)(lane + 1), x2n); }   // x^(-128 (l+1))
// bytes [lo, hi) of a little-endian word (lo, hi clamped to 0..4)
MSD_HD uint32_t crc_byte_mask(int lo, int hi) {
    lo = lo < 0 ? 0 : lo; hi = hi > 4 ? 4 : hi;
    if (hi <= lo) return 0u;
    return (0xffffffffu >> (8 * (4 - hi))) & (0xffffffffu << (8 * lo));
}
struct CrcWords { uint32_t x, y, z, w; };
// c * x^32: four byte steps with the classic table (T0 may be a shared-memory table or a host array)
template <class T0> MSD_HD uint32_t crc_times_x32(uint32_t c, const T0& t0) {
    c = (c >> 8) ^ t0(c & 0xffu); c = (c >> 8) ^ t0(c & 0xffu); c = (c >> 8) ^ t0(c & 0xffu); c = (c >> 8) ^ t0(c & 0xffu);
    return c;
}
// the words of row 0 that a lane has to cancel (bytes in front of the message: only lane 0 has any)
MSD_HD CrcWords crc_head_cancel(CrcWords v, int lane, uint32_t pad) {
    const int k = (int)pad - 16 * lane;       // leading bytes of this lane's 16 that are not message
    return CrcWords{v.x & ~crc_byte_mask(k, 4), v.y & ~crc_byte_mask(k - 4, 4), v.z & ~crc_byte_mask(k - 8, 4), v.w & ~crc_byte_mask(k - 12, 4)};
}
// a lane's 16 bytes of the tail at message' offset g0 = 512 R + 16 l, restricted to [pad, n2)
MSD_HD CrcWords crc_tail_keep(CrcWords v, uint32_t g0, uint32_t pad, uint32_t n2) {
    const int lo = (int)pad - (int)g0, hi = (int)n2 - (int)g0;
    return CrcWords{v.x & crc_byte_mask(lo, hi), v.y & crc_byte_mask(lo - 4, hi - 4), v.z & crc_byte_mask(lo - 8, hi - 8), v.w & crc_byte_mask(lo - 12, hi - 12)};
}
// state-0 CRC of a lane's (masked) tail words, then the fold G of the comment above
template <class T0> MSD_HD uint32_t crc_lane_fold(const uint32_t d[4], CrcWords t, const T0& t0) {
    uint32_t c = crc_times_x32(t.x, t0);
    c = crc_times_x32(c ^ t.y, t0); c = crc_times_x32(c ^ t.z, t0); c = crc_times_x32(c ^ t.w, t0);
    uint32_t g = crc_times_x32(d[0], t0) ^ d[1];
    g = crc_times_x32(g, t0) ^ d[2];
    g = crc_times_x32(g, t0) ^ d[3];
    return crc_times_x32(g, t0) ^ c;
}
// The constants of K1c, built once on the host (msd_create copies them to the device; the kernel replicates u[][] per lane).
struct CrcTables {
    uint32_t u[4][256];        // U[m][i]
    uint32_t t0[256];          // classic byte table
    uint32_t xa[kCrcXaN];      // x^(8 t)
    uint32_t xbi[kCrcXbN + 3]; // 0xffffffff * x^(4096 r)
    uint32_t xl[32];           // x^(-128 (lane + 1))
};
constexpr int kCrcTableWords = (int)(sizeof(CrcTables) / 4);
inline void crc_build_tables(CrcTables& t) {
    uint32_t x2n[32]; crc_build_x2n(x2n);
    for (uint32_t i = 0; i < 256; i++) t.t0[i] = crc_table0_entry(i);
    for (int m = 0; m < 4; m++) { const uint32_t f = crc_xpow(8u * (511u - (uint32_t)m), x2n); for (uint32_t i = 0; i < 256; i++) t.u[m][i] = crc_multmodp(f, t.t0[i]); }
    for (uint32_t k = 0; k < (uint32_t)kCrcXaN; k++) t.xa[k] = crc_xpow(8u * k, x2n);
    for (uint32_t r = 0; r < (uint32_t)kCrcXbN + 3; r++) t.xbi[r] = crc_multmodp(0xffffffffu, crc_xpow(4096u * r, x2n));
    for (int l = 0; l < 32; l++) t.xl[l] = crc_lane_weight(l, x2n);
}
// v = XOR over the lanes of G_l * x^(-128 (l+1)); xa_tail = x^(8 tail); xa_n = x^(8 (n mod 512)); xbi = XBI[n / 512]
MSD_HD uint32_t crc_finish(uint32_t v, uint32_t xa_tail, uint32_t xa_n, uint32_t xbi) { return ~(crc_multmodp(xa_tail, v) ^ crc_multmodp(xa_n, xbi)); }

}  // namespace inf
}  // namespace msd
