// bedgz_core.cuh -- K10 "bed_deflate": the per-base BED text of a contig, formatted AND deflated on the device (SURVEY
// section 8f rank 1: mosdepth.nim:783-793 writes `chrom \t start \t stop \t depth \n` per run of gen_depths through hts-nim's
// BGZF writer, int2str.nim:59-76 formats the numbers; in the reference this is the second-largest cost of a default run).
//
// A general deflate encoder searches for matches; this text needs no search, its redundancy is known by construction:
//   line i   = NAME \t S_i \t E_i \t D_i \n        with S_i == E_{i-1} (runs are contiguous)
//   token 1  = match(len NAME+1, dist = length of line i-1)                     the contig name, from the previous line
//   token 2  = match(len |S_i|+1, dist = |line i-1| - |S_{i-1}| - 1)            "S_i \t" is the previous line's "E_{i-1} \t"
//   token 3  = match(len c, dist = |S_i|+1), c = common digit prefix of E_i, S_i (taken when c >= 3)
//   the rest = literals (last digits of E_i, \t, D_i, \n)
// So one thread per line emits its tokens with no dependence on any other line but i-1, and lines are cut into independent
// BGZF members (the first line of a member has no previous line and is written as literals + token 3).  Token bits come from
// ONE dynamic Huffman code per contig, built on the host from a sample of the contig's own lines (every symbol keeps a
// code, so any line is encodable); its header bits are the same for every member.  Bit offsets are exclusive scans of the
// per-line bit counts; lines OR their bits into the zeroed output (first / last word with atomicOr, interior words stored).
// The output is valid RFC 1951/1952 + BGZF: any gunzip / htslib reads it; the compressed bytes differ from zlib's, which is
// allowed (parity is on decompressed content).
//
// Everything a kernel computes per line is in MSD_HD functions here, so tools/bedgz_model.cpp replays the exact arithmetic
// on the host against zlib (tests/test_bedgz_model.py).  Nothing in this file is a CPU fallback: the host replay is a test.
#pragma once
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace msd {

constexpr int kBzMaxName = 200;                       // contig name bytes the device path accepts
constexpr int kBzMaxLabels = 64, kBzMaxLabel = 48;    // quantized.bed: the 4th field is a label (mosdepth.nim:111-122) instead of a depth
constexpr uint32_t kBzMaxLine = kBzMaxName + 1 + 11 + 1 + 11 + 1 + kBzMaxLabel + 1;
constexpr uint32_t kBzBlockSpan = 32768 - 320;        // lines that START in [b*span, (b+1)*span) of the text form member b
constexpr uint32_t kBzMaxBlockText = kBzBlockSpan + kBzMaxLine;   // <= 32,768 bytes: 15 bits per byte still fit a 64 KB member
static_assert(kBzMaxBlockText <= 32768, "member text bound");
constexpr int kBzHdrMaxBytes = 512;
constexpr int kBzNumMax = 11 + 1 + 11 + 1 + kBzMaxLabel + 1 + 7;

struct BzCodes {
    uint32_t litlen[286];       // bit-reversed code | nbits << 16 (ready for LSB-first packing)
    uint32_t dist[30];
    uint32_t hdr_bits;          // BFINAL/BTYPE + code-length section, identical in every member
    uint32_t name_len;          // without the tab
    uint8_t name[kBzMaxName + 8];   // NAME \t
    uint8_t hdr[kBzHdrMaxBytes];
    uint32_t n_labels;          // 0: the 4th field is the decimal value (per-base.bed); else it is lab[value] (quantized.bed)
    uint8_t lab_len[kBzMaxLabels];
    uint8_t lab[kBzMaxLabels][kBzMaxLabel];
};

MSD_HD uint32_t bz_digits(uint32_t v) {
    uint32_t n = 1;
    if (v >= 100000u) { if (v >= 10000000u) { n = v >= 1000000000u ? 10 : v >= 100000000u ? 9 : 8; } else n = v >= 1000000u ? 7 : 6; }
    else { if (v >= 1000u) n = v >= 10000u ? 5 : 4; else n = v >= 100u ? 3 : v >= 10u ? 2 : 1; }
    return n;
}
MSD_HD uint32_t bz_idigits(int32_t v) { return v < 0 ? 1 + bz_digits((uint32_t)(-(int64_t)v)) : bz_digits((uint32_t)v); }
MSD_HD uint32_t bz_value_len(const BzCodes& C, int32_t v) { return C.n_labels ? ((uint32_t)v < C.n_labels ? C.lab_len[v] : 0u) : bz_idigits(v); }
MSD_HD uint32_t bz_line_len(const BzCodes& C, int32_t st, int32_t en, int32_t v) { return C.name_len + 1 + bz_idigits(st) + 1 + bz_idigits(en) + 1 + bz_value_len(C, v) + 1; }

MSD_HD char* bz_put_i32(char* p, int32_t v) {
    uint32_t u = (uint32_t)v;
    if (v < 0) { *p++ = '-'; u = (uint32_t)(-(int64_t)v); }
    const uint32_t n = bz_digits(u);
    for (uint32_t k = n; k-- > 0;) { p[k] = (char)('0' + u % 10u); u /= 10u; }
    return p + n;
}
// "S \t E \t D \n" (D: the decimal value, or its label) into num; returns its length, the lengths of S and E through ls / le
MSD_HD uint32_t bz_format_num(const BzCodes& C, char* num, int32_t st, int32_t en, int32_t v, uint32_t* ls, uint32_t* le) {
    char* p = bz_put_i32(num, st); *ls = (uint32_t)(p - num); *p++ = '\t';
    char* q = bz_put_i32(p, en); *le = (uint32_t)(q - p); *q++ = '\t';
    if (C.n_labels) { const uint32_t ll = bz_value_len(C, v); for (uint32_t k = 0; k < ll; k++) *q++ = (char)C.lab[v][k]; }
    else q = bz_put_i32(q, v);
    *q++ = '\n';
    return (uint32_t)(q - num);
}

// RFC 1951 3.2.5: length 3..258 -> symbol 257..285 + extra bits; distance 1..32768 -> code 0..29 + extra bits
MSD_HD uint32_t bz_ilog2(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return 31u - (uint32_t)__clz((int)v);
#else
    uint32_t n = 0; while (v >>= 1) n++; return n;
#endif
}
MSD_HD void bz_len_symbol(uint32_t len, uint32_t* sym, uint32_t* ebits, uint32_t* extra) {
    const uint32_t l = len - 3;
    if (l < 8) { *sym = 257 + l; *ebits = 0; *extra = 0; return; }
    if (l == 255) { *sym = 285; *ebits = 0; *extra = 0; return; }
    const uint32_t eb = bz_ilog2(l) - 2;
    *sym = 257 + 4 * (eb + 1) + ((l >> eb) & 3); *ebits = eb; *extra = l & ((1u << eb) - 1);
}
MSD_HD void bz_dist_symbol(uint32_t dist, uint32_t* code, uint32_t* ebits, uint32_t* extra) {
    const uint32_t d = dist - 1;
    if (d < 4) { *code = d; *ebits = 0; *extra = 0; return; }
    const uint32_t n = bz_ilog2(d), eb = n - 1;
    *code = 2 * n + ((d >> eb) & 1); *ebits = eb; *extra = d & ((1u << eb) - 1);
}

// What a line knows about the line in front of it (same member): nothing else crosses lines.
struct BzPrev { bool has; bool chained; uint32_t len, ls; uint32_t lab_dist; };   // chained: previous stop == this start; lab_dist: label mode, see bz_label_dist

// The token walk shared by the bit counter, the bit emitter and the host's sampler.  Sink: lit(byte), match(len, dist).
template <class Sink>
MSD_HD void bz_tokens(const uint8_t* name, uint32_t name_len, const char* num, uint32_t num_len, uint32_t ls, uint32_t le, const BzPrev pv, Sink& sink) {
    const uint32_t nl1 = name_len + 1;
    if (pv.has && nl1 >= 3 && nl1 <= 258) sink.match(nl1, pv.len);
    else for (uint32_t k = 0; k < nl1; k++) sink.lit(name[k]);
    uint32_t k0 = 0;
    if (pv.has && pv.chained && ls + 1 >= 3) { sink.match(ls + 1, pv.len - pv.ls - 1); k0 = ls + 1; }
    else for (; k0 < ls + 1; k0++) sink.lit((uint8_t)num[k0]);
    uint32_t c = 0;                                         // common prefix of E and S
    const uint32_t cmax = ls < le ? ls : le;
    while (c < cmax && num[c] == num[ls + 1 + c]) c++;
    if (c >= 3) { sink.match(c, ls + 1); k0 += c; }
    const uint32_t lab_start = ls + 1 + le + 1;              // label mode: "LABEL \n" copied from the nearest line of the member that has it
    const uint32_t lit_end = pv.lab_dist ? lab_start : num_len;
    for (; k0 < lit_end; k0++) sink.lit((uint8_t)num[k0]);
    if (pv.lab_dist) sink.match(num_len - lab_start, pv.lab_dist);
}

struct BzCountSink {
    const uint32_t* litlen; const uint32_t* dist; uint32_t bits;
    MSD_HD void lit(uint8_t b) { bits += litlen[b] >> 16; }
    MSD_HD void match(uint32_t len, uint32_t d) {
        uint32_t s, eb, ex; bz_len_symbol(len, &s, &eb, &ex); bits += (litlen[s] >> 16) + eb;
        bz_dist_symbol(d, &s, &eb, &ex); bits += (dist[s] >> 16) + eb;
    }
};

// LSB-first bit writer into a zeroed array of 32-bit words that other lines write at the same time: the first and the last
// word of a line are shared with its neighbours (OR), the words in between belong to the line alone (store).
struct BzBitWriter {
    uint32_t* out; uint64_t word; uint64_t acc; uint32_t nacc; bool first;
    MSD_HD void init(uint32_t* o, uint64_t bitpos) { out = o; word = bitpos >> 5; nacc = (uint32_t)(bitpos & 31); acc = 0; first = true; }
    MSD_HD void or_word(uint64_t w, uint32_t v) {
#if defined(__CUDA_ARCH__)
        if (v) atomicOr(out + w, v);
#else
        out[w] |= v;
#endif
    }
    MSD_HD void put(uint32_t v, uint32_t n) {             // n <= 32, v < 2^n
        acc |= (uint64_t)v << nacc; nacc += n;
        if (nacc >= 32) {
            if (first) { or_word(word, (uint32_t)acc); first = false; } else out[word] = (uint32_t)acc;
            word++; acc >>= 32; nacc -= 32;
        }
    }
    MSD_HD void finish() { if (nacc) or_word(word, (uint32_t)acc); }
};

struct BzEmitSink {
    const uint32_t* litlen; const uint32_t* dist; BzBitWriter w;
    MSD_HD void lit(uint8_t b) { const uint32_t e = litlen[b]; w.put(e & 0xffffu, e >> 16); }
    MSD_HD void match(uint32_t len, uint32_t d) {
        uint32_t s, eb, ex; bz_len_symbol(len, &s, &eb, &ex);
        uint32_t e = litlen[s]; w.put(e & 0xffffu, e >> 16); if (eb) w.put(ex, eb);
        bz_dist_symbol(d, &s, &eb, &ex);
        e = dist[s]; w.put(e & 0xffffu, e >> 16); if (eb) w.put(ex, eb);
    }
};

// ---- per-line steps of the kernels (bedgz.cuh) -- the host replay calls exactly these ---------------------------------
MSD_HD uint32_t bz_block_of(uint64_t text_off) { return (uint32_t)(text_off / kBzBlockSpan); }

// Label mode (quantized.bed): distance from the label of line i back to the label of the nearest of the 8 lines in front of it
// that carries the same label (and, when `bounded`, starts in the same member); 0 if there is none.  No search structure: the
// candidates are the lines themselves, their lengths are digit counts.
MSD_HD uint32_t bz_label_dist(const BzCodes& C, const int32_t* st, const int32_t* en, const int32_t* dp, uint64_t i, bool bounded, uint64_t text_off) {
    if (!C.n_labels || bz_value_len(C, dp[i]) + 1 < 3) return 0;
    const uint32_t blk = bounded ? bz_block_of(text_off) : 0;
    uint64_t back = 0, off = text_off;
    for (uint64_t j = 1; j <= 8 && j <= i; j++) {
        const uint32_t lj = bz_line_len(C, st[i - j], en[i - j], dp[i - j]);
        back += lj;
        if (bounded) { if (off < lj) break; off -= lj; if (bz_block_of(off) != blk) break; }
        if (dp[i - j] == dp[i]) {
            const uint64_t d = back + bz_idigits(st[i]) + bz_idigits(en[i]) - bz_idigits(st[i - j]) - bz_idigits(en[i - j]);
            return d <= 32768 ? (uint32_t)d : 0u;
        }
    }
    return 0;
}

struct BzLineCtx { bool first, last; BzPrev pv; };
// line i of n: is it the first / last line of its member, and what does it know of line i-1
MSD_HD BzLineCtx bz_line_ctx(const BzCodes& C, const int32_t* st, const int32_t* en, const int32_t* dp, uint64_t i, uint64_t n, uint64_t text_off) {
    BzLineCtx c; c.pv.has = false; c.pv.chained = false; c.pv.len = 0; c.pv.ls = 0; c.pv.lab_dist = bz_label_dist(C, st, en, dp, i, true, text_off);
    const uint32_t blk = bz_block_of(text_off);
    c.first = true;
    if (i > 0) {
        const uint32_t plen = bz_line_len(C, st[i - 1], en[i - 1], dp[i - 1]);
        if (bz_block_of(text_off - plen) == blk) { c.first = false; c.pv.has = true; c.pv.len = plen; c.pv.ls = bz_idigits(st[i - 1]); c.pv.chained = en[i - 1] == st[i]; }
    }
    c.last = i + 1 == n || bz_block_of(text_off + bz_line_len(C, st[i], en[i], dp[i])) != blk;
    return c;
}
// bits of line i (its tokens; plus the end-of-block code when it closes a member)
MSD_HD uint32_t bz_line_bits(const BzCodes& C, const uint32_t* litlen, const uint32_t* dist, const int32_t* st, const int32_t* en, const int32_t* dp, uint64_t i, const BzLineCtx& lc) {
    char num[kBzNumMax]; uint32_t ls, le;
    const uint32_t nn = bz_format_num(C, num, st[i], en[i], dp[i], &ls, &le);
    BzCountSink s; s.litlen = litlen; s.dist = dist; s.bits = 0;
    bz_tokens(C.name, C.name_len, num, nn, ls, le, lc.pv, s);
    return s.bits + (lc.last ? (litlen[256] >> 16) : 0u);
}
// text bytes of line i at text + text_off, token bits at bit position bitpos of out32
MSD_HD void bz_line_emit(const BzCodes& C, const uint32_t* litlen, const uint32_t* dist, const int32_t* st, const int32_t* en, const int32_t* dp, uint64_t i, const BzLineCtx& lc,
                         uint8_t* text, uint64_t text_off, uint32_t* out32, uint64_t bitpos) {
    char num[kBzNumMax]; uint32_t ls, le;
    const uint32_t nn = bz_format_num(C, num, st[i], en[i], dp[i], &ls, &le);
    uint8_t* t = text + text_off;
    for (uint32_t k = 0; k <= C.name_len; k++) t[k] = C.name[k];
    t += C.name_len + 1;
    for (uint32_t k = 0; k < nn; k++) t[k] = (uint8_t)num[k];
    BzEmitSink s; s.litlen = litlen; s.dist = dist; s.w.init(out32, bitpos);
    bz_tokens(C.name, C.name_len, num, nn, ls, le, lc.pv, s);
    if (lc.last) { const uint32_t e = litlen[256]; s.w.put(e & 0xffffu, e >> 16); }
    s.w.finish();
}
// member b: compressed payload bytes and member length from the bit count of its lines
MSD_HD uint32_t bz_payload_bytes(uint32_t hdr_bits, uint64_t data_bits) { return (uint32_t)((hdr_bits + data_bits + 7) >> 3); }
MSD_HD void bz_frame(uint8_t* o, const BzCodes& C, uint32_t payload, uint32_t crc, uint32_t isize) {
    const uint32_t blen = payload + 26;
    const uint8_t h[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    for (int k = 0; k < 16; k++) o[k] = h[k];
    o[16] = (uint8_t)((blen - 1) & 0xff); o[17] = (uint8_t)((blen - 1) >> 8);
    const uint32_t hb = C.hdr_bits >> 3;
    for (uint32_t k = 0; k < hb; k++) o[18 + k] = C.hdr[k];
    if (C.hdr_bits & 7) o[18 + hb] |= C.hdr[hb];             // the byte the first line's bits share with the header
    uint8_t* t = o + 18 + payload;
    for (int k = 0; k < 4; k++) { t[k] = (uint8_t)(crc >> (8 * k)); t[4 + k] = (uint8_t)(isize >> (8 * k)); }
}

// ---- host: the contig's Huffman code ----------------------------------------------------------------------------------
struct BzHistSink {
    uint64_t* litlen; uint64_t* dist;
    MSD_HD void lit(uint8_t b) { litlen[b]++; }
    MSD_HD void match(uint32_t len, uint32_t d) { uint32_t s, eb, ex; bz_len_symbol(len, &s, &eb, &ex); litlen[s]++; bz_dist_symbol(d, &s, &eb, &ex); dist[s]++; }
};

// code lengths of a complete prefix code for freq[0..n) (all > 0), none longer than `limit`
inline void bz_huffman_lengths(const std::vector<uint64_t>& freq_in, int limit, std::vector<uint8_t>& len) {
    const int n = (int)freq_in.size();
    std::vector<uint64_t> freq = freq_in;
    len.assign((size_t)n, 0);
    for (;;) {
        // two-queue Huffman over symbols sorted by frequency
        std::vector<int> order((size_t)n); for (int i = 0; i < n; i++) order[(size_t)i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return freq[(size_t)a] < freq[(size_t)b]; });
        std::vector<uint64_t> w((size_t)(2 * n)); std::vector<int> parent((size_t)(2 * n), -1);
        for (int i = 0; i < n; i++) w[(size_t)i] = freq[(size_t)order[(size_t)i]];
        int leaf = 0, node = n, made = n;
        auto take = [&]() { int r; if (leaf < n && (node >= made || w[(size_t)leaf] <= w[(size_t)node])) r = leaf++; else r = node++; return r; };
        while (made < 2 * n - 1) { const int a = take(), b = take(); w[(size_t)made] = w[(size_t)a] + w[(size_t)b]; parent[(size_t)a] = parent[(size_t)b] = made; made++; }
        int maxlen = 0;
        for (int i = 0; i < n; i++) { int d = 0; for (int p = i; parent[(size_t)p] >= 0; p = parent[(size_t)p]) d++; len[(size_t)order[(size_t)i]] = (uint8_t)d; maxlen = std::max(maxlen, d); }
        if (maxlen <= limit) return;
        for (auto& f : freq) f = (f + 1) / 2;                    // flatter frequencies -> shallower tree; all-equal ends it
    }
}
// canonical codes (RFC 1951 3.2.2), bit-reversed for LSB-first packing: code | len << 16
inline void bz_canonical(const std::vector<uint8_t>& len, uint32_t* out) {
    uint32_t bl_count[16] = {0}, next_code[16] = {0};
    for (uint8_t l : len) bl_count[l]++;
    bl_count[0] = 0;
    uint32_t code = 0;
    for (int b = 1; b < 16; b++) { code = (code + bl_count[b - 1]) << 1; next_code[b] = code; }
    for (size_t i = 0; i < len.size(); i++) {
        const uint32_t l = len[i]; if (!l) { out[i] = 0; continue; }
        const uint32_t c = next_code[l]++; uint32_t r = 0;
        for (uint32_t k = 0; k < l; k++) r |= ((c >> k) & 1u) << (l - 1 - k);
        out[i] = r | (l << 16);
    }
}

// Builds the code of a contig from a sample of its own lines and the member header bits (BFINAL=1, BTYPE=2, HLIT/HDIST/HCLEN,
// code-length code, the 316 code lengths with repeat symbol 16).  Returns false if the name is too long for the device path.
// labels (n_labels > 0): dp[i] indexes them (quantized.bed); every dp[i] must be in range.
inline bool bz_build_codes(const char* chrom, const int32_t* st, const int32_t* en, const int32_t* dp, uint64_t n, BzCodes* C, const char* const* labels = nullptr, int n_labels = 0) {
    const size_t nl = strlen(chrom);
    if (nl > (size_t)kBzMaxName || n_labels < 0 || n_labels > kBzMaxLabels) return false;
    memset(C, 0, sizeof *C);
    C->name_len = (uint32_t)nl; memcpy(C->name, chrom, nl); C->name[nl] = '\t';
    for (int k = 0; k < n_labels; k++) { const size_t ll = strlen(labels[k]); if (ll > (size_t)kBzMaxLabel) return false; C->lab_len[k] = (uint8_t)ll; memcpy(C->lab[k], labels[k], ll); }
    C->n_labels = (uint32_t)n_labels;
    if (n_labels) for (uint64_t i = 0; i < n; i++) if ((uint32_t)dp[i] >= (uint32_t)n_labels) return false;
    std::vector<uint64_t> fl(286, 0), fd(30, 0);
    BzHistSink hs; hs.litlen = fl.data(); hs.dist = fd.data();
    const uint64_t want = 1u << 14, step = n > want ? n / want : 1;      // (16,384 lines: 1.6 ms of host time per contig; 65,536 cost 6.4 ms for a code that is no better)
    uint64_t lines = 0;
    for (uint64_t i = 0; i < n; i += step) {                       // evenly spaced lines, each with its true predecessor
        char num[kBzNumMax]; uint32_t ls, le;
        const uint32_t nn = bz_format_num(*C, num, st[i], en[i], dp[i], &ls, &le);
        BzPrev pv; pv.has = i > 0; pv.chained = false; pv.len = 0; pv.ls = 0; pv.lab_dist = bz_label_dist(*C, st, en, dp, i, false, 0);
        if (i > 0) { pv.len = bz_line_len(*C, st[i - 1], en[i - 1], dp[i - 1]); pv.ls = bz_idigits(st[i - 1]); pv.chained = en[i - 1] == st[i]; }
        bz_tokens(C->name, C->name_len, num, nn, ls, le, pv, hs);
        lines++;
    }
    // a member's first line is all literals: about one line in (32 KB / 20 B); and one end-of-block code per member
    for (size_t k = 0; k <= nl; k++) fl[C->name[k]] += lines / 1024 + 1;
    fl[256] += lines / 1024 + 1;
    for (auto& f : fl) f = f * 16 + 1;                             // every symbol keeps a code
    for (auto& f : fd) f = f * 16 + 1;
    std::vector<uint8_t> ll, dl;
    bz_huffman_lengths(fl, 15, ll); bz_huffman_lengths(fd, 15, dl);
    bz_canonical(ll, C->litlen); bz_canonical(dl, C->dist);
    // code-length sequence with repeats (symbol 16: copy the previous length 3..6 times, 2 extra bits)
    std::vector<uint8_t> seq(ll); seq.insert(seq.end(), dl.begin(), dl.end());
    struct Cl { uint8_t sym, extra; };
    std::vector<Cl> cl;
    for (size_t i = 0; i < seq.size();) {
        size_t run = 1; while (i + run < seq.size() && seq[i + run] == seq[i]) run++;
        cl.push_back(Cl{seq[i], 0}); size_t left = run - 1;
        while (left >= 3) { const size_t r = std::min<size_t>(left, 6); cl.push_back(Cl{16, (uint8_t)(r - 3)}); left -= r; }
        while (left--) cl.push_back(Cl{seq[i], 0});
        i += run;
    }
    std::vector<uint64_t> fc(19, 0); for (const Cl& c : cl) fc[c.sym]++;
    std::vector<int> used; for (int s = 0; s < 19; s++) if (fc[(size_t)s]) used.push_back(s);
    std::vector<uint8_t> cll(19, 0);
    if (used.size() == 1) { cll[(size_t)used[0]] = 1; used.push_back(used[0] == 0 ? 1 : 0); cll[(size_t)used[1]] = 1; }   // a lone code needs a sibling to be complete
    else { std::vector<uint64_t> f; for (int s : used) f.push_back(fc[(size_t)s]); std::vector<uint8_t> l; bz_huffman_lengths(f, 7, l); for (size_t k = 0; k < used.size(); k++) cll[(size_t)used[k]] = l[k]; }
    uint32_t clc[19]; bz_canonical(cll, clc);
    static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19; while (hclen > 4 && cll[(size_t)order[hclen - 1]] == 0) hclen--;
    std::vector<uint32_t> words(kBzHdrMaxBytes / 4 + 2, 0);
    BzBitWriter w; w.init(words.data(), 0);
    w.put(1, 1); w.put(2, 2);                                       // BFINAL = 1, BTYPE = 10 (dynamic)
    w.put(286 - 257, 5); w.put(30 - 1, 5); w.put((uint32_t)(hclen - 4), 4);
    for (int k = 0; k < hclen; k++) w.put(cll[(size_t)order[k]], 3);
    uint64_t bits = 17 + 3ull * (uint64_t)hclen;
    for (const Cl& c : cl) { bits += (clc[c.sym] >> 16) + (c.sym == 16 ? 2 : 0); }
    if (bits > 8ull * kBzHdrMaxBytes) return false;
    for (const Cl& c : cl) { w.put(clc[c.sym] & 0xffffu, clc[c.sym] >> 16); if (c.sym == 16) w.put(c.extra, 2); }
    w.finish();
    C->hdr_bits = (uint32_t)bits;
    memcpy(C->hdr, words.data(), kBzHdrMaxBytes);
    return true;
}

}  // namespace msd
