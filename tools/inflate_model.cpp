// inflate_model.cpp -- host replay of the arithmetic of the two inflate kernels (mosdepth_b200/csrc/inflate.cuh):
// same table-entry encodings, canonical bookkeeping, limit-compare code search, token format and the
// 32-byte-window resolve schedule, executed lane by lane on the CPU and compared with zlib.  A design/debug aid and a
// CPU test of inflate_core.cuh (tests/test_inflate_model.py); it is NOT a fallback and nothing in the library uses it.
//   usage: inflate_model file.bam|file.gz [max_blocks]      (any BGZF file)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <zlib.h>
#include "../mosdepth_b200/csrc/inflate_core.cuh"

using namespace msd::inf;

struct Tables {
    uint16_t lit_sym[289], dist_sym[31]; int16_t lit_base[17], dist_base[17];
    uint32_t llim[15], dlim[15];
};

static uint32_t brev32(uint32_t v) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }

// mirrors coop_build(): ranks by (length, symbol order), canonical bookkeeping, symbols stored as entries
static uint32_t build(const uint8_t* lens, int hlit, int hdist, Tables& T) {
    uint32_t status = kOk;
    memset(&T, 0xA5, sizeof T);   // stale entries of an earlier block
    for (int pass = 0; pass < 2; pass++) {
        const int n = pass ? hdist : hlit;
        const uint8_t* L = pass ? lens + hlit : lens;
        uint32_t cnt[16] = {0};
        std::vector<uint32_t> rank(n);
        for (int s = 0; s < n; s++) rank[s] = cnt[L[s]]++;
        msd::inf::Canon c; canon_from_counts(cnt, c);
        if (c.oversubscribed) status = kOversubscribed;
        for (int s = 0; s < n; s++) if (L[s]) {
            const uint32_t at = c.offs[L[s]] + rank[s];
            if (pass) T.dist_sym[at] = (uint16_t)dist_entry_of(s); else T.lit_sym[at] = (uint16_t)lit_entry_of(s);
        }
        for (int k = 1; k < 16; k++) {
            if (pass) { T.dlim[k - 1] = canon_limit(c, k); T.dist_base[k] = (int16_t)canon_base(c, k); }
            else { T.llim[k - 1] = canon_limit(c, k); T.lit_base[k] = (int16_t)canon_base(c, k); }
        }
        if (pass) { T.dist_base[16] = kDistBadSlot; T.dist_sym[kDistBadSlot] = kDistNibBad << 4; }
        else { T.lit_base[16] = kLitBadSlot; T.lit_sym[kLitBadSlot] = kEntBad; }
    }
    return status;
}

// the definition (number of limits <= c) checked against the kernels' packed evaluation (canon_len_minus1, inflate_core.cuh)
static uint32_t len_minus1(uint32_t c, const uint32_t* lim) {
    uint32_t n = 0; for (int k = 0; k < 15; k++) n += c >= lim[k];
    uint32_t limp[8];
    for (int j = 0; j < 8; j++) limp[j] = std::min(lim[2 * j], 0x8000u) | ((j == 7 ? 0x8000u : std::min(lim[2 * j + 1], 0x8000u)) << 16);
    if (canon_len_minus1(c, limp) != n) { fprintf(stderr, "canon_len_minus1 disagrees with the definition (c = %u)\n", c); abort(); }
    return n;
}

struct BitReader {      // the kernel's first reader without its ring (the ring only stages the words)
    uint64_t bb = 0; int bc = 0; uint32_t widx = 0, a0 = 0, a1 = 0;
    const uint32_t* words; uint32_t wlast;
    BitReader(const uint8_t* comp_aligned, uint32_t total_words) : words(reinterpret_cast<const uint32_t*>(comp_aligned)), wlast(total_words - 1) {}
    void seed(uint32_t byte_pos) { bb = 0; bc = 0; widx = byte_pos >> 2; a0 = words[std::min(widx, wlast)]; a1 = words[std::min(widx + 1, wlast)]; refill(); drop((int)(byte_pos & 3u) * 8); refill(); }
    void refill() { if (bc <= 32) { bb |= (uint64_t)a0 << bc; bc += 32; widx++; a0 = a1; a1 = words[std::min(widx + 1, wlast)]; } }
    void refill_before_distance() { refill(); }
    uint32_t lo() const { return (uint32_t)bb; }
    void drop(int n) { bb >>= n; bc -= n; }
    uint32_t take(int n) { const uint32_t v = (uint32_t)bb & ((1u << n) - 1u); drop(n); return v; }
    uint32_t bits_to_byte() const { return (uint32_t)bc & 7u; }
    uint32_t byte_pos() const { return widx * 4u - (uint32_t)bc / 8u; }
    uint32_t word_pos() const { return widx; }
};
// BitReader2 of inflate_core.cuh (the code the kernel runs with MSD_BITREADER2) over a host ring of 16 words
struct RingReader {
    BitReader2<1> r; uint32_t ringbuf[16]; const uint4* chunks; uint32_t clast;
    RingReader(const uint8_t* comp_aligned, uint32_t total_words) : chunks(reinterpret_cast<const uint4*>(comp_aligned)), clast((total_words * 4u + 15u) / 16u - 1u) {
        memset(ringbuf, 0xA5, sizeof ringbuf); r.ring.base = ringbuf; r.bitpos = 0; r.cur = 0; r.p0 = r.p1 = r.p2 = r.p3 = 0;
    }
    void seed(uint32_t byte_pos) { r.seed(chunks, clast, ringbuf, byte_pos); }
    void refill() { r.refill(chunks, clast, ringbuf); }
    void refill_before_distance() {}           // the kernel skips it: the ring holds the next two chunks
    uint32_t lo() const { return r.lo(); }
    void drop(int n) { r.drop(n); }
    uint32_t take(int n) { return r.take(n); }
    uint32_t bits_to_byte() const { return r.bits_to_byte(); }
    uint32_t byte_pos() const { return r.byte_pos(); }
    uint32_t word_pos() const { return r.word_pos(); }
};

static const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// K1a for one block: literals, packed in stream order, into `lits` (the kernel: downwards from the top of the block's token region),
// matches into tok, a final skip token for the literals behind the last match.  Returns status.
template <class Reader>
static uint32_t decode_block(const uint8_t* comp_aligned, uint32_t lead, uint32_t clen, uint32_t isize, std::vector<uint8_t>& lits, std::vector<uint32_t>& tok, bool packed, uint8_t* dst) {
    const uint32_t total_words = std::max((lead + clen + 3) / 4, 1u);
    Reader br(comp_aligned, total_words);
    br.seed(lead);
    uint32_t outp = 0, run_base = 0; int bfinal = 0;
    Tables T;
    while (!bfinal) {
        br.refill();
        if (br.word_pos() > total_words + 4u) return kTruncated;
        bfinal = (int)br.take(1);
        const int btype = (int)br.take(2);
        uint8_t lens[320]; int hlit = 288, hdist = 30;
        if (btype == 0) {
            br.drop((int)br.bits_to_byte()); br.refill();
            const uint32_t len = br.take(16), nlen = br.take(16);
            const uint32_t sp = br.byte_pos();
            if ((len ^ 0xffffu) != nlen || sp + len > lead + clen) return kBadStored;
            if (outp + len > isize) return kOverflow;
            if (packed) lits.insert(lits.end(), comp_aligned + sp, comp_aligned + sp + len); else memcpy(dst + outp, comp_aligned + sp, len);
            outp += len;
            br.seed(sp + len);
            continue;
        } else if (btype == 1) {
            for (int i = 0; i < 288; i++) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
            for (int i = 0; i < 30; i++) lens[288 + i] = 5;
        } else if (btype == 2) {
            hlit = (int)br.take(5) + 257; hdist = (int)br.take(5) + 1;
            const int hclen = (int)br.take(4) + 4;
            if (hlit > 286 || hdist > 30) return kBadCodeLengths;
            uint64_t clp = 0;
            for (int i = 0; i < hclen; i++) { br.refill(); clp |= (uint64_t)br.take(3) << (3 * cl_order[i]); }
            uint64_t cntp = 0;
            for (int s = 0; s < 19; s++) cntp += 1ull << (5 * (int)((clp >> (3 * s)) & 7));
            cntp &= ~31ull;
            int left = 1; uint32_t code = 0; uint64_t nextp = 0;
            for (int q = 1; q < 8; q++) {
                const uint32_t c = (uint32_t)(cntp >> (5 * q)) & 31u;
                left = (left << 1) - (int)c; code <<= 1; nextp |= (uint64_t)(code & 0xffu) << (8 * q); code += c;
            }
            if (left < 0) return kOversubscribed;
            uint8_t cl_tab[128] = {0};
            for (int s = 0; s < 19; s++) {
                const uint32_t q = (uint32_t)(clp >> (3 * s)) & 7u;
                if (!q) continue;
                const uint32_t c = (uint32_t)(nextp >> (8 * q)) & 0xffu; nextp += 1ull << (8 * q);
                const uint32_t rev = brev32(c) >> (32 - q);
                for (uint32_t k = rev; k < 128; k += (1u << q)) cl_tab[k] = (uint8_t)((s << 3) | q);
            }
            const int total = hlit + hdist; int idx = 0;
            while (idx < total) {
                br.refill();
                const uint32_t e = cl_tab[br.lo() & 127u], q = e & 7u, sym = e >> 3;
                if (!q) return kBadCodeLengths;
                br.drop((int)q);
                if (sym < 16) { lens[idx++] = (uint8_t)sym; continue; }
                uint32_t rep, val = 0;
                if (sym == 16) { if (!idx) return kBadCodeLengths; val = lens[idx - 1]; rep = 3 + br.take(2); }
                else if (sym == 17) rep = 3 + br.take(3);
                else rep = 11 + br.take(7);
                if (idx + (int)rep > total) return kBadCodeLengths;
                while (rep--) lens[idx++] = (uint8_t)val;
            }
            if (lens[256] == 0) return kBadCodeLengths;
        } else return kBadBlockType;
        const uint32_t st = build(lens, hlit, hdist, T);
        if (st) return st;
        for (;;) {
            br.refill();
            uint32_t lo = br.lo();
            uint32_t c = brev32(lo) >> 17;
            uint32_t l1 = len_minus1(c, T.llim);
            uint32_t idx = (uint32_t)(T.lit_base[1 + l1] + (int)(c >> ((14u - l1) & 31u)));
            const uint32_t e = T.lit_sym[std::min(idx, (uint32_t)kLitBadSlot)];
            const uint32_t x = (e >> 4) & 7u;
            const uint32_t ext = (lo >> (l1 + 1)) & ((1u << x) - 1u);
            br.drop((int)(l1 + 1 + x));
            if ((e & 0x83u) == kEntNotLen) { if (outp >= isize) return kOverflow; if (packed) lits.push_back((uint8_t)(e >> 8)); else dst[outp] = (uint8_t)(e >> 8); outp++; continue; }
            if (e & kEntNotLen) { if (e == kEntEob) break; return kBadSymbol; }
            const uint32_t len = 3u + (e >> 8) + ext;
            br.refill_before_distance();
            lo = br.lo();
            c = brev32(lo) >> 17;
            l1 = len_minus1(c, T.dlim);
            idx = (uint32_t)(T.dist_base[1 + l1] + (int)(c >> ((14u - l1) & 31u)));
            const uint32_t d = T.dist_sym[std::min(idx, (uint32_t)kDistBadSlot)];
            const uint32_t dx = (d >> 4) & 15u;
            const uint32_t dist = 1u + (((d >> 8) & 3u) << dx) + ((lo >> (l1 + 1)) & ((1u << dx) - 1u));
            br.drop((int)(l1 + 1 + dx));
            if (dx == kDistNibBad || dist > outp) return kBadDistance;
            if (outp + len > isize) return kOverflow;
            uint32_t lit_run = outp - run_base;
            if (lit_run > 510u) { tok.push_back((kTokSkip << 23) | lit_run); lit_run = 0; }
            tok.push_back(make_token(lit_run, len, dist)); outp += len; run_base = outp;
        }
    }
    if (outp != isize) return kShort;
    if (packed && outp > run_base) tok.push_back((kTokSkip << 23) | (outp - run_base));      // the literals behind the last match
    return kOk;
}

// K1b for one block, lane by lane with the kernel's batch/window/round schedule (31 tokens in lanes 1..31, lane 0 a
// sentinel).  `mis` plays the output misalignment.
static void resolve_block(uint8_t* dst, const std::vector<uint32_t>& tok, const std::vector<uint8_t>& lits, int mis, bool packed) {
    const int nt = (int)tok.size();
    int pos_base = 0, lit_base = 0;
    for (int t0 = 0; t0 < nt; t0 += 31) {
        int lr[32], dist[32], adv[32], incl[32], start[32], mstart[32], head_win[32], ldelta[32]; uint32_t head_bit[32];
        int run = 0, lrun = 0;
        for (int lane = 0; lane < 32; lane++) {
            lr[lane] = dist[lane] = adv[lane] = 0; bool is_match = false;
            if (lane >= 1 && t0 + lane - 1 < nt) {
                const uint32_t w = tok[t0 + lane - 1];
                lr[lane] = (int)(w >> 23);
                if (lr[lane] == (int)kTokSkip) { lr[lane] = (int)(w & 0x7fffffu); adv[lane] = lr[lane]; }
                else { dist[lane] = (int)((w >> 8) & 0x7fffu) + 1; adv[lane] = lr[lane] + (int)(w & 0xffu) + 3; is_match = true; }
            }
            run += adv[lane]; incl[lane] = run;
            start[lane] = pos_base + incl[lane] - adv[lane]; mstart[lane] = is_match ? start[lane] + lr[lane] : 0x3fffffff;
            ldelta[lane] = (lit_base + lrun) - start[lane];        // literal index of byte p of this token's literal run = p + ldelta
            lrun += lr[lane];
        }
        const int total = incl[31], hi = pos_base + total;
        const int wb0 = ((pos_base + mis) & ~31) - mis;
        for (int lane = 0; lane < 32; lane++) { head_win[lane] = adv[lane] > 0 ? (start[lane] - wb0) >> 5 : -1; head_bit[lane] = 1u << ((start[lane] - wb0) & 31); }
        int before = 0, j = 0;
        for (int wb = wb0; wb < hi; wb += 32, j++) {
            uint32_t heads = 0;
            for (int lane = 0; lane < 32; lane++) if (head_win[lane] == j) heads |= head_bit[lane];
            bool far[32], near[32], is_lit[32]; int back[32]; uint8_t v[32];
            for (int lane = 0; lane < 32; lane++) {
                const uint32_t le = 0xffffffffu >> (31 - lane);
                const int p = wb + lane, m = before + __builtin_popcount(heads & le);
                const int ms = mstart[m], md = dist[m];
                const int k = p - ms;
                const bool is_copy = k >= 0 && p < hi;
                is_lit[lane] = packed && !is_copy && p >= pos_base && p < hi;
                if (is_lit[lane]) v[lane] = lits[(size_t)(p + ldelta[m])];
                back[lane] = md;
                if (is_copy && k >= md) back[lane] = k - k % md + md;
                far[lane] = is_copy && back[lane] > lane; near[lane] = is_copy && back[lane] <= lane;
                if (far[lane]) v[lane] = dst[p - back[lane]];
            }
            before += __builtin_popcount(heads);
            bool any_near = false;
            for (int lane = 0; lane < 32; lane++) any_near |= near[lane];
            if (any_near) {
                int root[32];
                for (int lane = 0; lane < 32; lane++) {
                    const int p = wb + lane;
                    if (!far[lane] && !near[lane] && !is_lit[lane] && p >= 0 && p < hi) v[lane] = dst[p];      // a byte of an earlier batch
                    root[lane] = near[lane] ? lane - back[lane] : lane;
                }
                for (int r = 0; r < 5; r++) { int nr[32]; for (int lane = 0; lane < 32; lane++) nr[lane] = root[root[lane]]; memcpy(root, nr, sizeof nr); }
                uint8_t nv[32];
                for (int lane = 0; lane < 32; lane++) nv[lane] = v[root[lane]];
                memcpy(v, nv, sizeof nv);
            }
            for (int lane = 0; lane < 32; lane++) if (far[lane] || near[lane] || is_lit[lane]) dst[wb + lane] = v[lane];
        }
        pos_base = hi; lit_base += lrun;
    }
}

// ---- K1c replay: bgzf_crc_kernel's arithmetic, lane by lane (32 lanes, four word streams each) -----------------------
static CrcTables g_crc; static bool g_crc_built = false;
static uint32_t crc_block_model(const uint8_t* p, uint32_t n) {
    if (!g_crc_built) { crc_build_tables(g_crc); g_crc_built = true; }
    const uint32_t pad = (uint32_t)((uintptr_t)p & 15u), n2 = n + pad, R = n2 / kCrcRow, tail = n2 % kCrcRow;
    const uint8_t* base = p - pad;
    auto ldw = [&](uint32_t off, bool any_valid) -> CrcWords {   // the kernel's 16-byte load (bytes outside the message are masked by the callers)
        CrcWords w{0, 0, 0, 0}; if (!any_valid) return w;
        uint32_t v[4]; memcpy(v, base + off, 16); w.x = v[0]; w.y = v[1]; w.z = v[2]; w.w = v[3]; return w;
    };
    auto step = [&](uint32_t w) { return g_crc.u[0][w & 0xffu] ^ g_crc.u[1][(w >> 8) & 0xffu] ^ g_crc.u[2][(w >> 16) & 0xffu] ^ g_crc.u[3][w >> 24]; };
    auto t0 = [&](uint32_t i) { return g_crc.t0[i]; };
    uint32_t v = 0;
    for (int lane = 0; lane < 32; lane++) {
        uint32_t d[4] = {0, 0, 0, 0};
        if (R && pad && lane == 0) { const CrcWords c = crc_head_cancel(ldw(0, true), 0, pad); d[0] = c.x; d[1] = c.y; d[2] = c.z; d[3] = c.w; }
        for (uint32_t r = 0; r < R; r++) {
            const CrcWords w = ldw(r * kCrcRow + 16u * lane, true);
            d[0] = step(w.x ^ d[0]); d[1] = step(w.y ^ d[1]); d[2] = step(w.z ^ d[2]); d[3] = step(w.w ^ d[3]);
        }
        CrcWords t{0, 0, 0, 0};
        const uint32_t g0 = R * kCrcRow + 16u * lane;
        if (g0 < n2) t = crc_tail_keep(ldw(g0, true), g0, pad, n2);
        v ^= crc_multmodp(g_crc.xl[lane], crc_lane_fold(d, t, t0));
    }
    return crc_finish(v, g_crc.xa[tail], g_crc.xa[n % kCrcRow], g_crc.xbi[n / kCrcRow]);
}
// every length class x every alignment of the block inside the inflated buffer, against zlib's crc32
static int crc_selftest() {
    std::vector<uint8_t> buf(65536 + 64 + 512);
    uint64_t x = 0x9e3779b97f4a7c15ull;
    for (auto& b : buf) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = (uint8_t)(x >> 32); }
    uint8_t* al = buf.data() + ((16 - ((uintptr_t)buf.data() & 15)) & 15) + 16;
    const uint32_t sizes[] = {1, 2, 3, 4, 5, 15, 16, 17, 31, 33, 496, 497, 511, 512, 513, 527, 528, 1023, 1024, 1025, 4096, 12345, 65279, 65280, 65281, 65519, 65520, 65535, 65536};
    long bad = 0, n = 0;
    for (uint32_t sz : sizes) for (int pad = 0; pad < 16; pad++) {
        const uint32_t want = (uint32_t)crc32(0L, al + pad, sz), got = crc_block_model(al + pad, sz);
        n++; if (want != got) { if (bad++ < 5) fprintf(stderr, "crc model: size %u pad %d: %08x != %08x\n", sz, pad, got, want); }
    }
    std::fill(buf.begin(), buf.end(), 0);                    // all-zero and all-ones messages (start-state term on its own)
    for (int pad = 0; pad < 16; pad += 5) { n++; if (crc_block_model(al + pad, 65280) != (uint32_t)crc32(0L, al + pad, 65280)) bad++; }
    std::fill(buf.begin(), buf.end(), 0xff);
    for (int pad = 0; pad < 16; pad += 5) { n++; if (crc_block_model(al + pad, 777) != (uint32_t)crc32(0L, al + pad, 777)) bad++; }
    printf("{\"crc_cases\": %ld, \"bad\": %ld}\n", n, bad);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: inflate_model file [max_blocks] | inflate_model --crc-selftest\n"); return 2; }
    if (!strcmp(argv[1], "--crc-selftest")) return crc_selftest();
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    const long maxb = argc > 2 ? atol(argv[2]) : 1L << 40;
    long nblk = 0, bad = 0, ntok = 0; unsigned long long bytes = 0;
    std::vector<uint8_t> raw(1 << 17), want(65536 + 64), got(65536 + 64 + 32);
    while (nblk < maxb) {
        uint8_t h[12];
        if (fread(h, 1, 12, f) != 12) break;
        if (h[0] != 31 || h[1] != 139) { fprintf(stderr, "not BGZF at block %ld\n", nblk); return 1; }
        const int xlen = h[10] | (h[11] << 8);
        std::vector<uint8_t> extra(xlen);
        if (fread(extra.data(), 1, xlen, f) != (size_t)xlen) break;
        int bsize = -1;
        for (int i = 0; i + 4 <= xlen;) { const int sl = extra[i + 2] | (extra[i + 3] << 8); if (extra[i] == 'B' && extra[i + 1] == 'C') bsize = extra[i + 4] | (extra[i + 5] << 8); i += 4 + sl; }
        if (bsize < 0) { fprintf(stderr, "no BC field\n"); return 1; }
        const int clen = bsize + 1 - 12 - xlen - 8;
        for (int lead = 0; lead < 4; lead += 3) {   // exercise two alignments of the stream inside its word buffer
            if (lead && (nblk & 7)) continue;
            if (lead == 0) { if (fread(raw.data() + 8, 1, clen + 8, f) != (size_t)clen + 8) return 1; }
            else memmove(raw.data() + 8 + lead, raw.data() + 8, clen + 8);
            const uint8_t* cp = raw.data() + 8 + lead;
            const uint32_t isize = cp[clen + 4] | (cp[clen + 5] << 8) | (cp[clen + 6] << 16) | ((uint32_t)cp[clen + 7] << 24);
            z_stream zs; memset(&zs, 0, sizeof zs);
            inflateInit2(&zs, -15);
            zs.next_in = const_cast<uint8_t*>(cp); zs.avail_in = clen; zs.next_out = want.data(); zs.avail_out = 65536;
            const int zr = inflate(&zs, Z_FINISH); inflateEnd(&zs);
            if (zr != Z_STREAM_END || zs.total_out != isize) { fprintf(stderr, "zlib failed on block %ld\n", nblk); return 1; }
            std::vector<uint32_t> tok;
            const int mis = (int)(nblk % 32);
            std::fill(got.begin(), got.end(), 0xEE);
            uint32_t st = kOk;
            for (int packed = 0; packed < 2; packed++) {       // 0: literals in place (the two-kernel pipeline); 1: packed literals (the fused kernel)
                tok.clear();
                std::fill(got.begin(), got.end(), 0xEE);
                std::vector<uint8_t> lits;
                st = decode_block<BitReader>(raw.data() + 8, lead, clen, isize, lits, tok, packed != 0, got.data() + mis);
                {   // the second reader must produce the same literals, tokens and status
                    std::vector<uint32_t> tok2; std::vector<uint8_t> lits2; static std::vector<uint8_t> got2(65536 + 64 + 32);
                    std::fill(got2.begin(), got2.end(), 0xEE);
                    const uint32_t st2 = decode_block<RingReader>(raw.data() + 8, lead, clen, isize, lits2, tok2, packed != 0, got2.data() + mis);
                    if (st2 != st || tok2 != tok || lits2 != lits || memcmp(got2.data(), got.data(), got2.size()) != 0) { bad++; fprintf(stderr, "block %ld (lead %d): BitReader2 disagrees (status %u vs %u, %zu vs %zu tokens)\n", nblk, lead, st2, st, tok2.size(), tok.size()); if (bad > 5) return 1; }
                }
                if (st == kOk) resolve_block(got.data() + mis, tok, lits, mis, packed != 0);
                if (st == kOk && lits.size() + 4 * tok.size() > (size_t)kTokenCap * 4) { fprintf(stderr, "literals and tokens do not fit the block's token region\n"); return 1; }
                if (st != kOk || memcmp(got.data() + mis, want.data(), isize) != 0) {
                    bad++;
                    size_t d = 0; while (d < isize && got[mis + d] == want[d]) d++;
                    fprintf(stderr, "block %ld (lead %d, packed %d): status %u, first difference at %zu of %u\n", nblk, lead, packed, st, d, isize);
                    if (bad > 5) return 1;
                }
            }
            if ((int)tok.size() > kTokenCap) { fprintf(stderr, "token cap exceeded\n"); return 1; }
            if (lead == 0) {     // K1c: the block's CRC-32 by the row-strided formulation against the gzip trailer, at a varying alignment
                static std::vector<uint8_t> crcbuf(65536 + 128);
                uint8_t* al = crcbuf.data() + ((16 - ((uintptr_t)crcbuf.data() & 15)) & 15) + 16 + (nblk % 16);
                memcpy(al, want.data(), isize);
                const uint32_t trailer = cp[clen] | (cp[clen + 1] << 8) | (cp[clen + 2] << 16) | ((uint32_t)cp[clen + 3] << 24);
                if (crc_block_model(al, isize) != trailer) { bad++; fprintf(stderr, "block %ld: CRC model disagrees with the gzip trailer\n", nblk); if (bad > 5) return 1; }
            }
            ntok += tok.size(); bytes += isize;
            if (lead) memmove(raw.data() + 8, raw.data() + 8 + lead, clen + 8);
        }
        nblk++;
    }
    printf("{\"blocks\": %ld, \"bad\": %ld, \"bytes\": %llu, \"tokens\": %ld}\n", nblk, bad, bytes, ntok);
    return bad ? 1 : 0;
}
