// inflate_model.cpp -- host replay of the arithmetic of the two inflate kernels (mosdepth_b200/csrc/inflate.cuh):
// same table-entry encodings, canonical bookkeeping, root-by-index fill, long-code search, token format and the
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
    uint16_t lit_root[256], dist_root[128], long_lit[286];
    uint8_t long_dist[30];
    uint32_t llim[7], lbase[7], dlim[8], dbase[8];
};

static uint32_t brev32(uint32_t v) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }

// mirrors coop_build(): ranks by (length, symbol order), canonical bookkeeping, root filled by index
static uint32_t build(const uint8_t* lens, int hlit, int hdist, Tables& T) {
    uint32_t status = kOk;
    for (int pass = 0; pass < 2; pass++) {
        const int n = pass ? hdist : hlit, root = pass ? kDistRootBits : kLitRootBits;
        const uint8_t* L = pass ? lens + hlit : lens;
        uint32_t cnt[16] = {0};
        std::vector<uint32_t> rank(n);
        for (int s = 0; s < n; s++) rank[s] = cnt[L[s]]++;
        msd::inf::Canon c; canon_from_counts(cnt, c);
        if (c.oversubscribed) status = kOversubscribed;
        uint16_t sorted[320];
        uint32_t nsyms = 0;
        for (int s = 0; s < n; s++) if (L[s]) { sorted[c.offs[L[s]] + rank[s]] = (uint16_t)s; nsyms++; }
        for (uint32_t i = 0; i < (1u << root); i++) {
            const uint32_t cw = brev32(i) >> (32 - root);
            uint32_t len = 1;
            for (int k = 1; k <= root; k++) len += cw >= canon_limit(c, k, root);
            uint32_t ent = 0;
            if (len <= (uint32_t)root) {
                const uint32_t sym = sorted[c.offs[len] + (cw >> (root - len)) - c.first[len]];
                ent = (pass ? dist_entry_of(sym) : lit_entry_of(sym)) | len;
            }
            if (pass) T.dist_root[i] = (uint16_t)ent; else T.lit_root[i] = (uint16_t)ent;
        }
        const uint32_t offs_long = c.offs[root + 1];
        for (uint32_t i = 0; i + offs_long < nsyms; i++) {
            if (pass) T.long_dist[i] = (uint8_t)(dist_entry_of(sorted[offs_long + i]) >> 4); else T.long_lit[i] = (uint16_t)lit_entry_of(sorted[offs_long + i]);
        }
        for (int k = root + 1; k < 16; k++) {
            if (pass) { T.dlim[k - 8] = canon_limit(c, k, 15); T.dbase[k - 8] = canon_long_base(c, k, root); }
            else { T.llim[k - 9] = canon_limit(c, k, 15); T.lbase[k - 9] = canon_long_base(c, k, root); }
        }
    }
    return status;
}

struct BitReader {
    uint64_t bb = 0; int bc = 0; uint32_t widx = 0, ahead = 0;
    const uint32_t* words; uint32_t wlast;
    void seed(uint32_t byte_pos) { bb = 0; bc = 0; widx = byte_pos >> 2; ahead = words[std::min(widx, wlast)]; refill(); drop((int)(byte_pos & 3u) * 8); refill(); }
    void refill() { if (bc <= 32) { bb |= (uint64_t)ahead << bc; bc += 32; widx++; ahead = words[std::min(widx, wlast)]; } }
    uint32_t lo() const { return (uint32_t)bb; }
    void drop(int n) { bb >>= n; bc -= n; }
    uint32_t take(int n) { const uint32_t v = (uint32_t)bb & ((1u << n) - 1u); drop(n); return v; }
};

static const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// K1a for one block: literals into dst, matches into tok.  Returns status.
static uint32_t decode_block(const uint8_t* comp_aligned, uint32_t lead, uint32_t clen, uint32_t isize, uint8_t* dst, std::vector<uint32_t>& tok) {
    BitReader br; br.words = reinterpret_cast<const uint32_t*>(comp_aligned);
    const uint32_t total_words = std::max((lead + clen + 3) / 4, 1u); br.wlast = total_words - 1;
    br.seed(lead);
    uint32_t outp = 0, lit_run = 0; int bfinal = 0;
    Tables T;
    while (!bfinal) {
        br.refill();
        if (br.widx > total_words + 4u) return kTruncated;
        bfinal = (int)br.take(1);
        const int btype = (int)br.take(2);
        uint8_t lens[320]; int hlit = 288, hdist = 30;
        if (btype == 0) {
            br.drop(br.bc & 7); br.refill();
            const uint32_t len = br.take(16), nlen = br.take(16);
            const uint32_t sp = br.widx * 4u - (uint32_t)br.bc / 8u;
            if ((len ^ 0xffffu) != nlen || sp + len > lead + clen) return kBadStored;
            if (outp + len > isize) return kOverflow;
            memcpy(dst + outp, comp_aligned + sp, len);
            outp += len; lit_run += len;
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
            uint32_t e = T.lit_root[lo & 255u];
            if ((e & 15u) == 0) {
                const uint32_t c = brev32(lo) >> 17;
                uint32_t n = 0, base = T.lbase[0];
                for (int k = 0; k < 6; k++) if (c >= T.llim[k]) { n = k + 1; base = T.lbase[k + 1]; }
                if (c >= T.llim[6]) return kBadSymbol;
                const uint32_t l = 9 + n; e = T.long_lit[(base + (c >> (15 - l))) & 0xffffu] | l;
            }
            br.drop((int)(e & 15u));
            const uint32_t nib = (e >> 4) & 15u;
            if (nib & 8u) { if (outp >= isize) return kOverflow; dst[outp++] = (uint8_t)(e >> 8); lit_run++; continue; }
            if (nib >= kNibEob) { if (nib == kNibEob) break; return kBadSymbol; }
            const uint32_t len = 3u + ((e >> 8) & 0xffu) + br.take((int)nib);
            br.refill();
            lo = br.lo();
            uint32_t d = T.dist_root[lo & 127u];
            if ((d & 15u) == 0) {
                const uint32_t c = brev32(lo) >> 17;
                uint32_t n = 0, base = T.dbase[0];
                for (int k = 0; k < 7; k++) if (c >= T.dlim[k]) { n = k + 1; base = T.dbase[k + 1]; }
                if (c >= T.dlim[7]) return kBadDistance;
                const uint32_t l = 8 + n; d = ((uint32_t)T.long_dist[(base + (c >> (15 - l))) & 0xffffu] << 4) | l;
            }
            br.drop((int)(d & 15u));
            const uint32_t x = (d >> 4) & 15u;
            const uint32_t dist = 1u + (((d >> 8) & 3u) << x) + br.take((int)x);
            if (x == kDistNibBad || dist > outp) return kBadDistance;
            if (outp + len > isize) return kOverflow;
            if (lit_run > 510u) { tok.push_back((kTokSkip << 23) | lit_run); lit_run = 0; }
            tok.push_back(make_token(lit_run, len, dist)); lit_run = 0; outp += len;
        }
    }
    return outp == isize ? kOk : kShort;
}

// K1b for one block, lane by lane with the kernel's window/round schedule.  `mis` plays the output misalignment.
static void resolve_block(uint8_t* dst, const std::vector<uint32_t>& tok, int mis) {
    const uint32_t nt = (uint32_t)tok.size();
    int pos_base = 0;
    for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
        int lr[32], len[32], dist[32], adv[32], incl[32], start[32], mstart[32];
        int run = 0;
        for (int lane = 0; lane < 32; lane++) {
            lr[lane] = len[lane] = dist[lane] = adv[lane] = 0;
            const uint32_t t = t0 + lane;
            if (t < nt) {
                const uint32_t w = tok[t];
                lr[lane] = (int)(w >> 23);
                if (lr[lane] == (int)kTokSkip) { lr[lane] = (int)(w & 0x7fffffu); adv[lane] = lr[lane]; }
                else { len[lane] = (int)(w & 0xffu) + 3; dist[lane] = (int)((w >> 8) & 0x7fffu) + 1; adv[lane] = lr[lane] + len[lane]; }
            }
            run += adv[lane]; incl[lane] = run;
            start[lane] = pos_base + incl[lane] - adv[lane]; mstart[lane] = start[lane] + lr[lane];
        }
        const int total = incl[31], hi = pos_base + total;
        int before = 0;
        for (int wb = ((pos_base + mis) & ~31) - mis; wb < hi; wb += 32) {
            uint32_t heads = 0;
            for (int lane = 0; lane < 32; lane++) if (adv[lane] > 0 && start[lane] >= wb && start[lane] < wb + 32) heads |= 1u << (start[lane] - wb);
            bool pend[32]; int src[32];
            for (int lane = 0; lane < 32; lane++) {
                const uint32_t le = 0xffffffffu >> (31 - lane);
                const int p = wb + lane, m = before + __builtin_popcount(heads & le) - 1;
                const int ms = mstart[m & 31], md = dist[m & 31];
                pend[lane] = p >= pos_base && p < hi && m >= 0 && md > 0 && p >= ms;
                src[lane] = p - md;
                if (pend[lane] && p - ms >= md) src[lane] = ms - md + (p - ms) % md;
            }
            before += __builtin_popcount(heads);
            for (;;) {
                uint32_t done = 0;
                for (int lane = 0; lane < 32; lane++) if (!pend[lane]) done |= 1u << lane;
                if (done == 0xffffffffu) break;
                uint8_t v[32]; bool ready[32];
                for (int lane = 0; lane < 32; lane++) {
                    ready[lane] = pend[lane] && (src[lane] < wb || ((done >> (src[lane] - wb)) & 1u));
                    if (ready[lane]) v[lane] = dst[src[lane]];
                }
                for (int lane = 0; lane < 32; lane++) if (ready[lane]) { dst[wb + lane] = v[lane]; pend[lane] = false; }
            }
        }
        pos_base = hi;
    }
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: inflate_model file [max_blocks]\n"); return 2; }
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
            const uint32_t st = decode_block(raw.data() + 8, lead, clen, isize, got.data() + mis, tok);
            if (st == kOk) resolve_block(got.data() + mis, tok, mis);
            if (st != kOk || memcmp(got.data() + mis, want.data(), isize) != 0) {
                bad++;
                size_t d = 0; while (d < isize && got[mis + d] == want[d]) d++;
                fprintf(stderr, "block %ld (lead %d): status %u, first difference at %zu of %u\n", nblk, lead, st, d, isize);
                if (bad > 5) return 1;
            }
            if ((int)tok.size() > kTokenCap) { fprintf(stderr, "token cap exceeded\n"); return 1; }
            ntok += tok.size(); bytes += isize;
            if (lead) memmove(raw.data() + 8, raw.data() + 8 + lead, clen + 8);
        }
        nblk++;
    }
    printf("{\"blocks\": %ld, \"bad\": %ld, \"bytes\": %llu, \"tokens\": %ld}\n", nblk, bad, bytes, ntok);
    return bad ? 1 : 0;
}
