// gen_bam.cpp -- deterministic synthetic BAM (+BAI, +BED) generator for the configs of BASELINE.json
// (SURVEY.md section 8d).  Test/bench infrastructure: produces coordinate-sorted, BGZF-compressed BAM files with
// the record shapes of short-read paired-end WGS (C2/C3), exome capture (C4) and nanopore long reads (C5).
//
//   gen_bam --config c3 --genome-fraction 0.05 --threads 16 --level 6 -o /dev/shm/c3.bam [--straddle] [--bed out.bed] [--cg-every N [--cg-short]] [--dense]
//
// Everything is a pure function of (seed, contig, tile, index): threads generate disjoint position tiles and the
// output is identical for any thread count.  BGZF blocks are written htslib-style (a record never straddles a block
// if it fits) unless --straddle is given (htsjdk-style: blocks are filled to the brim).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

static inline uint64_t mix64(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(mix64(seed)) {}
    uint64_t next() { s += 0x9E3779B97F4A7C15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    double normal() { double u1 = uni(), u2 = uni(); if (u1 < 1e-300) u1 = 1e-300; return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
};

struct Contig { std::string name; uint32_t len; };
static const Contig kGRCh38[] = {
    {"chr1", 248956422}, {"chr2", 242193529}, {"chr3", 198295559}, {"chr4", 190214555}, {"chr5", 181538259}, {"chr6", 170805979},
    {"chr7", 159345973}, {"chr8", 145138636}, {"chr9", 138394717}, {"chr10", 133797422}, {"chr11", 135086622}, {"chr12", 133275309},
    {"chr13", 114364328}, {"chr14", 107043718}, {"chr15", 101991189}, {"chr16", 90338345}, {"chr17", 83257441}, {"chr18", 80373285},
    {"chr19", 58617616}, {"chr20", 64444167}, {"chr21", 46709983}, {"chr22", 50818468}, {"chrX", 156040895}, {"chrY", 57227415}, {"chrM", 16569}};

struct Record {   // one alignment, before serialisation
    int32_t pos, end;          // reference span [pos, end)
    uint64_t order;            // tie-break inside equal pos (pair index * 2 + mate)
    std::vector<uint8_t> bytes; // block_size + payload
};

static void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; i++) v.push_back((uint8_t)(x >> (8 * i))); }
static void put16(std::vector<uint8_t>& v, uint32_t x) { v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); }

static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

struct ReadSpec {
    int32_t tid, pos, mpos, tlen; uint16_t flag; uint8_t mapq; std::string qname; std::vector<uint32_t> cigar; int32_t l_seq; uint64_t seed;
};

static int32_t cigar_reflen(const std::vector<uint32_t>& c) { int32_t r = 0; for (uint32_t x : c) { int op = x & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) r += x >> 4; } return r; }

// --cg-every N: every Nth record (and every record with more than 65,535 operations, always) keeps its CIGAR in a CG:B,I
// tag behind the <l_seq>S<reflen>N placeholder (SAM spec 4.2.2; htslib's bam_read1 swaps it back in)
static int g_cg_every = 0;
static bool g_cg_short = false;
static bool g_dense = false;   // --dense: one long read in six gets an operation every ~2 bp (more than 65,535 on a long read)

static void serialise(const ReadSpec& r, bool long_read, Record& out) {
    Rng g(r.seed);
    std::vector<uint8_t>& b = out.bytes; b.clear();
    const int32_t reflen = cigar_reflen(r.cigar);
    // htslib leaves the placeholder alone when the tag holds fewer operations than the placeholder's two: one-operation
    // CIGARs are only moved with --cg-short (they then count as <l_seq>S<reflen>N, in the reference as well)
    const bool cg = r.cigar.size() > 65535 || (g_cg_every > 0 && r.cigar.size() >= (g_cg_short ? 1u : 2u) && (r.seed >> 7) % (uint64_t)g_cg_every == 0);
    out.pos = r.pos; out.end = r.pos + std::max(reflen, 1);
    put32(b, 0);   // block_size patched below
    put32(b, (uint32_t)r.tid); put32(b, (uint32_t)r.pos);
    b.push_back((uint8_t)(r.qname.size() + 1)); b.push_back(r.mapq); put16(b, (uint32_t)reg2bin(out.pos, out.end));
    put16(b, cg ? 2u : (uint32_t)r.cigar.size()); put16(b, r.flag); put32(b, (uint32_t)r.l_seq);
    put32(b, (uint32_t)((r.flag & 1) ? r.tid : -1)); put32(b, (uint32_t)((r.flag & 1) ? r.mpos : -1)); put32(b, (uint32_t)r.tlen);
    b.insert(b.end(), r.qname.begin(), r.qname.end()); b.push_back(0);
    if (cg) { put32(b, ((uint32_t)r.l_seq << 4) | 4u); put32(b, ((uint32_t)reflen << 4) | 3u); }
    else for (uint32_t c : r.cigar) put32(b, c);
    static const uint8_t code[4] = {1, 2, 4, 8};
    for (int32_t i = 0; i < (r.l_seq + 1) / 2; i++) { uint32_t x = (uint32_t)g.next(); b.push_back((uint8_t)((code[x & 3] << 4) | code[(x >> 2) & 3])); }
    if (!long_read) {   // 4-level binned qualities with runs (NovaSeq-like)
        static const uint8_t q4[4] = {37, 25, 11, 2};
        uint8_t cur = 37;
        for (int32_t i = 0; i < r.l_seq; i++) { uint32_t x = (uint32_t)g.next(); if ((x & 0xff) < 40) { uint32_t y = (x >> 8) % 100; cur = q4[y < 70 ? 0 : y < 85 ? 1 : y < 95 ? 2 : 3]; } b.push_back(cur); }
    } else {            // nanopore: noisy qualities, poorly compressible
        for (int32_t i = 0; i < r.l_seq; i++) b.push_back((uint8_t)(3 + (g.next() % 28)));
    }
    const char* rg = "RGZgrp1"; b.insert(b.end(), rg, rg + 7); b.push_back(0);
    b.push_back('N'); b.push_back('M'); b.push_back('C'); b.push_back((uint8_t)(g.next() % 4));
    if (!long_read) { std::string md = "MDZ" + std::to_string(reflen); b.insert(b.end(), md.begin(), md.end()); b.push_back(0); }
    if (cg && (r.seed & 64)) { b.push_back('C'); b.push_back('G'); b.push_back('B'); b.push_back('I'); put32(b, (uint32_t)r.cigar.size()); for (uint32_t c : r.cigar) put32(b, c); }
    b.push_back('A'); b.push_back('S'); b.push_back('C'); b.push_back((uint8_t)(100 + g.next() % 50));
    if (cg && !(r.seed & 64)) { b.push_back('C'); b.push_back('G'); b.push_back('B'); b.push_back('I'); put32(b, (uint32_t)r.cigar.size()); for (uint32_t c : r.cigar) put32(b, c); }
    uint32_t bs = (uint32_t)b.size() - 4; memcpy(b.data(), &bs, 4);
}

struct Config {
    std::string name = "c3"; double fraction = 0.01; double coverage = 30; uint64_t seed = 20261019; int threads = 8; int level = 6; bool straddle = false;
    std::string out = "out.bam", bed; int read_len = 150; double paired_frac_long = 0.02; int only_contig = -1;
};

static const int kTile = 1 << 16;

// short-read pairs whose left read starts in tile `tile` of contig `tid`; appends both mates that start inside [lo, hi)
static void gen_short_tile(const Config& cfg, int tid, uint32_t clen, int64_t tile, int64_t lo, int64_t hi, double cov, std::vector<Record>& out,
                           const std::vector<std::pair<uint32_t, uint32_t>>* targets) {
    const int64_t t0 = tile * kTile, t1 = std::min<int64_t>(t0 + kTile, clen);
    if (t0 >= (int64_t)clen) return;
    const int L = cfg.read_len;
    double npairs_f = cov * (double)(t1 - t0) / (2.0 * L);
    if (targets) {   // exome: coverage applies to target bases in the tile (+ background set by caller through cov)
        npairs_f = 0;
        auto it = std::lower_bound(targets->begin(), targets->end(), std::make_pair((uint32_t)std::max<int64_t>(0, t0 - 12001), 0u));
        for (; it != targets->end() && (int64_t)it->first < t1; ++it) { int64_t a = std::max<int64_t>(it->first, t0), b = std::min<int64_t>(it->second, t1); if (b > a) npairs_f += cov * (double)(b - a) / (2.0 * L); }
        npairs_f += 0.5 * (double)(t1 - t0) / (2.0 * L);
    }
    Rng tg(mix64(cfg.seed ^ mix64((uint64_t)tid * 1000003ull + (uint64_t)tile)));
    int64_t npairs = (int64_t)npairs_f; if (tg.uni() < npairs_f - (double)npairs) npairs++;
    for (int64_t i = 0; i < npairs; i++) {
        Rng g(mix64(cfg.seed * 31 + (uint64_t)tid) ^ mix64(((uint64_t)tile << 24) + (uint64_t)i));
        int insert = (int)std::lround(400 + 90 * g.normal()); insert = std::max(L, std::min(1000, insert));
        int64_t lpos = t0 + (int64_t)g.below((uint32_t)(t1 - t0));
        if (lpos + insert + 2 >= (int64_t)clen) continue;
        int64_t rpos = lpos + insert - L;
        ReadSpec r[2];
        const bool fwd_first = g.below(2) == 0;
        const uint32_t kind1 = g.below(1000), kind2 = g.below(1000);
        const bool dup = g.below(100) == 0, mq0 = g.below(100) == 0, improper = g.below(200) == 0, supp = g.below(200) == 0;
        char qn[64]; snprintf(qn, sizeof qn, "SYN:%d:%lld:%lld:%u", 1 + tid, (long long)tile, (long long)i, g.below(100000));
        for (int m = 0; m < 2; m++) {
            ReadSpec& x = r[m]; x.tid = tid; x.qname = qn; x.l_seq = L; x.mapq = mq0 ? 0 : 60;
            x.seed = mix64(cfg.seed ^ ((uint64_t)tid << 48) ^ ((uint64_t)tile << 20) ^ (uint64_t)(2 * i + m));
            uint32_t kind = m == 0 ? kind1 : kind2;
            int64_t p = m == 0 ? lpos : rpos;
            if (kind < 850) x.cigar = {(uint32_t)L << 4};
            else if (kind < 930) { uint32_t k = 1 + g.below(60); if (m == 0) x.cigar = {(k << 4) | 4, ((uint32_t)L - k) << 4}; else x.cigar = {((uint32_t)L - k) << 4, (k << 4) | 4}; }
            else if (kind < 970) { uint32_t a = 20 + g.below(L - 40), d = 1 + g.below(10); x.cigar = {a << 4, (d << 4) | 2, ((uint32_t)L - a) << 4}; }
            else { uint32_t a = 20 + g.below(L - 40), ins = 1 + g.below(8); x.cigar = {a << 4, (ins << 4) | 1, ((uint32_t)L - a - ins) << 4}; }
            x.pos = (int32_t)p;
        }
        if (supp) { r[1].cigar = {(50u << 4) | 5, ((uint32_t)L - 50) << 4}; r[1].l_seq = L - 50; }
        const int64_t e0 = r[0].pos + cigar_reflen(r[0].cigar), e1 = r[1].pos + cigar_reflen(r[1].cigar);
        if (e1 + 1 >= (int64_t)clen || e0 + 1 >= (int64_t)clen) continue;
        const int32_t tl = (int32_t)(std::max(e0, e1) - r[0].pos);
        for (int m = 0; m < 2; m++) {
            ReadSpec& x = r[m];
            x.mpos = r[1 - m].pos; x.tlen = m == 0 ? tl : -tl;
            uint16_t f = 1 | (improper ? 0 : 2);
            const bool rev = (m == 0) ? !fwd_first : fwd_first;
            if (rev) f |= 16; else f |= 32;
            f |= ((m == 0) == fwd_first) ? 64 : 128;
            if (dup) f |= 1024;
            if (supp && m == 1) f |= 2048;
            x.flag = f;
            if (x.pos >= lo && x.pos < hi) { out.emplace_back(); out.back().order = (uint64_t)(((uint64_t)tile << 24) + i) * 2 + m; serialise(x, false, out.back()); }
        }
    }
}

static void gen_long_tile(const Config& cfg, int tid, uint32_t clen, int64_t tile, int64_t lo, int64_t hi, std::vector<Record>& out) {
    const int64_t t0 = tile * kTile, t1 = std::min<int64_t>(t0 + kTile, clen);
    if (t0 >= (int64_t)clen) return;
    const double mean_len = (100000.0 - 10000.0) / std::log(10.0);
    double n_f = cfg.coverage * (double)(t1 - t0) / mean_len;
    Rng tg(mix64(cfg.seed ^ mix64(0xABCDull + (uint64_t)tid * 7919ull + (uint64_t)tile)));
    int64_t n = (int64_t)n_f; if (tg.uni() < n_f - (double)n) n++;
    for (int64_t i = 0; i < n; i++) {
        Rng g(mix64(cfg.seed * 131 + (uint64_t)tid) ^ mix64(((uint64_t)tile << 24) + (uint64_t)i + 77));
        int64_t len = (int64_t)std::lround(10000.0 * std::pow(10.0, g.uni()));
        int64_t p = t0 + (int64_t)g.below((uint32_t)(t1 - t0));
        if (p + 1000 >= (int64_t)clen) continue;
        len = std::min<int64_t>(len, (int64_t)clen - p - 2);
        if (p < lo || p >= hi) continue;
        ReadSpec x; x.tid = tid; x.pos = (int32_t)p; x.mpos = -1; x.tlen = 0; x.mapq = (uint8_t)(1 + g.below(60));
        const bool supp = g.below(20) == 0;
        x.flag = (uint16_t)((g.below(2) ? 16 : 0) | (supp ? 2048 : 0));
        char qn[64]; snprintf(qn, sizeof qn, "%08x-%04x-%04x-%04x-%012llx", (uint32_t)g.next(), (uint32_t)(g.next() & 0xffff), (uint32_t)(g.next() & 0xffff), (uint32_t)(g.next() & 0xffff), (unsigned long long)(g.next() & 0xffffffffffffull));
        x.qname = qn;
        // op every ~10 bp: M 50 % / D 28 % / I 21 % (tests/nanopore.bam histogram), geometric lengths
        int64_t ref = 0, qlen = 0;
        if (supp) { uint32_t h = 100 + g.below(2000); x.cigar.push_back((h << 4) | 5); }
        const bool dense = g_dense && g.below(6) == 0;
        while (ref < len && x.cigar.size() < (g_dense ? 250000u : 60000u)) {
            uint32_t m = 1 + (uint32_t)(-std::log(1.0 - g.uni() * 0.999999) * (dense ? 1.2 : 18.0));
            m = (uint32_t)std::min<int64_t>(m, len - ref);
            if (m) { x.cigar.push_back(m << 4); ref += m; qlen += m; }
            if (ref >= len) break;
            uint32_t u = g.below(49);
            uint32_t l = 1 + (uint32_t)(-std::log(1.0 - g.uni() * 0.999999) * 1.5);
            if (u < 28) { l = (uint32_t)std::min<int64_t>(l, len - ref - 1); if (l) { x.cigar.push_back((l << 4) | 2); ref += l; } }
            else { x.cigar.push_back((l << 4) | 1); qlen += l; }
        }
        if ((x.cigar.back() & 15) != 0) x.cigar.push_back(1u << 4), qlen += 1;
        x.l_seq = (int32_t)qlen;
        x.seed = mix64(cfg.seed ^ 0x5555ull ^ ((uint64_t)tid << 48) ^ ((uint64_t)tile << 20) ^ (uint64_t)i);
        out.emplace_back(); out.back().order = (((uint64_t)tile << 24) + (uint64_t)i) * 2 + (1ull << 62); serialise(x, true, out.back());
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct BgzfOut {
    std::vector<uint8_t> data;     // compressed bytes
    std::vector<uint8_t> cur;      // pending uncompressed bytes
    int level; bool straddle;
    z_stream zs; bool zinit = false;
    void flush_block() {
        if (cur.empty()) return;
        if (!zinit) { memset(&zs, 0, sizeof zs); deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY); zinit = true; } else deflateReset(&zs);
        uint8_t buf[65536 + 1024];
        zs.next_in = cur.data(); zs.avail_in = (uInt)cur.size(); zs.next_out = buf + 18; zs.avail_out = sizeof buf - 18 - 8;
        int r = deflate(&zs, Z_FINISH);
        if (r != Z_STREAM_END) { fprintf(stderr, "deflate failed\n"); exit(1); }
        uint32_t clen = (uint32_t)zs.total_out, blen = clen + 26;
        static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
        memcpy(buf, hdr, 16); buf[16] = (uint8_t)(blen - 1); buf[17] = (uint8_t)((blen - 1) >> 8);
        uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), cur.data(), (uInt)cur.size()), isz = (uint32_t)cur.size();
        memcpy(buf + 18 + clen, &crc, 4); memcpy(buf + 22 + clen, &isz, 4);
        data.insert(data.end(), buf, buf + blen);
        cur.clear();
    }
    uint64_t tell() const { return ((uint64_t)data.size() << 16) | (uint64_t)cur.size(); }   // local virtual offset
    void write(const uint8_t* p, size_t n) {
        const size_t kMax = 0xff00;
        if (!straddle && n <= kMax && cur.size() + n > kMax) flush_block();
        while (n) { size_t k = std::min(n, kMax - cur.size()); cur.insert(cur.end(), p, p + k); p += k; n -= k; if (cur.size() == kMax) flush_block(); }
    }
    ~BgzfOut() { if (zinit) deflateEnd(&zs); }
};

struct IndexPiece {   // per thread, local virtual offsets
    std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
    std::vector<uint64_t> linear; uint64_t first = ~0ull, last = 0; uint64_t n_mapped = 0;
};

static uint64_t rebase(uint64_t v, uint64_t base) { return (((v >> 16) + base) << 16) | (v & 0xffff); }

int main(int argc, char** argv) {
    Config cfg;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return std::string(argv[++i]); };
        if (a == "--config") cfg.name = val(); else if (a == "--genome-fraction") cfg.fraction = atof(val().c_str());
        else if (a == "--coverage") cfg.coverage = atof(val().c_str()); else if (a == "--seed") cfg.seed = strtoull(val().c_str(), nullptr, 10);
        else if (a == "--threads") cfg.threads = atoi(val().c_str()); else if (a == "--level") cfg.level = atoi(val().c_str());
        else if (a == "--straddle") cfg.straddle = true; else if (a == "--cg-every") g_cg_every = atoi(val().c_str()); else if (a == "--dense") g_dense = true; else if (a == "--cg-short") g_cg_short = true; else if (a == "-o") cfg.out = val(); else if (a == "--bed") cfg.bed = val();
        else if (a == "--contig") cfg.only_contig = atoi(val().c_str());
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    const bool is_long = cfg.name == "c5", is_exome = cfg.name == "c4";
    if (cfg.name == "c4" && cfg.coverage == 30) cfg.coverage = 100;
    std::vector<Contig> contigs;
    if (cfg.name == "c2") contigs.push_back({"chr20", (uint32_t)std::max(4000.0, 64444167 * cfg.fraction)});
    else for (auto& c : kGRCh38) contigs.push_back({c.name, (uint32_t)std::max(4000.0, c.len * cfg.fraction)});
    if (cfg.only_contig >= 0) { for (size_t i = 0; i < contigs.size(); i++) if ((int)i != cfg.only_contig) contigs[i].len = std::max<uint32_t>(contigs[i].len, 4000); }

    // exome targets (C4): clustered, log-normal lengths, overlapping allowed, unsorted in the BED file
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> targets(contigs.size());
    if (is_exome) {
        FILE* fb = cfg.bed.empty() ? nullptr : fopen(cfg.bed.c_str(), "w");
        const double n_total = 1195764 * cfg.fraction; uint64_t G = 0; for (auto& c : contigs) G += c.len;
        std::vector<std::string> lines;
        for (size_t t = 0; t < contigs.size(); t++) {
            Rng g(mix64(cfg.seed ^ 0xE0E0ull ^ t));
            int64_t n = (int64_t)std::llround(n_total * (double)contigs[t].len / (double)G);
            int64_t genes = std::max<int64_t>(1, n / 20);
            for (int64_t k = 0; k < n; k++) {
                Rng gg(mix64(cfg.seed ^ 0xE1E1ull ^ (t << 40) ^ (uint64_t)(k % genes)));
                int64_t centre = (int64_t)gg.below(contigs[t].len);
                int64_t s = centre + (int64_t)std::llround(g.normal() * 15000);
                int64_t len = (int64_t)std::llround(std::exp(std::log(130.0) + 0.8 * g.normal())); len = std::max<int64_t>(10, std::min<int64_t>(12000, len));
                s = std::max<int64_t>(0, std::min<int64_t>(s, (int64_t)contigs[t].len - len - 1)); if (s < 0) continue;
                targets[t].push_back({(uint32_t)s, (uint32_t)(s + len)});
                char ln[256]; snprintf(ln, sizeof ln, "%s\t%lld\t%lld\texon_%zu_%lld\n", contigs[t].name.c_str(), (long long)s, (long long)(s + len), t, (long long)k); lines.push_back(ln);
            }
            std::sort(targets[t].begin(), targets[t].end());
        }
        if (fb) { Rng g(cfg.seed ^ 0xBEDull); for (size_t i = lines.size(); i > 1; i--) std::swap(lines[i - 1], lines[g.below((uint32_t)i)]); for (auto& l : lines) fputs(l.c_str(), fb); fclose(fb); }
    }

    FILE* fo = fopen(cfg.out.c_str(), "wb");
    if (!fo) { perror(cfg.out.c_str()); return 1; }
    uint64_t file_off = 0;
    {   // header
        BgzfOut h; h.level = cfg.level; h.straddle = true;
        std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
        for (auto& c : contigs) text += "@SQ\tSN:" + c.name + "\tLN:" + std::to_string(c.len) + "\n";
        text += "@RG\tID:grp1\tSM:synthetic\n@PG\tID:gen_bam\tCL:" + cfg.name + "\n";
        std::vector<uint8_t> b; b.insert(b.end(), {'B', 'A', 'M', 1}); put32(b, (uint32_t)text.size()); b.insert(b.end(), text.begin(), text.end()); put32(b, (uint32_t)contigs.size());
        for (auto& c : contigs) { put32(b, (uint32_t)c.name.size() + 1); b.insert(b.end(), c.name.begin(), c.name.end()); b.push_back(0); put32(b, c.len); }
        h.write(b.data(), b.size()); h.flush_block();
        fwrite(h.data.data(), 1, h.data.size(), fo); file_off = h.data.size();
    }
    std::vector<uint8_t> bai; bai.insert(bai.end(), {'B', 'A', 'I', 1}); put32(bai, (uint32_t)contigs.size());
    uint64_t total_reads = 0;
    for (size_t t = 0; t < contigs.size(); t++) {
        const uint32_t clen = contigs[t].len;
        const bool skip = cfg.only_contig >= 0 && (int)t != cfg.only_contig;
        const int64_t ntiles = ((int64_t)clen + kTile - 1) / kTile;
        const int nth = (int)std::max<int64_t>(1, std::min<int64_t>(cfg.threads, ntiles));
        std::vector<BgzfOut> outs(nth); std::vector<IndexPiece> idx(nth); std::vector<uint64_t> nreads(nth, 0);
        std::vector<std::thread> th;
        for (int w = 0; w < nth; w++) th.emplace_back([&, w]() {
            BgzfOut& o = outs[w]; o.level = cfg.level; o.straddle = cfg.straddle;
            if (skip) return;
            IndexPiece& ip = idx[w]; ip.linear.assign((clen >> 14) + 1, 0);
            const int64_t ta = ntiles * w / nth, tb = ntiles * (w + 1) / nth;
            std::vector<Record> recs;
            for (int64_t tile = ta; tile < tb; tile++) {
                recs.clear();
                const int64_t lo = tile * kTile, hi = std::min<int64_t>(lo + kTile, clen);
                if (is_long) { gen_long_tile(cfg, (int)t, clen, tile, lo, hi, recs); }
                if (!is_long || cfg.paired_frac_long > 0) {
                    const double cov = is_long ? cfg.coverage * cfg.paired_frac_long : cfg.coverage;
                    const auto* tg = is_exome ? &targets[t] : nullptr;
                    if (tile > 0) gen_short_tile(cfg, (int)t, clen, tile - 1, lo, hi, cov, recs, tg);   // right mates spilling in
                    gen_short_tile(cfg, (int)t, clen, tile, lo, hi, cov, recs, tg);
                }
                std::sort(recs.begin(), recs.end(), [](const Record& a, const Record& b) { return a.pos != b.pos ? a.pos < b.pos : a.order < b.order; });
                for (auto& r : recs) {
                    if (!o.straddle && r.bytes.size() <= 0xff00 && o.cur.size() + r.bytes.size() > 0xff00) o.flush_block();
                    const uint64_t v0 = o.tell();
                    o.write(r.bytes.data(), r.bytes.size());
                    const uint64_t v1 = o.tell();
                    auto& ch = ip.bins[(uint32_t)reg2bin(r.pos, r.end)];
                    if (!ch.empty() && ch.back().second == v0) ch.back().second = v1; else ch.push_back({v0, v1});
                    for (int64_t wdw = r.pos >> 14; wdw <= (r.end - 1) >> 14; wdw++) if (ip.linear[wdw] == 0 || v0 + 1 < ip.linear[wdw]) ip.linear[wdw] = v0 + 1;   // +1: 0 means empty
                    if (ip.first == ~0ull) ip.first = v0;
                    ip.last = v1; ip.n_mapped++; nreads[w]++;
                }
            }
            o.flush_block();
        });
        for (auto& x : th) x.join();
        // stitch: thread outputs are whole BGZF blocks, so local virtual offsets only need their coffset rebased
        std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins; std::vector<uint64_t> linear((clen >> 14) + 1, 0);
        uint64_t ref_beg = 0, ref_end = 0, n_mapped = 0; bool any = false;
        for (int w = 0; w < nth; w++) {
            const uint64_t base = file_off;
            // an offset that sits exactly at the end of this piece's last block belongs to the next block start
            auto fix = [&](uint64_t v) { return rebase(v, base); };
            for (auto& kv : idx[w].bins) for (auto& c : kv.second) bins[kv.first].push_back({fix(c.first), fix(c.second)});
            for (size_t i = 0; i < idx[w].linear.size(); i++) if (idx[w].linear[i]) { uint64_t v = fix(idx[w].linear[i] - 1); if (!linear[i] || v < linear[i]) linear[i] = v; }
            if (idx[w].n_mapped) { if (!any) { ref_beg = fix(idx[w].first); any = true; } ref_end = fix(idx[w].last); n_mapped += idx[w].n_mapped; }
            fwrite(outs[w].data.data(), 1, outs[w].data.size(), fo); file_off += outs[w].data.size();
            total_reads += nreads[w];
        }
        // BAI for this contig
        if (!any) { put32(bai, 0); put32(bai, 0); continue; }
        put32(bai, (uint32_t)bins.size() + 1);
        for (auto& kv : bins) { put32(bai, kv.first); put32(bai, (uint32_t)kv.second.size()); for (auto& c : kv.second) { for (int k = 0; k < 8; k++) bai.push_back((uint8_t)(c.first >> (8 * k))); for (int k = 0; k < 8; k++) bai.push_back((uint8_t)(c.second >> (8 * k))); } }
        put32(bai, 37450); put32(bai, 2);
        for (uint64_t v : {ref_beg, ref_end, n_mapped, (uint64_t)0}) for (int k = 0; k < 8; k++) bai.push_back((uint8_t)(v >> (8 * k)));
        size_t nl = linear.size(); while (nl && !linear[nl - 1]) nl--;
        for (size_t i = 1; i < nl; i++) if (!linear[i]) linear[i] = linear[i - 1];
        put32(bai, (uint32_t)nl);
        for (size_t i = 0; i < nl; i++) for (int k = 0; k < 8; k++) bai.push_back((uint8_t)(linear[i] >> (8 * k)));
    }
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, fo); fclose(fo);
    FILE* fi = fopen((cfg.out + ".bai").c_str(), "wb"); fwrite(bai.data(), 1, bai.size(), fi); fclose(fi);
    fprintf(stderr, "[gen_bam] %s: %llu reads, %llu bytes\n", cfg.out.c_str(), (unsigned long long)total_reads, (unsigned long long)(file_off + 28));
    printf("{\"reads\": %llu, \"bytes\": %llu}\n", (unsigned long long)total_reads, (unsigned long long)(file_off + 28));
    return 0;
}
