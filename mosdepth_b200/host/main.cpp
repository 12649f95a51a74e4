// mosdepth-b200: C++ stand-in for the reference's Nim host (mosdepth.nim:864-989 CLI, :589-840 main loop).
//
// The reference's host cannot be compiled in this image (no Nim, no hts-nim/htslib), so this program mirrors it
// one-to-one in C++ above the C ABI of libmosdepth_b200.so: same flags, same per-contig emission order, same text
// grammars.  What stays on the host (as in the reference): flag parsing, BAM header and BAI/CSI parsing, BED
// loading, label/precision environment variables, and writing the output files.  Everything numeric comes from
// the library (GPU); there is no CPU depth computation in this file.
//
// Output files: <prefix>.mosdepth.{global,region}.dist.txt, <prefix>.mosdepth.summary.txt and the bed.gz files
// (BGZF-compressed with system zlib, each with a .csi index like hts-nim's BGZI writer; per-base and quantized runs are
// formatted and deflated by a pool of host threads, bgzf_writer.hpp.  The reference's own writer stays untouched when the
// Nim host binds the C ABI).
#include <algorithm>
#include <cerrno>
#include <cinttypes>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <numeric>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>
#include <zlib.h>

#include "mosdepth_b200.h"

static void die(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3), noreturn));
#include <cstdarg>
static void die(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); exit(code); }

// ---- BGZF text writer + CSI index (replaces hts-nim wopen_bgzi / write_interval for the stand-in): bgzf_writer.hpp ------
#include "bgzf_writer.hpp"
#include "lines.hpp"
using bgz::BgzWriter;

// ---- BAM header + index parsing (what hts-nim does for the reference's host) -----------------------------------
struct Target { std::string name; uint32_t length; };
struct Header { std::vector<Target> targets; uint64_t first_voff = 0; };

static std::vector<uint8_t> read_file(const std::string& p, bool required) {
    FILE* f = fopen(p.c_str(), "rb"); if (!f) { if (required) die(1, "[mosdepth] cannot open %s", p.c_str()); return {}; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> v((size_t)n); if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) die(1, "[mosdepth] read error on %s", p.c_str()); fclose(f); return v;
}
static bool file_exists(const std::string& p) { FILE* f = fopen(p.c_str(), "rb"); if (!f) return false; fclose(f); return true; }

static Header read_header(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb"); if (!f) die(1, "[mosdepth] cannot open %s", path.c_str());
    std::vector<uint8_t> inflated; std::vector<std::pair<uint64_t, uint64_t>> starts;   // (uncompressed start, coffset)
    std::vector<uint32_t> isizes; uint64_t coff = 0; bool eof = false;
    auto need = [&](size_t n) {
        while (inflated.size() < n && !eof) {
            uint8_t h[18]; if (fread(h, 1, 18, f) != 18) { eof = true; break; }
            if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) die(1, "[mosdepth] %s is not BGZF", path.c_str());
            uint32_t xlen = h[10] | (h[11] << 8);
            std::vector<uint8_t> extra(xlen); memcpy(extra.data(), h + 12, std::min<uint32_t>(xlen, 6)); if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, f) != xlen - 6) die(1, "truncated");
            uint32_t bsize = 0; for (uint32_t q = 0; q + 4 <= xlen;) { uint32_t sl = extra[q + 2] | (extra[q + 3] << 8); if (extra[q] == 66 && extra[q + 1] == 67 && sl == 2) bsize = extra[q + 4] | (extra[q + 5] << 8); q += 4 + sl; }
            uint32_t blen = bsize + 1;
            if (xlen < 6 || blen < xlen + 20 + 2) die(1, "[mosdepth] %s: invalid BGZF block in the header", path.c_str());
            uint32_t clen = blen - xlen - 20;
            std::vector<uint8_t> c(clen + 8); if (fread(c.data(), 1, clen + 8, f) != clen + 8) die(1, "[mosdepth] truncated BAM header");
            uint32_t isz; memcpy(&isz, c.data() + clen + 4, 4);
            if (isz > 65536) die(1, "[mosdepth] %s: invalid BGZF block in the header", path.c_str());
            std::vector<uint8_t> o(isz ? isz : 1);
            z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15); zs.next_in = c.data(); zs.avail_in = clen; zs.next_out = o.data(); zs.avail_out = isz;
            int r = inflate(&zs, Z_FINISH); const bool whole = zs.total_out == isz; inflateEnd(&zs); if (r != Z_STREAM_END || !whole) die(1, "[mosdepth] corrupt BAM header block");
            starts.push_back({inflated.size(), coff}); isizes.push_back(isz);
            inflated.insert(inflated.end(), o.begin(), o.begin() + isz); coff += blen;
        }
        if (inflated.size() < n) die(1, "[mosdepth] truncated BAM header");
    };
    need(12);
    if (memcmp(inflated.data(), "BAM\1", 4)) die(1, "[mosdepth] %s is not a BAM file", path.c_str());
    int32_t l_text; memcpy(&l_text, inflated.data() + 4, 4);
    if (l_text < 0) die(1, "[mosdepth] %s: corrupt BAM header", path.c_str());
    need(12 + (size_t)l_text);
    size_t off = 8 + (size_t)l_text; int32_t n_ref; memcpy(&n_ref, inflated.data() + off, 4); off += 4;
    if (n_ref < 0) die(1, "[mosdepth] %s: corrupt BAM header", path.c_str());
    Header H;
    for (int i = 0; i < n_ref; i++) {
        need(off + 4); int32_t l_name; memcpy(&l_name, inflated.data() + off, 4);
        if (l_name < 1 || l_name > (1 << 20)) die(1, "[mosdepth] %s: corrupt BAM header", path.c_str());
        need(off + 8 + (size_t)l_name);
        Target t; t.name.assign((const char*)inflated.data() + off + 4, (size_t)l_name - 1); int32_t l; memcpy(&l, inflated.data() + off + 4 + l_name, 4); t.length = (uint32_t)l;
        H.targets.push_back(t); off += 8 + (size_t)l_name;
    }
    bool found = false;
    for (size_t b = 0; b < starts.size(); b++) if (off >= starts[b].first && off < starts[b].first + isizes[b]) { H.first_voff = (starts[b].second << 16) | (off - starts[b].first); found = true; }
    if (!found) H.first_voff = coff << 16;   // header ends exactly at a block boundary
    fclose(f);
    return H;
}

struct ContigIndex { bool has = false; uint64_t beg = 0, end = 0; std::vector<uint64_t> linear; };   // linear: BAI ioffset per 16,384-bp window (empty for CSI)
static std::vector<uint8_t> gunzip_all(const std::vector<uint8_t>& raw) {
    if (raw.size() < 2 || raw[0] != 0x1f || raw[1] != 0x8b) return raw;
    std::vector<uint8_t> out; size_t p = 0;
    while (p < raw.size()) {
        z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, 31); zs.next_in = (Bytef*)raw.data() + p; zs.avail_in = (uInt)(raw.size() - p);
        int r; do { size_t o = out.size(); out.resize(o + 65536); zs.next_out = out.data() + o; zs.avail_out = 65536; r = inflate(&zs, Z_NO_FLUSH); out.resize(o + 65536 - zs.avail_out); } while (r == Z_OK);
        if (r != Z_STREAM_END) die(1, "[mosdepth] corrupt index");
        p = raw.size() - zs.avail_in; inflateEnd(&zs);
    }
    return out;
}
static bool load_index(const std::string& bam, std::vector<ContigIndex>& out) {
    std::string p; bool csi = false;
    if (file_exists(bam + ".csi")) { p = bam + ".csi"; csi = true; } else if (file_exists(bam + ".bai")) p = bam + ".bai";
    else if (bam.size() > 4 && file_exists(bam.substr(0, bam.size() - 4) + ".bai")) p = bam.substr(0, bam.size() - 4) + ".bai"; else return false;
    const std::vector<uint8_t> d = gunzip_all(read_file(p, true)); size_t q = 0; uint32_t pseudo = 37450;
    // every read is bounds-checked: a truncated or corrupt index ends the run with a message, like hts_idx_load failing (:969-971)
    auto need = [&](size_t n) { if (n > d.size() - std::min(q, d.size())) die(1, "[mosdepth] %s: truncated or corrupt index", p.c_str()); };
    auto i32 = [&]() { need(4); int32_t v; memcpy(&v, d.data() + q, 4); q += 4; return v; };
    auto u64 = [&]() { need(8); uint64_t v; memcpy(&v, d.data() + q, 8); q += 8; return v; };
    auto count = [&](size_t item_bytes) { const int32_t v = i32(); if (v < 0 || (uint64_t)v * item_bytes > d.size() - q) die(1, "[mosdepth] %s: truncated or corrupt index", p.c_str()); return v; };
    need(4);
    if (memcmp(d.data(), csi ? "CSI\1" : "BAI\1", 4)) die(1, "[mosdepth] %s: not a %s index", p.c_str(), csi ? "CSI" : "BAI");
    q = 4;
    if (csi) { i32(); const int32_t depth = i32(), l_aux = i32(); if (depth < 0 || depth > 9 || l_aux < 0) die(1, "[mosdepth] %s: truncated or corrupt index", p.c_str()); need((size_t)l_aux); q += (size_t)l_aux; pseudo = (uint32_t)(((1u << ((depth + 1) * 3)) - 1) / 7 + 1); }
    const int32_t n_ref = count(4); out.assign((size_t)n_ref, ContigIndex());
    for (int r = 0; r < n_ref; r++) {
        const int32_t n_bin = count(csi ? 16 : 8); uint64_t mn = UINT64_MAX, mx = 0;
        for (int b = 0; b < n_bin; b++) {
            const uint32_t bin = (uint32_t)i32(); if (csi) u64();
            const int32_t nc = count(16);
            for (int c = 0; c < nc; c++) { const uint64_t cb = u64(), ce = u64(); if (bin != pseudo) { mn = std::min(mn, cb); mx = std::max(mx, ce); } }
        }
        if (!csi) { const int32_t ni = count(8); out[r].linear.resize((size_t)ni); for (int k = 0; k < ni; k++) out[r].linear[(size_t)k] = u64(); }
        if (mn != UINT64_MAX) { out[r].has = true; out[r].beg = mn; out[r].end = mx; }
    }
    return true;
}

// ---- regions (bed_to_table mosdepth.nim:362-380, bed_line_to_region :195-209) --------------------------------------
struct Region { uint32_t start, stop; std::string name; };
static int64_t nim_parse_int(const std::string& s) { if (s.empty()) die(1, "invalid integer: %s", s.c_str()); char* e; errno = 0; long long v = strtoll(s.c_str(), &e, 10); if (*e || errno) die(1, "invalid integer: %s", s.c_str()); return v; }
static std::string strip(const std::string& s) { size_t a = 0, b = s.size(); while (a < b && isspace((unsigned char)s[a])) a++; while (b > a && isspace((unsigned char)s[b - 1])) b--; return s.substr(a, b - a); }
static std::vector<std::pair<std::string, std::vector<Region>>> bed_to_table(const std::string& path) {
    std::vector<uint8_t> d = gunzip_all(read_file(path, true));
    std::vector<std::pair<std::string, std::vector<Region>>> tab;
    size_t p = 0;
    while (p < d.size()) {
        size_t e = p; while (e < d.size() && d[e] != '\n') e++;
        std::string line((const char*)d.data() + p, e - p); p = e + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) break;                                   // hts_getline(...) > 0 (:366)
        if (line[0] == 't' && line.rfind("track ", 0) == 0) continue;
        if (line[0] == '#') continue;
        std::vector<std::string> f; size_t a = 0;
        while (f.size() < 5) { size_t t = line.find('\t', a); if (t == std::string::npos) break; f.push_back(line.substr(a, t - a)); a = t + 1; }
        f.push_back(line.substr(a));
        if (f.size() < 3) { fprintf(stderr, "[mosdepth] skipping bad bed line:%s\n", strip(line).c_str()); continue; }
        int64_t s = nim_parse_int(f[1]), en = nim_parse_int(f[2]);
        if (s > en) die(1, "[slivar] ERROR: start > end in bed line:%s", line.c_str());
        Region r; r.start = (uint32_t)s; r.stop = (uint32_t)en; if (f.size() > 3) r.name = strip(f[3]);
        auto it = std::find_if(tab.begin(), tab.end(), [&](auto& kv) { return kv.first == f[0]; });
        if (it == tab.end()) { tab.push_back({f[0], {}}); it = tab.end() - 1; }
        it->second.push_back(r);
    }
    for (auto& kv : tab) std::stable_sort(kv.second.begin(), kv.second.end(), [](const Region& a, const Region& b) { return (int64_t)a.start - (int64_t)b.start < 0; });
    return tab;
}

// ---- text (write_distribution :439-458, write_summary :460-480, make_lookup :111-122) -----------------------------
static int g_precision = 2; static bool g_summary_header = true;
static void write_distribution(const std::string& chrom, const std::vector<int64_t>& d, FILE* fh) {
    int64_t sum = 0; for (int64_t v : d) sum += v;
    if (sum < 1) return;
    double cum = 0;
    for (size_t i = 0; i < d.size(); i++) {
        int64_t irev = (int64_t)d.size() - (int64_t)i - 1, v = d[(size_t)irev];
        if (irev > 300 && v == 0) continue;
        cum += (double)v / (double)sum;
        if (cum < 8e-5) continue;
        fprintf(fh, "%s\t%" PRId64 "\t%.*f\n", chrom.c_str(), irev, g_precision, cum);
    }
}
static void write_summary(const std::string& region, const msd_depth_stat& st, FILE* fh) {
    double mean = st.cum_length > 0 ? (double)st.cum_depth / (double)st.cum_length : 0.0;
    uint32_t mn = st.min_depth == UINT32_MAX ? 0 : st.min_depth;
    if (g_summary_header) { fputs("chrom\tlength\tbases\tmean\tmin\tmax\n", fh); g_summary_header = false; }
    fprintf(fh, "%s\t%" PRId64 "\t%" PRIu64 "\t%.*f\t%u\t%u\n", region.c_str(), st.cum_length, st.cum_depth, g_precision, mean, mn, st.max_depth);
}
static void stat_clear(msd_depth_stat& s) { s.cum_length = 0; s.cum_depth = 0; s.min_depth = UINT32_MAX; s.max_depth = 0; }
static msd_depth_stat stat_add(const msd_depth_stat& a, const msd_depth_stat& b) { msd_depth_stat r; r.cum_length = a.cum_length + b.cum_length; r.cum_depth = a.cum_depth + b.cum_depth; r.min_depth = std::min(a.min_depth, b.min_depth); r.max_depth = std::max(a.max_depth, b.max_depth); return r; }
static void sum_into(const std::vector<int64_t>& from, std::vector<int64_t>& to) { if (from.size() > to.size()) to.resize(from.size(), 0); for (size_t i = 0; i < from.size(); i++) to[i] += from[i]; }

static std::vector<int64_t> get_quantize_args(const char* qa) {   // :501-519
    std::vector<int64_t> q; if (!qa) return q;
    std::string a = qa;
    if (a.find(':') == std::string::npos) a = ":" + a + ":";
    if (a[0] == ':') a = "0" + a;
    if (a.back() == ':') a += std::to_string(INT64_MAX);
    size_t p = 0;
    for (;;) { size_t c = a.find(':', p); std::string tok = a.substr(p, c == std::string::npos ? std::string::npos : c - p); char* e; errno = 0; long long v = strtoll(tok.c_str(), &e, 10);
        if (tok.empty() || *e || errno) { fprintf(stderr, "[mosdepth] invalid quantize string: '%s'\n", a.c_str()); exit(2); } q.push_back(v); if (c == std::string::npos) break; p = c + 1; }
    std::sort(q.begin(), q.end());
    return q;
}
static std::vector<std::string> make_lookup(const std::vector<int64_t>& q) {
    std::vector<std::string> L;
    for (size_t i = 0; i + 1 < q.size(); i++) { std::string env = "MOSDEPTH_Q" + std::to_string(i); const char* t = getenv(env.c_str());
        if (!t || !*t) L.push_back(q[i + 1] == INT64_MAX ? std::to_string(q[i]) + ":inf" : std::to_string(q[i]) + ":" + std::to_string(q[i + 1])); else L.push_back(t); }
    return L;
}
static inline char* fmt_u32(char* p, uint32_t v) { char t[12]; int n = 0; do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v); while (n) *p++ = t[--n]; return p; }
static inline char* fmt_i32(char* p, int32_t v) { if (v < 0) { *p++ = '-'; return fmt_u32(p, (uint32_t)(-(int64_t)v)); } return fmt_u32(p, (uint32_t)v); }


// ---- multi-GPU sharding (nothing in the single-process reference corresponds to it) -----------------------------------
// A Piece is what one rank owns of one contig: the records with lo <= pos < hi, read from the virtual-offset span [vb, ve).
// cover = the position from which on the span is known to hold every record (linear-index window start), see msd_set_contig_cover.
struct Piece { int tid; uint32_t lo, hi; uint64_t vb, ve; uint32_t cover; bool cut; };

// N contiguous BGZF offset ranges of equal compressed size; the cuts fall inside contigs at BAI linear-index entries, at multiples
// of lcm(32, window) so that no --by window straddles a cut.  A rank's pieces are adjacent in the file.
static std::vector<std::vector<Piece>> offset_shards(const std::vector<ContigIndex>& index, const std::vector<uint32_t>& tlen, const std::vector<uint8_t>& want,
                                                     int n, uint32_t window, uint32_t margin_bp, uint32_t beyond_windows = 8) {
    std::vector<int> have; for (int t = 0; t < (int)index.size() && t < (int)tlen.size(); t++) if (want[(size_t)t] && index[(size_t)t].has) have.push_back(t);
    std::vector<std::vector<Piece>> out((size_t)n);
    if (have.empty()) return out;
    const uint64_t align = window ? std::lcm<uint64_t>(32, window) : 32;
    const uint64_t c_beg = index[(size_t)have.front()].beg >> 16, c_end = index[(size_t)have.back()].end >> 16;
    std::vector<std::pair<int, uint64_t>> cuts;      // (contig, position); position 0 = at the start of the contig
    for (int k = 1; k < n; k++) {
        const uint64_t target = c_beg + (c_end - c_beg) * (uint64_t)k / (uint64_t)n;
        int t = have.back(); for (int u : have) if ((index[(size_t)u].end >> 16) > target) { t = u; break; }
        const auto& lin = index[(size_t)t].linear;
        size_t w = lin.size(); for (size_t j = 0; j < lin.size(); j++) if (lin[j] && (lin[j] >> 16) >= target) { w = j; break; }
        uint64_t pos = ((uint64_t)w * 16384 + align / 2) / align * align;
        if (pos >= tlen[(size_t)t]) { int nx = -1; for (int u : have) if (u > t) { nx = u; break; } if (nx < 0) continue; t = nx; pos = 0; }
        if (!cuts.empty() && std::make_pair(t, pos) <= cuts.back()) continue;
        cuts.push_back({t, pos});
    }
    std::vector<std::pair<int, uint64_t>> bounds; bounds.push_back({have.front(), 0}); for (auto& c : cuts) bounds.push_back(c); bounds.push_back({-1, 0});
    for (size_t r = 0; r + 1 < bounds.size(); r++) {
        const int t0 = bounds[r].first, t1 = bounds[r + 1].first; const uint64_t p0 = bounds[r].second, p1 = bounds[r + 1].second;
        for (int t : have) {
            if (t < t0 || (t1 >= 0 && (t > t1 || (t == t1 && p1 == 0)))) continue;
            const ContigIndex& ci = index[(size_t)t]; const uint32_t L = tlen[(size_t)t];
            Piece pc; pc.tid = t; pc.lo = t == t0 ? (uint32_t)p0 : 0; pc.hi = (t1 >= 0 && t == t1) ? (uint32_t)p1 : L; pc.cut = pc.lo > 0 || pc.hi < L;
            // span start: the linear-index entry of a window at least margin_bp in front of lo (an empty entry: the nearest one before it)
            pc.cover = 0; pc.vb = ci.beg;
            if (pc.lo > margin_bp) {
                int64_t wi = std::min<int64_t>(((int64_t)pc.lo - (int64_t)margin_bp) / 16384, (int64_t)ci.linear.size() - 1);
                while (wi > 0 && !ci.linear[(size_t)wi]) wi--;
                if (wi > 0) { pc.vb = ci.linear[(size_t)wi]; pc.cover = (uint32_t)wi * 16384u; }
            }
            // span end: beyond_windows index windows after hi (an empty entry: the next one after it), so that a record at or beyond hi is seen
            pc.ve = ci.end;
            if (pc.hi < L) { size_t wi = (size_t)pc.hi / 16384 + beyond_windows; while (wi < ci.linear.size() && !ci.linear[wi]) wi++; if (wi < ci.linear.size()) pc.ve = ci.linear[wi]; }
            out[r].push_back(pc);
        }
    }
    return out;
}

// whole contigs, longest-processing-time bin packing on compressed span size (modes whose queries need the whole array on one GPU)
static std::vector<std::vector<Piece>> contig_shards(const std::vector<ContigIndex>& index, const std::vector<uint32_t>& tlen, const std::vector<uint8_t>& want, int n) {
    std::vector<int> order; for (int t = 0; t < (int)index.size() && t < (int)tlen.size(); t++) if (want[(size_t)t] && index[(size_t)t].has) order.push_back(t);
    auto weight = [&](int t) { return (index[(size_t)t].end >> 16) - (index[(size_t)t].beg >> 16) + 1; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weight(a) > weight(b); });
    std::vector<uint64_t> load((size_t)n, 0); std::vector<std::vector<Piece>> out((size_t)n);
    for (int t : order) { size_t r = 0; for (size_t k = 1; k < load.size(); k++) if (load[k] < load[r]) r = k; load[r] += weight(t); out[r].push_back(Piece{t, 0, tlen[(size_t)t], index[(size_t)t].beg, index[(size_t)t].end, 0, false}); }
    for (auto& v : out) std::sort(v.begin(), v.end(), [](const Piece& a, const Piece& b) { return a.tid < b.tid; });
    return out;
}

// neighbouring spans of one genuine BAM file merge when the bytes between them are few: they are whole records of other contigs,
// which the library's target mask drops
static void merge_spans(std::vector<uint64_t>& sb, std::vector<uint64_t>& se) {
    std::vector<size_t> ord(sb.size()); std::iota(ord.begin(), ord.end(), 0); std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return sb[a] < sb[b]; });
    std::vector<uint64_t> mb, me;
    for (size_t i : ord) { if (!mb.empty() && sb[i] >= me.back() && (sb[i] >> 16) - (me.back() >> 16) <= (1u << 20)) me.back() = std::max(me.back(), se[i]); else if (!mb.empty() && sb[i] < me.back()) me.back() = std::max(me.back(), se[i]); else { mb.push_back(sb[i]); me.push_back(se[i]); } }
    sb = mb; se = me;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#define MSD(call) do { int _rc = (int)(call); if (_rc < 0) die(1, "[mosdepth] %s", msd_last_error(ctx)); } while (0)

int main(int argc, char** argv) {
    const double t_main = now_s();
    const char *chrom_arg = nullptr, *by = nullptr, *quant_arg = nullptr, *thr_arg = nullptr, *rg_arg = nullptr;
    int print_shards = 0, no_per_base = 0, fast_mode = 0, fragment_mode = 0, use_median = 0, mapq = 0, device = 0, plain = 0, n_gpus = getenv("MOSDEPTH_B200_GPUS") ? atoi(getenv("MOSDEPTH_B200_GPUS")) : 1; int64_t min_len = -1, max_len = -1;
    uint16_t eflag = 1796, iflag = 0; std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) die(1, "error parsing arguments"); return argv[++i]; };
        auto is = [&](const char* s, const char* l) { return a == s || a == l; };
        if (is("-t", "--threads")) (void)nim_parse_int(val());            // BAM decompression threads: inflate runs on the GPU
        else if (is("-c", "--chrom")) chrom_arg = val();
        else if (is("-b", "--by")) by = val();
        else if (is("-n", "--no-per-base")) no_per_base = 1;
        else if (is("-f", "--fasta")) (void)val();
        else if (is("-F", "--flag")) eflag = (uint16_t)nim_parse_int(val());
        else if (is("-i", "--include-flag")) iflag = (uint16_t)nim_parse_int(val());
        else if (is("-x", "--fast-mode")) fast_mode = 1;
        else if (is("-a", "--fragment-mode")) fragment_mode = 1;
        else if (is("-q", "--quantize")) quant_arg = val();
        else if (is("-Q", "--mapq")) mapq = (int)nim_parse_int(val());
        else if (is("-l", "--min-frag-len")) min_len = nim_parse_int(val());
        else if (is("-u", "--max-frag-len")) max_len = nim_parse_int(val());
        else if (is("-T", "--thresholds")) thr_arg = val();
        else if (is("-m", "--use-median")) use_median = 1;
        else if (is("-R", "--read-groups")) rg_arg = val();
        else if (a == "--device") device = (int)nim_parse_int(val());
        else if (a == "--gpus") n_gpus = (int)nim_parse_int(val());         // shard the genome over GPUs device .. device + gpus - 1 (one context and one host thread each)
        else if (a == "--print-shards") print_shards = 1;                  // test hook: print the multi-GPU plan as JSON and exit (no GPU is touched)
        else if (a == "--plain") plain = 1;                                // write *.bed (uncompressed) instead of *.bed.gz
        else if (a.size() > 1 && a[0] == '-') die(1, "error parsing arguments");
        else pos.push_back(a);
    }
    if (pos.size() != 2) die(1, "error parsing arguments");
    const std::string prefix = pos[0], bam_path = pos[1];
    { const char* e = getenv("MOSDEPTH_PRECISION"); if (e && *e) { char* end; long v = strtol(e, &end, 10); g_precision = *end ? 2 : (int)v; } }
    if (max_len < 0) max_len = INT64_MAX;
    if (max_len < min_len) { fprintf(stderr, "[mosdepth] error --max-frag-len was lower than --min-frag-len.\n"); return 2; }
    std::vector<int32_t> thr;
    if (thr_arg) { std::string s = thr_arg; size_t p = 0; for (;;) { size_t c = s.find(',', p); thr.push_back((int32_t)nim_parse_int(s.substr(p, c == std::string::npos ? std::string::npos : c - p))); if (c == std::string::npos) break; p = c + 1; } std::sort(thr.begin(), thr.end()); }
    if (fragment_mode && fast_mode) { fprintf(stderr, "[mosdepth] error only one of --fast-mode and --fragment-mode can be specified\n"); return 2; }
    if (!by && !thr.empty()) { fprintf(stderr, "[mosdepth] error --thresholds can only be used when --by is specified.\n"); return 2; }
    if (bam_path.size() > 5 && bam_path.substr(bam_path.size() - 5) == ".cram") { fprintf(stderr, "[mosdepth] ERROR: specify a reference file (or set REF_PATH env var) for decoding CRAM\n"); return 1; }

    Header H = read_header(bam_path);
    std::vector<ContigIndex> index;
    if (!load_index(bam_path, index)) { fprintf(stderr, "[mosdepth] error alignment file must be indexed\n"); return 2; }
    const int nt = (int)H.targets.size();
    std::string chrom;
    if (chrom_arg && *chrom_arg && strcmp(chrom_arg, "nil")) {
        chrom = chrom_arg;
        if (chrom.find(':') != std::string::npos) { int seps = 0; size_t cut = std::string::npos; for (size_t k = chrom.size(); k-- > 0;) if (chrom[k] == ':' || chrom[k] == '-') { if (++seps == 2) { cut = k; break; } } if (cut == std::string::npos) cut = chrom.rfind(':'); chrom = chrom.substr(0, cut); }
        bool ok = false; for (auto& t : H.targets) if (t.name == chrom) ok = true;
        if (!ok) { fprintf(stderr, "[mosdepth] chromosome %s not found\n", chrom.c_str()); return 1; }
    }
    std::vector<int64_t> quant = get_quantize_args(quant_arg);
    std::vector<std::string> lookup = quant.empty() ? std::vector<std::string>() : make_lookup(quant);
    uint32_t window = 0; bool by_digit = false; std::vector<std::pair<std::string, std::vector<Region>>> bed; bool have_bed = false;
    if (by) { by_digit = *by != 0; for (const char* p = by; *p; p++) if (*p < '0' || *p > '9') by_digit = false; if (by_digit) window = (uint32_t)nim_parse_int(by); else { bed = bed_to_table(by); have_bed = true; } }

    // ---- which contigs does the main loop visit (get_targets :482-488, skip rule :703-705)? ------------------------
    std::vector<uint8_t> want(nt, 0);
    auto bed_find = [&](const std::string& name) -> std::vector<Region>* { for (auto& kv : bed) if (kv.first == name) return &kv.second; return nullptr; };
    for (int t = 0; t < nt; t++) {
        if (!chrom.empty() && H.targets[t].name != chrom) continue;
        if (no_per_base && thr.empty() && quant.empty() && have_bed && !bed_find(H.targets[t].name)) continue;
        want[t] = 1;
    }
    // ---- contexts: one per GPU; the genome is cut into one shard per GPU ---------------------------------------------------
    const double t_start = now_s();
    std::vector<uint32_t> tl(nt); for (int t = 0; t < nt; t++) tl[t] = H.targets[t].length;
    if (n_gpus < 1) n_gpus = 1;
    const int mode = fast_mode ? MSD_MODE_FAST : fragment_mode ? MSD_MODE_FRAGMENT : MSD_MODE_DEFAULT;
    // ranges inside contigs need: window / whole-contig outputs only (msd_regions and the run calls want the whole array on one GPU),
    // no fragment mode (a fragment starts before its read 1), and a linear index (BAI)
    bool any_linear = false; for (auto& ci : index) if (!ci.linear.empty()) any_linear = true;
    const bool cut_contigs = n_gpus > 1 && no_per_base && quant.empty() && thr.empty() && !have_bed && !fragment_mode && any_linear && !getenv("MOSDEPTH_B200_SHARD_CONTIGS");
    std::vector<std::vector<Piece>> shards;
    std::vector<msd_ctx*> ctxs((size_t)n_gpus, nullptr);
    std::vector<std::string> rank_err((size_t)n_gpus);
    std::vector<const char*> rgs; std::vector<std::string> rg_store;
    if (rg_arg) { std::string s = rg_arg; size_t p = 0; for (;;) { size_t c = s.find(',', p); rg_store.push_back(s.substr(p, c == std::string::npos ? std::string::npos : c - p)); if (c == std::string::npos) break; p = c + 1; } for (auto& r : rg_store) rgs.push_back(r.c_str()); }
    if (print_shards) {
        const auto plan = (n_gpus > 1 && cut_contigs) ? offset_shards(index, tl, want, n_gpus, window, 65536) : contig_shards(index, tl, want, n_gpus);
        printf("{\"sharding\": \"%s\", \"ranks\": [", (n_gpus > 1 && cut_contigs) ? "offsets" : "contigs");
        for (size_t r = 0; r < plan.size(); r++) { printf("%s[", r ? ", " : ""); for (size_t k = 0; k < plan[r].size(); k++) { const Piece& pc = plan[r][k]; printf("%s[%d, %u, %u, %llu, %llu, %u]", k ? ", " : "", pc.tid, pc.lo, pc.hi, (unsigned long long)pc.vb, (unsigned long long)pc.ve, pc.cover); } printf("]"); }
        printf("]}\n");
        return 0;
    }
    if (n_gpus > 1) {      // N ingest pumps share the host: each gets its share of the cores and a smaller pinned ring (cudaMallocHost is serialised by the driver: 0.45 s per GB)
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        if (!getenv("MSD_INGEST_THREADS")) setenv("MSD_INGEST_THREADS", std::to_string(std::max(2u, hw / (2u * (unsigned)n_gpus))).c_str(), 0);
        if (!getenv("MSD_BOUNCE_MB")) setenv("MSD_BOUNCE_MB", "16", 0);
        if (!getenv("MSD_BOUNCE_PIECES")) setenv("MSD_BOUNCE_PIECES", "4", 0);
    }
    uint8_t comm_id[MSD_COMM_ID_BYTES];
    if (n_gpus > 1 && msd_comm_unique_id(comm_id) != 0) die(1, "[mosdepth] %s", msd_last_error(nullptr));
    std::vector<std::thread> comm_threads((size_t)n_gpus); std::vector<int> comm_rc((size_t)n_gpus, 0);
    auto join_comm = [&]() { for (int r = 0; r < n_gpus; r++) if (comm_threads[(size_t)r].joinable()) { comm_threads[(size_t)r].join(); if (comm_rc[(size_t)r] < 0) die(1, "[mosdepth] msd_comm_init: %s", msd_last_error(ctxs[(size_t)r])); } };
    uint32_t margin_bp = 65536;
    double wall_create = 0, wall_run = 0;      // rank 0's msd_create (CUDA context) and msd_run (first-run allocations + the run), for MOSDEPTH_B200_STATS
    for (int attempt = 0;; attempt++) {
        if (n_gpus == 1) { shards.assign(1, {}); for (int t = 0; t < nt; t++) if (want[t] && t < (int)index.size() && index[t].has) shards[0].push_back(Piece{t, 0, tl[t], index[t].beg, index[t].end, 0, false}); }
        else shards = cut_contigs ? offset_shards(index, tl, want, n_gpus, window, margin_bp) : contig_shards(index, tl, want, n_gpus);
        bool any_cut = false; for (auto& v : shards) for (auto& pc : v) any_cut |= pc.cut;
        std::vector<int> rank_rc((size_t)n_gpus, 0);
        auto run_rank = [&](int r) {
            msd_ctx*& ctx = ctxs[(size_t)r];
            auto fail = [&](const char* what) { rank_rc[(size_t)r] = -1; rank_err[(size_t)r] = std::string(what) + ": " + msd_last_error(ctx); };
            if (!ctx) {
                const double tc0 = now_s();
                if (msd_create(device + r, &ctx) != 0) { rank_rc[(size_t)r] = -1; rank_err[(size_t)r] = msd_last_error(nullptr); ctx = nullptr; return; }
                if (r == 0) wall_create = now_s() - tc0;
                if (msd_open(ctx, bam_path.c_str(), nt, tl.data(), H.first_voff) < 0) return fail("msd_open");
                if (msd_set_filters(ctx, mapq, min_len, max_len, eflag, iflag, rgs.empty() ? nullptr : rgs.data(), (int)rgs.size(), mode) < 0) return fail("msd_set_filters");
                // joining the NCCL communicator takes a few hundred milliseconds and is only needed after msd_run (spill exchange, totals):
                // it runs on a thread of its own beside open / plan / run (it touches the context's communicator fields only)
                if (n_gpus > 1) comm_threads[(size_t)r] = std::thread([&, r] { comm_rc[(size_t)r] = msd_comm_init(ctxs[(size_t)r], r, n_gpus, comm_id); });
            }
            std::vector<uint8_t> mine((size_t)nt, 0); std::vector<uint64_t> sb, se;
            if (attempt > 0) for (int t = 0; t < nt; t++) msd_set_contig_range(ctx, t, 0, 0);   // a retry with a longer margin: drop the previous ranges
            for (auto& pc : shards[(size_t)r]) {
                mine[(size_t)pc.tid] = 1; sb.push_back(pc.vb); se.push_back(pc.ve);
                if (pc.cut) { if (msd_set_contig_range(ctx, pc.tid, pc.lo, pc.hi) < 0 || msd_set_contig_cover(ctx, pc.tid, pc.cover) < 0) return fail("msd_set_contig_range"); }
            }
            // with one GPU every visited contig is selected even when the index has nothing for it (coverage() then returns -2)
            if (n_gpus == 1) mine = want;
            if (msd_set_targets(ctx, mine.data()) < 0) return fail("msd_set_targets");
            merge_spans(sb, se);
            if (sb.empty()) { sb.push_back(1ull << 16); se.push_back(1ull << 16); }   // nothing indexed for this rank: read nothing
            if (msd_set_spans(ctx, (int)sb.size(), sb.data(), se.data()) < 0) return fail("msd_set_spans");
            if (by_digit && window && !use_median && msd_prefetch_windows(ctx, window) < 0) return fail("msd_prefetch_windows");   // --by <int>: every contig's window loop in one launch
            const double tr0 = now_s();
            const int rc = msd_run(ctx);                                              // coverage() + to_coverage() for this rank's pieces
            if (r == 0) wall_run = now_s() - tr0;
            if (rc < 0) { rank_rc[(size_t)r] = rc; rank_err[(size_t)r] = msd_last_error(ctx); }
        };
        if (n_gpus == 1) run_rank(0);
        else { std::vector<std::thread> th; for (int r = 0; r < n_gpus; r++) th.emplace_back(run_rank, r); for (auto& t : th) t.join(); }
        bool retry = false;
        for (int r = 0; r < n_gpus; r++) {
            if (rank_rc[(size_t)r] == MSD_ESPAN && margin_bp < (1u << 30)) retry = true;       // a mate lies in front of a shard's span: reach further back
            else if (rank_rc[(size_t)r] < 0) { fprintf(stderr, "[mosdepth] %s\n", rank_err[(size_t)r].c_str()); fflush(nullptr); _exit(1); }   // (other ranks may still be joining the communicator: no unwinding)
        }
        if (retry) { margin_bp = margin_bp >= (1u << 26) ? 0xfffffff0u : margin_bp * 16; if (attempt > 6) die(1, "[mosdepth] %s", rank_err[0].c_str()); continue; }
        join_comm();
        if (any_cut) {   // the depth that reads leave beyond a cut goes to the next rank (NCCL send / recv), then every shard gets its stats
            std::vector<std::thread> th;
            for (int r = 0; r < n_gpus; r++) th.emplace_back([&, r] { if (msd_exchange_spills(ctxs[(size_t)r]) < 0) { rank_rc[(size_t)r] = -1; rank_err[(size_t)r] = msd_last_error(ctxs[(size_t)r]); } });
            for (auto& t : th) t.join();
            for (int r = 0; r < n_gpus; r++) if (rank_rc[(size_t)r] < 0) die(1, "[mosdepth] %s", rank_err[(size_t)r].c_str());
        }
        break;
    }
    const double t_run = now_s();
    // which ranks hold which piece of every contig, in genome (= rank) order
    struct Owner { int rank; Piece pc; };
    std::vector<std::vector<Owner>> owners((size_t)nt);
    for (int r = 0; r < n_gpus; r++) for (auto& pc : shards[(size_t)r]) owners[(size_t)pc.tid].push_back(Owner{r, pc});
    msd_ctx* ctx = ctxs[0];
    // coverage()'s return value for a contig: the records of all its pieces
    auto contig_records = [&](int ti, int64_t* nrec) -> int {
        if (n_gpus == 1) return msd_contig_records(ctx, ti, nrec);
        int64_t total = 0;
        for (auto& o : owners[(size_t)ti]) { int64_t n = 0; const int rc = msd_contig_records(ctxs[(size_t)o.rank], ti, &n); if (rc < -2) { ctx = ctxs[(size_t)o.rank]; return rc; } total += n; }
        if (nrec) *nrec = total;
        return total > 0 ? ti : -2;
    };
    auto owner_ctx = [&](int ti) -> msd_ctx* { return (n_gpus == 1 || owners[(size_t)ti].empty()) ? ctxs[0] : ctxs[(size_t)owners[(size_t)ti][0].rank]; };
    auto is_cut = [&](int ti) { return n_gpus > 1 && (owners[(size_t)ti].size() > 1 || (owners[(size_t)ti].size() == 1 && owners[(size_t)ti][0].pc.cut)); };

    { uint64_t mx = 0; for (auto& t : H.targets) mx = std::max<uint64_t>(mx, t.length); bgz::csi_set_depth(bgz::csi_min_levels(mx)); }   // levels = get_min_levels(targets) :636
    // ---- outputs (mosdepth.nim:641-682) -------------------------------------------------------------------------------
    const std::string gz = plain ? "" : ".gz";
    BgzWriter fbase, fquant, fthr, fregion; FILE *fh_gd, *fh_sum, *fh_rd = nullptr;
    const bool device_bgzf = !getenv("MOSDEPTH_B200_DEVICE_BGZF") || atoi(getenv("MOSDEPTH_B200_DEVICE_BGZF")) != 0;   // per-base / quantized .bed.gz formatted + deflated on the device (K10): the default for .gz output since r2 (0: host threads)
    const int wthreads = getenv("MOSDEPTH_B200_WRITER_THREADS") ? atoi(getenv("MOSDEPTH_B200_WRITER_THREADS")) : 0;   // 0: one per core, at most 32
    if (!no_per_base && !fbase.open(prefix + ".per-base.bed" + gz, 1, plain, wthreads, !plain)) die(1, "[mosdepth] could not open per-base file");
    if (!quant.empty() && !fquant.open(prefix + ".quantized.bed" + gz, 1, plain, wthreads, !plain)) die(1, "[mosdepth] could not open quantized file");
    if (!thr.empty()) { if (!fthr.open(prefix + ".thresholds.bed" + gz, 1, plain, wthreads, !plain)) die(1, "[mosdepth] could not open thresholds file"); std::string h = "#chrom\tstart\tend\tregion"; for (int32_t t : thr) h += "\t" + std::to_string(t) + "X"; h += "\n"; fthr.write(h); }
    if (!(fh_gd = fopen((prefix + ".mosdepth.global.dist.txt").c_str(), "w"))) die(1, "[mosdepth] could not open file:%s.mosdepth.global.dist.txt", prefix.c_str());
    if (!(fh_sum = fopen((prefix + ".mosdepth.summary.txt").c_str(), "w"))) die(1, "[mosdepth] could not open file:%s.mosdepth.summary.txt", prefix.c_str());
    if (by) { if (!(fh_rd = fopen((prefix + ".mosdepth.region.dist.txt").c_str(), "w"))) die(1, "could not open region dist"); if (!fregion.open(prefix + ".regions.bed" + gz, 6, plain, wthreads, !plain)) die(1, "[mosdepth] could not open regions file"); }

    msd_depth_stat global_stat, global_region_stat; stat_clear(global_stat); stat_clear(global_region_stat);
    std::vector<int64_t> region_distribution(512, 0), global_distribution(512, 0);
    std::vector<double> values; std::vector<int64_t> thr_counts; std::vector<uint32_t> rs, re; std::string line; line.reserve(1 << 20);
    char num[64];

    for (int ti = 0; ti < nt; ti++) {
        const Target& target = H.targets[ti];
        if (!chrom.empty() && target.name != chrom) continue;
        std::vector<Region>* bc = have_bed ? bed_find(target.name) : nullptr;
        if (no_per_base && thr.empty() && quant.empty() && have_bed && !bc) continue;       // :703-705
        int64_t nrec = 0;
        ctx = owner_ctx(ti);                                                                   // the context the MSD() macro reports errors of
        int tid = contig_records(ti, &nrec);                                                   // coverage() return value
        if (tid < -2) die(1, "[mosdepth] %s", msd_last_error(ctx));
        std::vector<int64_t> chrom_global(512, 0), chrom_region(512, 0);
        msd_depth_stat chrom_stat, chrom_region_stat; stat_clear(chrom_stat); stat_clear(chrom_region_stat);
        if (by) {
            // region_gen :390-397
            std::vector<Region> wins; const std::vector<Region>* regs = bc;
            size_t n_regions = 0;
            if (!have_bed) n_regions = (size_t)std::max<int64_t>(0, msd_n_windows(target.length, window)); else n_regions = bc ? bc->size() : 0;
            values.assign(n_regions ? n_regions : 1, 0.0);
            if (!thr.empty()) thr_counts.assign((n_regions ? n_regions : 1) * thr.size(), 0);
            if (tid != -2 && n_regions) {
                if (!have_bed) {
                    int64_t d512[512];
                    if (!thr.empty() || false) { /* thresholds in window mode go through the region call below */ }
                    if (!is_cut(ti)) { MSD(msd_windows(ctx, ti, window, use_median, values.data(), (int64_t)values.size(), &chrom_region_stat, d512)); chrom_region.assign(d512, d512 + 512); }
                    else {   // the contig's windows, shard by shard in genome order (cuts are multiples of the window)
                        int64_t at = 0;
                        for (auto& o : owners[(size_t)ti]) {
                            ctx = ctxs[(size_t)o.rank]; msd_depth_stat st1;
                            const int64_t got = msd_windows(ctx, ti, window, use_median, values.data() + at, (int64_t)values.size() - at, &st1, d512);
                            if (got < 0) die(1, "[mosdepth] %s", msd_last_error(ctx));
                            if ((uint64_t)at * window != o.pc.lo) die(1, "[mosdepth] internal error: the shards of %s do not meet at a window boundary", target.name.c_str());
                            at += got; chrom_region_stat = stat_add(chrom_region_stat, st1);
                            for (int k = 0; k < 512; k++) chrom_region[(size_t)k] += d512[k];
                        }
                        if (at != (int64_t)n_regions) die(1, "[mosdepth] internal error: the shards of %s returned %lld of %zu windows", target.name.c_str(), (long long)at, n_regions);
                    }
                    if (!thr.empty()) {   // thresholds need per-window counts: same windows as explicit regions
                        rs.resize(n_regions); re.resize(n_regions);
                        for (size_t k = 0; k < n_regions; k++) { rs[k] = (uint32_t)(k * (uint64_t)window); re[k] = (uint32_t)std::min<uint64_t>((k + 1) * (uint64_t)window, target.length); }
                        std::vector<double> dummy(n_regions);
                        MSD(msd_regions(ctx, ti, (int64_t)n_regions, rs.data(), re.data(), use_median, dummy.data(), nullptr, nullptr, 0, nullptr, thr.data(), (int)thr.size(), thr_counts.data()));
                    }
                } else {
                    rs.resize(n_regions); re.resize(n_regions);
                    for (size_t k = 0; k < n_regions; k++) { rs[k] = (*regs)[k].start; re[k] = (*regs)[k].stop; }
                    int64_t dl = 0; chrom_region.assign(400001, 0);
                    MSD(msd_regions(ctx, ti, (int64_t)n_regions, rs.data(), re.data(), use_median, values.data(), &chrom_region_stat, chrom_region.data(), (int64_t)chrom_region.size(), &dl,
                                    thr.empty() ? nullptr : thr.data(), (int)thr.size(), thr.empty() ? nullptr : thr_counts.data()));
                    chrom_region.resize((size_t)std::max<int64_t>(dl, 512));
                }
            }
            // region lines (:725-735) + thresholds (:743 -> :522-557); formatted (and, for .gz, deflated) by the writer's threads
            auto interval = [&](int64_t k, int64_t* s, int64_t* e) {
                if (!have_bed) lines::window_interval(k, window, target.length, s, e);
                else { *s = (*regs)[(size_t)k].start; *e = (*regs)[(size_t)k].stop; }
            };
            fregion.write_lines(target.name, (int64_t)n_regions, [&](int64_t k, std::string& out) {
                int64_t s, e; interval(k, &s, &e);
                lines::region_row(out, target.name, (uint32_t)s, (uint32_t)e, have_bed ? &(*regs)[(size_t)k].name : nullptr, tid != -2 ? values[(size_t)k] : 0.0, g_precision);
            }, interval);
            if (!thr.empty()) fthr.write_lines(target.name, (int64_t)n_regions, [&](int64_t k, std::string& out) {
                int64_t s, e; interval(k, &s, &e);
                lines::threshold_row(out, target.name, (uint32_t)s, (uint32_t)e, have_bed ? &(*regs)[(size_t)k].name : nullptr, tid == -2 ? nullptr : &thr_counts[(size_t)k * thr.size()], thr.size());
            }, interval);
            if (bc) bc->clear(), bed.erase(std::find_if(bed.begin(), bed.end(), [&](auto& kv) { return kv.first == target.name; }));   // bed_regions.del (:397)
        }
        if (tid != -2) {                                                                       // :745-753
            int64_t dl = 0; std::vector<int64_t> d(400001);
            if (!is_cut(ti)) { MSD(msd_contig_stats(ctx, ti, &chrom_stat, d.data(), (int64_t)d.size(), &dl)); d.resize((size_t)std::max<int64_t>(dl, 512)); chrom_global = d; }
            else for (auto& o : owners[(size_t)ti]) {   // a sharded contig: lengths / sums add up, extremes combine, bins add up (depth_stat `+`, sum_into)
                ctx = ctxs[(size_t)o.rank]; msd_depth_stat st1;
                MSD(msd_contig_stats(ctx, ti, &st1, d.data(), (int64_t)d.size(), &dl));
                chrom_stat = stat_add(chrom_stat, st1);
                sum_into(std::vector<int64_t>(d.begin(), d.begin() + dl), chrom_global);
            }
            global_stat = stat_add(global_stat, chrom_stat);
            write_summary(target.name, chrom_stat, fh_sum);
            if (by) write_summary(target.name + "_region", chrom_region_stat, fh_sum);
            global_region_stat = stat_add(global_region_stat, chrom_region_stat);
        }
        write_distribution(target.name, chrom_global, fh_gd);                                 // :756-758
        if (by) write_distribution(target.name, chrom_region, fh_rd);
        if (tid >= 0) { if (by) sum_into(chrom_region, region_distribution); sum_into(chrom_global, global_distribution); }   // :761-764
        if (!no_per_base) {                                                                    // :766-793
            if (tid == -2) { line = target.name + "\t0\t" + std::to_string(target.length) + "\t0\n"; fbase.write_interval(line, target.name, 0, target.length); }
            else {
                const int32_t *st, *en, *dp; int64_t nr = 0;
                std::string pre = target.name + "\t"; line.clear();
                if (!plain && device_bgzf && target.name.size() <= 200) {                     // K10: formatted + deflated on the device, appended as is
                    const uint8_t* zb; int64_t zn = 0, nm = 0; const msd_bgzf_member* mem;
                    MSD(msd_per_base_bgzf(ctx, ti, target.name.c_str(), &zb, &zn, &mem, &nm, &st, &en, &dp, &nr));
                    if (nm > 0) { fbase.write_members(target.name, zb, zn, mem, nm, st, en, dp, nr); nr = 0; }      // (a short contig comes back as runs only)
                } else MSD(msd_per_base_runs(ctx, ti, &st, &en, &dp, &nr));
                if (!plain && nr) { fbase.write_runs(target.name, st, en, dp, nr); nr = 0; }      // .gz: formatted + deflated by the writer's threads
                for (int64_t k = 0; k < nr; k++) {
                    line += pre; char* p = fmt_i32(num, st[k]); *p++ = '\t'; p = fmt_i32(p, en[k]); *p++ = '\t'; p = fmt_i32(p, dp[k]); *p++ = '\n'; line.append(num, p - num);
                    if (line.size() > (1 << 19)) { fbase.write(line); line.clear(); }
                }
                fbase.write(line);
            }
        }
        if (!quant.empty()) {                                                                  // :794-805
            if (tid == -2 && quant[0] == 0) { line = target.name + "\t0\t" + std::to_string(target.length) + "\t" + lookup[0] + "\n"; fquant.write_interval(line, target.name, 0, target.length); }
            else if (tid != -2) {
                const int32_t *st, *en, *bn; int64_t nr = 0;
                bool labels_fit = true; for (auto& l : lookup) if (l.size() > 48) labels_fit = false;
                if (!plain && device_bgzf && target.name.size() <= 200 && labels_fit && lookup.size() <= 64) {     // K10, label mode
                    std::vector<const char*> lab; for (auto& l : lookup) lab.push_back(l.c_str());
                    const uint8_t* zb; int64_t zn = 0, nm = 0; const msd_bgzf_member* mem;
                    MSD(msd_quantized_bgzf(ctx, ti, quant.data(), (int)quant.size(), target.name.c_str(), lab.data(), (int)lab.size(), &zb, &zn, &mem, &nm, &st, &en, &bn, &nr));
                    if (nm > 0) { fquant.write_members(target.name, zb, zn, mem, nm, st, en, bn, nr, &lookup); nr = 0; }
                } else MSD(msd_quantized_runs(ctx, ti, quant.data(), (int)quant.size(), &st, &en, &bn, &nr));
                line.clear();
                if (!plain) { fquant.write_runs(target.name, st, en, bn, nr, &lookup); nr = 0; }
                for (int64_t k = 0; k < nr; k++) { line += target.name; line += '\t'; line += std::to_string(st[k]); line += '\t'; line += std::to_string(en[k]); line += '\t'; line += lookup[(size_t)bn[k]]; line += '\n'; if (line.size() > (1 << 19)) { fquant.write(line); line.clear(); } }
                fquant.write(line);
            }
        }
    }
    const double t_emit = now_s();
    if (n_gpus > 1) {
        // the `total` rows across GPUs: every context has accumulated what it was asked for (sum_into, depth_stat `+`); all-reduce them
        // over NCCL and check them against the sums this host kept in header order like the reference (:761-764)
        std::vector<int> arc((size_t)n_gpus, 0); std::vector<std::thread> th;
        for (int r = 0; r < n_gpus; r++) th.emplace_back([&, r] { arc[(size_t)r] = msd_totals_allreduce(ctxs[(size_t)r]); });
        for (auto& t : th) t.join();
        for (int r = 0; r < n_gpus; r++) if (arc[(size_t)r] < 0) die(1, "[mosdepth] %s", msd_last_error(ctxs[(size_t)r]));
        msd_depth_stat tg, tr; std::vector<int64_t> gd(400001), rd(400001); int64_t gl = 0, rl = 0;
        ctx = ctxs[0];
        MSD(msd_totals(ctx, &tg, &tr, gd.data(), (int64_t)gd.size(), &gl, rd.data(), (int64_t)rd.size(), &rl));
        auto same_stat = [](const msd_depth_stat& a, const msd_depth_stat& b) { return a.cum_length == b.cum_length && a.cum_depth == b.cum_depth && a.min_depth == b.min_depth && a.max_depth == b.max_depth; };
        auto same_dist = [](const std::vector<int64_t>& host, const std::vector<int64_t>& dev, int64_t n) { for (size_t i = 0; i < std::max(host.size(), (size_t)n); i++) { const int64_t a = i < host.size() ? host[i] : 0, b = (int64_t)i < n ? dev[i] : 0; if (a != b) return false; } return true; };
        if (!same_stat(tg, global_stat) || !same_dist(global_distribution, gd, gl) || (by && (!same_stat(tr, global_region_stat) || !same_dist(region_distribution, rd, rl))))
            die(1, "[mosdepth] internal error: the all-reduced totals of the %d GPUs differ from the per-contig sums", n_gpus);
        global_stat = tg; if (by) global_region_stat = tr;
    }
    write_summary("total", global_stat, fh_sum);                                               // :807-813
    if (by) write_summary("total_region", global_region_stat, fh_sum);
    write_distribution("total", global_distribution, fh_gd);
    if (by) { write_distribution("total", region_distribution, fh_rd); fclose(fh_rd); }
    if (have_bed && chrom.empty()) for (auto& kv : bed) fprintf(stderr, "[mosdepth] warning chromosome:%s from bed with %zu regions not found\n", kv.first.c_str(), kv.second.size());
    if (fregion.f && fregion.close() != 0) { fprintf(stderr, "[mosdepth] error writing region file\n\n"); return 1; }
    if (fquant.f && fquant.close() != 0) { fprintf(stderr, "[mosdepth] error writing quantize file\n\n"); return 1; }
    if (fthr.f && fthr.close() != 0) { fprintf(stderr, "[mosdepth] error writing thresholds file\n\n"); return 1; }
    if (fbase.f && fbase.close() != 0) { fprintf(stderr, "[mosdepth] error writing per-base file\n\n"); return 1; }
    fclose(fh_gd); fclose(fh_sum);
    const double t_close = now_s();
    if (getenv("MOSDEPTH_B200_STATS")) {
        unsigned long long recs = 0, blocks = 0; double dev_ms = 0;
        for (int r = 0; r < n_gpus; r++) { msd_run_info ri; msd_last_run_info(ctxs[(size_t)r], &ri); recs += ri.n_records; blocks += ri.n_blocks; dev_ms = std::max(dev_ms, ri.ms_total); }
        fprintf(stderr, "[mosdepth-b200] {\"gpus\": %d, \"sharding\": \"%s\", \"records\": %llu, \"blocks\": %llu, \"device_ms_max\": %.3f, \"wall_s\": {\"parse_header_index\": %.3f, \"create_open_run\": %.3f, \"msd_create_rank0\": %.3f, \"msd_run_rank0\": %.3f, \"emit\": %.3f, \"totals_close\": %.3f}}\n",
                n_gpus, n_gpus == 1 ? "none" : cut_contigs ? "offsets" : "contigs", recs, blocks, dev_ms, t_start - t_main, t_run - t_start, wall_create, wall_run, t_emit - t_run, t_close - t_emit);
    }
    // every output file is closed.  Tearing the contexts down by hand (cudaFree of ~20 GB, unpinning the ingest ring, destroying the CUDA
    // context) costs 0.3-0.7 s of wall clock for nothing: the process ends here and the driver reclaims everything.
    if (getenv("MOSDEPTH_B200_CLEAN_EXIT")) { for (msd_ctx* c : ctxs) msd_destroy(c); return 0; }
    fflush(nullptr);
    _exit(0);
}
