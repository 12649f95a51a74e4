// bgzf_writer.hpp -- BGZF text writer of the host CLI stand-in (replaces hts-nim's wopen_bgzi / write_interval,
// mosdepth.nim:650-678,733,793,803, for the stand-in only: the reference's writer stays when the Nim host binds the C ABI).
//
// Paths:
//   write() / write_interval()  serial: text is cut into blocks and deflated one block at a time (headers, the single rows of contigs
//                 without reads);
//   write_lines() parallel: the lines of one contig (region / threshold rows, per-base and quantized runs) are formatted by a callback
//                 AND deflated on worker threads, 8,192 lines per task, each task a whole number of BGZF blocks; the main thread writes
//                 the tasks in order, so the bytes do not depend on the number of threads.  per-base.bed.gz is the reference's own
//                 second-largest cost (README.md:157-159: 24.8 s with per-base output vs 5.9 s without) and regions.bed.gz (level 6,
//                 6.2 M window rows for a genome at --by 500) is the largest host cost of the north-star run; both are pure host work;
//   write_runs()  = write_lines over the (start, stop, value) runs that msd_per_base_runs / msd_quantized_runs return;
//   write_members() members deflated on the device (msd_per_base_bgzf) are appended as they are, the index is built from the runs.
// With `plain` the text is written as is (parity tests compare decompressed content).
// A CSI index (.csi, tabix preset bed: min_shift 14, depth 5, as hts-nim's BGZI writes next to every bed.gz) is built when
// index_path is given: every line's (contig, start, stop) is registered with the virtual offsets of its first and
// last byte.  [ext] CSIv1 layout: SAM/BAM spec "CSIv1.pdf"; tabix aux header: tabix.pdf.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

namespace bgz {

inline char* fmt_u32(char* p, uint32_t v) { char t[12]; int n = 0; do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v); while (n) *p++ = t[--n]; return p; }
inline char* fmt_i32(char* p, int32_t v) { if (v < 0) { *p++ = '-'; return fmt_u32(p, (uint32_t)(-(int64_t)v)); } return fmt_u32(p, (uint32_t)v); }

constexpr size_t kBlockIn = 0xff00;          // uncompressed bytes per BGZF block (htslib's BGZF_BLOCK_SIZE)
constexpr int64_t kRunsPerTask = 8192;

// one BGZF member for in[0..n), appended to out; returns false if deflate fails
inline bool deflate_block(const char* in, size_t n, int level, std::vector<uint8_t>& out) {
    const size_t at = out.size();
    out.resize(at + 18 + compressBound((uLong)n) + 64 + 8);
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    zs.next_in = (Bytef*)in; zs.avail_in = (uInt)n; zs.next_out = out.data() + at + 18; zs.avail_out = (uInt)(out.size() - at - 26);
    const int r = deflate(&zs, Z_FINISH);
    const uint32_t clen = (uint32_t)zs.total_out; deflateEnd(&zs);
    if (r != Z_STREAM_END || clen + 26 > 65536) return false;
    const uint32_t blen = clen + 26;
    static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    uint8_t* o = out.data() + at;
    memcpy(o, hdr, 16); o[16] = (uint8_t)(blen - 1); o[17] = (uint8_t)((blen - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)in, (uInt)n), isz = (uint32_t)n;
    memcpy(o + 18 + clen, &crc, 4); memcpy(o + 22 + clen, &isz, 4);
    out.resize(at + blen);
    return true;
}

// ---- CSI index of a BGZF-compressed, position-sorted, tab-separated text (tabix "bed" preset) -----------------------
struct CsiBuilder {
    static constexpr int kMinShift = 14;
    // levels of the binning scheme: the reference derives them from the longest target (get_min_levels, mosdepth.nim:565-577, passed to
    // wopen_bgzi :636-678); 5 levels reach 2^29 bp.  Set once per process (csi_set_depth) before any writer is opened.
    static int& depth_ref() { static int d = 5; return d; }
    struct Chunk { uint64_t beg, end; };
    struct Ref {
        std::map<uint32_t, std::vector<Chunk>> bins;
        std::vector<uint64_t> lin;                                // per 16 KB window: smallest offset of a line that overlaps it
        uint32_t last_bin = 0xffffffffu; std::vector<Chunk>* last_vec = nullptr;
    };
    std::vector<std::string> names; std::vector<Ref> refs; std::map<std::string, int> tid_of;
    std::string last_chrom; int last_tid = -1;
    uint64_t n_lines = 0;
    static uint32_t reg2bin(int64_t beg, int64_t end) {         // CSIv1 spec, section 3
        int l = depth_ref(), s = kMinShift; uint32_t t = ((1u << (depth_ref() * 3)) - 1) / 7;
        for (--end; l > 0; --l, s += 3, t -= 1u << (l * 3)) if (beg >> s == end >> s) return t + (uint32_t)(beg >> s);
        return 0;
    }
    void add(const std::string& chrom, int64_t beg, int64_t end, uint64_t voff_beg, uint64_t voff_end) {
        if (beg < 0) beg = 0;
        if (end <= beg) end = beg + 1;
        add_seg(chrom, reg2bin(beg, end), (size_t)(beg >> kMinShift), (size_t)((end - 1) >> kMinShift), voff_beg, voff_end);
    }
    // a run of consecutive lines of one bin that covers the 16 KB windows w0..w1 and the virtual offsets [voff_beg, voff_end)
    void add_seg(const std::string& chrom, uint32_t bin, size_t w0, size_t w1, uint64_t voff_beg, uint64_t voff_end) {
        if (last_tid < 0 || chrom != last_chrom) {
            auto it = tid_of.find(chrom);
            if (it == tid_of.end()) { it = tid_of.emplace(chrom, (int)names.size()).first; names.push_back(chrom); refs.emplace_back(); }
            last_tid = it->second; last_chrom = chrom;
        }
        Ref& r = refs[(size_t)last_tid];
        if (bin != r.last_bin) { r.last_vec = &r.bins[bin]; r.last_bin = bin; }
        std::vector<Chunk>& v = *r.last_vec;
        if (!v.empty() && v.back().end == voff_beg) v.back().end = voff_end;                   // next line of the same bin: one chunk
        else v.push_back(Chunk{voff_beg, voff_end});
        if (r.lin.size() <= w1) r.lin.resize(w1 + 1, UINT64_MAX);
        for (size_t w = w0; w <= w1; w++) if (voff_beg < r.lin[w]) r.lin[w] = voff_beg;
        n_lines++;
    }
    static void put32(std::vector<uint8_t>& o, uint32_t v) { for (int i = 0; i < 4; i++) o.push_back((uint8_t)(v >> (8 * i))); }
    static void put64(std::vector<uint8_t>& o, uint64_t v) { for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i))); }
    std::vector<uint8_t> serialise() const {
        std::vector<uint8_t> o;
        o.insert(o.end(), {'C', 'S', 'I', 1});
        put32(o, (uint32_t)kMinShift); put32(o, (uint32_t)depth_ref());
        std::vector<uint8_t> aux;                                 // tabix header: format=0x10000 (UCSC/BED: 0-based, half-open), col_seq 1, col_beg 2, col_end 3, meta '#', skip 0
        put32(aux, 0x10000u); put32(aux, 1); put32(aux, 2); put32(aux, 3); put32(aux, (uint32_t)'#'); put32(aux, 0);
        std::string nm; for (const auto& n : names) { nm += n; nm.push_back('\0'); }
        put32(aux, (uint32_t)nm.size()); aux.insert(aux.end(), nm.begin(), nm.end());
        put32(o, (uint32_t)aux.size()); o.insert(o.end(), aux.begin(), aux.end());
        put32(o, (uint32_t)refs.size());
        for (const Ref& r : refs) {
            std::vector<uint64_t> lin = r.lin;                    // windows without a line take the offset of the next line (htslib does the same)
            for (size_t w = lin.size(); w-- > 0;) if (lin[w] == UINT64_MAX) lin[w] = w + 1 < lin.size() ? lin[w + 1] : 0;
            put32(o, (uint32_t)r.bins.size());
            for (const auto& kv : r.bins) {
                // loffset of a bin = linear-index entry of its first window: a lower bound for every line that overlaps the bin
                int l = 0; uint32_t first = 0;
                for (int k = depth_ref(); k >= 0; k--) { const uint32_t f = ((1u << (3 * k)) - 1) / 7; if (kv.first >= f) { l = k; first = f; break; } }
                const size_t w = (size_t)(kv.first - first) << (3 * (depth_ref() - l));
                put32(o, kv.first); put64(o, w < lin.size() ? lin[w] : 0); put32(o, (uint32_t)kv.second.size());
                for (const Chunk& c : kv.second) { put64(o, c.beg); put64(o, c.end); }
            }
        }
        put64(o, 0);                                              // n_no_coor
        return o;
    }
};

// get_min_levels (mosdepth.nim:565-577): how many levels the longest target needs above the 16 KB leaves
inline int csi_min_levels(uint64_t max_len) { int r = 0; uint64_t s = 1ull << 14; while (max_len > s) { r++; s <<= 3; } return r; }
inline void csi_set_depth(int d) { CsiBuilder::depth_ref() = d < 0 ? 0 : d > 9 ? 9 : d; }

// Worker-side aggregation of the index entries of sorted lines: consecutive lines that lie in the same 16 KB window (so in the same
// leaf bin) and follow each other without a gap become ONE segment, which is what CsiBuilder::add would have merged them into anyway
// (same chunks, same linear-index minima); a line that spans windows stays a segment of its own.  Offsets are (block index, offset in
// block) pairs relative to the worker's piece; the thread that writes the file turns them into virtual offsets.
struct CsiSeg { uint32_t bin; uint32_t w0, w1; uint32_t b0, u0, b1, u1; };
struct CsiSegAgg {
    std::vector<CsiSeg> segs; bool have = false; CsiSeg cur{};
    void line(int64_t beg, int64_t end, uint32_t b0, uint32_t u0, uint32_t b1, uint32_t u1) {
        if (beg < 0) beg = 0;
        if (end <= beg) end = beg + 1;
        const uint32_t w0 = (uint32_t)(beg >> CsiBuilder::kMinShift), w1 = (uint32_t)((end - 1) >> CsiBuilder::kMinShift);
        if (have && w0 == w1 && cur.w0 == w0 && cur.w1 == w0 && cur.b1 == b0 && cur.u1 == u0) { cur.b1 = b1; cur.u1 = u1; return; }
        if (have) segs.push_back(cur);
        cur = CsiSeg{CsiBuilder::reg2bin(beg, end), w0, w1, b0, u0, b1, u1}; have = true;
    }
    void finish() { if (have) segs.push_back(cur); have = false; }
};

struct BgzWriter {
    FILE* f = nullptr; std::string buf; int level = 1; bool plain = false; int threads = 1; bool failed = false;
    uint64_t coff = 0;                       // compressed bytes written so far (virtual offsets of the index)
    CsiBuilder* csi = nullptr; std::string csi_path;
    bool open(const std::string& path, int lvl, bool plain_text, int n_threads = 0, bool with_index = false) {
        f = fopen(path.c_str(), "wb"); level = lvl; plain = plain_text; buf.reserve(1 << 17);
        threads = n_threads > 0 ? n_threads : (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
        if (with_index && !plain) { csi = new CsiBuilder(); csi_path = path + ".csi"; }
        return f != nullptr;
    }
    uint64_t voff() const { return (coff << 16) | (uint64_t)buf.size(); }     // of the next byte written through write()
    void put(const std::vector<uint8_t>& o) { if (fwrite(o.data(), 1, o.size(), f) != o.size()) failed = true; coff += o.size(); }
    void flush_block(size_t n) {
        if (!n) return;
        std::vector<uint8_t> o;
        if (!deflate_block(buf.data(), n, level, o)) { failed = true; return; }
        put(o);
        buf.erase(0, n);
    }
    void write(const char* s, size_t n) {
        if (plain) { if (fwrite(s, 1, n, f) != n) failed = true; return; }
        buf.append(s, n);
        while (buf.size() >= kBlockIn) flush_block(kBlockIn);
    }
    void write(const std::string& s) { write(s.data(), s.size()); }
    // one tab-separated line `chrom \t beg \t end ...` that should be findable through the index
    void write_interval(const std::string& line, const std::string& chrom, int64_t beg, int64_t end) {
        if (csi && buf.size() + line.size() > kBlockIn && !buf.empty()) flush_block(buf.size());   // keep a line inside one block (like bgzf_flush_try)
        const uint64_t v0 = voff();
        write(line);
        if (csi) csi->add(chrom, beg, end, v0, voff());
    }
    // chrom \t start \t stop \t (value | labels[value]) \n for every run, formatted and deflated by `threads` workers
    void write_runs(const std::string& chrom, const int32_t* st, const int32_t* en, const int32_t* val, int64_t n, const std::vector<std::string>* labels = nullptr) {
        write_lines(chrom, n,
            [&](int64_t k, std::string& out) {
                char num[48];
                out += chrom; out += '\t';
                char* p = fmt_i32(num, st[k]); *p++ = '\t'; p = fmt_i32(p, en[k]); *p++ = '\t';
                if (!labels) { p = fmt_i32(p, val[k]); *p++ = '\n'; out.append(num, (size_t)(p - num)); }
                else { out.append(num, (size_t)(p - num)); out += (*labels)[(size_t)val[k]]; out += '\n'; }
            },
            [&](int64_t k, int64_t* beg, int64_t* end) { *beg = st[k]; *end = en[k]; });
    }
    // n lines of one contig: fmt(k, out) appends line k (newline included) and is called from the worker threads (it must only read
    // shared state); interval(k, &beg, &end) is what the index records for it.  kRunsPerTask lines per task, each task a whole number
    // of BGZF blocks, tasks written in order: the bytes do not depend on the number of threads.
    template <class Fmt, class Iv>
    void write_lines(const std::string& chrom, int64_t n, Fmt fmt, Iv interval) {
        if (n <= 0) return;
        auto format = [&](int64_t a, int64_t b, std::string& out, std::vector<uint32_t>* line_off) {
            for (int64_t k = a; k < b; k++) {
                if (line_off) line_off->push_back((uint32_t)out.size());
                fmt(k, out);
            }
        };
        if (plain) {
            std::string line; line.reserve(1 << 20);
            for (int64_t a = 0; a < n; a += kRunsPerTask) { line.clear(); format(a, std::min(n, a + kRunsPerTask), line, nullptr); write(line); }
            return;
        }
        flush_block(buf.size());                                   // tasks start on a block boundary
        const int64_t n_tasks = (n + kRunsPerTask - 1) / kRunsPerTask;
        struct Task { std::vector<uint8_t> out; std::vector<uint32_t> line_off; std::vector<uint32_t> blk_in, blk_out; std::vector<CsiSeg> segs; size_t text = 0; bool ok = true; };
        const int64_t wave = (int64_t)threads * 4;
        for (int64_t w0 = 0; w0 < n_tasks; w0 += wave) {
            const int64_t w1 = std::min(n_tasks, w0 + wave);
            std::vector<Task> tasks((size_t)(w1 - w0));
            std::atomic<int64_t> next(w0);
            auto work = [&]() {
                std::string text; text.reserve(1 << 19);
                for (;;) {
                    const int64_t t = next.fetch_add(1);
                    if (t >= w1) break;
                    Task& T = tasks[(size_t)(t - w0)];
                    text.clear();
                    format(t * kRunsPerTask, std::min(n, (t + 1) * kRunsPerTask), text, csi ? &T.line_off : nullptr);
                    T.text = text.size();
                    // blocks end on line boundaries when an index is built (a line never straddles two blocks), else every 0xff00 bytes
                    size_t pos = 0, li = 0;
                    while (pos < text.size()) {
                        size_t len = std::min(kBlockIn, text.size() - pos);
                        if (csi && pos + len < text.size()) {
                            while (li + 1 < T.line_off.size() && T.line_off[li + 1] <= pos + len) li++;
                            if (T.line_off[li] > pos) len = T.line_off[li] - pos;
                        }
                        const size_t before = T.out.size();
                        if (!deflate_block(text.data() + pos, len, level, T.out)) { T.ok = false; break; }
                        T.blk_in.push_back((uint32_t)len); T.blk_out.push_back((uint32_t)(T.out.size() - before));
                        pos += len;
                    }
                    if (csi && T.ok) {      // index entries of the task's lines, as (block, offset) pairs: walk its blocks
                        CsiSegAgg agg; size_t b = 0, blk_start_text = 0;
                        const int64_t a = t * kRunsPerTask;
                        for (size_t i = 0; i < T.line_off.size(); i++) {
                            while (T.line_off[i] >= blk_start_text + T.blk_in[b]) { blk_start_text += T.blk_in[b]; b++; }
                            const size_t end_text = i + 1 < T.line_off.size() ? T.line_off[i + 1] : T.text;
                            // the line ends inside the same block (blocks end on line boundaries); the end offset of the last line of a block points at the next block
                            const bool last_in_block = end_text == blk_start_text + T.blk_in[b];
                            int64_t ib = 0, ie = 0; interval(a + (int64_t)i, &ib, &ie);
                            agg.line(ib, ie, (uint32_t)b, (uint32_t)(T.line_off[i] - blk_start_text), last_in_block ? (uint32_t)b + 1 : (uint32_t)b, last_in_block ? 0u : (uint32_t)(end_text - blk_start_text));
                        }
                        agg.finish(); T.segs.swap(agg.segs);
                    }
                }
            };
            const int nt = (int)std::min<int64_t>(threads, w1 - w0);
            std::vector<std::thread> pool;
            for (int i = 1; i < nt; i++) pool.emplace_back(work);
            work();
            for (auto& th : pool) th.join();
            for (int64_t t = w0; t < w1; t++) {
                Task& T = tasks[(size_t)(t - w0)];
                if (!T.ok) { failed = true; return; }
                if (csi) {
                    std::vector<uint64_t> bc(T.blk_out.size() + 1, coff);
                    for (size_t j = 0; j < T.blk_out.size(); j++) bc[j + 1] = bc[j] + T.blk_out[j];
                    for (const CsiSeg& g : T.segs) csi->add_seg(chrom, g.bin, g.w0, g.w1, (bc[g.b0] << 16) | g.u0, (bc[g.b1] << 16) | g.u1);
                }
                put(T.out);
            }
        }
    }
    // BGZF members that were deflated elsewhere (the device: msd_per_base_bgzf) holding the lines `chrom \t st \t en \t val \n` of the
    // n runs; Member = {offset, first_run, csize, isize}.  The bytes are appended as they are; the index entries come from the runs
    // (a line's length is a digit count) and the member table.  A line never straddles two members.
    template <class Member>
    void write_members(const std::string& chrom, const uint8_t* bytes, int64_t n_bytes, const Member* mem, int64_t n_mem, const int32_t* st, const int32_t* en, const int32_t* val, int64_t n,
                       const std::vector<std::string>* labels = nullptr) {
        if (n_bytes <= 0 || plain) { if (plain) failed = true; return; }
        flush_block(buf.size());
        if (csi) {
            auto digits = [](int64_t v) { int d = v < 0 ? 2 : 1; if (v < 0) v = -v; while (v >= 10) { v /= 10; d++; } return d; };
            const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n_mem / 64));
            std::vector<std::vector<CsiSeg>> segs((size_t)nt); std::vector<char> bad((size_t)nt, 0);
            auto work = [&](int w) {      // members [k0, k1): their lines' index entries as (member, offset) pairs
                const int64_t k0 = n_mem * w / nt, k1 = n_mem * (w + 1) / nt;
                CsiSegAgg agg;
                for (int64_t k = k0; k < k1; k++) {
                    const int64_t a = (int64_t)mem[k].first_run, b = k + 1 < n_mem ? (int64_t)mem[k + 1].first_run : n;
                    uint64_t uoff = 0;
                    for (int64_t i = a; i < b; i++) {
                        const uint64_t len = chrom.size() + 4 + (uint64_t)digits(st[i]) + (uint64_t)digits(en[i]) + (labels ? (uint64_t)(*labels)[(size_t)val[i]].size() : (uint64_t)digits(val[i]));
                        const bool last = i + 1 == b;
                        agg.line(st[i], en[i], (uint32_t)k, (uint32_t)uoff, last ? (uint32_t)k + 1 : (uint32_t)k, last ? 0u : (uint32_t)(uoff + len));
                        uoff += len;
                    }
                    if (uoff != mem[k].isize) { bad[(size_t)w] = 1; return; }
                }
                agg.finish(); segs[(size_t)w].swap(agg.segs);
            };
            std::vector<std::thread> pool;
            for (int w = 1; w < nt; w++) pool.emplace_back(work, w);
            work(0);
            for (auto& th : pool) th.join();
            for (char b : bad) if (b) { failed = true; return; }
            auto mc = [&](uint32_t k) { return coff + ((int64_t)k < n_mem ? mem[k].offset : (uint64_t)n_bytes); };
            for (const auto& vs : segs) for (const CsiSeg& g : vs) csi->add_seg(chrom, g.bin, g.w0, g.w1, (mc(g.b0) << 16) | g.u0, (mc(g.b1) << 16) | g.u1);
        }
        if (fwrite(bytes, 1, (size_t)n_bytes, f) != (size_t)n_bytes) failed = true;
        coff += (uint64_t)n_bytes;
    }
    int close() {
        if (!f) return 0;
        if (!plain) {
            flush_block(buf.size());
            static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (fwrite(eof, 1, 28, f) != 28) failed = true;
        }
        int r = fclose(f); f = nullptr;
        if (csi) {      // the .csi is itself BGZF-compressed
            const std::vector<uint8_t> raw = csi->serialise();
            BgzWriter ix; if (!ix.open(csi_path, 6, false, 1, false)) r = -1;
            else { ix.write((const char*)raw.data(), raw.size()); if (ix.close() != 0) r = -1; }
            delete csi; csi = nullptr;
        }
        return failed ? -1 : r;
    }
};

}  // namespace bgz
