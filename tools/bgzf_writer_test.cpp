// bgzf_writer_test.cpp -- CPU harness for mosdepth_b200/host/bgzf_writer.hpp (tests/test_bgzf_writer.py): writes the same
// deterministic runs and intervals through the parallel / indexed BGZF paths and through the plain path.
//   usage: bgzf_writer_test <prefix> <n_runs_per_contig> <threads>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../mosdepth_b200/host/bgzf_writer.hpp"

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: bgzf_writer_test prefix n_runs threads\n"); return 2; }
    const std::string prefix = argv[1]; const int64_t n = atoll(argv[2]); const int threads = atoi(argv[3]);
    uint64_t x = 0x9e3779b97f4a7c15ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    const char* contigs[] = {"chr1", "chr2_random_with_a_longer_name", "chrM"};
    const std::vector<std::string> labels = {"NO_COVERAGE", "LOW_COVERAGE", "CALLABLE", "HIGH_COVERAGE"};
    struct Member { uint64_t offset, first_run; uint32_t csize, isize; };      // layout of msd_bgzf_member (include/mosdepth_b200.h)
    bgz::BgzWriter gz, plain, qz, qplain, rz, rplain, mz, lz, lplain;
    if (!lz.open(prefix + ".lines.bed.gz", 6, false, threads, true) || !lplain.open(prefix + ".lines.bed", 6, true)) return 1;
    if (!mz.open(prefix + ".members.bed.gz", 1, false, threads, true)) return 1;
    if (!gz.open(prefix + ".runs.bed.gz", 1, false, threads, true) || !plain.open(prefix + ".runs.bed", 1, true) ||
        !qz.open(prefix + ".quant.bed.gz", 1, false, threads, true) || !qplain.open(prefix + ".quant.bed", 1, true) ||
        !rz.open(prefix + ".regions.bed.gz", 6, false, threads, true) || !rplain.open(prefix + ".regions.bed", 6, true)) return 1;
    for (int c = 0; c < 3; c++) {
        const int64_t m = c == 2 ? std::min<int64_t>(n, 50) : n;
        std::vector<int32_t> st((size_t)m), en((size_t)m), dp((size_t)m), bn((size_t)m);
        int32_t pos = 0;
        for (int64_t k = 0; k < m; k++) {
            const uint64_t r = rnd();
            int32_t len = 1 + (int32_t)(r % 40); if ((r >> 20) % 97 == 0) len += (int32_t)((r >> 8) % 70000);     // a few runs span many 16 KB windows
            st[(size_t)k] = pos; en[(size_t)k] = pos + len; pos += len;
            dp[(size_t)k] = (int32_t)((r >> 32) % 60); bn[(size_t)k] = (int32_t)((r >> 40) % 4);
        }
        gz.write_runs(contigs[c], st.data(), en.data(), dp.data(), m); plain.write_runs(contigs[c], st.data(), en.data(), dp.data(), m);
        qz.write_runs(contigs[c], st.data(), en.data(), bn.data(), m, &labels); qplain.write_runs(contigs[c], st.data(), en.data(), bn.data(), m, &labels);
        {   // members deflated "elsewhere" (here: zlib over groups of 700 lines), appended through write_members
            std::vector<uint8_t> bytes; std::vector<Member> mem; std::string text; char num[64];
            for (int64_t a = 0; a < m; a += 700) {
                text.clear();
                for (int64_t k = a; k < std::min(m, a + 700); k++) { text += contigs[c]; text += '\t'; char* p = bgz::fmt_i32(num, st[(size_t)k]); *p++ = '\t'; p = bgz::fmt_i32(p, en[(size_t)k]); *p++ = '\t'; p = bgz::fmt_i32(p, dp[(size_t)k]); *p++ = '\n'; text.append(num, (size_t)(p - num)); }
                const size_t before = bytes.size();
                if (!bgz::deflate_block(text.data(), text.size(), 1, bytes)) return 1;
                mem.push_back(Member{before, (uint64_t)a, (uint32_t)(bytes.size() - before), (uint32_t)text.size()});
            }
            mz.write_members(contigs[c], bytes.data(), (int64_t)bytes.size(), mem.data(), (int64_t)mem.size(), st.data(), en.data(), dp.data(), m);
        }
        {   // the same kind of rows through write_lines (formatted on the worker threads)
            const int64_t nw = (pos + 499) / 500; const std::string chrom = contigs[c]; const int32_t top = pos;
            auto fmt = [&](int64_t k, std::string& out) { char b[128]; const int len = snprintf(b, sizeof b, "%s\t%lld\t%lld\t%.2f\n", chrom.c_str(), (long long)k * 500, (long long)std::min<int64_t>(top, (k + 1) * 500), (double)((k * 2654435761u) % 4000) / 100.0); out.append(b, (size_t)len); };
            auto iv = [&](int64_t k, int64_t* s, int64_t* e) { *s = k * 500; *e = std::min<int64_t>(top, (k + 1) * 500); };
            lz.write_lines(chrom, nw, fmt, iv); lplain.write_lines(chrom, nw, fmt, iv);
        }
        // regions: 500-base windows over the contig through the serial, line-indexed path
        for (int32_t s = 0; s < pos; s += 500) {
            const int32_t e = std::min(pos, s + 500);
            char buf[128]; const int len = snprintf(buf, sizeof buf, "%s\t%d\t%d\t%.2f\n", contigs[c], s, e, (double)(rnd() % 4000) / 100.0);
            rz.write_interval(std::string(buf, (size_t)len), contigs[c], s, e); rplain.write_interval(std::string(buf, (size_t)len), contigs[c], s, e);
        }
    }
    int rc = 0;
    rc |= gz.close(); rc |= plain.close(); rc |= qz.close(); rc |= qplain.close(); rc |= rz.close(); rc |= rplain.close(); rc |= mz.close(); rc |= lz.close(); rc |= lplain.close();
    return rc ? 1 : 0;
}
