// lines.hpp -- the rows of regions.bed and thresholds.bed exactly as the reference prints them (host CLI stand-in).
//   regions   mosdepth.nim:725-735   chrom \t start \t end \t [name \t] mean-or-median with MOSDEPTH_PRECISION digits (%.*f)
//   thresholds mosdepth.nim:522-557  chrom \t start \t end \t name|"unknown" { \t count }      (all counts 0 for a contig without reads)
//   windows   mosdepth.nim:382-388   [k*w, min((k+1)*w, tlen))
// Kept apart from main.cpp so that tools/lines_test.cpp can check them on the CPU (tests/test_host_lines.py).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>

namespace lines {

inline char* put_u32(char* p, uint32_t v) { char t[12]; int n = 0; do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v); while (n) *p++ = t[--n]; return p; }

inline void window_interval(int64_t k, uint32_t window, uint32_t tlen, int64_t* s, int64_t* e) {
    *s = (int64_t)((uint64_t)k * (uint64_t)window);
    *e = (int64_t)std::min<uint64_t>(((uint64_t)k + 1) * (uint64_t)window, tlen);
}

inline void head(std::string& out, const std::string& chrom, uint32_t s, uint32_t e) {
    char nb[16];
    out += chrom; out += '\t';
    char* p = put_u32(nb, s); out.append(nb, (size_t)(p - nb)); out += '\t';
    p = put_u32(nb, e); out.append(nb, (size_t)(p - nb)); out += '\t';
}

inline void region_row(std::string& out, const std::string& chrom, uint32_t s, uint32_t e, const std::string* name, double value, int precision) {
    head(out, chrom, s, e);
    if (name && !name->empty()) { out += *name; out += '\t'; }
    char nb[512];
    const int n = snprintf(nb, sizeof nb, "%.*f", precision, value);
    out.append(nb, (size_t)std::min<int>(std::max(n, 0), (int)sizeof nb - 1));
    out += '\n';
}

// counts == nullptr: the contig has no reads, every count prints as 0 (mosdepth.nim:538-542)
inline void threshold_row(std::string& out, const std::string& chrom, uint32_t s, uint32_t e, const std::string* name, const int64_t* counts, size_t n_thr) {
    head(out, chrom, s, e);
    if (name && !name->empty()) out += *name; else out += "unknown";
    char nb[24];
    for (size_t j = 0; j < n_thr; j++) {
        out += '\t';
        if (!counts) out += '0';
        else { const int n = snprintf(nb, sizeof nb, "%lld", (long long)counts[j]); out.append(nb, (size_t)n); }
    }
    out += '\n';
}

}  // namespace lines
