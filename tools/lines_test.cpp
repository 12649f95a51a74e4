// lines_test.cpp -- prints rows through mosdepth_b200/host/lines.hpp for tests/test_host_lines.py (CPU).
//   stdin: one case per line:  R chrom s e name|- value precision   |   T chrom s e name|- n c1..cn (n = -1: no counts)   |   W k window tlen
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <vector>
#include "../mosdepth_b200/host/lines.hpp"

int main() {
    std::string ln, out;
    while (std::getline(std::cin, ln)) {
        std::istringstream is(ln); std::string kind; is >> kind;
        if (kind == "R") { std::string c, nm, v; uint32_t s, e; int p; is >> c >> s >> e >> nm >> v >> p; std::string name = nm == "-" ? "" : nm; lines::region_row(out, c, s, e, nm == "." ? nullptr : &name, strtod(v.c_str(), nullptr), p); }
        else if (kind == "T") { std::string c, nm; uint32_t s, e; int n; is >> c >> s >> e >> nm >> n; std::vector<int64_t> cnt; for (int i = 0; i < n; i++) { long long x; is >> x; cnt.push_back(x); }
            std::string name = nm == "-" ? "" : nm; size_t nt = n < 0 ? 3 : (size_t)n; lines::threshold_row(out, c, s, e, nm == "." ? nullptr : &name, n < 0 ? nullptr : cnt.data(), nt); }
        else if (kind == "W") { long long k; uint32_t w, t; is >> k >> w >> t; int64_t s, e; lines::window_interval(k, w, t, &s, &e); out += std::to_string(s) + " " + std::to_string(e) + "\n"; }
    }
    fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}
