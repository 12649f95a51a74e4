// ingest_test.cpp -- CPU check of mosdepth_b200/csrc/ingest.hpp (tests/test_abi_and_host.py): parallel_copy == memcpy for sizes around
// the piece boundaries and for every thread count.
#include <cstdio>
#include <cstdint>
#include "../mosdepth_b200/csrc/ingest.hpp"
int main() {
    const size_t sizes[] = {0, 1, 4095, (4u << 20) - 1, 4u << 20, (8u << 20) + 1, (33u << 20) + 4097, (64u << 20) - 3};
    std::vector<uint8_t> src((64u << 20) + 64), dst(src.size());
    uint64_t x = 88172645463325252ull;
    for (auto& b : src) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = (uint8_t)x; }
    int bad = 0, n = 0;
    for (size_t sz : sizes) for (int t : {1, 2, 3, 7, 16, 64}) {
        std::fill(dst.begin(), dst.end(), 0xEE);
        msd::parallel_copy(dst.data() + 3, src.data() + 5, sz, t);
        n++;
        if (memcmp(dst.data() + 3, src.data() + 5, sz) != 0 || dst[2] != 0xEE || dst[3 + sz] != 0xEE) { bad++; fprintf(stderr, "size %zu threads %d differs\n", sz, t); }
    }
    printf("{\"cases\": %d, \"bad\": %d}\n", n, bad);
    return bad ? 1 : 0;
}
