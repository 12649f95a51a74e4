// ingest_test.cpp -- CPU check of mosdepth_b200/csrc/ingest.hpp (tests/test_abi_and_host.py):
//   parallel_copy / CopyWorkers == memcpy for sizes around the piece boundaries and for every thread count;
//   chase_blocks (all threads at once, slices stitched) == the serial BSIZE walk, on a real BGZF file (argv[1]) from several start
//   offsets and for several thread counts, and on the same file with "false" block signatures planted inside a block's payload.
#include <cstdio>
#include <cstdint>
#include "../mosdepth_b200/csrc/ingest.hpp"
int main(int argc, char** argv) {
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
    for (int t : {1, 2, 5, 8}) {
        msd::CopyWorkers w(t);
        for (size_t sz : sizes) {
            std::fill(dst.begin(), dst.end(), 0xEE);
            w.copy(dst.data() + 3, src.data() + 5, sz);
            n++;
            if (memcmp(dst.data() + 3, src.data() + 5, sz) != 0 || dst[2] != 0xEE || dst[3 + sz] != 0xEE) { bad++; fprintf(stderr, "CopyWorkers size %zu threads %d differs\n", sz, t); }
        }
    }
    int chase_cases = 0;
    if (argc > 1) {
        FILE* f = fopen(argv[1], "rb"); if (!f) { perror(argv[1]); return 2; }
        fseek(f, 0, SEEK_END); const long sz = ftell(f); fseek(f, 0, SEEK_SET);
        std::vector<uint8_t> file((size_t)sz); if (fread(file.data(), 1, (size_t)sz, f) != (size_t)sz) return 2; fclose(f);
        for (int variant = 0; variant < 2; variant++) {
            std::vector<msd::BlockInfo> ref; uint64_t badat = 0;
            if (msd::chase_serial(file.data(), file.size(), 0, file.size(), ref, &badat) == ~0ull) { fprintf(stderr, "serial chase failed at %llu\n", (unsigned long long)badat); return 2; }
            if (variant == 1) {
                // plant the 16 fixed header bytes (with a plausible BSIZE) inside the payload of every 50th block: the parallel walk must not be
                // fooled -- a slice that starts on a planted signature is rejected when its predecessor's walk does not land on it.  (The CRC of
                // those blocks is now wrong; chase_blocks only reads headers and ISIZE.)
                static const uint8_t sig[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x40, 0x00};
                for (size_t b = 0; b < ref.size(); b += 50) if (ref[b].blen > 200) memcpy(file.data() + ref[b].coff + 60, sig, 18);
            }
            for (int t : {1, 2, 3, 8, 16, 64}) for (size_t start_blk : {(size_t)0, ref.size() / 3, ref.size() - 2}) for (int cut_end : {0, 1}) {
                const uint64_t beg = ref[start_blk].coff, stop = cut_end ? ref[ref.size() * 2 / 3].coff + 1 : (uint64_t)file.size();
                std::vector<msd::BlockInfo> got; uint64_t b2 = 0;
                // small slices so that even a test-sized file is cut: the library's 8 MB minimum would leave everything to one thread
                const bool ok = msd::chase_blocks(file.data(), file.size(), beg, stop, t, got, &b2);
                size_t want_n = 0; for (size_t i = start_blk; i < ref.size() && ref[i].coff < stop; i++) want_n++;
                chase_cases++;
                bool same = ok && got.size() == want_n;
                for (size_t i = 0; same && i < want_n; i++) same = got[i].coff == ref[start_blk + i].coff && got[i].blen == ref[start_blk + i].blen && got[i].isize == ref[start_blk + i].isize && got[i].xlen == ref[start_blk + i].xlen;
                if (!same) { bad++; fprintf(stderr, "chase_blocks variant %d threads %d start %zu cut %d: %zu blocks, expected %zu\n", variant, t, start_blk, cut_end, got.size(), want_n); }
            }
        }
    }
    printf("{\"cases\": %d, \"chase_cases\": %d, \"bad\": %d}\n", n, chase_cases, bad);
    return bad ? 1 : 0;
}
