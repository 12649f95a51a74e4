// ingest.hpp -- host side of the compressed-byte ingest: when the BAM is a mapped file (msd_open) or any
// pageable buffer, every chunk crosses a pinned bounce buffer before its H2D copy.  One thread copying out of the page cache
// moves a few GB/s while the device consumes ~50 GB/s of compressed BAM, so the bounce copy is spread over several threads
// (MSD_INGEST_THREADS, default min(16, cores / 2)); pieces are page aligned and at least 4 MB, small copies stay on the caller.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace msd {

inline int ingest_threads_default() {
    const char* e = getenv("MSD_INGEST_THREADS");
    if (e && *e) { const int v = atoi(e); return v < 1 ? 1 : v > 64 ? 64 : v; }
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min<unsigned>(16u, std::max<unsigned>(1u, hw / 2));
}

inline void parallel_copy(void* dst, const void* src, size_t n, int threads) {
    constexpr size_t kMinPiece = 4u << 20;
    const size_t t = std::min<size_t>((size_t)std::max(threads, 1), n / kMinPiece);
    if (t <= 1) { memcpy(dst, src, n); return; }
    const size_t piece = ((n + t - 1) / t + 4095) & ~(size_t)4095;
    std::vector<std::thread> pool;
    for (size_t i = 1; i < t; i++) {
        const size_t a = i * piece;
        if (a >= n) break;
        const size_t len = std::min(piece, n - a);
        pool.emplace_back([=]() { memcpy(static_cast<char*>(dst) + a, static_cast<const char*>(src) + a, len); });
    }
    memcpy(dst, src, std::min(piece, n));
    for (auto& th : pool) th.join();
}

}  // namespace msd
