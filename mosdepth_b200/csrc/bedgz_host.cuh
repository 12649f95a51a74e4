// bedgz_host.cuh -- the host sequence of K10 (msd_per_base_bgzf): buffers, launches, the three read-backs.  Kept in a header of
// its own, written against the CUDA runtime API plus one launch macro, so that the very same code runs (a) in libmosdepth_b200.so
// under nvcc and (b) in tests/emu/bedgz_emu.cpp under g++ with tests/emu/cuda_runtime.h standing in for the runtime (kernels on
// host threads, copies as memcpy) -- under AddressSanitizer if one wishes.  The CRC step is passed in: K1c (bgzf_crc_kernel) in
// the library, zlib in the emulation (K1c is inline PTX and has GPU tests of its own).
#pragma once
#include <chrono>
#include <cstdarg>
#include <vector>
#include "../../include/mosdepth_b200.h"
#include "runs.cuh"
#include "bedgz.cuh"

#if defined(__CUDACC__)
#define BZ_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#else
#define BZ_LAUNCH(kernel, grid, block, stream, ...) emu::launch((grid), (block), [&]() { kernel(__VA_ARGS__); })
#endif

namespace msd {

struct BzBuffers {      // grow-only, owned by the context
    void *d_line = nullptr, *d_blk = nullptr, *d_out = nullptr; uint64_t line_cap = 0, blk_cap = 0, out_cap = 0;
    uint8_t* h_out = nullptr; uint64_t h_cap = 0;      // pinned copy of the members
    std::vector<msd_bgzf_member> members;
    uint64_t out_bytes = 0; uint32_t launches = 0;
    char err[384] = "";
};
inline void bz_release(BzBuffers& B) {
    if (B.d_line) cudaFree(B.d_line); if (B.d_blk) cudaFree(B.d_blk); if (B.d_out) cudaFree(B.d_out); if (B.h_out) cudaFreeHost(B.h_out);
    B.d_line = B.d_blk = B.d_out = nullptr; B.h_out = nullptr; B.line_cap = B.blk_cap = B.out_cap = B.h_cap = 0;
}
inline int bz_fail(BzBuffers& B, int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(B.err, sizeof B.err, fmt, ap); va_end(ap); return code; }
inline int bz_grow(BzBuffers& B, void** p, uint64_t* cap, uint64_t bytes) {
    if (bytes <= *cap) return 0;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    bytes += bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return bz_fail(B, MSD_ECUDA, "cudaMalloc(%llu bytes) for msd_per_base_bgzf: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    *cap = bytes;
    return 0;
}
// carves 256-byte aligned pieces out of one allocation; with a null base it only measures
struct BzCarve { uint8_t* base; uint64_t used = 0; template <class T> T* take(uint64_t n) { T* r = base ? reinterpret_cast<T*>(base + used) : nullptr; used += (n * sizeof(T) + 255) & ~255ull; return r; } };
inline unsigned bz_grid_for(uint64_t n, int threads, int max_ctas) { uint64_t g = (n + threads - 1) / threads; if (g < 1) g = 1; if (g > (uint64_t)max_ctas) g = (uint64_t)max_ctas; return (unsigned)g; }

// exclusive scan uint32 -> uint64 (+ total) in three launches
inline void bz_scan(cudaStream_t s, const uint32_t* in, uint64_t n, uint32_t* tile_sum, unsigned long long* tile_base, unsigned long long* total, unsigned long long* out) {
    const unsigned nt = (unsigned)((n + kBzScanTile - 1) / kBzScanTile);
    BZ_LAUNCH(bz_tile_sum_kernel, nt, kBzThreads, s, in, n, tile_sum);
    BZ_LAUNCH(runs_scan_kernel, 1, 1024, s, tile_sum, nt, tile_base, total);
    BZ_LAUNCH(bz_tile_scan_kernel, nt, kBzThreads, s, in, n, tile_base, out);
}

#define BZ_CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return bz_fail(B, MSD_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); } while (0)

// The n runs (start, stop, depth) of a contig sit in device memory (d_*) and in host memory (h_*, for the code sample).  On success
// B.h_out holds B.out_bytes bytes of BGZF members and B.members describes them.  max_ctas caps the grid of the per-line kernels.
// labels (n_labels > 0): the third array holds label indices and the 4th field of a line is labels[value] (quantized.bed).
// crc(stream, d_text, d_desc, nb, d_ticket, d_expected, d_status, d_errors, d_crc_out) must leave the CRC-32 of every member's text in d_crc_out.
template <class CrcFn>
int bz_run(BzBuffers& B, cudaStream_t sc, const char* chrom, const int32_t* d_st, const int32_t* d_en, const int32_t* d_dp, const int32_t* h_st, const int32_t* h_en,
           const int32_t* h_dp, uint64_t n, int max_ctas, CrcFn crc, const char* const* labels = nullptr, int n_labels = 0) {
    typedef unsigned long long ull;
    B.members.clear(); B.out_bytes = 0; B.launches = 0; B.err[0] = 0;
    if (n == 0) return 0;
    int rc;
    const bool trace = getenv("MSD_TRACE") != nullptr;
    const auto w0 = std::chrono::steady_clock::now(); double w_last = 0;
    auto lap = [&](const char* what) {      // MSD_TRACE: host wall clock between the synchronisation points of the sequence
        if (!trace) return;
        const double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
        fprintf(stderr, "[msd trace] bgzf: %-34s %9.3f ms\n", what, now - w_last); w_last = now;
    };
    // the contig's Huffman code, from a sample of its own lines (host: 16,384 lines at most)
    static_assert(sizeof(BzCodes) % 4 == 0, "BzCodes is copied as words");
    BzCodes codes;
    if (!bz_build_codes(chrom, h_st, h_en, h_dp, n, &codes, labels, n_labels))
        return bz_fail(B, MSD_EINVAL, "cannot build a code for contig '%s' (name longer than %d bytes, more than %d labels, a label longer than %d bytes, or a value without a label)", chrom,
                       kBzMaxName, kBzMaxLabels, kBzMaxLabel);
    lap("code from a sample of the lines");
    // ---- per line: text length -> text offset ------------------------------------------------------------------------
    const uint64_t n_tiles = (n + kBzScanTile - 1) / kBzScanTile;
    BzCarve lay{nullptr};
    lay.take<BzCodes>(1); lay.take<uint32_t>(n); lay.take<ull>(n); lay.take<uint32_t>(n); lay.take<ull>(n);
    lay.take<uint32_t>(n_tiles); lay.take<ull>(n_tiles); lay.take<ull>(4);
    if ((rc = bz_grow(B, &B.d_line, &B.line_cap, lay.used))) return rc;
    lay.base = (uint8_t*)B.d_line; lay.used = 0;
    BzCodes* d_codes = lay.take<BzCodes>(1);
    uint32_t* d_text_len = lay.take<uint32_t>(n); ull* d_text_off = lay.take<ull>(n);
    uint32_t* d_bit_len = lay.take<uint32_t>(n); ull* d_bit_off = lay.take<ull>(n);
    uint32_t* d_tsum = lay.take<uint32_t>(n_tiles); ull* d_tbase = lay.take<ull>(n_tiles);
    ull* d_totals = lay.take<ull>(4);      // text bytes, bits, output bytes
    BZ_CUDA_OK(cudaMemcpyAsync(d_codes, &codes, sizeof codes, cudaMemcpyHostToDevice, sc));
    const unsigned g_lines = bz_grid_for(n, kBzThreads, max_ctas);
    BZ_LAUNCH(bz_len_kernel, g_lines, kBzThreads, sc, d_codes, d_st, d_en, d_dp, n, d_text_len);
    bz_scan(sc, d_text_len, n, d_tsum, d_tbase, d_totals + 0, d_text_off);
    B.launches += 4;
    ull text_total = 0, last_off = 0;
    BZ_CUDA_OK(cudaGetLastError());
    BZ_CUDA_OK(cudaMemcpyAsync(&text_total, d_totals + 0, 8, cudaMemcpyDeviceToHost, sc));
    BZ_CUDA_OK(cudaMemcpyAsync(&last_off, d_text_off + (n - 1), 8, cudaMemcpyDeviceToHost, sc));
    BZ_CUDA_OK(cudaStreamSynchronize(sc));
    lap("line lengths + scan");
    const uint64_t nb = (uint64_t)bz_block_of(last_off) + 1;
    if (text_total == 0 || last_off >= text_total || nb > (1ull << 30)) return bz_fail(B, MSD_ECUDA, "msd_per_base_bgzf: text scan out of range (%llu bytes, last line at %llu)", text_total, last_off);
    // ---- per member: first line, text range, payload bytes -> member offset -----------------------------------------------
    const uint64_t nb_tiles = (nb + kBzScanTile - 1) / kBzScanTile;
    BzCarve lb{nullptr};
    lb.take<ull>(nb); lb.take<BlockDesc>(nb); lb.take<uint32_t>(nb); lb.take<uint32_t>(nb); lb.take<ull>(nb);
    lb.take<uint32_t>(nb); lb.take<uint32_t>(nb); lb.take<uint32_t>(nb); lb.take<uint32_t>(nb_tiles); lb.take<ull>(nb_tiles); lb.take<uint32_t>(4);
    lb.take<uint8_t>(text_total + 512);
    if ((rc = bz_grow(B, &B.d_blk, &B.blk_cap, lb.used))) return rc;
    lb.base = (uint8_t*)B.d_blk; lb.used = 0;
    ull* d_blk_first = lb.take<ull>(nb); BlockDesc* d_desc = lb.take<BlockDesc>(nb);
    uint32_t* d_payload = lb.take<uint32_t>(nb); uint32_t* d_member_len = lb.take<uint32_t>(nb); ull* d_member_off = lb.take<ull>(nb);
    uint32_t* d_crc = lb.take<uint32_t>(nb); uint32_t* d_crc_expected = lb.take<uint32_t>(nb); uint32_t* d_status = lb.take<uint32_t>(nb);
    uint32_t* d_btsum = lb.take<uint32_t>(nb_tiles); ull* d_btbase = lb.take<ull>(nb_tiles); uint32_t* d_ticket = lb.take<uint32_t>(4);
    uint8_t* d_text = lb.take<uint8_t>(text_total + 512);
    BZ_CUDA_OK(cudaMemsetAsync(d_blk_first, 0xff, 8 * nb, sc));
    BZ_LAUNCH(bz_mark_kernel, g_lines, kBzThreads, sc, d_codes, d_st, d_en, d_dp, n, d_text_off, d_bit_len, d_blk_first);
    bz_scan(sc, d_bit_len, n, d_tsum, d_tbase, d_totals + 1, d_bit_off);
    ull bit_total = 0;
    BZ_CUDA_OK(cudaMemcpyAsync(&bit_total, d_totals + 1, 8, cudaMemcpyDeviceToHost, sc));
    BZ_CUDA_OK(cudaStreamSynchronize(sc));          // bz_blocks_kernel takes the totals by value
    lap("members of lines, bits + scan");
    BZ_LAUNCH(bz_blocks_kernel, (unsigned)((nb + kBzThreads - 1) / kBzThreads), kBzThreads, sc, d_codes, (uint32_t)nb, n, d_blk_first, d_text_off, text_total, d_bit_off, bit_total,
              d_desc, d_payload, d_member_len);
    bz_scan(sc, d_member_len, nb, d_btsum, d_btbase, d_totals + 2, d_member_off);
    B.launches += 8;
    ull out_total = 0;
    BZ_CUDA_OK(cudaGetLastError());
    BZ_CUDA_OK(cudaMemcpyAsync(&out_total, d_totals + 2, 8, cudaMemcpyDeviceToHost, sc));
    BZ_CUDA_OK(cudaStreamSynchronize(sc));
    lap("member sizes + scan");
    if (out_total < 27 * nb || out_total > 65536ull * nb) return bz_fail(B, MSD_ECUDA, "msd_per_base_bgzf: member sizes out of range (%llu bytes in %llu members)", out_total, (ull)nb);
    // ---- output: zeroed, lines OR their bits in, the CRC step sums the text, members get their frames -----------------------
    if ((rc = bz_grow(B, &B.d_out, &B.out_cap, out_total + 64))) return rc;
    if (out_total + 64 > B.h_cap) {
        if (B.h_out) { cudaFreeHost(B.h_out); B.h_out = nullptr; B.h_cap = 0; }
        const uint64_t cap = out_total + out_total / 8 + 4096;
        BZ_CUDA_OK(cudaMallocHost((void**)&B.h_out, cap));
        B.h_cap = cap;
    }
    lap("output buffers");
    uint8_t* d_out = (uint8_t*)B.d_out;
    BZ_CUDA_OK(cudaMemsetAsync(d_out, 0, out_total + 64, sc));
    BZ_CUDA_OK(cudaMemsetAsync(d_text + text_total, 0, 512, sc));            // K1c reads whole 16-byte words
    BZ_CUDA_OK(cudaMemsetAsync(d_crc_expected, 0, 4 * nb, sc)); BZ_CUDA_OK(cudaMemsetAsync(d_status, 0, 4 * nb, sc)); BZ_CUDA_OK(cudaMemsetAsync(d_ticket, 0, 16, sc));
    BZ_LAUNCH(bz_emit_kernel, g_lines, kBzThreads, sc, d_codes, d_st, d_en, d_dp, n, d_text_off, d_bit_off, d_blk_first, d_member_off, d_text, reinterpret_cast<uint32_t*>(d_out));
    crc(sc, d_text, d_desc, (int)nb, d_ticket, d_crc_expected, d_status, d_ticket + 1, d_crc);      // expected = 0: the status words are scratch here
    BZ_LAUNCH(bz_frame_kernel, (unsigned)((nb + kBzThreads - 1) / kBzThreads), kBzThreads, sc, d_codes, (uint32_t)nb, d_desc, d_payload, d_member_off, d_crc, d_out);
    B.launches += 3;
    BZ_CUDA_OK(cudaGetLastError());
    std::vector<ull> h_first(nb), h_off(nb); std::vector<uint32_t> h_len(nb); std::vector<BlockDesc> h_desc(nb);
    BZ_CUDA_OK(cudaMemcpyAsync(B.h_out, d_out, out_total, cudaMemcpyDeviceToHost, sc));
    BZ_CUDA_OK(cudaMemcpyAsync(h_first.data(), d_blk_first, 8 * nb, cudaMemcpyDeviceToHost, sc));
    BZ_CUDA_OK(cudaMemcpyAsync(h_off.data(), d_member_off, 8 * nb, cudaMemcpyDeviceToHost, sc));
    BZ_CUDA_OK(cudaMemcpyAsync(h_len.data(), d_member_len, 4 * nb, cudaMemcpyDeviceToHost, sc));
    BZ_CUDA_OK(cudaMemcpyAsync(h_desc.data(), d_desc, sizeof(BlockDesc) * nb, cudaMemcpyDeviceToHost, sc));
    BZ_CUDA_OK(cudaStreamSynchronize(sc));
    lap("emit + CRC + frames + copy to host");
    B.members.resize(nb);
    for (uint64_t b = 0; b < nb; b++) {
        msd_bgzf_member& m = B.members[b];
        m.offset = h_off[b]; m.first_run = h_first[b]; m.csize = h_len[b]; m.isize = h_desc[b].isize;
        const uint64_t next_first = b + 1 < nb ? h_first[b + 1] : n;
        if (m.csize > 65536 || m.csize < 27 || m.isize == 0 || m.isize > kBzMaxBlockText || m.first_run >= next_first || (b == 0 && m.first_run != 0) ||
            m.offset + m.csize != (b + 1 < nb ? h_off[b + 1] : out_total))
            return bz_fail(B, MSD_ECUDA, "msd_per_base_bgzf: member %llu of %llu is inconsistent (offset %llu, %u -> %u bytes, first run %llu)", (ull)b, (ull)nb, (ull)m.offset, m.isize,
                           m.csize, (ull)m.first_run);
    }
    B.out_bytes = out_total;
    lap("member table");
    return 0;
}

}  // namespace msd
