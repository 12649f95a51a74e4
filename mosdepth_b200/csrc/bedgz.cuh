// bedgz.cuh -- K10 kernels: per-base BED text -> BGZF members on the device (arithmetic in bedgz_core.cuh).
// Pipeline for the n runs (start, stop, depth) that K8 left in device memory, all on one stream, no spin-waits anywhere:
//   bz_len_kernel      text bytes of every line                         -> scan -> text_off[i]
//   bz_mark_kernel     member of every line, first line of every member, bits of every line  -> scan -> bit_off[i]
//   bz_blocks_kernel   per member: text range (BlockDesc for K1c), payload bytes, member bytes  -> scan -> member offsets
//   bz_emit_kernel     per line: text bytes (for the CRC) + token bits ORed into the zeroed output
//   bgzf_crc_kernel    K1c, unchanged: CRC-32 of every member's text
//   bz_frame_kernel    per member: BGZF header, Huffman header bits, CRC32 + ISIZE trailer
// HBM traffic per line (~20 text bytes): 12 B runs x 3 reads, 24 B of offsets written + read, 20 B text written + read by
// the CRC, ~6 B of output: ~130 B per line, so 25 M lines (30X chr20) move ~3 GB -- under a millisecond at HBM speed;
// the cost that matters is the D2H copy of the ~6 B/line result instead of 12 B/line of runs plus host deflate.
#pragma once
#include "bedgz_core.cuh"

namespace msd {

constexpr int kBzThreads = 256;

// ---- exclusive scan uint32 -> uint64 in three launches (tile sums, runs_scan_kernel over the tiles, tile scan) ---------
constexpr int kBzScanItems = 8;
constexpr int kBzScanTile = kBzThreads * kBzScanItems;

__global__ void __launch_bounds__(kBzThreads)
bz_tile_sum_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t s_w[kBzThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kBzScanTile;
    uint32_t c = 0;
    for (int k = 0; k < kBzScanItems; k++) { const uint64_t i = base + (uint64_t)k * kBzThreads + threadIdx.x; if (i < n) c += in[i]; }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t s = 0; for (int w = 0; w < kBzThreads / 32; w++) s += s_w[w]; tile_sum[blockIdx.x] = s; }
}

// thread t owns items [t*kBzScanItems, (t+1)*kBzScanItems) of its tile
__global__ void __launch_bounds__(kBzThreads)
bz_tile_scan_kernel(const uint32_t* __restrict__ in, uint64_t n, const unsigned long long* __restrict__ tile_base, unsigned long long* __restrict__ out) {
    __shared__ uint32_t s_w[kBzThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kBzScanTile + (uint64_t)threadIdx.x * kBzScanItems;
    uint32_t v[kBzScanItems]; uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kBzScanItems; k++) { v[k] = base + k < n ? in[base + k] : 0u; c += v[k]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    uint32_t wex = 0;
    for (int w = 0; w < warp; w++) wex += s_w[w];
    unsigned long long o = tile_base[blockIdx.x] + wex + x - c;
#pragma unroll
    for (int k = 0; k < kBzScanItems; k++) { if (base + k < n) out[base + k] = o; o += v[k]; }
}

// ---- K10 --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBzThreads)
bz_len_kernel(const BzCodes* __restrict__ C, const int32_t* __restrict__ st, const int32_t* __restrict__ en, const int32_t* __restrict__ dp, uint64_t n,
              uint32_t* __restrict__ text_len) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        text_len[i] = bz_line_len(*C, st[i], en[i], dp[i]);
}

struct BzTables { uint32_t litlen[286]; uint32_t dist[30]; };
__device__ __forceinline__ void bz_load_tables(const BzCodes* __restrict__ C, BzTables* s) {
    for (int k = threadIdx.x; k < 286; k += blockDim.x) s->litlen[k] = C->litlen[k];
    for (int k = threadIdx.x; k < 30; k += blockDim.x) s->dist[k] = C->dist[k];
    __syncthreads();
}

__global__ void __launch_bounds__(kBzThreads)
bz_mark_kernel(const BzCodes* __restrict__ C, const int32_t* __restrict__ st, const int32_t* __restrict__ en, const int32_t* __restrict__ dp, uint64_t n,
               const unsigned long long* __restrict__ text_off, uint32_t* __restrict__ bit_len, unsigned long long* __restrict__ blk_first) {
    __shared__ BzTables T;
    bz_load_tables(C, &T);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t to = text_off[i];
        const BzLineCtx lc = bz_line_ctx(*C, st, en, dp, i, n, to);
        if (lc.first) blk_first[bz_block_of(to)] = i;
        bit_len[i] = bz_line_bits(*C, T.litlen, T.dist, st, en, dp, i, lc);
    }
}

// member b: lines [blk_first[b], blk_first[b+1]); desc (text range for K1c), payload, member length
__global__ void __launch_bounds__(kBzThreads)
bz_blocks_kernel(const BzCodes* __restrict__ C, uint32_t nb, uint64_t n, const unsigned long long* __restrict__ blk_first,
                 const unsigned long long* __restrict__ text_off, unsigned long long text_total,
                 const unsigned long long* __restrict__ bit_off, unsigned long long bit_total,
                 BlockDesc* __restrict__ desc, uint32_t* __restrict__ payload, uint32_t* __restrict__ member_len) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint64_t f = blk_first[b];
    const bool last = b + 1 == nb;
    const uint64_t f1 = last ? n : blk_first[b + 1];
    const uint64_t t0 = text_off[f], t1 = last ? text_total : text_off[f1];
    const uint64_t b0 = bit_off[f], b1 = last ? bit_total : bit_off[f1];
    desc[b].src = 0; desc[b].clen = 0; desc[b].dst = t0; desc[b].isize = (uint32_t)(t1 - t0);
    const uint32_t p = bz_payload_bytes(C->hdr_bits, b1 - b0);
    payload[b] = p; member_len[b] = p + 26;
}

__global__ void __launch_bounds__(kBzThreads)
bz_emit_kernel(const BzCodes* __restrict__ C, const int32_t* __restrict__ st, const int32_t* __restrict__ en, const int32_t* __restrict__ dp, uint64_t n,
               const unsigned long long* __restrict__ text_off, const unsigned long long* __restrict__ bit_off, const unsigned long long* __restrict__ blk_first,
               const unsigned long long* __restrict__ member_off, uint8_t* __restrict__ text, uint32_t* __restrict__ out32) {
    __shared__ BzTables T;
    bz_load_tables(C, &T);
    const uint32_t hdr_bits = C->hdr_bits;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t to = text_off[i];
        const BzLineCtx lc = bz_line_ctx(*C, st, en, dp, i, n, to);
        const uint32_t b = bz_block_of(to);
        const uint64_t bitpos = (member_off[b] + 18) * 8 + hdr_bits + (bit_off[i] - bit_off[blk_first[b]]);
        bz_line_emit(*C, T.litlen, T.dist, st, en, dp, i, lc, text, to, out32, bitpos);
    }
}

__global__ void __launch_bounds__(kBzThreads)
bz_frame_kernel(const BzCodes* __restrict__ C, uint32_t nb, const BlockDesc* __restrict__ desc, const uint32_t* __restrict__ payload,
                const unsigned long long* __restrict__ member_off, const uint32_t* __restrict__ crc, uint8_t* __restrict__ out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    bz_frame(out + member_off[b], *C, payload[b], crc[b], desc[b].isize);
}

}  // namespace msd
