// frame.cuh -- K2 bam_frame: find every BAM record start in the inflated stream of one chunk.
//
// Replaces htslib's bam_read1 record framing behind `for rec in bam.query(...)` (mosdepth.nim:174-193) [ext].
// Record starts are only discoverable by chasing block_size, which is serial.  The chunk is therefore cut at
// BGZF block boundaries and every block is chased speculatively from its first byte (htslib never lets a record
// straddle a block when it fits, so the guess is right for short-read files).  A single-CTA link kernel then
// validates the guesses in parallel (block b is right iff the chain of block b-1 lands exactly on it) and
// walks serially only where a record really straddles blocks (long reads, htsjdk-written files, the carried
// tail of the previous chunk).  "Block 0" is the carry: the unfinished record bytes of the previous chunk,
// copied in front of this chunk's first block, whose first byte is a record start by construction.
#pragma once
#include "common.cuh"

namespace msd {

constexpr uint32_t kNoOff = 0xffffffffu;
constexpr uint32_t kMaxRecordBytes = 1u << 28;   // sanity bound on block_size

// device-resident state that flows from chunk to chunk on the stream (no host round trip)
struct ChunkState {
    uint32_t s0;          // absolute offset in this chunk's buffer where framing starts (start of carry)
    uint32_t limit;       // absolute end of valid inflated bytes
    uint32_t tail_start;  // first byte of the first incomplete record (== limit when none)
    uint32_t n_rec;       // records framed in this chunk
    uint32_t error;       // non-zero: corrupt record chain
    uint32_t carry_len;   // bytes to carry into the next chunk
    uint32_t pad[2];
};

// bstart[0] = s0 (carry), bstart[1+b] = base + rel[b], bstart[nb+1] = limit
__global__ void frame_setup_kernel(ChunkState* __restrict__ cs, const uint32_t* __restrict__ rel, int nb, uint32_t base,
                                   uint32_t first_uoff, uint32_t total, int first_of_span, uint32_t* __restrict__ bstart) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        uint32_t s0 = first_of_span ? base + first_uoff : base - cs->carry_len;
        cs->s0 = s0; cs->limit = base + total; cs->tail_start = base + total; cs->n_rec = 0;
        if (first_of_span) cs->error = 0;
        bstart[0] = s0;
        bstart[nb + 1] = base + total;
    }
    if (i < nb) bstart[1 + i] = base + ((i == 0 && first_of_span) ? first_uoff : rel[i]);
}

// speculative chase: one thread per block (0..nb inclusive of the carry pseudo-block)
__global__ void frame_spec_kernel(const uint8_t* __restrict__ buf, const ChunkState* __restrict__ cs, const uint32_t* __restrict__ bstart,
                                  int nblk, uint32_t* __restrict__ land, uint32_t* __restrict__ count) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint32_t limit = cs->limit;
    uint32_t p = bstart[b], end = bstart[b + 1], c = 0;
    while (p < end) {
        if (p + 4 > limit) break;
        uint32_t bs = ld_u32(buf, p);
        if (bs < 32 || bs > kMaxRecordBytes) { p = kNoOff; break; }
        if ((uint64_t)p + 4 + bs > limit) break;
        c++; p += 4 + bs;
    }
    land[b] = p; count[b] = c;
}

// single CTA: validate / repair the speculative chains, then exclusive-scan the per-block record counts
__global__ void __launch_bounds__(1024, 1)
frame_link_kernel(const uint8_t* __restrict__ buf, ChunkState* __restrict__ cs, const uint32_t* __restrict__ bstart, int nblk,
                  const uint32_t* __restrict__ land, const uint32_t* __restrict__ count, uint32_t* __restrict__ first,
                  uint32_t* __restrict__ cnt, uint32_t* __restrict__ rec_base) {
    __shared__ uint32_t s_p, s_b, s_done, s_err, s_min, s_tail;
    __shared__ uint32_t s_scan[32];
    const int t = threadIdx.x, nt = blockDim.x;
    const uint32_t limit = cs->limit;
    for (int b = t; b < nblk; b += nt) { first[b] = kNoOff; cnt[b] = 0; }
    if (t == 0) { s_p = bstart[0]; s_b = 0; s_done = 0; s_err = 0; s_tail = limit; }
    __syncthreads();
    for (;;) {
        // phase A (thread 0): walk serially until the chain stands exactly on a block start
        if (t == 0) {
            uint32_t p = s_p, b = s_b;
            for (;;) {
                while (b < (uint32_t)nblk && bstart[b + 1] <= p && !(bstart[b] == p && bstart[b + 1] == p)) b++;
                if (b >= (uint32_t)nblk || p >= limit) { s_done = 1; s_tail = p < limit ? p : limit; break; }
                if (p == bstart[b]) break;   // aligned: hand over to the parallel phase
                // p is inside block b: chase to the end of the block by hand
                uint32_t end = bstart[b + 1], c = 0, q = p; bool incomplete = false;
                while (q < end) {
                    if (q + 4 > limit) { incomplete = true; break; }
                    uint32_t bs = ld_u32(buf, q);
                    if (bs < 32 || bs > kMaxRecordBytes) { s_err = 1; incomplete = true; break; }
                    if ((uint64_t)q + 4 + bs > limit) { incomplete = true; break; }
                    c++; q += 4 + bs;
                }
                first[b] = p; cnt[b] = c;
                p = q;
                if (incomplete) { s_done = 1; s_tail = q; break; }
                b++;
            }
            s_p = p; s_b = b; s_min = (uint32_t)nblk;
        }
        __syncthreads();
        if (s_done) break;
        // phase B (all threads): first block >= s_b whose speculative chain does not land on its successor
        const uint32_t b0 = s_b;
        uint32_t my = (uint32_t)nblk;
        for (uint32_t b = b0 + t; b < (uint32_t)nblk; b += nt) if (land[b] != bstart[b + 1]) { my = b; break; }
        my = __reduce_min_sync(0xffffffffu, my);
        if ((t & 31) == 0 && my < (uint32_t)nblk) atomicMin(&s_min, my);
        __syncthreads();
        const uint32_t bad = s_min;
        const uint32_t hi = bad < (uint32_t)nblk ? bad + 1 : (uint32_t)nblk;   // blocks [b0, hi) start on their first byte
        for (uint32_t b = b0 + t; b < hi; b += nt) { first[b] = bstart[b]; cnt[b] = count[b]; }
        __syncthreads();
        if (t == 0) {
            if (bad >= (uint32_t)nblk) { s_done = 1; s_tail = limit; }
            else {
                uint32_t l = land[bad];
                if (l == kNoOff) { s_err = 1; s_done = 1; s_tail = limit; }
                else if (l < bstart[bad + 1]) { s_done = 1; s_tail = l; }   // incomplete record at the end of the chunk
                else { s_p = l; s_b = bad + 1; }
            }
        }
        __syncthreads();
        if (s_done) break;
    }
    __syncthreads();
    // exclusive scan of cnt -> rec_base
    uint32_t running = 0;
    for (int base = 0; base < nblk; base += nt) {
        int b = base + t;
        uint32_t v = b < nblk ? cnt[b] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((t & 31) >= o) x += y; }
        if ((t & 31) == 31) s_scan[t >> 5] = x;
        __syncthreads();
        if (t < 32) {
            uint32_t w = t < (nt >> 5) ? s_scan[t] : 0, z = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, z, o); if (t >= o) z += y; }
            s_scan[t] = z - w;
        }
        __syncthreads();
        uint32_t excl = running + s_scan[t >> 5] + x - v;
        if (b < nblk) rec_base[b] = excl;
        __syncthreads();
        if (t == nt - 1) s_scan[0] = excl + v;
        __syncthreads();
        running = s_scan[0];
        __syncthreads();
    }
    if (t == 0) {
        cs->n_rec = running; cs->tail_start = s_tail; cs->carry_len = limit - s_tail;
        if (s_err) cs->error = 1;
    }
}

// write the record offsets: one thread per block
__global__ void frame_emit_kernel(const uint8_t* __restrict__ buf, const uint32_t* __restrict__ first, const uint32_t* __restrict__ cnt,
                                  const uint32_t* __restrict__ rec_base, int nblk, uint32_t* __restrict__ rec_off) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    uint32_t p = first[b], n = cnt[b], o = rec_base[b];
    for (uint32_t i = 0; i < n; i++) { rec_off[o + i] = p; p += 4 + ld_u32(buf, p); }
}

// copy the unfinished tail in front of the next chunk's first block
__global__ void carry_copy_kernel(const uint8_t* __restrict__ src_buf, uint8_t* __restrict__ dst_buf, const ChunkState* __restrict__ cs,
                                  uint32_t base, uint32_t carry_cap, uint32_t* __restrict__ error) {
    const uint32_t len = cs->carry_len, from = cs->tail_start;
    if (len > carry_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(error, 2u); return; }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) dst_buf[base - len + i] = src_buf[from + i];
}

}  // namespace msd
