// frame.cuh -- K2 bam_frame: find every BAM record start in the inflated stream of one chunk.
//
// Replaces htslib's bam_read1 record framing behind `for rec in bam.query(...)` (mosdepth.nim:174-193) [ext].
// Record starts are only discoverable by chasing block_size, which is serial.  The chunk is therefore cut into
// SEGMENTS, each chased speculatively from its first byte; a single-CTA link kernel then validates the guesses in
// parallel (segment s is right iff the chain of segment s-1 lands exactly on it) and walks serially only where a
// guess was wrong.  Segment 0 starts at the carry: the unfinished record bytes of the previous chunk, copied in
// front of this chunk's first block, whose first byte is a record start by construction.
//
// Where the segments start (r2): r1 started one at the first byte of every BGZF block -- right for htslib-written
// short-read files (a record never straddles a block when it fits), wrong for every block of a long-read file or an
// htsjdk-written one, where the single thread of the link kernel then walked the whole chunk record by record
// (C5: 249 ms of framing per 280 ms step, profiles/README.md r2c).  Now a warp per block looks for the first
// position in its block that LOOKS like a record start -- the test split-BAM readers use: block_size, refID, pos,
// l_read_name, l_seq, next_refID, next_pos in range, the fixed part fits block_size, the read name ends in NUL, and
// the record behind it passes the same test -- and the segments start there (frame_guess_kernel,
// frame_segments_kernel).  A block that holds no record start has no segment; a wrong guess is caught by the link
// kernel exactly like a wrong block start was, so the guess can only cost time, never change the result.
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
    uint32_t n_seg;       // segments of this chunk (frame_segments_kernel)
    uint32_t pad;
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

// does a BAM record start at p?  (the fixed-length core must lie inside the valid bytes)
MSD_D bool rec_plausible(const uint8_t* __restrict__ buf, uint32_t p, uint32_t limit, int n_targets) {
    if ((uint64_t)p + 36 > limit) return false;
    const uint32_t bs = ld_u32(buf, p);
    if (bs < 32 || bs > kMaxRecordBytes) return false;
    const int32_t tid = (int32_t)ld_u32(buf, p + 4), pos = (int32_t)ld_u32(buf, p + 8);
    if (tid < -1 || tid >= n_targets || pos < -1) return false;
    const uint32_t l_qname = ld_u32(buf, p + 12) & 0xffu, n_cigar = ld_u32(buf, p + 16) & 0xffffu;
    const int32_t l_seq = (int32_t)ld_u32(buf, p + 20), mtid = (int32_t)ld_u32(buf, p + 24), mpos = (int32_t)ld_u32(buf, p + 28);
    if (l_qname < 1 || l_seq < 0 || mtid < -1 || mtid >= n_targets || mpos < -1) return false;
    if (32ull + l_qname + 4ull * n_cigar + ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq > bs) return false;
    const uint64_t nul = (uint64_t)p + 36 + l_qname - 1;
    if (nul < limit && buf[nul] != 0) return false;
    return true;
}

// one warp per BGZF block b = 1..nb (block 0 is the carry, a record start by construction): guess[b] = the first position in the
// block where a record seems to start and the record behind it seems to start as well, or kNoOff
__global__ void frame_guess_kernel(const uint8_t* __restrict__ buf, const ChunkState* __restrict__ cs, const uint32_t* __restrict__ bstart, int nblk,
                                   int n_targets, uint32_t* __restrict__ guess) {
    const int b = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (b >= nblk) return;
    if (b == 0) { if (lane == 0) guess[0] = bstart[0]; return; }
    const uint32_t limit = cs->limit, beg = bstart[b], end = bstart[b + 1];
    uint32_t found = kNoOff;
    for (uint32_t p0 = beg; p0 < end; p0 += 32) {
        const uint32_t p = p0 + lane;
        bool ok = p < end && rec_plausible(buf, p, limit, n_targets);
        if (ok) { const uint64_t q = (uint64_t)p + 4 + ld_u32(buf, p); if (q + 36 <= limit) ok = rec_plausible(buf, (uint32_t)q, limit, n_targets); }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) { found = p0 + (uint32_t)(__ffs((int)m) - 1); break; }
    }
    if (lane == 0) guess[b] = found;
}

// single CTA: the guessed starts, compacted in order -> seg[0 .. n_seg), seg[n_seg] = limit, cs->n_seg
__global__ void __launch_bounds__(1024, 1)
frame_segments_kernel(ChunkState* __restrict__ cs, const uint32_t* __restrict__ guess, int nblk, uint32_t* __restrict__ seg) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_run;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
        const int b = base + t;
        const uint32_t g = b < nblk ? guess[b] : kNoOff;
        const bool keep = g != kNoOff;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        uint32_t wex = 0, tot = 0;
        for (int w = 0; w < 32; w++) { const uint32_t x = s_w[w]; if (w < warp) wex += x; tot += x; }
        if (keep) seg[s_run + wex + __popc(m & ((1u << lane) - 1u))] = g;
        __syncthreads();
        if (t == 0) s_run += tot;
        __syncthreads();
    }
    if (t == 0) { seg[s_run] = cs->limit; cs->n_seg = s_run; }
}

// speculative chase: one thread per segment (segment 0 starts at the carry)
__global__ void frame_spec_kernel(const uint8_t* __restrict__ buf, const ChunkState* __restrict__ cs, const uint32_t* __restrict__ bstart,
                                  int nblk_cap, uint32_t* __restrict__ land, uint32_t* __restrict__ count) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int nblk = (int)cs->n_seg;
    if (b >= nblk || b >= nblk_cap) return;
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
frame_link_kernel(const uint8_t* __restrict__ buf, ChunkState* __restrict__ cs, const uint32_t* __restrict__ bstart, int nblk_cap,
                  const uint32_t* __restrict__ land, const uint32_t* __restrict__ count, uint32_t* __restrict__ first,
                  uint32_t* __restrict__ cnt, uint32_t* __restrict__ rec_base) {
    __shared__ uint32_t s_p, s_b, s_done, s_err, s_min, s_tail;
    __shared__ uint32_t s_scan[32];
    const int t = threadIdx.x, nt = blockDim.x;
    const int nblk = (int)cs->n_seg < nblk_cap ? (int)cs->n_seg : nblk_cap;      // "blocks" below are segments
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
__global__ void frame_emit_kernel(const uint8_t* __restrict__ buf, const ChunkState* __restrict__ cs, const uint32_t* __restrict__ first, const uint32_t* __restrict__ cnt,
                                  const uint32_t* __restrict__ rec_base, int nblk_cap, uint32_t* __restrict__ rec_off) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (int)cs->n_seg || b >= nblk_cap) return;
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
