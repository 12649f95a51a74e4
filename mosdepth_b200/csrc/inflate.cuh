// inflate.cuh -- K1 bgzf_inflate: raw DEFLATE (RFC 1951) decoder for BGZF blocks, as two kernels.
//
// Replaces htslib's bgzf inflate (+ its thread pool, `open(bam, threads=)`, mosdepth.nim:888,968) [ext].
// Each BGZF block is an independent DEFLATE stream (the LZ77 window never crosses blocks), so the parallelism is
// across blocks: a 30X WGS BAM has ~3.3 M of them.
//
// Design (r1g).  The earlier kernels (one decoder per warp, r1a; four 8-lane decoders per warp, r1c-r1f) spent ~30
// issued warp instructions per Huffman symbol because every instruction of the serial chain served 1 or 4 of 32
// lanes, and tied an L2/HBM round trip per match copy to that chain (profiles/README.md).  Now:
//
//   K1a huffman_decode_kernel -- ONE LANE PER BGZF BLOCK, 32 independent decoders per warp, kept in lock step by
//       construction: the symbol loop is a fixed-trip loop whose body is predicated on the lane's phase and ends in
//       __syncwarp(), so every issued instruction advances up to 32 chains.  A lane writes literal bytes straight to
//       their final position and appends one 4-byte token per match; it never reads its own output, so no memory
//       round trip sits on the chain (the next compressed word is prefetched one refill ahead).
//       Tables: NO root lookup tables.  In lock step a warp pays for the slowest lane, and with 32 lanes some lane
//       always has a code longer than any affordable root table (r1g first cut: 85 % of iterations), so every symbol
//       is decoded the same, branch-free way: the 15 per-length limits of the canonical code live in REGISTERS, the
//       code length is the number of limits <= the next 15 bits (15 independent compares), and shared memory holds
//       only the symbols in canonical order as 16-bit entries (literal byte | length base, extra-bit count) plus
//       the 16 index bases: 704 B per lane, halfword-interleaved across the warp -> 9 warps = 288 decoders per SM,
//       42,624 per B200 (the first cut, with 8/7-bit root tables, needed 1376 B: 5 warps per SM, latency bound).
//       Dynamic headers are read by their lane (lanes in the header phase run together); the tables of one lane are
//       then built by all 32 lanes (__match_any_sync ranks, redundant uniform canonical bookkeeping).
//       Lanes re-synchronise their phases every kBurst symbols.
//   K1b lz77_resolve_kernel -- ONE WARP PER BLOCK.  32 tokens per batch (warp scan -> output offsets); the batch's
//       output range is walked in 32-byte windows, one output byte per lane: a position-indexed bit mask of token
//       starts (redux.or) + popc gives every byte its token, match bytes load dst[p - dist] and store dst[p]
//       (one 32-byte sector per store instruction; 64-byte windows with two bytes per lane were tried and lost:
//       a third of them contain a source of their own and had to be redone in halves).  Bytes whose source lies inside their own window find their
//       root byte by pointer doubling over shuffles (5 steps; period folding for dist < len).
#pragma once
#include <cstdlib>
#include "inflate_core.cuh"

namespace msd {
using namespace inf;

constexpr int kDecThreads = 32;             // one warp per CTA
constexpr int kDecCtasPerSm = 9;
constexpr int kBurst = 256;                 // symbol iterations between phase re-synchronisations
constexpr int kScratchWords = 32;           // cnt[16] offs[16]
constexpr int kRingWords = 16;              // compressed words staged per lane
constexpr size_t kDecSmem = (size_t)kLaneHw * 32 * 2 + kRingWords * 32 * 4 + kScratchWords * 4;
static_assert(kDecCtasPerSm * (kDecSmem + 1024) <= 233472, "9 decoder CTAs must fit one SM's shared memory");

__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// output of K1a is written once and read by another kernel: streaming stores, so that the few KB of L1 left beside the
// tables keep the compressed words (at full occupancy the L1 hit rate of the refill loads was 1 %, profiles/README.md)
MSD_D void st_stream_u8(uint8_t* p, uint32_t v) { asm volatile("st.global.cs.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
MSD_D void st_stream_u32(uint32_t* p, uint32_t v) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// one lane's table region: halfword i at base[i * 32]
struct LaneMem {
    uint16_t* base;
    MSD_D uint32_t ld16(uint32_t hw) const { return base[hw * 32]; }
    MSD_D int lds16(uint32_t hw) const { return reinterpret_cast<const int16_t*>(base)[hw * 32]; }
    MSD_D void st16(uint32_t hw, uint32_t v) const { base[hw * 32] = (uint16_t)v; }
    MSD_D uint32_t ld8(uint32_t b) const { return reinterpret_cast<const uint8_t*>(base + (b >> 1) * 32)[b & 1]; }
    MSD_D void st8(uint32_t b, uint32_t v) const { reinterpret_cast<uint8_t*>(base + (b >> 1) * 32)[b & 1] = (uint8_t)v; }
};

// per-lane bit reader.  64-bit bit buffer; the next two compressed words wait in registers (a refill right after a
// refill -- a long lit/len code followed by its distance -- must not wait for a load issued a few instructions earlier).
// The words come from a 16-word ring per lane in shared memory that is topped up 16 bytes at a time: one 128-bit global
// load per 4 words, issued three 16-byte chunks ahead and parked in four registers until the lane enters the next
// chunk.  (r1g/r1h read the stream with one 32-bit global load per word: with 214 KB of the SM's 256 KB carved out as
// shared memory the L1 hit rate of those loads was 1 %, every word cost an L2 round trip that the slowest lane of the
// warp exposed, 39 % of all stall samples; profiles/README.md.)
struct BitReader {
    uint64_t bb;
    int bc;
    uint32_t widx;    // index of the word held in a0 == words already shifted into bb
    uint32_t a0, a1;  // words widx, widx + 1
    uint32_t p0, p1, p2, p3;   // chunk (widx / 4) + 3, in flight or landed, not yet in the ring
    static MSD_D uint32_t ring_ld(const uint32_t* ring, uint32_t w) { return ring[(w & 15u) * 32]; }
    static MSD_D void ring_st(uint32_t* ring, uint32_t chunk, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
        uint32_t* r = ring + (chunk & 3u) * 128;
        r[0] = x; r[32] = y; r[64] = z; r[96] = w;
    }
    MSD_D void seed(const uint4* __restrict__ chunks, uint32_t clast, uint32_t* ring, uint32_t byte_pos) {
        bb = 0; bc = 0; widx = byte_pos >> 2;
        const uint32_t c0 = widx >> 2;
#pragma unroll
        for (uint32_t i = 0; i < 3; i++) { const uint4 v = __ldg(chunks + min(c0 + i, clast)); ring_st(ring, c0 + i, v.x, v.y, v.z, v.w); }
        { const uint4 v = __ldg(chunks + min(c0 + 3, clast)); p0 = v.x; p1 = v.y; p2 = v.z; p3 = v.w; }
        a0 = ring_ld(ring, widx); a1 = ring_ld(ring, widx + 1);
        refill(chunks, clast, ring); drop((int)(byte_pos & 3u) * 8); refill(chunks, clast, ring);
    }
    MSD_D void refill(const uint4* __restrict__ chunks, uint32_t clast, uint32_t* ring) {
        if (bc <= 32) {
            bb |= (uint64_t)a0 << bc; bc += 32; widx++;
            a0 = a1;
            const bool entered = (widx & 3u) == 0;          // first word of a new chunk c: park chunk c + 2, fetch c + 3
            if (entered) ring_st(ring, (widx >> 2) + 2, p0, p1, p2, p3);
            // predicated PTX load whose destinations ARE p0..p3: left to the compiler a conditional load becomes a load
            // into temporaries plus moves in the same iteration, i.e. a wait for the full L2 latency
            const uint4* addr = chunks + min((widx >> 2) + 3, clast);
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
                         : "+r"(p0), "+r"(p1), "+r"(p2), "+r"(p3) : "l"(addr), "r"((uint32_t)entered));
            a1 = ring_ld(ring, widx + 1);
        }
    }
    MSD_D uint32_t lo() const { return (uint32_t)bb; }
    MSD_D void drop(int n) { bb >>= n; bc -= n; }
    MSD_D uint32_t take(int n) { const uint32_t v = (uint32_t)bb & ((1u << n) - 1u); drop(n); return v; }
    MSD_D uint32_t bits_to_byte() const { return (uint32_t)bc & 7u; }
    MSD_D uint32_t byte_pos() const { return widx * 4u - (uint32_t)bc / 8u; }   // byte index in the aligned stream
    MSD_D uint32_t word_pos() const { return widx; }
};
// which reader the decode kernel uses: 0 = the bit buffer above, 1 = BitReader2 (inflate_core.cuh: windows read from the ring)
#ifndef MSD_BITREADER2
#define MSD_BITREADER2 1
#endif

// canon_len_minus1(): inflate_core.cuh (shared with the host replay).

// Build the tables of lane `X` with all 32 lanes.  XM: X's region (same pointer in every lane).  Code lengths are
// read from the region first (they alias the tables).  `keep` is true in lane X only: it keeps the per-length limits
// in its registers.  Returns a Status (uniform across the warp).
MSD_D uint32_t coop_build(const LaneMem XM, uint32_t* scratch, int hlit, int hdist, int lane, bool keep,
                          uint32_t (&llim)[8], uint32_t (&dlim)[8]) {
    constexpr unsigned full = 0xffffffffu;
    uint32_t* cnt = scratch; uint32_t* offs = scratch + 16;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t status = kOk;

    uint32_t l[9];
#pragma unroll
    for (int j = 0; j < 9; j++) { const int s = lane + 32 * j; l[j] = s < hlit ? XM.ld8(kLensByte + s) : 0u; }
    const uint32_t dl = lane < hdist ? XM.ld8(kLensByte + hlit + lane) : 0u;
    __syncwarp();

    // ================= lit/len alphabet =================
    if (lane < 16) cnt[lane] = 0;
    __syncwarp();
    uint32_t rank[9];
#pragma unroll
    for (int j = 0; j < 9; j++) {
        const uint32_t m = __match_any_sync(full, l[j]);
        const uint32_t r = __popc(m & lt);
        const uint32_t pre = cnt[l[j]];
        __syncwarp();
        if (r == 0) cnt[l[j]] = pre + __popc(m);
        __syncwarp();
        rank[j] = pre + r;
    }
    {
        // canonical bookkeeping, computed redundantly by every lane (uniform) so that lane X simply keeps its copy
        uint32_t code = 0, idx = 0, lp[8];
#pragma unroll
        for (int j = 0; j < 8; j++) lp[j] = j == 7 ? 0x80000000u : 0u;      // the 16th "limit" is never reached
#pragma unroll
        for (int k = 1; k < 16; k++) {
            const uint32_t c = cnt[k];
            code <<= 1;
            if (lane == k) { offs[k] = idx; XM.st16(kLitBaseHw + k, (idx - code) & 0xffffu); }
            code += c; idx += c;
            if (code > (1u << k)) status = kOversubscribed;
            lp[(k - 1) >> 1] |= min(code << (15 - k), 0x8000u) << (16 * ((k - 1) & 1));
        }
        if (keep) {
#pragma unroll
            for (int j = 0; j < 8; j++) llim[j] = lp[j];
        }
        if (lane == 16) { XM.st16(kLitBaseHw + 16, kLitBadSlot); XM.st16(kLitSymHw + kLitBadSlot, kEntBad); }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 9; j++) if (l[j]) XM.st16(kLitSymHw + offs[l[j]] + rank[j], lit_entry_of((uint32_t)(lane + 32 * j)));
        __syncwarp();
    }

    // ================= distance alphabet =================
    if (lane < 16) cnt[lane] = 0;
    __syncwarp();
    {
        const uint32_t m = __match_any_sync(full, dl);
        const uint32_t r = __popc(m & lt);
        if (r == 0) cnt[dl] = __popc(m);
        __syncwarp();
        uint32_t code = 0, idx = 0, lp[8];
#pragma unroll
        for (int j = 0; j < 8; j++) lp[j] = j == 7 ? 0x80000000u : 0u;
#pragma unroll
        for (int k = 1; k < 16; k++) {
            const uint32_t c = cnt[k];
            code <<= 1;
            if (lane == k) { offs[k] = idx; XM.st16(kDistBaseHw + k, (idx - code) & 0xffffu); }
            code += c; idx += c;
            if (code > (1u << k)) status = kOversubscribed;
            lp[(k - 1) >> 1] |= min(code << (15 - k), 0x8000u) << (16 * ((k - 1) & 1));
        }
        if (keep) {
#pragma unroll
            for (int j = 0; j < 8; j++) dlim[j] = lp[j];
        }
        if (lane == 16) { XM.st16(kDistBaseHw + 16, kDistBadSlot); XM.st16(kDistSymHw + kDistBadSlot, kDistNibBad << 4); }
        __syncwarp();
        if (dl) XM.st16(kDistSymHw + offs[dl] + r, dist_entry_of((uint32_t)lane));
        __syncwarp();
    }
    return status;
}

enum DecPhase : int { kPhNeed = 0, kPhHeader = 1, kPhBuild = 2, kPhSym = 3, kPhFinish = 4, kPhFail = 5, kPhDone = 6 };

// K1a.  comp: compressed chunk buffer (4-byte aligned base); out: inflated chunk buffer.  Literals go to out, matches
// to tokens[b * kTokenCap ..]; n_tokens[b] and status[b] are written for every block.
// The second launch bound only caps the registers: shared memory allows 9 decoder warps per SM whatever it says, and what
// the decoder does not take of the register file is what lets resolve CTAs share its SM (MSD_DEC_MINBLOCKS 9 / 16 / 20 ->
// 156 / 122 / 96 registers, no spills worth the name; A/B in profiles/README.md, r1n).
#ifndef MSD_DEC_MINBLOCKS
#define MSD_DEC_MINBLOCKS 9
#endif
// One decode warp: 32 lanes = 32 blocks in lock step (see the header comment).  kPacked = false: literal bytes go straight to their final
// position in `out` (the two-kernel pipeline).  kPacked = true (the fused kernel): literals are written packed, downwards from the top of
// the block's own token region, so that the resolve warp is the only writer of `out` and can pick the block up while this kernel is still
// running (no other SM's L1 can hold a stale line of bytes that a decode lane wrote); a finished block is announced in `done_ring`.
template <bool kPacked>
__device__ __forceinline__ void
decode_warp(uint16_t* __restrict__ smem16, const uint8_t* __restrict__ comp, uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks, int n_blocks,
            uint32_t* __restrict__ ticket, uint32_t* __restrict__ tokens, uint32_t* __restrict__ n_tokens,
            uint32_t* __restrict__ status_out, uint32_t* __restrict__ error_count, uint32_t* __restrict__ crc_expected,
            unsigned long long* __restrict__ done_ring, uint32_t* __restrict__ done_tail) {
    constexpr unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const LaneMem M{smem16 + lane};
    uint32_t* const ring = reinterpret_cast<uint32_t*>(smem16 + kLaneHw * 32) + lane;   // word w of this lane: ring[(w & 15) * 32]
    uint32_t* const scratch = reinterpret_cast<uint32_t*>(smem16 + kLaneHw * 32) + kRingWords * 32;

    int phase = kPhNeed, b = 0, bfinal = 0, hlit = 0, hdist = 0;
    const uint4* __restrict__ chunks = nullptr;   // the block's DEFLATE stream from the 16-byte boundary at or before its first byte
    uint8_t* __restrict__ dst = nullptr;
    uint32_t* __restrict__ tok = nullptr;   // this block's token list
    uint8_t* __restrict__ lit_top = nullptr;   // literal k of the block goes to lit_top[-k]: the top of the block's token region, downwards (r2d)
    uint32_t ntok = 0, nlit = 0;
    uint32_t clast = 0, total_words = 0, isize = 0, lead = 0, clen = 0, outp = 0, run_base = 0, err = 0;
#if MSD_BITREADER2
    BitReader2<32> br; br.ring.base = ring; br.bitpos = 0; br.cur = 0; br.p0 = br.p1 = br.p2 = br.p3 = 0;
#else
    BitReader br; br.bb = 0; br.bc = 0; br.widx = 0; br.a0 = 0; br.a1 = 0; br.p0 = br.p1 = br.p2 = br.p3 = 0;
#endif
    uint32_t llim[8], dlim[8];     // per-length limits of the two canonical codes, packed two per register
#pragma unroll
    for (int k = 0; k < 8; k++) { llim[k] = 0; dlim[k] = 0; }

    for (;;) {
        // ---- blocks that ended during the last burst; new tickets --------------------------------------------------
        if (phase == kPhFinish) {
            if (outp != isize) { err = kShort; phase = kPhFail; }
            else {
                if (kPacked && outp > run_base) st_stream_u32(tok + ntok++, (kTokSkip << 23) | (outp - run_base));   // the literals behind the last match
                n_tokens[b] = ntok; status_out[b] = kOk; phase = kPhNeed;
                if (kPacked) { __threadfence(); done_ring[atomicAdd(done_tail, 1u)] = ((unsigned long long)ntok << 32) | (uint32_t)(b + 1); }   // literals + tokens are visible: a resolve warp may take the block
            }
        }
        if (phase == kPhFail) {
            n_tokens[b] = 0; status_out[b] = err ? err : (uint32_t)kBadSymbol; atomicAdd(error_count, 1u); phase = kPhNeed;
            if (kPacked) { __threadfence(); done_ring[atomicAdd(done_tail, 1u)] = (unsigned long long)(uint32_t)(b + 1); }    // nothing to resolve
        }
        if (phase == kPhNeed) {
            b = (int)atomicAdd(ticket, 1u);
            if (b >= n_blocks) phase = kPhDone;
            else {
                const BlockDesc bd = blocks[b];
                const uint8_t* first = comp + bd.src;
                lead = (uint32_t)((uintptr_t)first & 15u); clen = bd.clen; isize = bd.isize;
                chunks = reinterpret_cast<const uint4*>(first - lead);
                { const uint8_t* tr = first + bd.clen;    // gzip trailer: CRC32, ISIZE (the compressed buffer may be recycled before K1c runs)
                  crc_expected[b] = (uint32_t)tr[0] | ((uint32_t)tr[1] << 8) | ((uint32_t)tr[2] << 16) | ((uint32_t)tr[3] << 24); }
                total_words = (lead + clen + 3) / 4; clast = (lead + clen + 15) / 16; if (clast) clast--;
                dst = out + bd.dst;
                tok = tokens + (size_t)b * kTokenCap; ntok = 0; nlit = 0;
                lit_top = reinterpret_cast<uint8_t*>(tok + kTokenCap) - 1;
                outp = 0; run_base = 0; err = 0; bfinal = 0;
                br.seed(chunks, clast, ring, lead);
                phase = kPhHeader;
            }
        }
        if (__all_sync(full, phase == kPhDone)) break;

        // ---- DEFLATE block headers (lanes in this phase run together) ------------------------------------------------------
        if (phase == kPhHeader) {
            br.refill(chunks, clast, ring);
            if (br.word_pos() > total_words + 4u) err = kTruncated;
            bfinal = (int)br.take(1);
            const int btype = (int)br.take(2);
            if (err) {
            } else if (btype == 0) {
                br.drop((int)br.bits_to_byte());
                br.refill(chunks, clast, ring);
                const uint32_t len = br.take(16), nlen = br.take(16);
                const uint32_t sp = br.byte_pos();                             // byte index in the aligned stream
                if ((len ^ 0xffffu) != nlen || sp + len > lead + clen) err = kBadStored;
                else if (outp + len > isize) err = kOverflow;
                else {
                    const uint8_t* sb = reinterpret_cast<const uint8_t*>(chunks) + sp;
                    if (kPacked) { for (uint32_t k = 0; k < len; k++) *(lit_top - (nlit + k)) = sb[k]; }
                    else { for (uint32_t k = 0; k < len; k++) dst[outp + k] = sb[k]; }
                    outp += len; nlit += len;                               // stored bytes count as literals of the next token
                    br.seed(chunks, clast, ring, sp + len);
                    if (bfinal) phase = kPhFinish;
                }
            } else if (btype == 1) {
                hlit = 288; hdist = 30;
                for (int i = 0; i < 288; i++) M.st8(kLensByte + i, i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                for (int i = 0; i < 30; i++) M.st8(kLensByte + 288 + i, 5);
                phase = kPhBuild;
            } else if (btype == 2) {
                hlit = (int)br.take(5) + 257; hdist = (int)br.take(5) + 1;
                const int hclen = (int)br.take(4) + 4;
                if (hlit > 286 || hdist > 30) err = kBadCodeLengths;
                // the 19 code-length-code lengths, 3 bits each, packed; counts 5 bits x 8; next codes 8 bits x 8
                uint64_t clp = 0;
                for (int i = 0; i < hclen; i++) { br.refill(chunks, clast, ring); clp |= (uint64_t)br.take(3) << (3 * c_cl_order[i]); }
                uint64_t cntp = 0;
                for (int s = 0; s < 19; s++) cntp += 1ull << (5 * (int)((clp >> (3 * s)) & 7));
                cntp &= ~31ull;
                int left = 1; uint32_t code = 0; uint64_t nextp = 0;
                for (int q = 1; q < 8; q++) {
                    const uint32_t c = (uint32_t)(cntp >> (5 * q)) & 31u;
                    left = (left << 1) - (int)c;
                    code <<= 1;
                    nextp |= (uint64_t)(code & 0xffu) << (8 * q);
                    code += c;
                }
                if (left < 0) err = kOversubscribed;
                if (!err) {
                    for (int j = 0; j < 64; j++) M.st16(kClTabByte / 2 + j, 0u);
                    for (int s = 0; s < 19; s++) {
                        const uint32_t q = (uint32_t)(clp >> (3 * s)) & 7u;
                        if (!q) continue;
                        const uint32_t c = (uint32_t)(nextp >> (8 * q)) & 0xffu;
                        nextp += 1ull << (8 * q);
                        const uint32_t rev = __brev(c) >> (32 - q);
                        for (uint32_t k = rev; k < 128; k += (1u << q)) M.st8(kClTabByte + k, (s << 3) | q);
                    }
                    const int total = hlit + hdist;
                    int idx = 0;
                    while (idx < total) {
                        br.refill(chunks, clast, ring);
                        const uint32_t e = M.ld8(kClTabByte + (br.lo() & 127u));
                        const uint32_t q = e & 7u, sym = e >> 3;
                        if (!q) { err = kBadCodeLengths; break; }
                        br.drop((int)q);
                        if (sym < 16) { M.st8(kLensByte + idx, sym); idx++; continue; }
                        uint32_t rep, val = 0;
                        if (sym == 16) { if (idx == 0) { err = kBadCodeLengths; break; } val = M.ld8(kLensByte + idx - 1); rep = 3 + br.take(2); }
                        else if (sym == 17) rep = 3 + br.take(3);
                        else rep = 11 + br.take(7);
                        if (idx + (int)rep > total) { err = kBadCodeLengths; break; }
                        while (rep--) { M.st8(kLensByte + idx, val); idx++; }
                    }
                    if (!err && M.ld8(kLensByte + 256) == 0) err = kBadCodeLengths;   // no end-of-block code
                    if (!err) phase = kPhBuild;
                }
            } else {
                err = kBadBlockType;
            }
            if (err) phase = kPhFail;
        }
        __syncwarp();

        // ---- table construction: one lane's tables at a time, by the whole warp ----------------------------------------------
        {
            uint32_t bm = __ballot_sync(full, phase == kPhBuild);
            while (bm) {
                const int X = __ffs((int)bm) - 1; bm &= bm - 1;
                const int hl = __shfl_sync(full, hlit, X), hd = __shfl_sync(full, hdist, X);
                const uint32_t st = coop_build(LaneMem{smem16 + X}, scratch, hl, hd, lane, lane == X, llim, dlim);
                if (lane == X) { if (st) { err = st; phase = kPhFail; } else phase = kPhSym; }
            }
        }

        // ---- symbols: a burst of lock-step iterations ---------------------------------------------------------------------------
        if (!__any_sync(full, phase == kPhSym)) continue;
#pragma unroll 1
        for (int o = 0; o < kBurst / 32; o++) {
#pragma unroll 1
            for (int it = 0; it < 32; it++) {
                if (phase == kPhSym) {
                    br.refill(chunks, clast, ring);
                    uint32_t lo = br.lo();
                    uint32_t c = __brev(lo) >> 17;
                    uint32_t l1 = canon_len_minus1(c, llim);                       // code length - 1 (15: no such code)
                    uint32_t idx = (uint32_t)(M.lds16(kLitBaseHw + 1 + l1) + (int)(c >> ((14u - l1) & 31u)));
                    const uint32_t e = M.ld16(kLitSymHw + min(idx, (uint32_t)kLitBadSlot));
                    const uint32_t x = (e >> 4) & 7u;                              // extra bits of a length code
                    const uint32_t ext = (lo >> (l1 + 1)) & ((1u << x) - 1u);
                    br.drop((int)(l1 + 1 + x));
                    if ((e & 0x83u) == kEntNotLen) {                               // literal
                        if (outp < isize) { st_stream_u8(kPacked ? lit_top - nlit : dst + outp, e >> 8); outp++; nlit++; }
                        else err = kOverflow;
                    } else if (e & kEntNotLen) {
                        if (e == kEntEob) phase = bfinal ? kPhFinish : kPhHeader;
                        else err = kBadSymbol;
                    } else {
                        const uint32_t len = 3u + (e >> 8) + ext;
#if !MSD_BITREADER2
                        br.refill(chunks, clast, ring);        // (BitReader2: the ring already holds the next two chunks)
#endif
                        lo = br.lo();
                        c = __brev(lo) >> 17;
                        l1 = canon_len_minus1(c, dlim);
                        idx = (uint32_t)(M.lds16(kDistBaseHw + 1 + l1) + (int)(c >> ((14u - l1) & 31u)));
                        const uint32_t d = M.ld16(kDistSymHw + min(idx, (uint32_t)kDistBadSlot));
                        const uint32_t dx = (d >> 4) & 15u;
                        const uint32_t dist = 1u + (((d >> 8) & 3u) << dx) + ((lo >> (l1 + 1)) & ((1u << dx) - 1u));
                        br.drop((int)(l1 + 1 + dx));
                        if (dx == kDistNibBad || dist > outp) err = kBadDistance;
                        else if (outp + len > isize) err = kOverflow;
                        else {
                            uint32_t lit_run = outp - run_base;
                            if (lit_run > 510u) { st_stream_u32(tok + ntok++, (kTokSkip << 23) | lit_run); lit_run = 0; }
                            st_stream_u32(tok + ntok++, make_token(lit_run, len, dist));
                            outp += len; run_base = outp;
                        }
                    }
                    if (err) phase = kPhFail;
                }
                __syncwarp();
            }
            if (!__any_sync(full, phase == kPhSym)) break;
        }
    }
}

__global__ void __launch_bounds__(kDecThreads, MSD_DEC_MINBLOCKS)
huffman_decode_kernel(const uint8_t* __restrict__ comp, uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks, int n_blocks,
                      uint32_t* __restrict__ ticket, uint32_t* __restrict__ tokens, uint32_t* __restrict__ n_tokens,
                      uint32_t* __restrict__ status_out, uint32_t* __restrict__ error_count, uint32_t* __restrict__ crc_expected) {
    extern __shared__ __align__(16) uint16_t smem16[];
    decode_warp<false>(smem16, comp, out, blocks, n_blocks, ticket, tokens, n_tokens, status_out, error_count, crc_expected, nullptr, nullptr);
}

// K1b.  One warp per BGZF block: copies every match of the block's token list; literals are already in place.
// A batch is 31 tokens in lanes 1..31 (lane 0 is a sentinel "no token"), so that a byte's token index -- the number of
// token starts at or before it -- is used as a shuffle lane directly.
constexpr int kResolveWarps = 4;
constexpr int kResolveCtasPerSm = 16;
constexpr int kResolveBatch = 31;

// the matches (kPacked: and the literals) of ONE block, by one warp.  nt tokens at tok; kPacked: literal k of the block at lit_top[-k].
template <bool kPacked>
__device__ __forceinline__ void resolve_block(uint8_t* dst, const uint32_t* __restrict__ tok, int nt, int lane) {
    constexpr unsigned full = 0xffffffffu;
    const uint32_t le = 0xffffffffu >> (31 - lane);   // lanes <= this one
    {
        {
        asm volatile("" : "+l"(dst));                  // keep the block base as ONE 64-bit register pair
        const int mis = (int)((uintptr_t)dst & 31u);
        const uint8_t* lit_top = reinterpret_cast<const uint8_t*>(tok + kTokenCap) - 1;   // literal k of the block: lit_top[-k]
        int pos_base = 0, lit_base = 0;
        // kPacked: the tokens were written during THIS kernel (by a decode lane, before it announced the block): ordinary loads, not ld.global.nc
        auto ldtok = [&](int i) -> uint32_t { return kPacked ? tok[i] : __ldg(tok + i); };
        uint32_t w_next = (lane >= 1 && lane - 1 < nt) ? ldtok(lane - 1) : 0u;
        for (int t0 = 0; t0 < nt; t0 += kResolveBatch) {
            const uint32_t w = w_next;
            const bool have = lane >= 1 && t0 + lane - 1 < nt;
            { const int tn = t0 + kResolveBatch + lane - 1; w_next = (lane >= 1 && tn < nt) ? ldtok(tn) : 0u; }   // next batch: in flight during this one
            int lr = 0, dist = 0, adv = 0;
            bool is_match = false;
            if (have) {
                lr = (int)(w >> 23);
                if (lr == (int)kTokSkip) { lr = (int)(w & 0x7fffffu); adv = lr; }
                else { dist = (int)((w >> 8) & 0x7fffu) + 1; adv = lr + (int)(w & 0xffu) + 3; is_match = true; }
            }
            int incl = adv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(full, incl, o); if (lane >= o) incl += y; }
            const int total = __shfl_sync(full, incl, 31);
            const int start = pos_base + incl - adv;                 // first output byte of this token (its literals)
            int ldelta = 0;
            if (kPacked) {
                int lincl = lr;                                      // literals up to and including this token's run
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(full, lincl, o); if (lane >= o) lincl += y; }
                ldelta = (lit_base + lincl - lr) - start;            // byte p of this token's literal run is literal number p + ldelta of the block
                lit_base += __shfl_sync(full, lincl, 31);
            }
            const int mstart = is_match ? start + lr : 0x3fffffff;   // first byte of its match (sentinel / skip: never)
            const int hi = pos_base + total;
            const int wb0 = ((pos_base + mis) & ~31) - mis;          // windows are 32-byte aligned in memory
            const int head_win = adv > 0 ? (start - wb0) >> 5 : -1;  // window in which this token starts
            const uint32_t head_bit = 1u << ((start - wb0) & 31);
            int before = 0;                                          // tokens of the batch that start before the window
            int j = 0;
            uint8_t* pp = dst + wb0 + lane;                          // this lane's byte of the window
            for (int wb = wb0; wb < hi; wb += 32, j++, pp += 32) {
                const uint32_t heads = __reduce_or_sync(full, head_win == j ? head_bit : 0u);
                const int m = before + __popc(heads & le);          // lane that holds this byte's token (0: none yet)
                before += __popc(heads);
                const int ms = __shfl_sync(full, mstart, m);
                const int md = __shfl_sync(full, dist, m);
                const int ld = kPacked ? __shfl_sync(full, ldelta, m) : 0;
                const int p = wb + lane;
                const int k = p - ms;
                const bool is_copy = k >= 0 && p < hi;
                const bool is_lit = kPacked && !is_copy && p >= pos_base && p < hi;   // a literal of this batch: fetched from the packed literals; the window is written whole
                int back = md;                                       // source = p - back
                if (is_copy && k >= md) back = k - k % md + md;      // dist < len: fold onto the first period
                const bool far = is_copy && back > lane;            // source lies before this window: final already
                const bool near = is_copy && back <= lane;          // source is a byte of this window
                uint8_t v = 0;
                if (far) v = *(pp - back);
                if (is_lit) v = *(lit_top - (p + ld));
                if (__any_sync(full, near)) {
                    // sources inside the window: every lane follows its source lane to a root (a byte that is final:
                    // a literal, a byte of an earlier batch, or a far copy) by pointer doubling, then takes its value
                    if (!is_copy && !is_lit && p >= 0 && p < hi) v = *pp;
                    int root = near ? lane - back : lane;
#pragma unroll
                    for (int r = 0; r < 5; r++) root = __shfl_sync(full, root, root);
                    v = (uint8_t)__shfl_sync(full, (int)v, root);
                }
                if (is_copy || is_lit) *pp = v;
                __syncwarp();
            }
            pos_base = hi;
        }
        }
    }
}

__global__ void __launch_bounds__(kResolveWarps * 32, kResolveCtasPerSm)
lz77_resolve_kernel(uint8_t* out, const BlockDesc* __restrict__ blocks, int n_blocks, uint32_t* __restrict__ ticket,
                    const uint32_t* __restrict__ tokens, const uint32_t* __restrict__ n_tokens) {
    constexpr unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (;;) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(ticket, 1u);
        b = __shfl_sync(full, b, 0);
        if (b >= n_blocks) return;
        resolve_block<false>(out + blocks[b].dst, tokens + (size_t)b * kTokenCap, (int)n_tokens[b], lane);
    }
}

// K1 fused (r2h): decode and resolve in ONE kernel, warp-specialised.  A CTA is 4 decode warps (warpgroup 0: the tables of 128 blocks in
// shared memory, 152 registers per thread) + 12 resolve warps (warpgroups 1-3, 32 registers per thread); two CTAs per SM.  A decode lane
// announces a finished block in `done_ring`; resolve warps take ring slots in order and wait for their slot to be filled.  Decode never
// waits for resolve and every block in progress belongs to a resident CTA, so the waits end.  Why: as two kernels the pair runs as the
// SUM of its stand-alone times (16.8 + 10.6 ms per 85 k blocks) although decode leaves 65 % of the issue slots idle -- its CTAs hold the
// shared memory and most of the register file, so a concurrently launched resolve kernel only gets a few warps per SM; and between the
// two kernels ~9.5 k blocks x 64 KB overflow L2, so resolve's match sources come back from HBM.  Here resolve runs in decode's idle
// slots and picks a block up microseconds after it was decoded.
#ifndef MSD_FUSED_RES_WARPS
#define MSD_FUSED_RES_WARPS 12       // resolve warps per CTA (a multiple of 4: setmaxnreg works on warpgroups)
#endif
#ifndef MSD_FUSED_DEC_REGS
#define MSD_FUSED_DEC_REGS 152       // registers per decode thread after setmaxnreg.inc
#endif
constexpr int kFusedDecWarps = 4, kFusedResWarps = MSD_FUSED_RES_WARPS, kFusedThreads = (kFusedDecWarps + kFusedResWarps) * 32, kFusedCtasPerSm = 2;
// setmaxnreg.inc waits until the CTA's pool holds the registers: what the resolve warps give back (launch allocation - 32 each) must cover what
// the decode warps ask for, or the kernel hangs (r2i: 20 resolve warps + 96 decode registers = 40 at launch, 160 returned, 224 wanted).
constexpr int kFusedLaunchRegs = (65536 / (kFusedCtasPerSm * kFusedThreads)) / 8 * 8;
static_assert(kFusedResWarps * (kFusedLaunchRegs - 32) >= kFusedDecWarps * (MSD_FUSED_DEC_REGS - kFusedLaunchRegs), "fused inflate: the decode warps would wait for registers that are never released");
constexpr size_t kFusedSmem = kFusedDecWarps * ((kDecSmem + 127) / 128 * 128);
constexpr uint32_t kFusedSpinLimit = 1u << 22;      // x ~0.5 us: a resolve warp gives up after ~2 s without progress (status word, no hang)

__global__ void __launch_bounds__(kFusedThreads, kFusedCtasPerSm)
inflate_fused_kernel(const uint8_t* __restrict__ comp, uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks, int n_blocks,
                     uint32_t* __restrict__ ticket /* [0] decode ticket, [1] ring tail, [2] ring head */, uint32_t* __restrict__ tokens,
                     uint32_t* __restrict__ n_tokens, uint32_t* __restrict__ status_out, uint32_t* __restrict__ error_count,
                     uint32_t* __restrict__ crc_expected, unsigned long long* __restrict__ done_ring) {
    extern __shared__ __align__(128) uint16_t smem16[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < kFusedDecWarps) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(MSD_FUSED_DEC_REGS));
        decode_warp<true>(smem16 + (size_t)warp * (((kDecSmem + 127) / 128 * 128) / 2), comp, out, blocks, n_blocks, ticket, tokens, n_tokens, status_out, error_count,
                          crc_expected, done_ring, ticket + 1);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");
        constexpr unsigned full = 0xffffffffu;
        for (;;) {
            uint32_t slot = 0;
            if (lane == 0) slot = atomicAdd(ticket + 2, 1u);
            slot = __shfl_sync(full, slot, 0);
            if (slot >= (uint32_t)n_blocks) return;
            unsigned long long v = 0;
            if (lane == 0) {
                uint32_t spins = 0;
                while ((v = *((volatile unsigned long long*)(done_ring + slot))) == 0ull) {
                    __nanosleep(400);
                    if (++spins > kFusedSpinLimit) { v = ~0ull; break; }
                }
                __threadfence();        // the block's literals and tokens were written before the announcement
            }
            v = __shfl_sync(full, v, 0);
            if (v == ~0ull) { if (lane == 0) atomicAdd(error_count, 1u); return; }     // (cannot happen: see above; a bound instead of a hang)
            const int b = (int)(uint32_t)v - 1, nt = (int)(v >> 32);
            resolve_block<true>(out + blocks[b].dst, tokens + (size_t)b * kTokenCap, nt, lane);
        }
    }
}

// K1c.  CRC-32 of every inflated block against the gzip trailer (htslib's bgzf reader checks it [ext]; a corrupt stream can
// still inflate to ISIZE bytes).  One warp per block, row-strided formulation (inflate_core.cuh, "K1c"): every load is a
// coalesced 512-byte request, every table lookup goes to the lane's own bank (tables replicated per lane, 128 KB per
// CTA), so the kernel streams the inflated bytes at HBM speed instead of replaying bank conflicts (profiles/README.md, r1m).
constexpr int kCrcThreads = 512;
constexpr int kCrcUWords = 4 * 256 * 32;
constexpr int kCrcSmemWords = kCrcUWords + 256 + kCrcXaN + kCrcXbN + 3;
constexpr size_t kCrcSmem = (size_t)kCrcSmemWords * 4;
__device__ uint32_t g_crc_tables[kCrcTableWords];       // CrcTables, written by msd_create (crc_upload_tables)

inline cudaError_t crc_upload_tables() {
    static CrcTables host; static bool built = false;
    if (!built) { crc_build_tables(host); built = true; }
    cudaError_t e = cudaMemcpyToSymbol(g_crc_tables, &host, sizeof host);
    return e;
}

// MSD_CRC_DEPTH rows of a block in registers.  16 rows need 128 registers x 512 threads = the whole register file of an SM: alone the kernel streams at 79 % of the
// HBM peak (r2j); 8 rows fit 64 registers and reach 63 % (r2p).  In msd_run the kernel runs beside the persistent decode / resolve CTAs of the next chunks, where the
// footprint could matter: interleaved A/B at f = 0.06 (r2p) 70.5 ms per step with 8 rows, 71.0-71.3 with 16 -- inside the run-to-run noise, so the faster kernel ships.
// (An unguarded main loop for the rows that surely exist was slower: 70 % alone, r2o.)
#ifndef MSD_CRC_DEPTH
#define MSD_CRC_DEPTH 16
#endif
__global__ void __launch_bounds__(kCrcThreads, MSD_CRC_DEPTH > 8 ? 1 : 2)
bgzf_crc_kernel(const uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks, int n_blocks, uint32_t* __restrict__ ticket,
                const uint32_t* __restrict__ crc_expected, uint32_t* __restrict__ status, uint32_t* __restrict__ error_count,
                uint32_t* __restrict__ crc_out /* test hook: the computed CRCs (null in msd_run) */) {
    extern __shared__ __align__(16) uint32_t crc_sm[];
    uint32_t* const U = crc_sm;
    uint32_t* const T0 = U + kCrcUWords;
    uint32_t* const XA = T0 + 256;
    uint32_t* const XBI = XA + kCrcXaN;
    constexpr unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const CrcTables* const G = reinterpret_cast<const CrcTables*>(g_crc_tables);
    for (int e0 = warp * 32; e0 < 1024; e0 += kCrcThreads) {            // entry e0 + s of U goes to all 32 lane copies
        const uint32_t mine = (&G->u[0][0])[e0 + lane];
#pragma unroll 8
        for (int s = 0; s < 32; s++) U[(e0 + s) * 32 + lane] = __shfl_sync(full, mine, s);
    }
    for (int i = threadIdx.x; i < 256; i += kCrcThreads) T0[i] = G->t0[i];
    for (int i = threadIdx.x; i < kCrcXaN; i += kCrcThreads) XA[i] = G->xa[i];
    for (int i = threadIdx.x; i < kCrcXbN; i += kCrcThreads) XBI[i] = G->xbi[i];
    const uint32_t xl = G->xl[lane];
    __syncthreads();
    // one lookup = byte extract + multiply-add + LDS: entry (m, i) of this lane at Ub + 32768 m + 128 i
    const uint32_t Ub = (uint32_t)__cvta_generic_to_shared(U) + (uint32_t)lane * 4u;
    auto lk = [&](uint32_t byte, uint32_t m) -> uint32_t {
        uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(byte * 128u + (Ub + 32768u * m)));
        return v;
    };
    auto step = [&](uint32_t w) -> uint32_t {
        uint32_t b0, b1, b2, b3;       // PRMT for all four bytes (the compiler turns `w & 0xff` and `w >> 24` times 128 into shift + and + add)
        asm("prmt.b32 %0, %1, 0, 0x4440;" : "=r"(b0) : "r"(w)); asm("prmt.b32 %0, %1, 0, 0x4441;" : "=r"(b1) : "r"(w));
        asm("prmt.b32 %0, %1, 0, 0x4442;" : "=r"(b2) : "r"(w)); asm("prmt.b32 %0, %1, 0, 0x4443;" : "=r"(b3) : "r"(w));
        return lk(b0, 0) ^ lk(b1, 1) ^ lk(b2, 2) ^ lk(b3, 3);
    };
    auto t0 = [&](uint32_t i) -> uint32_t { return T0[i]; };
    // Software pipeline (r2i).  kCrcDepth rows of a block live in registers; a row's registers are loaded again (kCrcDepth rows further
    // on) right after the row is folded, so every warp keeps kCrcDepth x 512 bytes of loads in flight all the time (r1m issued eight
    // rows, folded them, and only then issued the next eight: the DRAM latency was paid once per 4 KB and the kernel stopped at 58-60 %
    // of the HBM rate).  The warp also holds two tickets: the descriptor of the next block is read while this one streams, its first
    // rows are requested before this block's serial epilogue (28 dependent table steps + three 32-step products), and the ticket
    // after that is in flight the whole time.
    struct Blk { int b; bool ok; uint32_t n, pad, n2, R, tail; const uint4* q; };
    auto fetch = [&](int b) -> Blk {
        Blk k; k.b = b; k.ok = false; k.n = 0; k.pad = 0; k.n2 = 0; k.R = 0; k.tail = 0; k.q = nullptr;
        if (b < n_blocks && status[b] == kOk) {                 // (a block that already failed in K1a is skipped)
            k.ok = true; k.n = blocks[b].isize;
            const uint8_t* p = out + blocks[b].dst;
            k.pad = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u); k.n2 = k.n + k.pad; k.R = k.n2 / kCrcRow; k.tail = k.n2 % kCrcRow;
            k.q = reinterpret_cast<const uint4*>(p - k.pad) + lane;
        }
        return k;
    };
    constexpr int kCrcDepth = MSD_CRC_DEPTH;
    uint4 buf[kCrcDepth];
    auto load_first = [&](const Blk& k) {
#pragma unroll
        for (int j = 0; j < kCrcDepth; j++) if ((uint32_t)j < k.R) buf[j] = __ldg(k.q + (size_t)j * 32);
    };
    uint32_t d[4];
    int b0 = 0, b1 = 0;
    if (lane == 0) { b0 = (int)atomicAdd(ticket, 1u); b1 = (int)atomicAdd(ticket, 1u); }
    b0 = __shfl_sync(full, b0, 0); b1 = __shfl_sync(full, b1, 0);
    Blk cur = fetch(b0);
    load_first(cur);
    while (cur.b < n_blocks) {                                  // a warp's tickets ascend: nothing is left once one is out of range
        const Blk nxt = fetch(b1);
        int b2 = 0;
        if (lane == 0) b2 = (int)atomicAdd(ticket, 1u);
        // the lane's 16 bytes of the tail row, requested now and used in the epilogue
        uint4 tv = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t g0 = cur.R * kCrcRow + (uint32_t)lane * 16u;
        if (cur.ok && g0 < cur.n2) tv = __ldg(cur.q + (size_t)cur.R * 32);
        d[0] = d[1] = d[2] = d[3] = 0u;
        if (cur.R && cur.pad && lane == 0) {                    // the bytes in front of the message cancel themselves (row 0 is buf[0])
            const CrcWords c = crc_head_cancel(CrcWords{buf[0].x, buf[0].y, buf[0].z, buf[0].w}, 0, cur.pad);
            d[0] = c.x; d[1] = c.y; d[2] = c.z; d[3] = c.w;
        }
        for (uint32_t r0 = 0; r0 < cur.R; r0 += kCrcDepth) {
#pragma unroll
            for (int j = 0; j < kCrcDepth; j++) {
                const uint32_t r = r0 + j;
                if (r < cur.R) { d[0] = step(buf[j].x ^ d[0]); d[1] = step(buf[j].y ^ d[1]); d[2] = step(buf[j].z ^ d[2]); d[3] = step(buf[j].w ^ d[3]); }
                if (r + kCrcDepth < cur.R) buf[j] = __ldg(cur.q + (size_t)(r + kCrcDepth) * 32);
            }
        }
        load_first(nxt);                                        // in flight during the epilogue
        if (cur.ok) {
            const CrcWords t = (g0 < cur.n2) ? crc_tail_keep(CrcWords{tv.x, tv.y, tv.z, tv.w}, g0, cur.pad, cur.n2) : CrcWords{0u, 0u, 0u, 0u};
            uint32_t v = crc_multmodp(xl, crc_lane_fold(d, t, t0));
            v = __reduce_xor_sync(full, v);
            const uint32_t crc = crc_finish(v, XA[cur.tail], XA[cur.n % kCrcRow], XBI[cur.n / kCrcRow]);
            if (lane == 0) {
                if (crc_out) crc_out[cur.b] = crc;
                if (crc != crc_expected[cur.b]) { status[cur.b] = kBadCrc; atomicAdd(error_count, 1u); }
            }
        }
        cur = nxt;
        b1 = __shfl_sync(full, b2, 0);
    }
}

inline void launch_crc(cudaStream_t s, const uint8_t* d_out, const BlockDesc* d_desc, int nb, uint32_t* ticket, const uint32_t* d_crc,
                       uint32_t* d_status, uint32_t* d_error_count, uint32_t* d_crc_out = nullptr) {
    const int cw = (nb + kCrcThreads / 32 - 1) / (kCrcThreads / 32);
    bgzf_crc_kernel<<<cw < kSMs ? cw : kSMs, kCrcThreads, kCrcSmem, s>>>(d_out, d_desc, nb, ticket, d_crc, d_status, d_error_count, d_crc_out);
}

// K1 = K1a + K1b (+ K1c) on one stream.  tickets: three zeroed words (K1a, K1b, K1c).  comp_read (may be null) is recorded
// after K1a: from then on the compressed chunk buffer may be overwritten.  resolved (may be null) is recorded after K1b.
// Tuning knobs (read at every launch so that one process can sweep them, tools/sweep_inflate.py): MSD_DEC_CTAS caps the decode
// warps per SM (default 9 = what shared memory allows), MSD_RESOLVE_CTAS the resolve CTAs per SM the grid is sized for.
inline int env_int(const char* name, int dflt, int lo, int hi) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const int x = atoi(v);
    return x < lo ? lo : x > hi ? hi : x;
}

// tickets: FOUR zeroed words (K1a / ring tail, K1b / ring head ... see below, K1c at [2], fused ring head at [3]).  d_ring (nb zeroed 64-bit words) selects
// the fused kernel when MSD_FUSED=1.
inline bool inflate_fused_enabled() { static const int on = env_int("MSD_FUSED", 0, 0, 1); return on != 0; }
inline void launch_inflate(cudaStream_t s, const uint8_t* d_comp, uint8_t* d_out, const BlockDesc* d_desc, int nb, uint32_t* tickets,
                           uint32_t* d_tokens, uint32_t* d_ntok, uint32_t* d_status, uint32_t* d_error_count, uint32_t* d_crc,
                           bool check_crc, cudaEvent_t comp_read, cudaEvent_t resolved = nullptr, cudaEvent_t decoded_trace = nullptr,
                           unsigned long long* d_ring = nullptr) {
    if (d_ring && inflate_fused_enabled()) {
        cudaMemsetAsync(d_ring, 0, 8ull * (size_t)nb, s);
        const int want = (nb + kFusedDecWarps * 32 - 1) / (kFusedDecWarps * 32), cap = kSMs * kFusedCtasPerSm;
        inflate_fused_kernel<<<want < cap ? want : cap, kFusedThreads, kFusedSmem, s>>>(d_comp, d_out, d_desc, nb, tickets, d_tokens, d_ntok, d_status, d_error_count, d_crc, d_ring);
        if (comp_read) cudaEventRecord(comp_read, s);
        if (decoded_trace) cudaEventRecord(decoded_trace, s);
        if (resolved) cudaEventRecord(resolved, s);
        if (check_crc) launch_crc(s, d_out, d_desc, nb, tickets + 3, d_crc, d_status, d_error_count);
        return;
    }
    const int dec_cap = kSMs * env_int("MSD_DEC_CTAS", kDecCtasPerSm, 1, kDecCtasPerSm);
    const int dctas = (nb + kDecThreads - 1) / kDecThreads < dec_cap ? (nb + kDecThreads - 1) / kDecThreads : dec_cap;
    huffman_decode_kernel<<<dctas, kDecThreads, kDecSmem, s>>>(d_comp, d_out, d_desc, nb, tickets, d_tokens, d_ntok, d_status, d_error_count, d_crc);
    if (comp_read) cudaEventRecord(comp_read, s);
    if (decoded_trace) cudaEventRecord(decoded_trace, s);
    const int want = (nb + kResolveWarps - 1) / kResolveWarps;
    const int res_cap = kSMs * env_int("MSD_RESOLVE_CTAS", kResolveCtasPerSm, 1, kResolveCtasPerSm);
    const int rctas = want < res_cap ? want : res_cap;
    lz77_resolve_kernel<<<rctas, kResolveWarps * 32, 0, s>>>(d_out, d_desc, nb, tickets + 1, d_tokens, d_ntok);
    if (resolved) cudaEventRecord(resolved, s);      // the inflated bytes are final: framing need not wait for the CRC check
    if (check_crc) launch_crc(s, d_out, d_desc, nb, tickets + 2, d_crc, d_status, d_error_count);
}

}  // namespace msd
