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
//       Tables: 8-bit lit/len root + 7-bit distance root (16-bit entries carrying length base / extra-bit count, so a
//       symbol costs one shared-memory load), canonical-order lists for longer codes whose per-length limits live in
//       REGISTERS (5 warps/SM leave 400 registers per lane).  1376 B of shared memory per lane, word-interleaved
//       across the warp (bank = lane: conflict free) -> 5 warps = 160 decoders per SM, 23,680 per B200.
//       Dynamic headers are read by their lane (lanes in the header phase run together); the tables of one lane are
//       then built by all 32 lanes (__match_any_sync ranks, redundant uniform canonical bookkeeping, root filled
//       by table index).  Lanes re-synchronise their phases every kBurst symbols.
//   K1b lz77_resolve_kernel -- ONE WARP PER BLOCK.  32 tokens per batch (warp scan -> output offsets); the batch's
//       output range is walked in 32-byte windows, one output byte per lane: a position-indexed bit mask of token
//       starts (redux.or) + popc gives every byte its token, match bytes load dst[p - dist] and store dst[p]
//       (one 32-byte sector per store instruction).  Bytes whose source lies inside their own window wait for it
//       (ballot rounds; period folding for dist < len).
#pragma once
#include "inflate_core.cuh"

namespace msd {
using namespace inf;

constexpr int kDecThreads = 32;             // one warp per CTA
constexpr int kDecCtasPerSm = 5;
constexpr int kBurst = 256;                 // symbol iterations between phase re-synchronisations
constexpr int kScratchWords = 48 + 160;     // cnt[16] first[16] offs[16] + 320 sorted halfwords
constexpr size_t kDecSmem = (size_t)(kLaneWords * 32 + kScratchWords) * 4;
static_assert(kDecCtasPerSm * (kDecSmem + 1024) <= 233472, "5 decoder CTAs must fit one SM's shared memory");

__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// one lane's table region: word j at base[j * 32]
struct LaneMem {
    uint32_t* base;
    MSD_D uint32_t ld16(uint32_t hw) const { return reinterpret_cast<const uint16_t*>(base)[(hw >> 1) * 64 + (hw & 1)]; }
    MSD_D void st16(uint32_t hw, uint32_t v) const { reinterpret_cast<uint16_t*>(base)[(hw >> 1) * 64 + (hw & 1)] = (uint16_t)v; }
    MSD_D uint32_t ld8(uint32_t b) const { return reinterpret_cast<const uint8_t*>(base)[(b >> 2) * 128 + (b & 3)]; }
    MSD_D void st8(uint32_t b, uint32_t v) const { reinterpret_cast<uint8_t*>(base)[(b >> 2) * 128 + (b & 3)] = (uint8_t)v; }
    MSD_D void stw(uint32_t j, uint32_t v) const { base[j * 32] = v; }
};

// per-lane bit reader: 64-bit buffer, next compressed word prefetched one refill ahead
struct BitReader {
    uint64_t bb;
    int bc;
    uint32_t widx;    // index of the word held in `ahead` == words already shifted into bb
    uint32_t ahead;
    MSD_D void seed(const uint32_t* __restrict__ words, uint32_t wlast, uint32_t byte_pos) {
        bb = 0; bc = 0; widx = byte_pos >> 2;
        ahead = __ldg(words + min(widx, wlast));
        refill(words, wlast); drop((int)(byte_pos & 3u) * 8); refill(words, wlast);
    }
    MSD_D void refill(const uint32_t* __restrict__ words, uint32_t wlast) {
        if (bc <= 32) {
            bb |= (uint64_t)ahead << bc; bc += 32; widx++;
            ahead = __ldg(words + min(widx, wlast));
        }
    }
    MSD_D uint32_t lo() const { return (uint32_t)bb; }
    MSD_D void drop(int n) { bb >>= n; bc -= n; }
    MSD_D uint32_t take(int n) { const uint32_t v = (uint32_t)bb & ((1u << n) - 1u); drop(n); return v; }
};

// Build the tables of lane `X` with all 32 lanes.  XM: X's region (same pointer in every lane).  Code lengths are
// read from the region first (they alias the tables).  `keep` is true in lane X only: it keeps the canonical
// limits/bases of the longer codes in its registers.  Returns a Status (uniform across the warp).
MSD_D uint32_t coop_build(const LaneMem XM, uint32_t* scratch, int hlit, int hdist, int lane, bool keep,
                          uint32_t (&llim)[7], uint32_t (&lbase)[7], uint32_t (&dlim)[8], uint32_t (&dbase)[8]) {
    constexpr unsigned full = 0xffffffffu;
    uint32_t* cnt = scratch; uint32_t* first = scratch + 16; uint32_t* offs = scratch + 32;
    uint16_t* sorted = reinterpret_cast<uint16_t*>(scratch + 48);
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
        uint32_t code = 0, idx = 0, lim8[9], offs9 = 0;
#pragma unroll
        for (int k = 1; k < 16; k++) {
            const uint32_t c = cnt[k];
            code <<= 1;
            const uint32_t f = code, o = idx;
            if (k == kLitRootBits + 1) offs9 = idx;
            code += c; idx += c;
            if (code > (1u << k)) status = kOversubscribed;
            if (lane == 0) { first[k] = f; offs[k] = o; }
            if (k <= kLitRootBits) lim8[k] = code << (kLitRootBits - k);
            else if (keep) { llim[k - 9] = code << (15 - k); lbase[k - 9] = (o - offs9 - f) & 0xffffu; }
        }
        const uint32_t n_syms = idx;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 9; j++) if (l[j]) sorted[offs[l[j]] + rank[j]] = (uint16_t)(lane + 32 * j);
        __syncwarp();
        // root table by index: 8 entries per lane
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const uint32_t i = (uint32_t)lane + 32u * t;
            const uint32_t c = __brev(i) >> 24;
            uint32_t len = 1;
#pragma unroll
            for (int k = 1; k <= kLitRootBits; k++) len += (c >= lim8[k]) ? 1u : 0u;
            uint32_t ent = 0;
            if (len <= (uint32_t)kLitRootBits) {
                const uint32_t sym = sorted[offs[len] + (c >> (kLitRootBits - len)) - first[len]];
                ent = lit_entry_of(sym) | len;
            }
            XM.st16(kLitRootHw + i, ent);
        }
        const uint32_t n_long = n_syms - offs9;
        for (uint32_t i = lane; i < n_long; i += 32) XM.st16(kLongLitHw + i, lit_entry_of(sorted[offs9 + i]));
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
        uint32_t code = 0, idx = 0, lim7[8], offs8 = 0;
#pragma unroll
        for (int k = 1; k < 16; k++) {
            const uint32_t c = cnt[k];
            code <<= 1;
            const uint32_t f = code, o = idx;
            if (k == kDistRootBits + 1) offs8 = idx;
            code += c; idx += c;
            if (code > (1u << k)) status = kOversubscribed;
            if (lane == 0) { first[k] = f; offs[k] = o; }
            if (k <= kDistRootBits) lim7[k] = code << (kDistRootBits - k);
            else if (keep) { dlim[k - 8] = code << (15 - k); dbase[k - 8] = (o - offs8 - f) & 0xffffu; }
        }
        const uint32_t n_syms = idx;
        __syncwarp();
        if (dl) sorted[offs[dl] + r] = (uint16_t)lane;
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const uint32_t i = (uint32_t)lane + 32u * t;
            const uint32_t c = __brev(i) >> 25;
            uint32_t len = 1;
#pragma unroll
            for (int k = 1; k <= kDistRootBits; k++) len += (c >= lim7[k]) ? 1u : 0u;
            uint32_t ent = 0;
            if (len <= (uint32_t)kDistRootBits) {
                const uint32_t sym = sorted[offs[len] + (c >> (kDistRootBits - len)) - first[len]];
                ent = dist_entry_of(sym) | len;
            }
            XM.st16(kDistRootHw + i, ent);
        }
        const uint32_t n_long = n_syms - offs8;
        if ((uint32_t)lane < n_long) XM.st8(kLongDistByte + lane, dist_entry_of(sorted[offs8 + lane]) >> 4);
        __syncwarp();
    }
    return status;
}

enum DecPhase : int { kPhNeed = 0, kPhHeader = 1, kPhBuild = 2, kPhSym = 3, kPhFinish = 4, kPhFail = 5, kPhDone = 6 };

// K1a.  comp: compressed chunk buffer (4-byte aligned base); out: inflated chunk buffer.  Literals go to out, matches
// to tokens[b * kTokenCap ..]; n_tokens[b] and status[b] are written for every block.
__global__ void __launch_bounds__(kDecThreads, kDecCtasPerSm)
huffman_decode_kernel(const uint8_t* __restrict__ comp, uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks, int n_blocks,
                      uint32_t* __restrict__ ticket, uint32_t* __restrict__ tokens, uint32_t* __restrict__ n_tokens,
                      uint32_t* __restrict__ status_out, uint32_t* __restrict__ error_count) {
    extern __shared__ __align__(16) uint32_t smem32[];
    constexpr unsigned full = 0xffffffffu;
    const int lane = threadIdx.x;
    const LaneMem M{smem32 + lane};
    uint32_t* const scratch = smem32 + kLaneWords * 32;

    int phase = kPhNeed, b = 0, bfinal = 0, hlit = 0, hdist = 0;
    const uint32_t* __restrict__ words = nullptr;
    uint8_t* __restrict__ dst = nullptr;
    uint32_t* __restrict__ tok = nullptr;
    uint32_t* tok_begin = nullptr;
    uint32_t wlast = 0, total_words = 0, isize = 0, lead = 0, clen = 0, outp = 0, lit_run = 0, err = 0;
    BitReader br; br.bb = 0; br.bc = 0; br.widx = 0; br.ahead = 0;
    uint32_t llim[7], lbase[7], dlim[8], dbase[8];
#pragma unroll
    for (int k = 0; k < 7; k++) { llim[k] = 0; lbase[k] = 0; }
#pragma unroll
    for (int k = 0; k < 8; k++) { dlim[k] = 0; dbase[k] = 0; }

    for (;;) {
        // ---- blocks that ended during the last burst; new tickets --------------------------------------------------
        if (phase == kPhFinish) {
            if (outp != isize) { err = kShort; phase = kPhFail; }
            else { n_tokens[b] = (uint32_t)(tok - tok_begin); status_out[b] = kOk; phase = kPhNeed; }
        }
        if (phase == kPhFail) {
            n_tokens[b] = 0; status_out[b] = err ? err : (uint32_t)kBadSymbol; atomicAdd(error_count, 1u); phase = kPhNeed;
        }
        if (phase == kPhNeed) {
            b = (int)atomicAdd(ticket, 1u);
            if (b >= n_blocks) phase = kPhDone;
            else {
                const BlockDesc bd = blocks[b];
                lead = (uint32_t)(bd.src & 3); clen = bd.clen; isize = bd.isize;
                words = reinterpret_cast<const uint32_t*>(comp + (bd.src - lead));
                total_words = max((lead + clen + 3) / 4, 1u); wlast = total_words - 1;
                dst = out + bd.dst;
                tok_begin = tokens + (size_t)b * kTokenCap; tok = tok_begin;
                outp = 0; lit_run = 0; err = 0; bfinal = 0;
                br.seed(words, wlast, lead);
                phase = kPhHeader;
            }
        }
        if (__all_sync(full, phase == kPhDone)) break;

        // ---- DEFLATE block headers (lanes in this phase run together) ------------------------------------------------------
        if (phase == kPhHeader) {
            br.refill(words, wlast);
            if (br.widx > total_words + 4u) err = kTruncated;
            bfinal = (int)br.take(1);
            const int btype = (int)br.take(2);
            if (err) {
            } else if (btype == 0) {
                br.drop(br.bc & 7);
                br.refill(words, wlast);
                const uint32_t len = br.take(16), nlen = br.take(16);
                const uint32_t sp = br.widx * 4u - (uint32_t)br.bc / 8u;   // byte index in the aligned stream
                if ((len ^ 0xffffu) != nlen || sp + len > lead + clen) err = kBadStored;
                else if (outp + len > isize) err = kOverflow;
                else {
                    const uint8_t* sb = reinterpret_cast<const uint8_t*>(words) + sp;
                    for (uint32_t k = 0; k < len; k++) dst[outp + k] = sb[k];
                    outp += len; lit_run += len;
                    br.seed(words, wlast, sp + len);
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
                for (int i = 0; i < hclen; i++) { br.refill(words, wlast); clp |= (uint64_t)br.take(3) << (3 * c_cl_order[i]); }
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
                    for (int j = 0; j < 32; j++) M.stw(kClTabByte / 4 + j, 0u);
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
                        br.refill(words, wlast);
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
                const uint32_t st = coop_build(LaneMem{smem32 + X}, scratch, hl, hd, lane, lane == X, llim, lbase, dlim, dbase);
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
                    br.refill(words, wlast);
                    uint32_t lo = br.lo();
                    uint32_t e = M.ld16(kLitRootHw + (lo & 255u));
                    if ((e & 15u) == 0) {
                        // code longer than the root: canonical search over the per-length limits (registers)
                        const uint32_t c = __brev(lo) >> 17;
                        uint32_t n = 0, base = lbase[0];
#pragma unroll
                        for (int k = 0; k < 6; k++) if (c >= llim[k]) { n = k + 1; base = lbase[k + 1]; }
                        if (c >= llim[6]) { err = kBadSymbol; e = (kNibBad << 4) | 1u; }
                        else { const uint32_t l = 9 + n; e = M.ld16(kLongLitHw + ((base + (c >> (15 - l))) & 0xffffu)) | l; }
                    }
                    br.drop((int)(e & 15u));
                    const uint32_t nib = (e >> 4) & 15u;
                    if (nib & 8u) {
                        if (outp < isize) { dst[outp] = (uint8_t)(e >> 8); outp++; lit_run++; }
                        else err = kOverflow;
                    } else if (nib >= kNibEob) {
                        if (nib == kNibEob) phase = bfinal ? kPhFinish : kPhHeader;
                        else if (!err) err = kBadSymbol;
                    } else {
                        const uint32_t len = 3u + ((e >> 8) & 0xffu) + br.take((int)nib);
                        br.refill(words, wlast);
                        lo = br.lo();
                        uint32_t d = M.ld16(kDistRootHw + (lo & 127u));
                        if ((d & 15u) == 0) {
                            const uint32_t c = __brev(lo) >> 17;
                            uint32_t n = 0, base = dbase[0];
#pragma unroll
                            for (int k = 0; k < 7; k++) if (c >= dlim[k]) { n = k + 1; base = dbase[k + 1]; }
                            if (c >= dlim[7]) d = (kDistNibBad << 4) | 1u;
                            else { const uint32_t l = 8 + n; d = (M.ld8(kLongDistByte + ((base + (c >> (15 - l))) & 0xffffu)) << 4) | l; }
                        }
                        br.drop((int)(d & 15u));
                        const uint32_t x = (d >> 4) & 15u;
                        const uint32_t dist = 1u + (((d >> 8) & 3u) << x) + br.take((int)x);
                        if (x == kDistNibBad || dist > outp) err = kBadDistance;
                        else if (outp + len > isize) err = kOverflow;
                        else {
                            if (lit_run > 510u) { *tok++ = (kTokSkip << 23) | lit_run; lit_run = 0; }
                            *tok++ = make_token(lit_run, len, dist);
                            lit_run = 0; outp += len;
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

// K1b.  One warp per BGZF block: copies every match of the block's token list; literals are already in place.
constexpr int kResolveWarps = 4;
constexpr int kResolveCtasPerSm = 12;

__global__ void __launch_bounds__(kResolveWarps * 32, kResolveCtasPerSm)
lz77_resolve_kernel(uint8_t* out, const BlockDesc* __restrict__ blocks, int n_blocks, uint32_t* __restrict__ ticket,
                    const uint32_t* __restrict__ tokens, const uint32_t* __restrict__ n_tokens) {
    constexpr unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t le = 0xffffffffu >> (31 - lane);   // lanes <= this one
    for (;;) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(ticket, 1u);
        b = __shfl_sync(full, b, 0);
        if (b >= n_blocks) return;
        uint8_t* dst = out + blocks[b].dst;
        const int mis = (int)((uintptr_t)dst & 31u);
        const uint32_t* __restrict__ tok = tokens + (size_t)b * kTokenCap;
        const uint32_t nt = n_tokens[b];
        int pos_base = 0;
        for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
            const uint32_t t = t0 + lane;
            int lr = 0, len = 0, dist = 0, adv = 0;
            if (t < nt) {
                const uint32_t w = __ldg(tok + t);
                lr = (int)(w >> 23);
                if (lr == (int)kTokSkip) { lr = (int)(w & 0x7fffffu); adv = lr; }
                else { len = (int)(w & 0xffu) + 3; dist = (int)((w >> 8) & 0x7fffu) + 1; adv = lr + len; }
            }
            int incl = adv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(full, incl, o); if (lane >= o) incl += y; }
            const int total = __shfl_sync(full, incl, 31);
            const int start = pos_base + incl - adv;        // first output byte of this token (its literals)
            const int mstart = start + lr;                  // first byte of its match
            const int hi = pos_base + total;
            int before = 0;                                  // tokens of the batch that start before the window
            for (int wb = ((pos_base + mis) & ~31) - mis; wb < hi; wb += 32) {
                const uint32_t my_head = (adv > 0 && start >= wb && start < wb + 32) ? (1u << (start - wb)) : 0u;
                const uint32_t heads = __reduce_or_sync(full, my_head);
                const int p = wb + lane;
                const int m = before + __popc(heads & le) - 1;
                before += __popc(heads);
                const int ms = __shfl_sync(full, mstart, m & 31);
                const int md = __shfl_sync(full, dist, m & 31);
                bool pend = p >= pos_base && p < hi && m >= 0 && md > 0 && p >= ms;
                int src = p - md;
                if (pend && p - ms >= md) src = ms - md + (p - ms) % md;   // dist < len: fold onto the first period
                uint32_t done = __ballot_sync(full, !pend);
                while (done != full) {
                    const bool ready = pend && (src < wb || ((done >> (src - wb)) & 1u));
                    if (ready) { dst[p] = dst[src]; pend = false; }
                    __syncwarp();
                    done = __ballot_sync(full, !pend);
                }
            }
            pos_base = hi;
        }
    }
}

// K1 = K1a + K1b on one stream.  tickets: two zeroed words (K1a, K1b).
inline void launch_inflate(cudaStream_t s, const uint8_t* d_comp, uint8_t* d_out, const BlockDesc* d_desc, int nb, uint32_t* tickets,
                           uint32_t* d_tokens, uint32_t* d_ntok, uint32_t* d_status, uint32_t* d_error_count) {
    const int dctas = (nb + kDecThreads - 1) / kDecThreads < kSMs * kDecCtasPerSm ? (nb + kDecThreads - 1) / kDecThreads : kSMs * kDecCtasPerSm;
    huffman_decode_kernel<<<dctas, kDecThreads, kDecSmem, s>>>(d_comp, d_out, d_desc, nb, tickets, d_tokens, d_ntok, d_status, d_error_count);
    const int want = (nb + kResolveWarps - 1) / kResolveWarps;
    const int rctas = want < kSMs * kResolveCtasPerSm ? want : kSMs * kResolveCtasPerSm;
    lz77_resolve_kernel<<<rctas, kResolveWarps * 32, 0, s>>>(d_out, d_desc, nb, tickets + 1, d_tokens, d_ntok);
}

}  // namespace msd
