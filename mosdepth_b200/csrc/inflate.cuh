// inflate.cuh -- K1 bgzf_inflate: raw DEFLATE (RFC 1951) decoder, one warp per BGZF block.
//
// Replaces htslib's bgzf inflate (+ its thread pool, `open(bam, threads=)`, mosdepth.nim:888,968) [ext].
// Each BGZF block is an independent DEFLATE stream (the LZ77 window never crosses blocks), so the only
// parallelism is across blocks: a 30X WGS BAM has ~3.3 M of them.  Design:
//   * one warp decodes one block; warps pull block indices from an atomic ticket (blocks differ in cost);
//   * per warp ~7 KB of shared memory: 10-bit lit/len and 8-bit distance lookup tables (one probe per
//     symbol; longer codes fall back to a canonical bit-by-bit walk), the code-length arrays, a 512-byte ring
//     of compressed words that all lanes refill with coalesced 128-byte loads, and a 32-entry symbol queue;
//   * lane 0 runs the serial Huffman chain for up to 32 symbols, then the whole warp resolves the queue:
//     literals are stored in parallel, matches are copied cooperatively straight in global memory
//     (the output block is the LZ77 window: sources are L2 hits).
//   * 8 warps per CTA, 4 CTAs per SM -> 32 concurrent decoders per SM, 4736 per B200.
#pragma once
#include "common.cuh"

namespace msd {

constexpr int kLitBits = 10;
constexpr int kDistBits = 8;
constexpr int kInflateWarps = 8;
constexpr int kRingWords = 96;           // compressed-word ring per warp; a round never needs more than 56 words
constexpr int kQueue = 32;               // symbols decoded per cooperative round

// table entry: [3:0] code length, [7:4] extra-bit count, [9:8] kind, [31:16] value
constexpr uint32_t kKindLit = 0u << 8, kKindLen = 1u << 8, kKindEob = 2u << 8, kKindLong = 3u << 8, kKindMask = 3u << 8;
constexpr uint32_t kMatchFlag = 0x80000000u;

enum InflateStatus : uint32_t {
    kInfOk = 0, kInfBadBlockType = 1, kInfBadStored = 2, kInfBadCodeLengths = 3, kInfBadSymbol = 4, kInfBadDistance = 5,
    kInfOverflow = 6, kInfShort = 7, kInfOversubscribed = 8
};

struct __align__(16) InflateWarpState {
    uint32_t lit_tab[1 << kLitBits];     // 4096 B
    uint32_t dist_tab[1 << kDistBits];   // 1024 B
    uint32_t ring[kRingWords];           //  384 B
    uint32_t q_sym[kQueue];              //  128 B
    uint32_t q_pos[kQueue];              //  128 B
    uint16_t lit_sorted[288];            //  576 B  symbols in canonical order (slow path / table fill)
    uint16_t dist_sorted[32];            //   64 B
    uint16_t lit_count[16];              //   32 B  codes per length
    uint16_t dist_count[16];             //   32 B
    uint16_t cl_tab[128];                //  256 B  code-length-code lookup (7 bits): [3:0] len, [15:4] symbol
    uint8_t lens[320];                   //  320 B  lit/len + distance code lengths
};
static_assert(sizeof(InflateWarpState) <= 7040, "8 warps x 7040 B + 1 KB reserve = 56 KB per CTA -> 4 CTAs (32 decoders) per SM");

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

MSD_D uint32_t litlen_entry(uint32_t sym, uint32_t len) {
    if (sym < 256) return (sym << 16) | kKindLit | len;
    if (sym == 256) return kKindEob | len;
    if (sym > 285) return 0;   // 286, 287 never appear in valid data
    return ((uint32_t)c_len_base[sym - 257] << 16) | kKindLen | ((uint32_t)c_len_extra[sym - 257] << 4) | len;
}
MSD_D uint32_t dist_entry(uint32_t sym, uint32_t len) {
    if (sym > 29) return 0;
    return ((uint32_t)c_dist_base[sym] << 16) | ((uint32_t)c_dist_extra[sym] << 4) | len;
}

// lane-0 bit reader over the shared ring
struct BitReader {
    uint64_t bb;
    int bc;
    uint32_t wpos;   // next ring word (absolute word index in the block's aligned stream)
    MSD_D void refill(const uint32_t* ring) {
        if (bc <= 32) { bb |= (uint64_t)ring[wpos % kRingWords] << bc; bc += 32; wpos++; }
    }
    MSD_D uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    MSD_D void drop(int n) { bb >>= n; bc -= n; }
    MSD_D uint32_t take(int n) { uint32_t v = peek(n); drop(n); return v; }
};

// Build one decode table from code lengths.  All 32 lanes call it.  Returns 0 ok, else InflateStatus.
// lens[0..n): code lengths; tab: 1<<bits entries; sorted/count: canonical order + per-length counts;
// scratch: 64 words of shared memory (the symbol queue, idle while tables are built).
template <bool kIsDist>
MSD_D uint32_t build_table(const uint8_t* lens, int n, uint32_t* tab, int bits, uint16_t* sorted, uint16_t* count,
                           uint32_t* scratch, int lane) {
    uint32_t* first_code = scratch;        // [16]
    uint32_t* first_idx = scratch + 16;    // [16]; first_idx[0] = number of coded symbols
    uint32_t* offs = scratch + 32;         // [16]
    uint32_t status = 0;
    if (lane < 16) count[lane] = 0;
    for (int i = lane; i < (1 << bits); i += 32) tab[i] = 0;
    __syncwarp();
    if (lane == 0) {
        // serial: ~300 iterations per DEFLATE block, <1% of the block's decode time
        for (int s = 0; s < n; s++) count[lens[s]]++;
        count[0] = 0;
        int left = 1;
        uint32_t code = 0, idx = 0;
        for (int l = 1; l < 16; l++) {
            left <<= 1; left -= (int)count[l]; if (left < 0) status = kInfOversubscribed;
            code = (code + count[l - 1]) << 1;
            first_code[l] = code; first_idx[l] = idx; offs[l] = idx; idx += count[l];
        }
        first_idx[0] = idx; first_code[0] = 0;
        for (int s = 0; s < n; s++) { int l = lens[s]; if (l) sorted[offs[l]++] = (uint16_t)s; }
    }
    status = __shfl_sync(0xffffffffu, status, 0);
    __syncwarp();
    if (status) return status;
    const int total = (int)first_idx[0];
    for (int i = lane; i < total; i += 32) {
        uint32_t sym = sorted[i];
        uint32_t len = lens[sym];
        uint32_t code = first_code[len] + ((uint32_t)i - first_idx[len]);
        uint32_t rev = __brev(code) >> (32 - len);   // DEFLATE packs Huffman codes MSB-first into an LSB-first stream
        if ((int)len <= bits) {
            uint32_t e = kIsDist ? dist_entry(sym, len) : litlen_entry(sym, len);
            for (uint32_t k = rev; k < (1u << bits); k += (1u << len)) tab[k] = e;
        } else {
            tab[rev & ((1u << bits) - 1u)] = kKindLong | 15u;   // resolved by the canonical walk
        }
    }
    __syncwarp();
    return 0;
}

// canonical bit-by-bit decode (codes longer than the root table); lane 0 only.  Returns symbol or 0xffff.
MSD_D uint32_t slow_decode(BitReader& br, const uint16_t* count, const uint16_t* sorted) {
    uint32_t code = 0, first = 0, index = 0;
    for (int len = 1; len < 16; len++) {
        code |= br.take(1);
        uint32_t c = count[len];
        if (code - first < c) return sorted[index + (code - first)];
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return 0xffffu;
}

// The kernel.  comp: compressed chunk buffer; out: inflated chunk buffer; status[b] receives InflateStatus.
__global__ void __launch_bounds__(kInflateWarps * 32, 4)
bgzf_inflate_kernel(const uint8_t* __restrict__ comp, uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks, int n_blocks,
                    uint32_t* __restrict__ ticket, uint32_t* __restrict__ status_out, uint32_t* __restrict__ error_count) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    InflateWarpState& S = reinterpret_cast<InflateWarpState*>(smem_raw)[warp];

    for (;;) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(ticket, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= n_blocks) break;
        const BlockDesc bd = blocks[b];
        const uint32_t lead = (uint32_t)(bd.src & 3);
        const uint32_t* __restrict__ src_words = reinterpret_cast<const uint32_t*>(comp + (bd.src - lead));
        const uint32_t total_words = (lead + bd.clen + 3) / 4;
        uint8_t* __restrict__ dst = out + bd.dst;
        const uint32_t isize = bd.isize;

        uint32_t loaded = 0;                 // words fetched into the ring so far (uniform)
        BitReader br; br.bb = 0; br.bc = 0; br.wpos = 0;
        uint32_t outp = 0;                   // lane 0: bytes produced
        uint32_t err = 0;                    // lane 0
        int state = 0;                       // lane 0: 0 header, 1 code lengths, 2 symbols, 3 stored, 4 done
        int bfinal = 0, hlit = 0, hdist = 0, cl_idx = 0;
        uint32_t stored_pos = 0, stored_len = 0;
        bool first_fill = true;
        uint32_t guard = 0;

        for (;;) {
            // ---- refill the ring (all lanes): keep > 64 words ahead of the reader --------------------------
            uint32_t consumed = __shfl_sync(0xffffffffu, br.wpos, 0);
            while (loaded < total_words && loaded + 32 - consumed <= (uint32_t)kRingWords) {
                uint32_t w = loaded + lane;
                if (w < total_words) S.ring[w % kRingWords] = __ldg(src_words + w);
                else S.ring[w % kRingWords] = 0;
                loaded += 32;
            }
            __syncwarp();
            if (lane == 0 && first_fill) { br.refill(S.ring); br.drop((int)lead * 8); br.refill(S.ring); }
            first_fill = false;

            int st = __shfl_sync(0xffffffffu, state, 0);
            uint32_t e_any = __shfl_sync(0xffffffffu, err, 0);
            if (st == 4 || e_any) break;
            if (++guard > 200000u) { if (lane == 0) err = kInfShort; break; }

            if (st == 0) {
                // ---- DEFLATE block header -----------------------------------------------------------------
                int btype = 0;
                if (lane == 0) {
                    br.refill(S.ring);
                    bfinal = (int)br.take(1);
                    btype = (int)br.take(2);
                    if (btype == 0) {
                        br.drop(br.bc & 7);
                        br.refill(S.ring);
                        uint32_t len = br.take(16), nlen = br.take(16);
                        if ((len ^ 0xffffu) != nlen) err = kInfBadStored;
                        stored_len = len;
                        stored_pos = br.wpos * 4u - (uint32_t)br.bc / 8u;   // byte index in the aligned stream
                        if (outp + len > isize) err = kInfOverflow;
                        if (stored_pos + len > lead + bd.clen) err = kInfBadStored;
                        state = 3;
                    } else if (btype == 2) {
                        hlit = (int)br.take(5) + 257; hdist = (int)br.take(5) + 1;
                        int hclen = (int)br.take(4) + 4;
                        if (hlit > 286 || hdist > 30) err = kInfBadCodeLengths;
                        for (int i = 0; i < 19; i++) S.lens[i] = 0;
                        for (int i = 0; i < hclen; i++) { br.refill(S.ring); S.lens[c_cl_order[i]] = (uint8_t)br.take(3); }
                        // code-length code table (7-bit root, serial: 19 symbols)
                        uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, next[8];
                        for (int i = 0; i < 19; i++) cnt[S.lens[i]]++;
                        cnt[0] = 0;
                        int left = 1;
                        for (int l = 1; l < 8; l++) { left <<= 1; left -= (int)cnt[l]; }
                        if (left < 0) err = kInfOversubscribed;
                        uint32_t code = 0;
                        for (int l = 1; l < 8; l++) { code = (code + cnt[l - 1]) << 1; next[l] = code; }
                        for (int i = 0; i < 128; i++) S.cl_tab[i] = 0;
                        for (int s = 0; s < 19; s++) {
                            uint32_t l = S.lens[s];
                            if (!l) continue;
                            uint32_t c = next[l]++;
                            uint32_t rev = __brev(c) >> (32 - l);
                            for (uint32_t k = rev; k < 128; k += (1u << l)) S.cl_tab[k] = (uint16_t)((s << 4) | l);
                        }
                        cl_idx = 0;
                        state = 1;
                    } else if (btype == 1) {
                        hlit = 288; hdist = 30;
                        state = 2;
                    } else {
                        err = kInfBadBlockType;
                    }
                }
                btype = __shfl_sync(0xffffffffu, btype, 0);
                if (btype == 1) {
                    // fixed Huffman code lengths, then the same table builder
                    for (int i = lane; i < 288; i += 32) S.lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                    if (lane < 30) S.lens[288 + lane] = 5;
                    __syncwarp();
                    uint32_t e1 = build_table<false>(S.lens, 288, S.lit_tab, kLitBits, S.lit_sorted, S.lit_count, S.q_sym, lane);
                    uint32_t e2 = build_table<true>(S.lens + 288, 30, S.dist_tab, kDistBits, S.dist_sorted, S.dist_count, S.q_sym, lane);
                    if (lane == 0 && (e1 | e2)) err = e1 ? e1 : e2;
                }
                __syncwarp();
            } else if (st == 1) {
                // ---- dynamic block: read up to 128 code lengths per round -----------------------------------
                int done = 0;
                if (lane == 0) {
                    const int total = hlit + hdist;
                    int budget = 128;
                    while (cl_idx < total && budget-- > 0 && !err) {
                        br.refill(S.ring);
                        uint32_t e = S.cl_tab[br.peek(7)];
                        uint32_t l = e & 15u, sym = e >> 4;
                        if (!l) { err = kInfBadCodeLengths; break; }
                        br.drop((int)l);
                        if (sym < 16) { S.lens[cl_idx++] = (uint8_t)sym; continue; }
                        uint32_t rep, val = 0;
                        if (sym == 16) { if (cl_idx == 0) { err = kInfBadCodeLengths; break; } val = S.lens[cl_idx - 1]; rep = 3 + br.take(2); }
                        else if (sym == 17) rep = 3 + br.take(3);
                        else rep = 11 + br.take(7);
                        if (cl_idx + (int)rep > total) { err = kInfBadCodeLengths; break; }
                        while (rep--) S.lens[cl_idx++] = (uint8_t)val;
                    }
                    done = (cl_idx >= total) && !err;
                    if (done && S.lens[256] == 0) { err = kInfBadCodeLengths; done = 0; }
                }
                done = __shfl_sync(0xffffffffu, done, 0);
                __syncwarp();
                if (done) {
                    int hl = __shfl_sync(0xffffffffu, hlit, 0), hd = __shfl_sync(0xffffffffu, hdist, 0);
                    uint32_t e1 = build_table<false>(S.lens, hl, S.lit_tab, kLitBits, S.lit_sorted, S.lit_count, S.q_sym, lane);
                    uint32_t e2 = build_table<true>(S.lens + hl, hd, S.dist_tab, kDistBits, S.dist_sorted, S.dist_count, S.q_sym, lane);
                    if (lane == 0) { if (e1 | e2) err = e1 ? e1 : e2; state = 2; }
                }
                __syncwarp();
            } else if (st == 2) {
                // ---- lane 0: up to kQueue symbols through the Huffman chain ---------------------------------
                int n = 0;
                if (lane == 0) {
                    while (n < kQueue) {
                        br.refill(S.ring);
                        uint32_t e = S.lit_tab[br.peek(kLitBits)];
                        if ((e & kKindMask) == kKindLong) {
                            uint32_t sym = slow_decode(br, S.lit_count, S.lit_sorted);
                            e = sym == 0xffffu ? 0u : litlen_entry(sym, 1);
                            if (!e) { err = kInfBadSymbol; break; }
                        } else {
                            if (!(e & 15u)) { err = kInfBadSymbol; break; }
                            br.drop((int)(e & 15u));
                        }
                        uint32_t kind = e & kKindMask;
                        if (kind == kKindLit) {
                            S.q_sym[n] = e >> 16; S.q_pos[n] = outp; n++; outp += 1;
                        } else if (kind == kKindEob) {
                            state = bfinal ? 4 : 0;
                            break;
                        } else {
                            uint32_t len = (e >> 16) + br.take((int)((e >> 4) & 15u));
                            br.refill(S.ring);
                            uint32_t d = S.dist_tab[br.peek(kDistBits)];
                            if ((d & kKindMask) == kKindLong) {
                                uint32_t sym = slow_decode(br, S.dist_count, S.dist_sorted);
                                d = sym == 0xffffu ? 0u : dist_entry(sym, 1);
                                if (!d) { err = kInfBadDistance; break; }
                                br.refill(S.ring);
                            } else {
                                if (!(d & 15u)) { err = kInfBadDistance; break; }
                                br.drop((int)(d & 15u));
                            }
                            uint32_t dist = (d >> 16) + br.take((int)((d >> 4) & 15u));
                            if (dist > outp) { err = kInfBadDistance; break; }
                            S.q_sym[n] = kMatchFlag | ((dist - 1u) << 16) | len; S.q_pos[n] = outp; n++; outp += len;
                        }
                        if (outp > isize) { err = kInfOverflow; break; }
                    }
                    if (err) n = 0;
                }
                n = __shfl_sync(0xffffffffu, n, 0);
                __syncwarp();
                // ---- whole warp: resolve the queue -----------------------------------------------------------
                uint32_t s = 0, p = 0;
                if (lane < n) { s = S.q_sym[lane]; p = S.q_pos[lane]; }
                bool is_match = (lane < n) && (s & kMatchFlag);
                if (lane < n && !is_match) dst[p] = (uint8_t)s;
                uint32_t mm = __ballot_sync(0xffffffffu, is_match);
                __syncwarp();
                while (mm) {
                    int i = __ffs((int)mm) - 1; mm &= mm - 1;
                    uint32_t ms = __shfl_sync(0xffffffffu, s, i), mp = __shfl_sync(0xffffffffu, p, i);
                    uint32_t len = ms & 0x1ffu, dist = ((ms >> 16) & 0x7fffu) + 1u;
                    const uint8_t* from = dst + mp - dist;
                    if (dist >= len) {
                        for (uint32_t k = lane; k < len; k += 32) dst[mp + k] = from[k];
                    } else {
                        for (uint32_t k = lane; k < len; k += 32) dst[mp + k] = from[k % dist];
                    }
                    __syncwarp();
                }
            } else {
                // ---- stored block: cooperative byte copy, then re-seed the bit reader -------------------------
                uint32_t sp = __shfl_sync(0xffffffffu, stored_pos, 0), sl = __shfl_sync(0xffffffffu, stored_len, 0);
                uint32_t op = __shfl_sync(0xffffffffu, outp, 0);
                const uint8_t* sb = reinterpret_cast<const uint8_t*>(src_words) + sp;
                for (uint32_t k = lane; k < sl; k += 32) dst[op + k] = sb[k];
                uint32_t np = sp + sl;
                loaded = np / 4;
                if (lane == 0) {
                    outp += sl;
                    br.bb = 0; br.bc = 0; br.wpos = np / 4;
                    state = bfinal ? 4 : 0;
                }
                __syncwarp();
                // refill now so lane 0 can drop the sub-word offset
                uint32_t consumed2 = np / 4;
                while (loaded < total_words && loaded + 32 - consumed2 <= (uint32_t)kRingWords) {
                    uint32_t w = loaded + lane;
                    S.ring[w % kRingWords] = (w < total_words) ? __ldg(src_words + w) : 0u;
                    loaded += 32;
                }
                __syncwarp();
                if (lane == 0) { br.refill(S.ring); br.drop((int)(np & 3u) * 8); br.refill(S.ring); }
                __syncwarp();
            }
        }
        if (lane == 0) {
            if (!err && outp != isize) err = kInfShort;
            status_out[b] = err;
            if (err) atomicAdd(error_count, 1u);
        }
        __syncwarp();
    }
}

}  // namespace msd
