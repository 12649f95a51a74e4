// inflate.cuh -- K1 bgzf_inflate: raw DEFLATE (RFC 1951) decoder for BGZF blocks.
//
// Replaces htslib's bgzf inflate (+ its thread pool, `open(bam, threads=)`, mosdepth.nim:888,968) [ext].
// Each BGZF block is an independent DEFLATE stream (the LZ77 window never crosses blocks), so the only
// parallelism is across blocks: a 30X WGS BAM has ~3.3 M of them.
//
// Design (r1c; the r1a profile showed one-decoder-per-warp to be issue bound with 1 active lane in 75 % of the
// instructions, profiles/README.md):
//   * a warp is four independent "octets" of 8 lanes; each octet decodes its own BGZF block, so every instruction of
//     the serial Huffman chain advances up to four symbols (lanes 0, 8, 16, 24 run four chains in SIMT lock step);
//   * octets pull block indices from an atomic ticket (blocks differ in cost) and never wait for each other
//     except at one warp-wide vote per round;
//   * 2.5 KB of shared memory per octet: 9-bit lit/len and 7-bit distance root tables with 16-bit entries (length /
//     distance bases are computed arithmetically), canonical-order symbol arrays + left-aligned limits for the ~5 % of
//     symbols with longer codes, a 64-word ring of compressed words refilled by the octet with 32-byte loads, and a
//     32-entry match queue; the code lengths of a dynamic header are parked inside the not-yet-built lit/len table;
//   * the octet leader stores literals straight to global memory while decoding; matches are queued and then
//     copied by the 8 lanes of the octet in stream order (the output block is the LZ77 window: sources hit L1/L2);
//   * 2 warps (8 octets) per CTA, 11 CTAs per SM -> 88 concurrent decoders per SM, 13024 per B200 (the kernel is
//     latency bound on the serial chains, so decoders per SM is what buys throughput).
#pragma once
#include "common.cuh"

namespace msd {

constexpr int kLitBits = 8;
constexpr int kDistBits = 6;
constexpr int kInflateWarps = 2;         // warps per CTA
constexpr int kOctetsPerWarp = 4;
constexpr int kRingWords = 64;           // compressed-word ring per octet (power of two)
constexpr int kRoundSymbols = 32;        // lit/len symbols decoded per round (bounds the ring words a round can eat: 48)
constexpr int kQueue = 32;               // matches queued per round

// 16-bit table entry: [3:0] code length (0 = invalid), [5:4] kind, [13:6] value
//   lit/len table: kind 0 literal (value = byte), 1 length (value = symbol - 257), 2 end of block, 3 longer code / invalid
//   distance table: kind 0 distance (value = symbol), 3 longer code / invalid
constexpr uint32_t kKindLit = 0u, kKindLen = 1u, kKindEob = 2u, kKindLong = 3u;

enum InflateStatus : uint32_t {
    kInfOk = 0, kInfBadBlockType = 1, kInfBadStored = 2, kInfBadCodeLengths = 3, kInfBadSymbol = 4, kInfBadDistance = 5,
    kInfOverflow = 6, kInfShort = 7, kInfOversubscribed = 8
};

struct __align__(16) InflateOctetState {
    uint16_t lit_tab[1 << kLitBits];     // 1024 B; while a dynamic header is read: [0,256) code-length-code table, [256,576) code lengths
    uint16_t dist_tab[1 << kDistBits];   //  256 B
    uint32_t ring[kRingWords];           //  256 B
    uint32_t q_sym[kQueue];              //  128 B  match queue: [8:0] length, [30:16] distance - 1   (also table-build scratch)
    uint16_t q_pos[kQueue];              //   64 B  output offset of the match inside the block        (also table-build scratch)
    uint16_t lit_sorted[288];            //  576 B  symbols in canonical order (longer codes / table fill)
    uint16_t dist_sorted[32];            //   64 B
    uint16_t lit_limit[16], lit_base[16];    // 64 B  canonical decode of codes longer than the root: left-aligned limits, index bases
    uint16_t dist_limit[16], dist_base[16];  // 64 B
};
static_assert(sizeof(InflateOctetState) == 1856, "8 octets x 1856 B + 1 KB reserve per CTA -> 14 CTAs (112 decoders) per SM");
constexpr int kInflateCtasPerSm = 14;

__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// RFC 1951 length code (symbol - 257): base and extra bits, computed instead of looked up
MSD_D void len_base_extra(uint32_t idx, uint32_t& base, uint32_t& extra) {
    if (idx < 8) { base = 3 + idx; extra = 0; }
    else if (idx == 28) { base = 258; extra = 0; }
    else { extra = (idx - 4) >> 2; base = 3 + ((4 + (idx & 3)) << extra); }
}
MSD_D void dist_base_extra(uint32_t sym, uint32_t& base, uint32_t& extra) {
    if (sym < 4) { base = 1 + sym; extra = 0; }
    else { extra = (sym >> 1) - 1; base = 1 + ((2 + (sym & 1)) << extra); }
}

MSD_D uint32_t litlen_entry(uint32_t sym, uint32_t len) {
    if (sym < 256) return (sym << 6) | (kKindLit << 4) | len;
    if (sym == 256) return (kKindEob << 4) | len;
    if (sym > 285) return 0;   // 286, 287 never appear in valid data
    return ((sym - 257) << 6) | (kKindLen << 4) | len;
}
MSD_D uint32_t dist_entry(uint32_t sym, uint32_t len) {
    if (sym > 29) return 0;
    return (sym << 6) | len;
}

// leader-lane bit reader over the octet's ring
struct BitReader {
    uint64_t bb;
    int bc;
    uint32_t wpos;   // next ring word (absolute word index in the block's aligned stream)
    MSD_D void refill(const uint32_t* ring) {
        if (bc <= 32) { bb |= (uint64_t)ring[wpos & (kRingWords - 1)] << bc; bc += 32; wpos++; }
    }
    MSD_D uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    MSD_D void drop(int n) { bb >>= n; bc -= n; }
    MSD_D uint32_t take(int n) { uint32_t v = peek(n); drop(n); return v; }
};

// Table construction, pass 1 (octet leader, serial: ~300 iterations per DEFLATE block, ~1 % of its decode time):
// canonical order of the coded symbols, per-length end indices, and the limit/base arrays of the canonical decoder.
// cnt: 16 words of scratch.  Returns 0 or an InflateStatus.
MSD_D uint32_t table_pass1(const uint8_t* lens, int n, uint16_t* sorted, uint16_t* limit, uint16_t* base, uint32_t* end, uint32_t* cnt) {
    uint32_t status = 0;
    for (int l = 0; l < 16; l++) cnt[l] = 0;
    for (int s = 0; s < n; s++) cnt[lens[s]]++;
    cnt[0] = 0;
    int left = 1;
    uint32_t code = 0, idx = 0;
    end[0] = 0; limit[0] = 0; base[0] = 0;
    for (int l = 1; l < 16; l++) {
        left <<= 1; left -= (int)cnt[l]; if (left < 0) status = kInfOversubscribed;
        code = (code + cnt[l - 1]) << 1;
        base[l] = (uint16_t)(idx - code);                              // sorted index = base + code (mod 2^16)
        limit[l] = (uint16_t)((code + cnt[l]) << (15 - l));            // exclusive, left aligned to 15 bits (0x8000 = complete)
        end[l] = idx; idx += cnt[l];
    }
    for (int s = 0; s < n; s++) { const int l = lens[s]; if (l) sorted[end[l]++] = (uint16_t)s; }   // end[l] now ends length l
    return status;
}

// pass 2 (8 lanes): fill the root table from the canonical order; codes longer than the root keep the "long" marker
template <bool kIsDist>
MSD_D void table_pass2(uint16_t* tab, int bits, const uint16_t* sorted, const uint16_t* base, const uint32_t* end, int gl) {
    for (int i = gl; i < (1 << bits); i += 8) tab[i] = (uint16_t)(kKindLong << 4);
    uint32_t start = 0;
    for (int l = 1; l <= bits; l++) {
        const uint32_t e = end[l];
        for (uint32_t i = start + gl; i < e; i += 8) {
            const uint32_t sym = sorted[i];
            const uint32_t code = (i - base[l]) & 0xffffu;
            const uint32_t rev = __brev(code) >> (32 - l);   // DEFLATE packs Huffman codes MSB-first into an LSB-first stream
            uint16_t ent = (uint16_t)(kIsDist ? dist_entry(sym, (uint32_t)l) : litlen_entry(sym, (uint32_t)l));
            if (!ent) ent = (uint16_t)((kKindLong << 4) | 15u);   // symbol that must never occur: flagged when met
            for (uint32_t k = rev; k < (1u << bits); k += (1u << l)) tab[k] = ent;
        }
        start = e;
    }
}

// canonical decode of a code longer than the root table (leader only): compare the next 15 bits, MSB first, with the
// left-aligned per-length limits.  Returns the symbol (0xffff: no such code) and consumes its bits.
MSD_D uint32_t slow_decode(BitReader& br, int root_bits, const uint16_t* limit, const uint16_t* base, const uint16_t* sorted) {
    const uint32_t c = __brev((uint32_t)br.bb) >> 17;
    for (int l = root_bits + 1; l < 16; l++) {
        if (c < limit[l]) { br.drop(l); return sorted[(base[l] + (c >> (15 - l))) & 0xffffu]; }
    }
    return 0xffffu;
}

enum OctetPhase : int { kPhIdle = 0, kPhHeader = 1, kPhLengths = 2, kPhSymbols = 3, kPhStored = 4, kPhFinished = 5 };

// The kernel.  comp: compressed chunk buffer; out: inflated chunk buffer; status[b] receives InflateStatus.
__global__ void __launch_bounds__(kInflateWarps * 32, kInflateCtasPerSm)
bgzf_inflate_kernel(const uint8_t* __restrict__ comp, uint8_t* __restrict__ out, const BlockDesc* __restrict__ blocks, int n_blocks,
                    uint32_t* __restrict__ ticket, uint32_t* __restrict__ status_out, uint32_t* __restrict__ error_count) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & 7;                       // lane inside the octet
    const int leader_lane = lane & 24;             // warp lane id of this octet's leader
    const unsigned gmask = 0xffu << leader_lane;
    const bool leader = gl == 0;
    InflateOctetState& S = reinterpret_cast<InflateOctetState*>(smem_raw)[warp * kOctetsPerWarp + (lane >> 3)];
    uint8_t* const lens = reinterpret_cast<uint8_t*>(S.lit_tab) + 256;   // code lengths live inside lit_tab until the tables are built

    // per-octet state; the decoder state proper lives in the leader's registers
    int phase = kPhIdle;
    int b = 0;
    const uint32_t* __restrict__ src_words = nullptr;
    uint8_t* __restrict__ dst = nullptr;
    uint32_t total_words = 0, isize = 0, lead = 0, clen = 0, loaded = 0;
    BitReader br; br.bb = 0; br.bc = 0; br.wpos = 0;
    uint32_t outp = 0, err = 0;
    int bfinal = 0, hlit = 0, hdist = 0, cl_idx = 0;
    uint32_t stored_pos = 0, stored_len = 0, guard = 0;
    bool first_fill = false;

    for (;;) {
        // ---- octets that finished a block take the next ticket ----------------------------------------------------
        if (phase == kPhIdle) {
            int nb = 0;
            if (leader) nb = (int)atomicAdd(ticket, 1u);
            b = __shfl_sync(gmask, nb, leader_lane);
            if (b >= n_blocks) phase = kPhFinished;
            else {
                const BlockDesc bd = blocks[b];
                lead = (uint32_t)(bd.src & 3); clen = bd.clen;
                src_words = reinterpret_cast<const uint32_t*>(comp + (bd.src - lead));
                total_words = (lead + bd.clen + 3) / 4;
                dst = out + bd.dst; isize = bd.isize;
                loaded = 0; br.bb = 0; br.bc = 0; br.wpos = 0; outp = 0; err = 0; guard = 0; bfinal = 0;
                first_fill = true;
                phase = kPhHeader;
            }
        }
        if (__all_sync(0xffffffffu, phase == kPhFinished)) break;   // also re-converges the four octets once per round
        if (phase == kPhFinished) continue;

        // ---- refill the ring (8 lanes): keep > 56 words ahead of the reader ---------------------------------------
        {
            const uint32_t consumed = __shfl_sync(gmask, br.wpos, leader_lane);
            while (loaded < total_words && loaded + 8 - consumed <= (uint32_t)kRingWords) {
                const uint32_t w = loaded + gl;
                S.ring[w & (kRingWords - 1)] = (w < total_words) ? __ldg(src_words + w) : 0u;
                loaded += 8;
            }
            __syncwarp(gmask);
            if (leader && first_fill) { br.refill(S.ring); br.drop((int)lead * 8); br.refill(S.ring); }
            first_fill = false;
        }
        if (++guard > 400000u && leader) err = kInfShort;
        {
            const uint32_t e_any = __shfl_sync(gmask, err, leader_lane);
            if (e_any) { if (leader) { status_out[b] = err; atomicAdd(error_count, 1u); } phase = kPhIdle; continue; }
        }

        if (phase == kPhSymbols) {
            // ---- leader: up to kRoundSymbols symbols through the Huffman chain; literals go straight to global -----
            int n = 0;          // matches queued
            int next_phase = kPhSymbols;
            const uint32_t round_start_outp = outp;
            if (leader) {
                for (int it = 0; it < kRoundSymbols; it++) {
                    br.refill(S.ring);
                    uint32_t e = S.lit_tab[br.peek(kLitBits)];
                    uint32_t kind = (e >> 4) & 3u;
                    if (kind == kKindLit) {
                        if (outp >= isize) { err = kInfOverflow; break; }
                        br.drop((int)(e & 15u));
                        dst[outp++] = (uint8_t)(e >> 6);
                        continue;
                    }
                    if (kind == kKindLong) {
                        if (e & 15u) { err = kInfBadSymbol; break; }
                        uint32_t sym = slow_decode(br, kLitBits, S.lit_limit, S.lit_base, S.lit_sorted);
                        e = sym == 0xffffu ? 0u : litlen_entry(sym, 1);
                        if (!e) { err = kInfBadSymbol; break; }
                        kind = (e >> 4) & 3u;
                        if (kind == kKindLit) {
                            if (outp >= isize) { err = kInfOverflow; break; }
                            dst[outp++] = (uint8_t)(e >> 6);
                            continue;
                        }
                    } else {
                        br.drop((int)(e & 15u));
                    }
                    if (kind == kKindEob) { next_phase = bfinal ? kPhIdle : kPhHeader; break; }
                    // length + distance
                    uint32_t lbase, lextra; len_base_extra((e >> 6) & 0xffu, lbase, lextra);
                    const uint32_t len = lbase + br.take((int)lextra);
                    br.refill(S.ring);
                    uint32_t d = S.dist_tab[br.peek(kDistBits)];
                    uint32_t dsym;
                    if (((d >> 4) & 3u) == kKindLong) {
                        if (d & 15u) { err = kInfBadDistance; break; }
                        dsym = slow_decode(br, kDistBits, S.dist_limit, S.dist_base, S.dist_sorted);
                        if (dsym > 29u) { err = kInfBadDistance; break; }
                        br.refill(S.ring);
                    } else {
                        br.drop((int)(d & 15u));
                        dsym = d >> 6;
                    }
                    uint32_t dbase, dextra; dist_base_extra(dsym, dbase, dextra);
                    const uint32_t dist = dbase + br.take((int)dextra);
                    if (dist > outp) { err = kInfBadDistance; break; }
                    if (outp + len > isize) { err = kInfOverflow; break; }
                    // 13k blocks are in flight per GPU, so their 64 KB windows (850 MB) overflow the 126 MB L2: start pulling
                    // the match source towards L2 now; the copy runs after the rest of the round has been decoded
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(dst + outp - dist));
                    S.q_sym[n] = ((dist - 1u) << 16) | len; S.q_pos[n] = (uint16_t)outp; n++;
                    outp += len;
                }
                if (err) n = 0;
            }
            n = __shfl_sync(gmask, n, leader_lane);
            next_phase = __shfl_sync(gmask, next_phase, leader_lane);
            const uint32_t round_start = __shfl_sync(gmask, round_start_outp, leader_lane);
            __syncwarp(gmask);   // literal stores and the queue are visible to the octet
            // ---- octet: copy the queued matches ---------------------------------------------------------------------
            // Pass A: matches whose source lies entirely before this round's output are independent of everything
            // decoded in this round; four of them are loaded before any is stored, so their L2 round trips overlap.
            // Pass B: the others (short distances into this round's own output) go one by one in stream order.
            uint32_t dep_mask = 0;
            for (int i0 = 0; i0 < n; i0 += 4) {
                uint32_t v[4]; uint32_t dstoff[4]; bool act[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int i = i0 + j;
                    act[j] = false;
                    if (i < n) {
                        const uint32_t ms = S.q_sym[i], mp = S.q_pos[i];
                        const uint32_t len = ms & 0x1ffu, dist = (ms >> 16) + 1u;
                        const uint32_t span = len < dist ? len : dist;
                        if (mp - dist + span <= round_start && len <= 8) {
                            act[j] = gl < len; dstoff[j] = mp + gl;
                            if (act[j]) v[j] = dst[mp - dist + (dist >= len ? gl : gl % dist)];
                        } else dep_mask |= 1u << i;
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) if (act[j]) dst[dstoff[j]] = (uint8_t)v[j];
            }
            __syncwarp(gmask);
            while (dep_mask) {
                const int i = __ffs((int)dep_mask) - 1; dep_mask &= dep_mask - 1;
                const uint32_t ms = S.q_sym[i], mp = S.q_pos[i];
                const uint32_t len = ms & 0x1ffu, dist = (ms >> 16) + 1u;
                const uint8_t* from = dst + mp - dist;
                if (dist >= len) {
                    // up to 32 bytes per step: four loads in flight per lane
                    for (uint32_t k0 = 0; k0 < len; k0 += 32) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) { const uint32_t k = k0 + 8 * j + gl; if (k < len) w[j] = from[k]; }
#pragma unroll
                        for (int j = 0; j < 4; j++) { const uint32_t k = k0 + 8 * j + gl; if (k < len) dst[mp + k] = (uint8_t)w[j]; }
                    }
                } else {
                    for (uint32_t k = gl; k < len; k += 8) dst[mp + k] = from[k % dist];
                }
                __syncwarp(gmask);
            }
            if (next_phase == kPhIdle) {   // final block done
                if (leader) { const uint32_t st = (outp != isize) ? (uint32_t)kInfShort : 0u; status_out[b] = st; if (st) atomicAdd(error_count, 1u); }
            }
            phase = next_phase;
        } else if (phase == kPhHeader) {
            // ---- DEFLATE block header ---------------------------------------------------------------------------------
            int btype = 0;
            if (leader) {
                br.refill(S.ring);
                bfinal = (int)br.take(1);
                btype = (int)br.take(2);
                if (btype == 0) {
                    br.drop(br.bc & 7);
                    br.refill(S.ring);
                    const uint32_t len = br.take(16), nlen = br.take(16);
                    if ((len ^ 0xffffu) != nlen) err = kInfBadStored;
                    stored_len = len;
                    stored_pos = br.wpos * 4u - (uint32_t)br.bc / 8u;   // byte index in the aligned stream
                    if (outp + len > isize) err = kInfOverflow;
                    if (stored_pos + len > lead + clen) err = kInfBadStored;
                } else if (btype == 2) {
                    hlit = (int)br.take(5) + 257; hdist = (int)br.take(5) + 1;
                    const int hclen = (int)br.take(4) + 4;
                    if (hlit > 286 || hdist > 30) err = kInfBadCodeLengths;
                    for (int i = 0; i < 19; i++) lens[i] = 0;
                    for (int i = 0; i < hclen; i++) { br.refill(S.ring); lens[c_cl_order[i]] = (uint8_t)br.take(3); }
                    // code-length code table (7-bit root, 19 symbols) in the first 256 B of lit_tab
                    uint16_t* cl_tab = S.lit_tab;
                    uint32_t* cnt = S.q_sym;          // [8]
                    uint32_t* next = S.q_sym + 8;     // [8]
                    for (int i = 0; i < 8; i++) cnt[i] = 0;
                    for (int i = 0; i < 19; i++) cnt[lens[i]]++;
                    cnt[0] = 0;
                    int left = 1;
                    for (int l = 1; l < 8; l++) { left <<= 1; left -= (int)cnt[l]; }
                    if (left < 0) err = kInfOversubscribed;
                    uint32_t code = 0;
                    for (int l = 1; l < 8; l++) { code = (code + cnt[l - 1]) << 1; next[l] = code; }
                    for (int i = 0; i < 128; i++) cl_tab[i] = 0;
                    for (int s = 0; s < 19; s++) {
                        const uint32_t l = lens[s];
                        if (!l) continue;
                        const uint32_t c = next[l]++;
                        const uint32_t rev = __brev(c) >> (32 - l);
                        for (uint32_t k = rev; k < 128; k += (1u << l)) cl_tab[k] = (uint16_t)((s << 4) | l);
                    }
                    cl_idx = 0;
                } else if (btype == 1) {
                    hlit = 288; hdist = 30;
                } else {
                    err = kInfBadBlockType;
                }
            }
            btype = __shfl_sync(gmask, btype, leader_lane);
            __syncwarp(gmask);
            if (btype == 0) phase = kPhStored;
            else if (btype == 2) phase = kPhLengths;
            else if (btype == 1) {
                // fixed Huffman code lengths, then the same table builder
                for (int i = gl; i < 288; i += 8) lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                for (int i = gl; i < 30; i += 8) lens[288 + i] = 5;
                __syncwarp(gmask);
                uint32_t e12 = 0;
                if (leader) {
                    e12 = table_pass1(lens, 288, S.lit_sorted, S.lit_limit, S.lit_base, S.q_sym, reinterpret_cast<uint32_t*>(S.q_pos));
                    e12 |= table_pass1(lens + 288, 30, S.dist_sorted, S.dist_limit, S.dist_base, S.q_sym + 16, reinterpret_cast<uint32_t*>(S.q_pos));
                    if (e12) err = e12;
                }
                __syncwarp(gmask);
                table_pass2<false>(S.lit_tab, kLitBits, S.lit_sorted, S.lit_base, S.q_sym, gl);
                table_pass2<true>(S.dist_tab, kDistBits, S.dist_sorted, S.dist_base, S.q_sym + 16, gl);
                __syncwarp(gmask);
                phase = kPhSymbols;
            }
        } else if (phase == kPhLengths) {
            // ---- dynamic block: read up to 100 code lengths per round (<= 175 ring bytes) ---------------------------------
            int done = 0;
            if (leader) {
                const uint16_t* cl_tab = S.lit_tab;
                const int total = hlit + hdist;
                int budget = 100;
                while (cl_idx < total && budget-- > 0 && !err) {
                    br.refill(S.ring);
                    const uint32_t e = cl_tab[br.peek(7)];
                    const uint32_t l = e & 15u, sym = e >> 4;
                    if (!l) { err = kInfBadCodeLengths; break; }
                    br.drop((int)l);
                    if (sym < 16) { lens[cl_idx++] = (uint8_t)sym; continue; }
                    uint32_t rep, val = 0;
                    if (sym == 16) { if (cl_idx == 0) { err = kInfBadCodeLengths; break; } val = lens[cl_idx - 1]; rep = 3 + br.take(2); }
                    else if (sym == 17) rep = 3 + br.take(3);
                    else rep = 11 + br.take(7);
                    if (cl_idx + (int)rep > total) { err = kInfBadCodeLengths; break; }
                    while (rep--) lens[cl_idx++] = (uint8_t)val;
                }
                done = (cl_idx >= total) && !err;
                if (done && lens[256] == 0) { err = kInfBadCodeLengths; done = 0; }
            }
            done = __shfl_sync(gmask, done, leader_lane);
            __syncwarp(gmask);
            if (done) {
                uint32_t e12 = 0;
                if (leader) {
                    e12 = table_pass1(lens, hlit, S.lit_sorted, S.lit_limit, S.lit_base, S.q_sym, reinterpret_cast<uint32_t*>(S.q_pos));
                    e12 |= table_pass1(lens + hlit, hdist, S.dist_sorted, S.dist_limit, S.dist_base, S.q_sym + 16, reinterpret_cast<uint32_t*>(S.q_pos));
                    if (e12) err = e12;
                }
                __syncwarp(gmask);   // code lengths are dead from here on: pass 2 overwrites them with the lit/len table
                table_pass2<false>(S.lit_tab, kLitBits, S.lit_sorted, S.lit_base, S.q_sym, gl);
                table_pass2<true>(S.dist_tab, kDistBits, S.dist_sorted, S.dist_base, S.q_sym + 16, gl);
                __syncwarp(gmask);
                phase = kPhSymbols;
            }
        } else if (phase == kPhStored) {
            // ---- stored block: octet byte copy, then re-seed the bit reader ------------------------------------------------
            const uint32_t sp = __shfl_sync(gmask, stored_pos, leader_lane), sl = __shfl_sync(gmask, stored_len, leader_lane);
            const uint32_t op = __shfl_sync(gmask, outp, leader_lane);
            const int bf = __shfl_sync(gmask, bfinal, leader_lane);
            const uint8_t* sb = reinterpret_cast<const uint8_t*>(src_words) + sp;
            for (uint32_t k = gl; k < sl; k += 8) dst[op + k] = sb[k];
            const uint32_t np = sp + sl;
            loaded = np / 4;
            if (leader) { outp += sl; br.bb = 0; br.bc = 0; br.wpos = np / 4; }
            __syncwarp(gmask);
            while (loaded < total_words && loaded + 8 - np / 4 <= (uint32_t)kRingWords) {
                const uint32_t w = loaded + gl;
                S.ring[w & (kRingWords - 1)] = (w < total_words) ? __ldg(src_words + w) : 0u;
                loaded += 8;
            }
            __syncwarp(gmask);
            if (leader) { br.refill(S.ring); br.drop((int)(np & 3u) * 8); br.refill(S.ring); }
            if (bf) {
                if (leader) { const uint32_t st = (outp != isize) ? (uint32_t)kInfShort : 0u; status_out[b] = st; if (st) atomicAdd(error_count, 1u); }
                phase = kPhIdle;
            } else phase = kPhHeader;
        }
    }
}

constexpr int kInflateThreads = kInflateWarps * 32;
constexpr size_t kInflateSmem = sizeof(InflateOctetState) * kInflateWarps * kOctetsPerWarp;
constexpr int kInflateOctetsPerCta = kInflateWarps * kOctetsPerWarp;

}  // namespace msd
