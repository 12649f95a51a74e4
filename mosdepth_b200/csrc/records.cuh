// records.cuh -- K3/K5 filter_emit + cigar_deltas, K4 mate_join.
//
// K3/K5 replace the body of coverage()'s record loop (mosdepth.nim:272-356): the read filters (:276-287), the
// fast-mode (:342-344) and fragment-mode (:346-353) events, and inc_coverage(gen_start_ends()) (:146-172).
// One thread per framed record; events are int32 atomicAdd into the contig's slice of the depth arena (the input
// is coordinate sorted, so a warp's atomics fall in a few cache lines).
//
// K4 replaces the `seen` table logic (mosdepth.nim:262,292-340).  The reference's semantics are sequential
// per read name (insert / overwrite / take in file order), so the kernel groups the records of one chunk by
// (tid, qname) in a device hash table and lets one thread replay each group in file order.  A mate that was
// inserted but not yet taken when the chunk ends is carried to the next chunk as a "live" entry (its qname and
// CIGAR copied to a small arena), so chunk boundaries are invisible.
#pragma once
#include "common.cuh"
#include "frame.cuh"

namespace msd {

struct FilterParams {
    int32_t mapq;
    int64_t min_len, max_len;
    uint32_t eflag, iflag;
    int32_t mode;           // MSD_MODE_*
    int32_t n_rg;           // read groups (0 = no RG filter)
    const char* rg_names;   // device: n_rg NUL-terminated strings back to back
    int32_t n_targets;
};

struct ContigTable {
    const uint64_t* depth_off;   // [n_targets] offset (in int32 cells) of the contig's array in the arena; ~0ull = not selected
    const uint32_t* tlen;        // [n_targets]
    // intra-contig sharding (msd_set_contig_range): this context OWNS the records with range_lo <= pos < range_hi.
    // Records outside still go through the filters and the mate join (an owned record may have to take its mate from
    // them), but emit nothing: every +1/-1 and every overlap correction is emitted by the owner of the record that is
    // being processed when the reference would emit it.  Unsharded contigs: [0, 0xffffffff).
    const uint32_t* range_lo;    // [n_targets]
    const uint32_t* range_hi;    // [n_targets]
    uint32_t* max_end;           // [n_targets] largest end position of an owned record (how far the depth spills past range_hi)
    unsigned long long* n_beyond;   // [n_targets] records with pos >= range_hi seen: proves that the span covered every owned record
    // soundness of the span START of a shard with range_lo > 0 (default mode): min_need = smallest mate position of an owned
    // record that looked for its mate in the `seen` table and did not find it.  The caller promises that its spans hold every
    // record from `covered_from` on (msd_set_contig_cover); min_need < covered_from means that a mate may sit in front of the
    // span: msd_run refuses the result (MSD_ESPAN) instead of silently skipping an overlap correction.  nullptr: not sharded.
    uint32_t* min_need;          // [n_targets]
    MSD_D bool owned(int32_t tid, int32_t pos) const { return (uint32_t)pos >= range_lo[tid] && (uint32_t)pos < range_hi[tid]; }
};

// only when some contig is sharded (max_end != nullptr): one atomicMax per warp and contig
MSD_D void note_end(const ContigTable& ct, int32_t tid, int64_t e, uint32_t tlen) {
    if (!ct.max_end) return;
    const uint32_t v = (uint32_t)(e < 0 ? 0 : e < (int64_t)tlen ? e : (int64_t)tlen);
    const unsigned m = __match_any_sync(__activemask(), tid);
    const uint32_t mx = __reduce_max_sync(m, v);
    if ((int)(__ffs(m) - 1) == (int)(threadIdx.x & 31)) atomicMax(ct.max_end + tid, mx);
}

constexpr uint8_t kClsNone = 0, kClsU = 1, kClsE = 2, kClsT = 3;

struct RecCore {
    uint32_t block_size; int32_t tid, pos; uint32_t l_qname, mapq, n_cigar, flag; int32_t l_seq, mtid, mpos, tlen;
};

MSD_D RecCore load_core(const uint8_t* __restrict__ buf, uint32_t p) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(buf + (p & ~3u));
    const uint32_t sh = (p & 3u) * 8u;
    uint32_t a[10];
#pragma unroll
    for (int i = 0; i < 10; i++) a[i] = w[i];
    uint32_t v[9];
#pragma unroll
    for (int i = 0; i < 9; i++) v[i] = __funnelshift_r(a[i], a[i + 1], sh);
    RecCore r;
    r.block_size = v[0]; r.tid = (int32_t)v[1]; r.pos = (int32_t)v[2];
    r.l_qname = v[3] & 0xffu; r.mapq = (v[3] >> 8) & 0xffu;
    r.n_cigar = v[4] & 0xffffu; r.flag = v[4] >> 16;
    r.l_seq = (int32_t)v[5]; r.mtid = (int32_t)v[6]; r.mpos = (int32_t)v[7]; r.tlen = (int32_t)v[8];
    return r;
}

// Where a record's CIGAR really is.  A CIGAR with more than 65,535 operations is stored as the placeholder
// <l_seq>S<reflen>N with the real operations in a CG:B,I tag; htslib's bam_read1 swaps it back in before hts-nim sees
// the record [ext: sam.c bam_tag2cigar; SURVEY Q13].  Same conditions: n_cigar > 0, tid >= 0, pos >= 0, first op = S of
// length l_seq, the first CG tag has type B and subtype I/i, its count is >= n_cigar and < 2^29.  Only records whose
// first operation soft-clips the whole read pay for the aux walk.
struct CigarRef { uint32_t off, n; };
MSD_D CigarRef rec_cigar(const uint8_t* __restrict__ buf, uint32_t p, const RecCore& r) {
    CigarRef c{p + 36 + r.l_qname, r.n_cigar};
    if (r.n_cigar == 0 || r.tid < 0 || r.pos < 0) return c;
    const uint32_t c0 = ld_u32(buf, c.off);
    if ((c0 & 15u) != 4u || (c0 >> 4) != (uint32_t)r.l_seq) return c;
    uint32_t a = c.off + 4 * r.n_cigar + (uint32_t)((r.l_seq + 1) / 2) + (uint32_t)r.l_seq;
    const uint32_t end = p + 4 + r.block_size;
    while (a + 3 <= end) {
        const uint8_t t0 = buf[a], t1 = buf[a + 1], ty = buf[a + 2];
        a += 3;
        const bool is_cg = (t0 == 'C' && t1 == 'G');
        switch (ty) {
            case 'A': case 'c': case 'C': a += 1; break;
            case 's': case 'S': a += 2; break;
            case 'i': case 'I': case 'f': a += 4; break;
            case 'd': a += 8; break;
            case 'Z': case 'H': while (a < end && buf[a]) a++; a++; break;
            case 'B': {
                if (a + 5 > end) return c;
                const uint8_t st = buf[a]; const uint32_t cnt = ld_u32(buf, a + 1);
                const uint32_t es = (st == 'c' || st == 'C') ? 1u : (st == 's' || st == 'S') ? 2u : 4u;
                if (is_cg) {
                    if ((st == 'I' || st == 'i') && cnt >= r.n_cigar && cnt < (1u << 29) && (uint64_t)a + 5 + 4ull * cnt <= end) { c.off = a + 5; c.n = cnt; }
                    return c;
                }
                if ((uint64_t)a + 5 + (uint64_t)es * cnt > end) return c;
                a += 5 + es * cnt; break;
            }
            default: return c;
        }
        if (is_cg) return c;       // the first CG tag is not an array: CIGAR untouched
    }
    return c;
}

// htslib bam_endpos: pos + reference length (unmapped: 0), a length of 0 counts as 1 -> hts-nim Record.stop
MSD_D int64_t rec_endpos(const uint8_t* __restrict__ buf, uint32_t cig, uint32_t n_cigar, uint32_t flag, int32_t pos) {
    int64_t rlen = 0;
    if (!(flag & kFlagUnmap))
        for (uint32_t i = 0; i < n_cigar; i++) { uint32_t c = ld_u32(buf, cig + 4 * i); if (cigar_type(c) & 2) rlen += c >> 4; }
    if (rlen == 0) rlen = 1;
    return (int64_t)pos + rlen;
}

MSD_D void depth_add(int32_t* __restrict__ arr, uint32_t tlen, int64_t pos, int32_t v, uint32_t& n_events) {
    // arr has tlen+1 cells (mosdepth.nim:274); the reference indexes unchecked, events outside only come from
    // invalid BAMs and are dropped (same rule in the oracle; DESIGN.md "Divergences")
    if (pos < 0 || pos > (int64_t)tlen) return;
    atomicAdd(arr + pos, v);
    n_events++;
}

// iterator over the aligned blocks of a CIGAR = the (+1, -1) pairs of gen_start_ends (mosdepth.nim:146-168)
struct BlockIter {
    const uint8_t* buf; uint32_t cig; uint32_t n, i; int64_t pos; int64_t cs, ce; bool have;
    MSD_D void init(const uint8_t* b, uint32_t cig_off, uint32_t n_cigar, int64_t start) { buf = b; cig = cig_off; n = n_cigar; i = 0; pos = start; have = false; cs = ce = 0; }
    MSD_D bool next(int64_t& s, int64_t& e) {
        while (i < n) {
            uint32_t c = ld_u32(buf, cig + 4 * i); i++;
            int t = cigar_type(c);
            if (!(t & 2)) continue;
            int64_t len = c >> 4;
            if (t & 1) {
                if (have && pos == ce) { ce = pos + len; pos += len; continue; }
                if (have) { s = cs; e = ce; cs = pos; ce = pos + len; pos += len; return true; }
                have = true; cs = pos; ce = pos + len;
            }
            pos += len;
        }
        if (have) { s = cs; e = ce; have = false; return true; }
        return false;
    }
};

MSD_D uint64_t hash_qname(const uint8_t* __restrict__ buf, uint32_t off, uint32_t len, int32_t tid) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ ((uint64_t)(uint32_t)tid * 0xD6E8FEB86659FD93ull);
    for (uint32_t i = 0; i < len; i++) { h ^= buf[off + i]; h *= 0x100000001B3ull; h ^= h >> 29; }
    h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 31;
    return h;
}

MSD_D bool rg_pass(const uint8_t* __restrict__ buf, uint32_t p, const RecCore& r, const FilterParams& fp) {
    // tag[string](rec, "RG") then read_groups.contains (mosdepth.nim:282-285)
    uint32_t a = p + 36 + r.l_qname + 4 * r.n_cigar + (uint32_t)((r.l_seq + 1) / 2) + (uint32_t)r.l_seq;
    const uint32_t end = p + 4 + r.block_size;
    while (a + 3 <= end) {
        uint8_t t0 = buf[a], t1 = buf[a + 1], ty = buf[a + 2];
        a += 3;
        bool is_rg = (t0 == 'R' && t1 == 'G');
        switch (ty) {
            case 'A': case 'c': case 'C': a += 1; break;
            case 's': case 'S': a += 2; break;
            case 'i': case 'I': case 'f': a += 4; break;
            case 'd': a += 8; break;
            case 'Z': case 'H': {
                if (is_rg && ty == 'Z') {
                    const char* names = fp.rg_names;
                    for (int g = 0; g < fp.n_rg; g++) {
                        uint32_t k = 0; bool eq = true;
                        while (true) {
                            char c = names[k]; uint8_t d = (a + k < end) ? buf[a + k] : 0;
                            if ((uint8_t)c != d) { eq = false; }
                            if (c == 0 || d == 0) break;
                            k++;
                        }
                        if (eq) return true;
                        while (names[0]) names++;
                        names++;
                    }
                    return false;
                }
                while (a < end && buf[a]) a++;
                a++;
                break;
            }
            case 'B': {
                if (a + 5 > end) return false;
                uint8_t st = buf[a]; uint32_t cnt = ld_u32(buf, a + 1);
                uint32_t es = (st == 'c' || st == 'C') ? 1u : (st == 's' || st == 'S') ? 2u : 4u;
                if ((uint64_t)a + 5 + (uint64_t)es * cnt > end) return false;
                a += 5 + es * cnt; break;
            }
            default: return false;
        }
        if (is_rg) return false;   // RG present but not a string
    }
    return false;
}

// per-chunk side arrays for the mate join
struct MateChunk {
    uint8_t* cls;        // [rec] kCls*
    uint64_t* qhash;     // [rec]
    int32_t* endpos;     // [rec]
    uint32_t* next;      // [rec] chunk-list link
    uint32_t* slot;      // [rec] table slot or kNoOff
};

struct RunCounters {
    unsigned long long n_events;
    unsigned long long n_records;
    uint32_t first_tid_valid; int32_t first_tid;
    uint32_t error;   // 1 table full, 2 live list full, 3 arena full, 4 record references another contig's array
    uint32_t pad;
};

// K5 for long reads: one CIGAR walked by the whole warp, 32 operations per step.  A thread that walks a 10-100 kb nanopore read alone
// runs thousands of dependent iterations while the few thousand reads of a chunk leave most of the GPU idle (r2f: 4.6 G events/s on C5, default
// mode); here lane i takes operation i0 + i: a warp scan of the reference-consuming lengths gives every operation its start, and the events are
// exactly those of gen_start_ends (mosdepth.nim:146-168) -- aligned operations that abut are one block: an operation emits its +1 unless it starts
// where the previous aligned one ended, and its -1 unless the next aligned one starts there (across steps the last end is carried).
// fast = true: only the reference length is needed (rec_endpos).  All 32 lanes call it with the same arguments; returns the end of the last
// block (fast: pos + max(reference length, 1)) in every lane.
constexpr uint32_t kCoopOps = 48;      // CIGARs at least this long are walked by the warp
MSD_D int64_t coop_cigar(const uint8_t* __restrict__ buf, uint32_t cig, uint32_t n, int64_t pos0, bool fast, int32_t* __restrict__ arr, uint32_t tlen, uint32_t& n_events) {
    constexpr unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int64_t pos = pos0, carry_end = pos0; bool have = false;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + (uint32_t)lane;
        const bool valid = i < n;
        const uint32_t op = valid ? ld_u32(buf, cig + 4 * i) : 0u;
        const int t = valid ? cigar_type(op) : 0;
        const int64_t len = (int64_t)(op >> 4);
        unsigned long long incl = (t & 2) ? (unsigned long long)len : 0ull;
        const unsigned long long mine = incl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(full, incl, o); if (lane >= o) incl += y; }
        const int64_t s = pos + (int64_t)(incl - mine), e = s + len;
        if (!fast) {
            const bool A = (t & 3) == 3;
            const unsigned mask = __ballot_sync(full, A);
            const unsigned below = mask & ((1u << lane) - 1u), above = lane == 31 ? 0u : (mask & ~((2u << lane) - 1u));
            const int64_t e_prev = __shfl_sync(full, e, below ? 31 - __clz((int)below) : 0);
            const int64_t s_next = __shfl_sync(full, s, above ? __ffs((int)above) - 1 : 0);
            if (A) {
                const bool has_prev = below ? true : have;
                const int64_t pe = below ? e_prev : carry_end;
                if (!has_prev || s != pe) {
                    if (!below && have) depth_add(arr, tlen, carry_end, -1, n_events);      // the block carried over from the last step ends here
                    depth_add(arr, tlen, s, 1, n_events);
                }
                if (above && s_next != e) depth_add(arr, tlen, e, -1, n_events);               // (the last aligned operation of the step is carried)
            }
            if (mask) { carry_end = __shfl_sync(full, e, 31 - __clz((int)mask)); have = true; }
        }
        pos += (int64_t)__shfl_sync(full, incl, 31);
    }
    if (fast) { const int64_t rlen = pos - pos0; return pos0 + (rlen ? rlen : 1); }
    if (have && lane == 0) depth_add(arr, tlen, carry_end, -1, n_events);
    return have ? carry_end : pos0;
}

__global__ void __launch_bounds__(256)
records_kernel(const uint8_t* __restrict__ buf, const uint32_t* __restrict__ rec_off, const ChunkState* __restrict__ cs,
               FilterParams fp, ContigTable ct, int32_t* __restrict__ arena, unsigned long long* __restrict__ n_prefilter,
               MateChunk mc, RunCounters* __restrict__ rc) {
    const uint32_t n_rec = cs->n_rec;
    uint32_t n_events = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_rec) atomicAdd(&rc->n_records, (unsigned long long)n_rec);
    // every lane of a warp makes the same number of trips (a lane beyond n_rec idles): the warp meets after each record for the long CIGARs
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_rec; base += gridDim.x * blockDim.x) {
        const uint32_t r = base + (threadIdx.x & 31u);
        // what this lane hands to the warp: a long CIGAR (coop_n != 0), the read's start and contig, and whether only its end is wanted (fast mode)
        uint32_t coop_cig = 0, coop_n = 0; int32_t coop_pos = 0, coop_tid = 0;
        if (r < n_rec) do {
        const uint32_t p = rec_off[r];
        const RecCore c = load_core(buf, p);
        if (fp.mode == MSD_MODE_DEFAULT) mc.cls[r] = kClsNone;
        if (r == 0) { rc->first_tid = c.tid; rc->first_tid_valid = 1; }
        if (c.tid < 0 || c.tid >= fp.n_targets) continue;                   // unplaced tail: never returned by a tid query
        const uint64_t doff = ct.depth_off[c.tid];
        if (doff == ~0ull) continue;                                          // contig not selected
        const uint32_t tlen = ct.tlen[c.tid];
        if (c.pos < 0 || (uint32_t)c.pos >= tlen) continue;                  // iterator overlap rule: pos < end
        const bool own = ct.owned(c.tid, c.pos);
        // `found` is set before the filters (mosdepth.nim:272-275): count every record the query would yield
        {
            unsigned m = __match_any_sync(__activemask(), (c.tid << 1) | (own ? 1 : 0));
            if ((int)(__ffs(m) - 1) == (int)(threadIdx.x & 31)) {
                if (own) atomicAdd(n_prefilter + c.tid, (unsigned long long)__popc(m));
                else if ((uint32_t)c.pos >= ct.range_hi[c.tid]) atomicAdd(ct.n_beyond + c.tid, (unsigned long long)__popc(m));
            }
        }
        if ((int32_t)c.mapq < fp.mapq) continue;                             // :276
        const int64_t ais = c.tlen < 0 ? -(int64_t)c.tlen : (int64_t)c.tlen;
        if (ais < fp.min_len || ais > fp.max_len) continue;                  // :277
        if ((c.flag & fp.eflag) != 0) continue;                              // :278-279
        if (fp.iflag != 0 && (c.flag & fp.iflag) == 0) continue;             // :280-281
        if (fp.n_rg > 0 && !rg_pass(buf, p, c, fp)) continue;                // :282-285
        int32_t* __restrict__ arr = arena + doff;
        const CigarRef cg = rec_cigar(buf, p, c);
        const uint32_t cig = cg.off;
        if (fp.mode == MSD_MODE_FAST) {                                      // :342-344
            if (!own) continue;
            if (cg.n >= kCoopOps && !(c.flag & kFlagUnmap)) { coop_cig = cig; coop_n = cg.n; coop_pos = c.pos; coop_tid = c.tid; continue; }
            const int64_t stop = rec_endpos(buf, cig, cg.n, c.flag, c.pos);
            depth_add(arr, tlen, c.pos, 1, n_events);
            depth_add(arr, tlen, stop, -1, n_events);
            note_end(ct, c.tid, stop, tlen);
        } else if (fp.mode == MSD_MODE_FRAGMENT) {                           // :346-353
            if (!own) continue;
            if ((c.flag & kFlagRead2) || !(c.flag & kFlagProper) || (c.flag & kFlagSupp)) continue;
            int64_t fs = c.pos < c.mpos ? c.pos : c.mpos;
            depth_add(arr, tlen, fs, 1, n_events);
            depth_add(arr, tlen, fs + ais, -1, n_events);
            note_end(ct, c.tid, fs + ais, tlen);
        } else {
            // classify for the mate join (:292-298) and emit the read's own blocks (:355-356)
            if ((c.flag & kFlagProper) && !(c.flag & kFlagSupp)) {
                int64_t stop = rec_endpos(buf, cig, cg.n, c.flag, c.pos);
                uint8_t cl = kClsT;
                if (c.tid == c.mtid && stop > (int64_t)c.mpos) { if (c.pos < c.mpos) cl = kClsU; else if (c.pos == c.mpos) cl = kClsE; }
                mc.cls[r] = cl;
                mc.endpos[r] = (int32_t)stop;
                mc.qhash[r] = hash_qname(buf, p + 36, c.l_qname ? c.l_qname - 1 : 0, c.tid);
                mc.slot[r] = kNoOff;
            }
            if (own) {
                if (cg.n >= kCoopOps) { coop_cig = cig; coop_n = cg.n; coop_pos = c.pos; coop_tid = c.tid; continue; }
                BlockIter it; it.init(buf, cig, cg.n, c.pos);
                int64_t s, e, last_e = c.pos;
                while (it.next(s, e)) { depth_add(arr, tlen, s, 1, n_events); depth_add(arr, tlen, e, -1, n_events); last_e = e; }
                note_end(ct, c.tid, last_e, tlen);
            }
        }
        } while (0);
        __syncwarp();
        for (unsigned todo = __ballot_sync(0xffffffffu, coop_n != 0); todo; todo &= todo - 1) {
            const int X = __ffs((int)todo) - 1;
            const uint32_t xc = __shfl_sync(0xffffffffu, coop_cig, X), xn = __shfl_sync(0xffffffffu, coop_n, X);
            const int32_t xp = __shfl_sync(0xffffffffu, coop_pos, X), xt = __shfl_sync(0xffffffffu, coop_tid, X);
            const uint32_t xl = ct.tlen[xt];
            const bool fast = fp.mode == MSD_MODE_FAST;
            const int64_t last = coop_cigar(buf, xc, xn, xp, fast, arena + ct.depth_off[xt], xl, n_events);
            if ((int)(threadIdx.x & 31) == X) {
                if (fast) { depth_add(arena + ct.depth_off[xt], xl, xp, 1, n_events); depth_add(arena + ct.depth_off[xt], xl, last, -1, n_events); }
                note_end(ct, xt, last, xl);
            }
        }
    }
    n_events = __reduce_add_sync(0xffffffffu, n_events);
    if ((threadIdx.x & 31) == 0 && n_events) atomicAdd(&rc->n_events, (unsigned long long)n_events);
}

// ------------------------------------------------------------------------------------------------------------
// mate join
// ------------------------------------------------------------------------------------------------------------
struct LiveEntry {
    uint64_t key;
    int32_t tid, start, stop;
    uint32_t n_cigar, l_qname;     // l_qname without the NUL
    uint32_t qname_off, cigar_off; // byte offsets into the arena (cigar_off 4-byte aligned)
    uint32_t state;                // 0 live, 1 dead (taken / overwritten)
};

struct MateTable {
    unsigned long long* word;   // (tag32 << 32) | ref ; empty = all ones.  ref: bit31 set -> live entry index, else chunk record
    uint32_t* head;             // chunk list head, kNoOff = empty
    uint32_t* live;             // index of the live entry that is this key's `seen` value, kNoOff = none
    uint32_t mask;
};

struct LiveList { LiveEntry* e; uint8_t* arena; uint32_t* n; uint32_t* arena_used; uint32_t cap, arena_cap; };

constexpr uint32_t kRefLive = 0x80000000u;

struct NameRef { const uint8_t* p; uint32_t len; int32_t tid; };
MSD_D NameRef name_of(uint32_t ref, const uint8_t* __restrict__ buf, const uint32_t* __restrict__ rec_off, const LiveList& L) {
    NameRef n;
    if (ref & kRefLive) { const LiveEntry& e = L.e[ref & ~kRefLive]; n.p = L.arena + e.qname_off; n.len = e.l_qname; n.tid = e.tid; }
    else { uint32_t p = rec_off[ref]; uint32_t lq = buf[p + kRecLName]; n.p = buf + p + 36; n.len = lq ? lq - 1 : 0; n.tid = (int32_t)ld_u32(buf, p + kRecRefID); }
    return n;
}
MSD_D bool names_equal(const NameRef& a, const NameRef& b) {
    if (a.tid != b.tid || a.len != b.len) return false;
    for (uint32_t i = 0; i < a.len; i++) if (a.p[i] != b.p[i]) return false;
    return true;
}

// find the slot of (tid, qname); insert it when `insert`.  Returns slot or kNoOff (not found / table full -> *full = true)
MSD_D uint32_t table_find(const MateTable& T, uint64_t key, uint32_t myref, bool insert, const uint8_t* __restrict__ buf,
                          const uint32_t* __restrict__ rec_off, const LiveList& L, bool& full) {
    const uint32_t tag = (uint32_t)(key >> 32);
    uint32_t i = (uint32_t)key & T.mask;
    const NameRef me = name_of(myref, buf, rec_off, L);
    for (uint32_t probes = 0; probes <= T.mask; probes++, i = (i + 1) & T.mask) {
        unsigned long long w = *((volatile unsigned long long*)&T.word[i]);
        if (w == ~0ull) {
            if (!insert) return kNoOff;
            unsigned long long mine = ((unsigned long long)tag << 32) | myref;
            unsigned long long prev = atomicCAS(&T.word[i], ~0ull, mine);
            if (prev == ~0ull) return i;
            w = prev;
        }
        if ((uint32_t)(w >> 32) == tag) {
            NameRef other = name_of((uint32_t)w, buf, rec_off, L);
            if (names_equal(me, other)) return i;
        }
    }
    full = true;
    return kNoOff;
}

// M0: put the live entries carried from the previous chunk back into the (cleared) table
__global__ void mate_reinsert_kernel(MateTable T, LiveList L, const uint8_t* __restrict__ buf, const uint32_t* __restrict__ rec_off,
                                     RunCounters* __restrict__ rc) {
    const uint32_t n = *L.n;
    const int32_t first_tid = rc->first_tid_valid ? rc->first_tid : INT32_MIN;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        LiveEntry& e = L.e[i];
        if (e.state) continue;
        if (first_tid >= 0 && e.tid < first_tid) { e.state = 1; continue; }   // its contig is finished: `seen` is per contig (:262)
        bool full = false;
        uint32_t s = table_find(T, e.key, kRefLive | i, true, buf, rec_off, L, full);
        if (full) { atomicExch(&rc->error, 1u); continue; }
        T.live[s] = i;
    }
}

// M1 (inserters: classes U and E) and M2 (takers: class T) -- two launches of the same kernel
__global__ void mate_link_kernel(MateTable T, LiveList L, const uint8_t* __restrict__ buf, const uint32_t* __restrict__ rec_off,
                                 const ChunkState* __restrict__ cs, MateChunk mc, int takers, RunCounters* __restrict__ rc, ContigTable ct) {
    const uint32_t n_rec = cs->n_rec;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        const uint8_t cl = mc.cls[r];
        if (cl == kClsNone) continue;
        if (takers != (cl == kClsT)) continue;
        bool full = false;
        uint32_t s = table_find(T, mc.qhash[r], r, !takers, buf, rec_off, L, full);
        if (full) { atomicExch(&rc->error, 1u); continue; }
        if (s == kNoOff) {
            if (takers && ct.min_need) {
                // a shard that starts inside its contig: this read found no mate in `seen`.  Fine when the mate does not reach
                // it (it was never inserted, :297) -- unless the mate lies in front of the span, where nobody looked
                const uint32_t p = rec_off[r];
                const int32_t tid = (int32_t)ld_u32(buf, p + kRecRefID), pos = (int32_t)ld_u32(buf, p + kRecPos), mpos = (int32_t)ld_u32(buf, p + kRecNextPos);
                if (ct.range_lo[tid] && ct.owned(tid, pos) && mpos >= 0 && mpos < pos) atomicMin(ct.min_need + tid, (uint32_t)mpos);
            }
            continue;
        }
        mc.slot[r] = s;
        mc.next[r] = atomicExch(&T.head[s], r);
    }
}

// subtract the intersection of the aligned blocks of two reads (the sweep of mosdepth.nim:323-338)
MSD_D void subtract_overlap(BlockIter& a, BlockIter& b, int32_t* __restrict__ arr, uint32_t tlen, uint32_t& n_events) {
    int64_t as, ae, bs, be;
    bool ha = a.next(as, ae), hb = b.next(bs, be);
    while (ha && hb) {
        int64_t lo = as > bs ? as : bs, hi = ae < be ? ae : be;
        if (lo < hi) { depth_add(arr, tlen, lo, -1, n_events); depth_add(arr, tlen, hi, 1, n_events); }
        if (ae < be) ha = a.next(as, ae); else hb = b.next(bs, be);
    }
}

// M3: one thread per name group (the group's first record in file order) replays insert / take in file order
__global__ void mate_replay_kernel(MateTable T, LiveList L, LiveList Lnew, const uint8_t* __restrict__ buf,
                                   const uint32_t* __restrict__ rec_off, const ChunkState* __restrict__ cs, MateChunk mc,
                                   ContigTable ct, int32_t* __restrict__ arena, RunCounters* __restrict__ rc) {
    const uint32_t n_rec = cs->n_rec;
    uint32_t n_events = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        if (mc.cls[r] == kClsNone) continue;
        const uint32_t s = mc.slot[r];
        if (s == kNoOff) continue;
        uint32_t mn = r;
        for (uint32_t x = T.head[s]; x != kNoOff; x = mc.next[x]) mn = x < mn ? x : mn;
        if (mn != r) continue;
        // state of seen[qname]: 0 none, 1 live entry, 2 chunk record
        int kind = 0; uint32_t sidx = 0;
        if (T.live[s] != kNoOff) { kind = 1; sidx = T.live[s]; }
        int64_t last = -1;
        for (;;) {
            uint32_t x = kNoOff;
            for (uint32_t y = T.head[s]; y != kNoOff; y = mc.next[y]) if ((int64_t)y > last && y < x) x = y;
            if (x == kNoOff) break;
            last = x;
            const uint8_t cl = mc.cls[x];
            if (cl == kClsU || (cl == kClsE && kind == 0)) {                  // seen[qname] = rec.copy()  (:299-300)
                if (kind == 1) L.e[sidx].state = 1;
                kind = 2; sidx = x;
            } else if (kind != 0) {                                           // seen.take(qname, mate)  (:302)
                const uint32_t px = rec_off[x];
                const RecCore cx = load_core(buf, px);
                const uint64_t doff = ct.depth_off[cx.tid];
                int32_t* __restrict__ arr = arena + doff;
                const uint32_t tlen = ct.tlen[cx.tid];
                uint32_t m_ncig; int64_t m_start, m_stop; const uint8_t* m_buf; uint32_t m_cig;
                if (kind == 1) { const LiveEntry& e = L.e[sidx]; m_ncig = e.n_cigar; m_start = e.start; m_stop = e.stop; m_buf = L.arena; m_cig = e.cigar_off; }
                else { const uint32_t pm = rec_off[sidx]; const RecCore cm = load_core(buf, pm); const CigarRef gm = rec_cigar(buf, pm, cm); m_ncig = gm.n; m_start = cm.pos; m_stop = mc.endpos[sidx]; m_buf = buf; m_cig = gm.off; }
                const CigarRef gx = rec_cigar(buf, px, cx);
                if (!ct.owned(cx.tid, cx.pos)) {
                    // the taker belongs to another shard: that shard applies the correction
                } else if (gx.n == 1 && m_ncig == 1) {                        // :307-309
                    depth_add(arr, tlen, cx.pos, -1, n_events);
                    depth_add(arr, tlen, m_stop, 1, n_events);
                    note_end(ct, cx.tid, m_stop, tlen);
                } else {                                                      // :323-338
                    BlockIter a, b;
                    a.init(buf, gx.off, gx.n, cx.pos);
                    b.init(m_buf, m_cig, m_ncig, m_start);
                    subtract_overlap(a, b, arr, tlen, n_events);
                }
                if (kind == 1) L.e[sidx].state = 1;
                kind = 0;
            }
        }
        if (kind == 2) {
            // still waiting for its mate when the chunk ends: carry it
            const uint32_t pm = rec_off[sidx];
            const RecCore cm = load_core(buf, pm);
            const uint32_t lq = cm.l_qname ? cm.l_qname - 1 : 0;
            const CigarRef gm = rec_cigar(buf, pm, cm);
            const uint32_t bytes = ((lq + 3u) & ~3u) + 4u * gm.n;
            const uint32_t idx = atomicAdd(Lnew.n, 1u);
            const uint32_t off = atomicAdd(Lnew.arena_used, bytes);
            if (idx >= Lnew.cap) { atomicExch(&rc->error, 2u); continue; }
            if ((uint64_t)off + bytes > Lnew.arena_cap) { atomicExch(&rc->error, 3u); continue; }
            LiveEntry& e = Lnew.e[idx];
            e.key = mc.qhash[sidx]; e.tid = cm.tid; e.start = cm.pos; e.stop = mc.endpos[sidx]; e.n_cigar = gm.n; e.l_qname = lq;
            e.qname_off = off; e.cigar_off = off + ((lq + 3u) & ~3u); e.state = 0;
            for (uint32_t i = 0; i < lq; i++) Lnew.arena[off + i] = buf[pm + 36 + i];
            uint32_t* cg = reinterpret_cast<uint32_t*>(Lnew.arena + e.cigar_off);
            for (uint32_t i = 0; i < gm.n; i++) cg[i] = ld_u32(buf, gm.off + 4 * i);
        }
    }
    n_events = __reduce_add_sync(0xffffffffu, n_events);
    if ((threadIdx.x & 31) == 0 && n_events) atomicAdd(&rc->n_events, (unsigned long long)n_events);
}

// M4: live entries that nobody took or overwrote move on to the next chunk's list
__global__ void mate_carry_kernel(LiveList L, LiveList Lnew, RunCounters* __restrict__ rc) {
    const uint32_t n = *L.n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const LiveEntry& o = L.e[i];
        if (o.state) continue;
        const uint32_t bytes = ((o.l_qname + 3u) & ~3u) + 4u * o.n_cigar;
        const uint32_t idx = atomicAdd(Lnew.n, 1u);
        const uint32_t off = atomicAdd(Lnew.arena_used, bytes);
        if (idx >= Lnew.cap) { atomicExch(&rc->error, 2u); continue; }
        if ((uint64_t)off + bytes > Lnew.arena_cap) { atomicExch(&rc->error, 3u); continue; }
        LiveEntry e = o;
        e.qname_off = off; e.cigar_off = off + ((o.l_qname + 3u) & ~3u);
        for (uint32_t k = 0; k < bytes; k++) Lnew.arena[off + k] = L.arena[o.qname_off + k];
        Lnew.e[idx] = e;
    }
}

}  // namespace msd
