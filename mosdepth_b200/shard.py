"""Host-side sharding of a BAM across GPUs/ranks: contigs are independent units (the reference
itself processes one contig at a time, mosdepth.nim:691), so each rank takes a set of contigs chosen by
longest-processing-time bin packing on their compressed span size, reads only those byte ranges of the file, and
only the small `total` histograms / stats are combined across ranks (sum_into, mosdepth.nim:761-764,807-813)."""
from __future__ import annotations

import os

import numpy as np


def contig_weights(index):
    """Compressed bytes spanned by each contig (from the BAI/CSI bin chunks)."""
    return [((c.ref_end >> 16) - (c.ref_beg >> 16) + 1) if c.has_data else 0 for c in index]


def lpt_shards(weights, n):
    """Longest-processing-time bin packing: returns n sorted lists of contig ids; every contig with weight > 0 is
    assigned exactly once."""
    loads = [0] * n
    assign = [[] for _ in range(n)]
    for tid in sorted(range(len(weights)), key=lambda t: (-weights[t], t)):
        if weights[tid] <= 0:
            continue
        r = min(range(n), key=lambda i: (loads[i], i))
        loads[r] += weights[tid]
        assign[r].append(tid)
    return [sorted(a) for a in assign]


def merge_adjacent_spans(begs, ends, max_gap_bytes=1 << 20):
    """Merge spans of one genuine BAM file whose compressed gap is small: the bytes between two record boundaries are
    whole records, and records of contigs that are not selected are dropped by the library's target mask."""
    out = []
    for b, e in sorted(zip(begs, ends)):
        if out and out[-1][1] != 0 and b >= out[-1][1] and (b >> 16) - (out[-1][1] >> 16) <= max_gap_bytes:
            out[-1][1] = e
        else:
            out.append([b, e])
    return [o[0] for o in out], [o[1] for o in out]


def split_contig(ci, tlen, n_parts, window=0, margin_bp=65536, beyond_windows=8, with_cover=False):
    """Intra-contig sharding: cut one contig into n_parts position ranges of roughly equal
    compressed size near BAI linear-index entries (16,384-bp windows), cut positions being multiples of
    lcm(32, window) so that no --by window straddles a cut.  Returns a list of (lo, hi, voff_beg, voff_end):
    part i owns the records with lo <= pos < hi; its span starts `margin_bp` before lo (default mode: an owned record may
    have to take its mate from there) and ends `beyond_windows` index windows after hi, so that it holds every owned
    record and at least one record at or beyond hi (the library checks that).  Fewer parts come back when the contig
    has too few possible cut positions.  with_cover: a fifth element, the position from which on the span is known to hold every
    record (msd_set_contig_cover: default mode checks the mates it could not find against it)."""
    import math
    lin = list(ci.linear)
    if n_parts <= 1 or not ci.has_data or len(lin) < 2:
        return [(0, tlen, ci.ref_beg, ci.ref_end, 0) if with_cover else (0, tlen, ci.ref_beg, ci.ref_end)]
    align = 32 if not window else 32 * window // math.gcd(32, window)   # scans are 128-bit aligned; no window straddles a cut

    def voff_at_window(w):
        if w <= 0:
            return ci.ref_beg
        if w >= len(lin):
            return ci.ref_end
        return lin[w] if lin[w] else ci.ref_beg

    beg_c, end_c = ci.ref_beg >> 16, ci.ref_end >> 16
    cuts = []
    for i in range(1, n_parts):
        target = beg_c + (end_c - beg_c) * i // n_parts
        w = next((k for k in range(len(lin)) if lin[k] and (lin[k] >> 16) >= target), len(lin))
        cut = (w * 16384 + align // 2) // align * align
        if cut <= (cuts[-1] if cuts else 0) or cut >= tlen:
            continue
        cuts.append(cut)
    bounds = [0] + cuts + [tlen]
    parts = []
    for i in range(len(bounds) - 1):
        lo, hi = bounds[i], bounds[i + 1]
        w0 = max(0, (lo - margin_bp) // 16384)
        vb = ci.ref_beg if lo == 0 else voff_at_window(w0)
        cover = w0 * 16384 if (lo > 0 and 0 < w0 < len(lin) and lin[w0]) else 0
        ve = ci.ref_end if hi >= tlen else voff_at_window(hi // 16384 + beyond_windows)
        parts.append((lo, hi, vb, ve, cover) if with_cover else (lo, hi, vb, ve))
    return parts


def offset_shards(index, lengths, n, window=0, margin_bp=65536, beyond_windows=8, with_cover=False):
    """Shard the genome into n contiguous BGZF offset ranges of equal compressed size (north_star: "index-derived BGZF
    offset ranges"): the cuts fall inside contigs, at BAI linear-index entries aligned like split_contig()'s.  Returns,
    per rank, a list of (tid, lo, hi, voff_beg, voff_end) in file order; lo == 0 and hi == tlen mark a whole contig.  A
    rank's pieces are adjacent in the file, so they merge into one span; only the (at most two) cut contigs of a rank
    need the spill exchange.  with_cover: a sixth element per piece, the position from which on the span is known to hold
    every record of the contig (the start of the linear-index window the span begins at; msd_set_contig_cover).  The same plan
    as the host CLI's (mosdepth_b200/host/main.cpp offset_shards, `--print-shards`)."""
    import math
    have = [t for t in range(len(index)) if index[t].has_data]
    if not have:
        return [[] for _ in range(n)]
    align = 32 if not window else 32 * window // math.gcd(32, window)   # scans are 128-bit aligned; no window straddles a cut
    c_beg, c_end = index[have[0]].ref_beg >> 16, index[have[-1]].ref_end >> 16
    # cut k (1 <= k < n) -> (tid, position); position 0 means "at the start of contig tid"
    cuts = []
    for k in range(1, n):
        target = c_beg + (c_end - c_beg) * k // n
        t = next((t for t in have if (index[t].ref_end >> 16) > target), have[-1])
        lin, tlen = index[t].linear, lengths[t]
        w = next((j for j in range(len(lin)) if lin[j] and (lin[j] >> 16) >= target), len(lin))
        pos = (w * 16384 + align // 2) // align * align
        if pos >= tlen:            # nearer to the end of the contig: cut at the next contig's start
            nxt = [u for u in have if u > t]
            if not nxt:
                continue
            t, pos = nxt[0], 0
        if cuts and (t, pos) <= cuts[-1]:
            continue
        cuts.append((t, pos))
    bounds = [(have[0], 0)] + cuts + [(None, 0)]
    out = []
    for r in range(len(bounds) - 1):
        (t0, p0), (t1, p1) = bounds[r], bounds[r + 1]
        pieces = []
        for t in have:
            if t < t0 or (t1 is not None and (t > t1 or (t == t1 and p1 == 0))):
                continue
            lo = p0 if t == t0 else 0
            hi = p1 if (t1 is not None and t == t1) else lengths[t]
            ci, lin = index[t], index[t].linear

            def voff_at(wi, ci=ci, lin=lin):
                if wi <= 0:
                    return ci.ref_beg
                if wi >= len(lin):
                    return ci.ref_end
                return lin[wi] if lin[wi] else ci.ref_beg

            vb, cover = ci.ref_beg, 0
            if lo > margin_bp:
                wi = min((lo - margin_bp) // 16384, len(lin) - 1)
                while wi > 0 and not lin[wi]:
                    wi -= 1
                if wi > 0:
                    vb, cover = lin[wi], wi * 16384
            ve = ci.ref_end
            if hi < lengths[t]:
                wi = hi // 16384 + beyond_windows
                while wi < len(lin) and not lin[wi]:
                    wi += 1
                if wi < len(lin):
                    ve = lin[wi]
            pieces.append((t, lo, hi, vb, ve, cover) if with_cover else (t, lo, hi, vb, ve))
        out.append(pieces)
    while len(out) < n:
        out.append([])
    return out


def exchange_spill(ctx, tid, rank_prev, rank_next, send, recv, lo, hi):
    """The carry across a shard boundary: copy out the depth this shard's reads leave beyond its hi, hand it to the next
    shard and add what the previous shard hands over (include/mosdepth_b200.h, "intra-contig sharding").
    send(dst_rank, array) / recv(src_rank) -> array move int32 numpy arrays between ranks (None: no such neighbour)."""
    start, n = ctx.contig_spill(tid)
    mine = ctx.depth_copy(tid, start, n) if (n and rank_next is not None) else np.zeros(0, dtype=np.int32)
    if rank_next is not None:
        send(rank_next, mine)
    if rank_prev is not None:
        got = recv(rank_prev)
        if got.size > hi - lo:
            raise ValueError("a shard is shorter than the reach of its predecessor's reads")
        if got.size:
            ctx.depth_add(tid, lo, got)
    ctx.contig_finalize(tid)


def _block_len(f, coff):
    f.seek(coff)
    h = f.read(18)
    assert len(h) == 18 and h[:2] == b"\x1f\x8b", "span end is not at a BGZF block"
    xlen = h[10] | (h[11] << 8)
    if xlen == 6 and h[12:14] == b"BC":
        return (h[16] | (h[17] << 8)) + 1
    f.seek(coff + 12)
    extra = f.read(xlen)
    q = 0
    while q + 4 <= xlen:
        slen = extra[q + 2] | (extra[q + 3] << 8)
        if extra[q] == 66 and extra[q + 1] == 67 and slen == 2:
            return (extra[q + 4] | (extra[q + 5] << 8)) + 1
        q += 4 + slen
    raise ValueError("BGZF block without BC subfield")


def load_compact(bam_path, begs, ends, alloc=None):
    """Read only the byte ranges that the spans [beg, end) (virtual offsets) need into one host buffer and rebase the
    virtual offsets onto it.  alloc(nbytes) -> writable uint8 numpy array (e.g. a view of pinned memory).
    Returns (buffer, new_begs, new_ends)."""
    fsize = os.path.getsize(bam_path)
    spans = sorted(zip(begs, ends))
    ranges = []   # [file_lo, file_hi): whole BGZF blocks only
    with open(bam_path, "rb") as f:
        for b, e in spans:
            lo = b >> 16
            hi = fsize if e == 0 else ((e >> 16) + _block_len(f, e >> 16) if (e & 0xFFFF) else (e >> 16))
            if ranges and lo <= ranges[-1][1]:
                ranges[-1][1] = max(ranges[-1][1], hi)
            else:
                ranges.append([lo, hi])
    total = sum(hi - lo for lo, hi in ranges)
    buf = alloc(total + 64) if alloc else np.zeros(total + 64, dtype=np.uint8)
    base, off = [], 0
    with open(bam_path, "rb") as f:
        for lo, hi in ranges:
            f.seek(lo)
            n = f.readinto(memoryview(buf)[off:off + (hi - lo)])
            assert n == hi - lo
            base.append((lo, hi, off))
            off += hi - lo
    buf[off:] = 0

    def rebase(v):
        c, u = v >> 16, v & 0xFFFF
        for lo, hi, o in base:
            if lo <= c < hi or (c == hi and u == 0):
                return ((c - lo + o) << 16) | u
        raise ValueError("virtual offset outside the loaded ranges")

    nb = [rebase(b) for b in begs]
    ne = [0 if e == 0 else rebase(e) for e in ends]
    return buf[:off + 64], nb, ne, off
