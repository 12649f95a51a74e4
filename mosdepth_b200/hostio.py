"""Host-side BAM header and BAI/CSI parsing for the Python tests and bench (the reference keeps this on the host
through hts-nim/htslib; formats in SURVEY.md Appendix B).  Pure stdlib + zlib."""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field
from pathlib import Path


@dataclass
class BamHeader:
    names: list
    lengths: list
    first_record_voff: int
    text: str = ""


def iter_bgzf_blocks(data: bytes, pos: int = 0):
    """Yield (coffset, block_len, xlen, isize) by chasing BSIZE."""
    n = len(data)
    while pos + 18 <= n:
        if data[pos:pos + 3] != b"\x1f\x8b\x08" or not (data[pos + 3] & 4):
            raise ValueError(f"not a BGZF block at {pos}")
        xlen = struct.unpack_from("<H", data, pos + 10)[0]
        q, bsize = pos + 12, None
        while q + 4 <= pos + 12 + xlen:
            si1, si2, slen = data[q], data[q + 1], struct.unpack_from("<H", data, q + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", data, q + 4)[0]
            q += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC subfield")
        blen = bsize + 1
        isize = struct.unpack_from("<I", data, pos + blen - 4)[0]
        yield pos, blen, xlen, isize
        pos += blen


def inflate_block(data: bytes, pos: int, blen: int, xlen: int) -> bytes:
    return zlib.decompress(data[pos + 12 + xlen: pos + blen - 8], -15)


def inflate_all(data: bytes) -> bytes:
    return b"".join(inflate_block(data, p, bl, xl) for p, bl, xl, _ in iter_bgzf_blocks(data))


def read_bam_header(path_or_bytes) -> BamHeader:
    """Parse the BAM header; returns target names/lengths and the virtual offset of the first record."""
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        data = bytes(path_or_bytes[: 64 << 20])
    else:
        with open(path_or_bytes, "rb") as f:
            data = f.read(64 << 20)
    buf = b""
    starts = []   # (uncompressed start, coffset, isize)
    blocks = iter_bgzf_blocks(data)

    def need(n):
        nonlocal buf
        while len(buf) < n:
            p, bl, xl, isz = next(blocks)
            starts.append((len(buf), p, isz))
            buf += inflate_block(data, p, bl, xl)

    need(12)
    if buf[:4] != b"BAM\x01":
        raise ValueError("not a BAM file")
    l_text = struct.unpack_from("<i", buf, 4)[0]
    need(12 + l_text)
    text = buf[8:8 + l_text].decode(errors="replace")
    off = 8 + l_text
    n_ref = struct.unpack_from("<i", buf, off)[0]
    off += 4
    names, lengths = [], []
    for _ in range(n_ref):
        need(off + 4)
        l_name = struct.unpack_from("<i", buf, off)[0]
        need(off + 4 + l_name + 4)
        names.append(buf[off + 4: off + 4 + l_name - 1].decode())
        lengths.append(struct.unpack_from("<i", buf, off + 4 + l_name)[0] & 0xFFFFFFFF)
        off += 8 + l_name
    # virtual offset of `off`
    voff = None
    for ustart, coff, isz in starts:
        if ustart <= off < ustart + isz:
            voff = (coff << 16) | (off - ustart)
    if voff is None:   # header ends exactly at a block boundary: first record starts the next block
        ustart, coff, isz = starts[-1]
        assert off == ustart + isz
        p, bl, _, _ = next(iter_bgzf_blocks(data, coff))
        voff = (p + bl) << 16
    return BamHeader(names, lengths, voff, text)


@dataclass
class ContigIndex:
    ref_beg: int = 0      # virtual offset of the first record of the contig
    ref_end: int = 0      # virtual offset just past its last record
    n_mapped: int = 0
    n_unmapped: int = 0
    linear: list = field(default_factory=list)   # BAI ioffset[] (16 kb windows)
    has_data: bool = False


def _maybe_gunzip(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        out, d = b"", raw
        while d:
            z = zlib.decompressobj(31)
            out += z.decompress(d)
            d = z.unused_data
        return out
    return raw


def read_index(bam_path) -> list:
    """Load <bam>.csi or <bam>.bai (or <stem>.bai): per-contig span + linear index. Raises if absent (mosdepth.nim:969-971)."""
    bam_path = str(bam_path)
    for cand, is_csi in ((bam_path + ".csi", True), (bam_path + ".bai", False), (bam_path[:-4] + ".bai", False)):
        if Path(cand).exists():
            d = _maybe_gunzip(Path(cand).read_bytes())
            break
    else:
        raise FileNotFoundError("[mosdepth] error alignment file must be indexed")
    q = 4
    if is_csi:
        assert d[:4] == b"CSI\x01"
        _min_shift, depth, l_aux = struct.unpack_from("<iii", d, 4)
        q = 16 + l_aux
        pseudo = ((1 << ((depth + 1) * 3)) - 1) // 7 + 1
    else:
        assert d[:4] == b"BAI\x01"
        pseudo = 37450
    n_ref = struct.unpack_from("<i", d, q)[0]
    q += 4
    out = []
    for _ in range(n_ref):
        ci = ContigIndex()
        n_bin = struct.unpack_from("<i", d, q)[0]
        q += 4
        mn, mx = None, 0
        for _b in range(n_bin):
            b = struct.unpack_from("<I", d, q)[0]
            q += 4
            if is_csi:
                q += 8
            n_chunk = struct.unpack_from("<i", d, q)[0]
            q += 4
            chunks = [struct.unpack_from("<QQ", d, q + 16 * c) for c in range(n_chunk)]
            q += 16 * n_chunk
            if b == pseudo:
                if n_chunk >= 2:
                    ci.n_mapped, ci.n_unmapped = chunks[1]
                continue
            for cb, ce in chunks:
                mn = cb if mn is None or cb < mn else mn
                mx = max(mx, ce)
        if not is_csi:
            n_intv = struct.unpack_from("<i", d, q)[0]
            q += 4
            ci.linear = list(struct.unpack_from(f"<{n_intv}Q", d, q))
            q += 8 * n_intv
        if mn is not None:
            ci.has_data, ci.ref_beg, ci.ref_end = True, mn, mx
        out.append(ci)
    return out
