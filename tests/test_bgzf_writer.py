"""CPU test of the host CLI's BGZF writer (mosdepth_b200/host/bgzf_writer.hpp): the parallel run writer and the line-indexed
writer must (1) decompress to exactly the text of the plain path, (2) write bytes that do not depend on the thread count,
(3) be valid BGZF (BC field, BSIZE, ISIZE <= 65536, EOF block), and (4) come with a CSI index through which htslib's query
algorithm (restated below: reg2bins, loffset pruning, chunk reads by virtual offset) finds exactly the overlapping lines."""
import gzip
import random
import struct
import subprocess
import zlib

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def outputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("bgzw")
    exe = d / "bgzf_writer_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-o", str(exe), str(ROOT / "tools" / "bgzf_writer_test.cpp"), "-lz"], check=True)
    subprocess.run([str(exe), str(d / "t1"), "60000", "1"], check=True)
    subprocess.run([str(exe), str(d / "t5"), "60000", "5"], check=True)
    return d


def bgzf_blocks(data):
    pos, out = 0, []
    while pos < len(data):
        assert data[pos:pos + 4] == b"\x1f\x8b\x08\x04"
        xlen = struct.unpack_from("<H", data, pos + 10)[0]
        assert data[pos + 12:pos + 14] == b"BC" and struct.unpack_from("<H", data, pos + 14)[0] == 2
        bsize = struct.unpack_from("<H", data, pos + 16)[0] + 1
        isize = struct.unpack_from("<I", data, pos + bsize - 4)[0]
        assert isize <= 65536 and xlen == 6
        out.append((pos, bsize, isize))
        pos += bsize
    assert pos == len(data)
    return out


@pytest.mark.parametrize("name", ["runs", "quant", "regions", "members", "lines"])
def test_content_structure_and_thread_independence(outputs, name):
    gz1 = (outputs / f"t1.{name}.bed.gz").read_bytes()
    gz5 = (outputs / f"t5.{name}.bed.gz").read_bytes()
    plain = (outputs / f"t1.{'runs' if name == 'members' else name}.bed").read_bytes()
    assert gz1 == gz5
    assert (outputs / f"t1.{name}.bed.gz.csi").read_bytes() == (outputs / f"t5.{name}.bed.gz.csi").read_bytes()
    assert gzip.decompress(gz1) == plain
    blocks = bgzf_blocks(gz1)
    assert blocks[-1][2] == 0 and blocks[-1][1] == 28              # the EOF marker block
    assert sum(b[2] for b in blocks) == len(plain)
    # an indexed file never splits a line across two blocks
    for pos, bsize, isize in blocks[:50]:
        if isize:
            text = zlib.decompress(gz1[pos + 18:pos + bsize - 8], -15)
            assert text.endswith(b"\n")


# ---- CSI reader + htslib's query (hts.c: hts_itr_query / reg2bins), restated ------------------------------------------
def parse_csi(raw):
    assert raw[:4] == b"CSI\x01"
    min_shift, depth, l_aux = struct.unpack_from("<iii", raw, 4)
    aux = raw[16:16 + l_aux]
    fmt, col_seq, col_beg, col_end, meta, skip, l_nm = struct.unpack_from("<7i", aux, 0)
    names = aux[28:28 + l_nm].split(b"\0")[:-1]
    p = 16 + l_aux
    n_ref = struct.unpack_from("<i", raw, p)[0]; p += 4
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", raw, p)[0]; p += 4
        bins = {}
        for _ in range(n_bin):
            b, loff, n_chunk = struct.unpack_from("<IQi", raw, p); p += 16
            chunks = [struct.unpack_from("<QQ", raw, p + 16 * i) for i in range(n_chunk)]; p += 16 * n_chunk
            bins[b] = (loff, chunks)
        refs.append(bins)
    return dict(min_shift=min_shift, depth=depth, conf=(fmt, col_seq, col_beg, col_end, meta, skip), names=names, refs=refs)


def reg2bins(beg, end, min_shift, depth):
    out, s, t = [], min_shift + depth * 3, 0
    end -= 1
    for l in range(depth + 1):
        b, e = t + (beg >> s), t + (end >> s)
        out.extend(range(b, e + 1))
        s -= 3; t += 1 << (l * 3)
    return out


def query(gz, idx, tid, beg, end):
    ms, dp = idx["min_shift"], idx["depth"]
    bidx = idx["refs"][tid]
    first_leaf = ((1 << (3 * dp)) - 1) // 7
    b = first_leaf + (beg >> ms)
    while True:                                          # min_off: loffset of the bin of `beg`, or of the nearest bin to its left / above
        if b in bidx:
            break
        parent = (b - 1) >> 3
        first = (parent << 3) + 1
        b = b - 1 if b > first else parent
        if b == 0:
            break
    min_off = bidx[b][0] if b in bidx else 0
    chunks = []
    for bn in reg2bins(beg, end, ms, dp):
        if bn in bidx:
            chunks += [c for c in bidx[bn][1] if c[1] > min_off]
    chunks.sort()
    lines = []
    for cb, ce in chunks:
        coff, uoff = cb >> 16, cb & 0xffff
        while (coff << 16 | uoff) < ce:
            bsize = struct.unpack_from("<H", gz, coff + 16)[0] + 1
            text = zlib.decompress(gz[coff + 18:coff + bsize - 8], -15)
            while uoff < len(text) and (coff << 16 | uoff) < ce:
                nl = text.index(b"\n", uoff)              # an indexed file never splits a line across two blocks
                lines.append(text[uoff:nl + 1]); uoff = nl + 1
            if uoff >= len(text):
                coff += bsize; uoff = 0
    hit = []
    for ln in lines:
        f = ln.split(b"\t")
        if int(f[1]) < end and int(f[2]) > beg:
            hit.append(ln)
    return hit, len(lines)


@pytest.mark.parametrize("name", ["runs", "regions", "quant", "members", "lines"])
def test_csi_queries_find_exactly_the_overlapping_lines(outputs, name):
    gz = (outputs / f"t5.{name}.bed.gz").read_bytes()
    idx = parse_csi(gzip.decompress((outputs / f"t5.{name}.bed.gz.csi").read_bytes()))
    assert idx["min_shift"] == 14 and idx["depth"] == 5 and idx["conf"] == (0x10000, 1, 2, 3, ord("#"), 0)
    assert idx["names"] == [b"chr1", b"chr2_random_with_a_longer_name", b"chrM"]
    plain = (outputs / f"t1.{'runs' if name == 'members' else name}.bed").read_bytes().splitlines(keepends=True)
    by_chrom = {}
    for ln in plain:
        f = ln.split(b"\t")
        by_chrom.setdefault(f[0], []).append((int(f[1]), int(f[2]), ln))
    rng = random.Random(7)
    scanned = 0
    for tid, chrom in enumerate(idx["names"]):
        recs = by_chrom[chrom]
        top = recs[-1][1]
        spots = [(0, 1), (0, top), (top - 1, top), (16383, 16385), (16384, 16384 + 1), (131071, 131073)]
        spots += [(b, b + rng.choice([1, 10, 500, 20000, 300000])) for b in (rng.randrange(0, top) for _ in range(60))]
        for beg, end in spots:
            beg = min(beg, top - 1); end = max(beg + 1, min(end, top))
            got, n_scanned = query(gz, idx, tid, beg, end)
            want = [ln for s, e, ln in recs if s < end and e > beg]
            assert got == want, (chrom, beg, end)
            scanned += n_scanned
    assert scanned < 40 * len(plain)      # the index prunes: queries do not scan the file
