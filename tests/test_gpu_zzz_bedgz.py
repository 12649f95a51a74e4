"""K10 on the GPU (msd_per_base_bgzf, mosdepth_b200/csrc/bedgz.cuh): the BGZF members the device writes for a contig must
decompress to exactly the per-base lines of the oracle (and of msd_per_base_runs), member by member, with a valid CRC; and the
host CLI with MOSDEPTH_B200_DEVICE_BGZF=1 must write a per-base.bed.gz (+ .csi) whose content equals the --plain output."""
import gzip
import struct
import subprocess
import zlib

import numpy as np
import pytest

from conftest import ROOT, run_tool

pytestmark = pytest.mark.gpu

import os
os.environ["MSD_BGZF_MIN_RUNS"] = "0"      # the fixtures are tiny: take the device path whatever the number of runs (library default: 50,000)

CLI = ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200"
GEN = ROOT / "tools" / "gen_bam"
EOF_BLOCK = bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])


@pytest.fixture(scope="module")
def ctx():
    import mosdepth_b200 as m
    c = m.Context(0)
    yield c
    c.close()


def check_contig(ctx, tid, chrom):
    data, mem, (s, e, d) = ctx.per_base_bgzf(tid, chrom)
    s2, e2, d2 = ctx.per_base_runs(tid)
    assert np.array_equal(s, s2) and np.array_equal(e, e2) and np.array_equal(d, d2)
    want = "".join(f"{chrom}\t{a}\t{b}\t{c}\n" for a, b, c in zip(s.tolist(), e.tolist(), d.tolist())).encode()
    assert gzip.decompress(data + EOF_BLOCK) == want
    assert len(mem) >= 1 and int(mem["offset"][0]) == 0 and int(mem["first_run"][0]) == 0
    assert int(mem["offset"][-1]) + int(mem["csize"][-1]) == len(data) and int(mem["isize"].sum()) == len(want)
    at = 0
    for k in range(len(mem)):
        off, cs, isz = int(mem["offset"][k]), int(mem["csize"][k]), int(mem["isize"][k])
        assert data[off:off + 4] == b"\x1f\x8b\x08\x04" and struct.unpack_from("<H", data, off + 16)[0] + 1 == cs
        piece = zlib.decompress(data[off + 18:off + cs - 8], -15)
        assert piece == want[at:at + isz] and piece.endswith(b"\n")
        assert struct.unpack_from("<II", data, off + cs - 8) == (zlib.crc32(piece), isz)
        first = int(mem["first_run"][k])
        assert piece.startswith(f"{chrom}\t{int(s[first])}\t".encode())
        at += isz
    assert ctx.debug_inflate(data + EOF_BLOCK, len(want)) == want          # and the product's own inflate kernels read it back
    # label mode (msd_quantized_bgzf): the members hold `chrom start stop labels[bin]` of the runs msd_quantized_runs returns
    quants = [0, 1, 4, 100, 2**63 - 1]; labels = ["NO_COVERAGE", "1:4", "CALLABLE", "100:inf"]
    zdata, zmem, (qs, qe, qb) = ctx.quantized_bgzf(tid, quants, chrom, labels)
    qs2, qe2, qb2 = ctx.quantized_runs(tid, quants)
    assert np.array_equal(qs, qs2) and np.array_equal(qe, qe2) and np.array_equal(qb, qb2)
    qwant = "".join(f"{chrom}\t{a}\t{b}\t{labels[c]}\n" for a, b, c in zip(qs.tolist(), qe.tolist(), qb.tolist())).encode()
    assert gzip.decompress(zdata + EOF_BLOCK) == qwant and int(zmem["isize"].sum()) == len(qwant)
    return want


@pytest.mark.parametrize("name,mode", [("big.bam", 1), ("ovl.bam", 0), ("nanopore.bam", 0), ("empty-tids.bam", 0)])
def test_device_bgzf_equals_oracle_per_base(ctx, oracle_bin, fixtures, tmp_path, name, mode):
    bam = fixtures / name
    run_tool(oracle_bin, ["o"] + (["-x"] if mode == 1 else []) + [bam], tmp_path)
    per = {}
    for line in (tmp_path / "o.per-base.bed").read_bytes().splitlines(keepends=True):
        per.setdefault(line.split(b"\t")[0].decode(), []).append(line)
    hdr = ctx.open(bam)
    ctx.set_filters(mode=mode)
    ctx.run()
    n_with = 0
    for tid, cname in enumerate(hdr.names):
        rc, n = ctx.contig_records(tid)
        if rc == -2:
            continue
        n_with += 1
        assert check_contig(ctx, tid, cname) == b"".join(per[cname]), cname
    assert n_with >= 1


def test_device_bgzf_many_members_and_cli(ctx, tmp_path):
    """A synthetic 30X BAM (C2 shape, small): tens of thousands of lines per contig, many members; then the CLI path."""
    bam = tmp_path / "c2.bam"
    subprocess.run([str(GEN), "--config", "c3", "--genome-fraction", "0.0005", "--threads", "8", "-o", str(bam)], check=True, capture_output=True)
    hdr = ctx.open(bam)
    ctx.set_filters(mode=0)
    ctx.run()
    total_members = 0
    for tid, cname in enumerate(hdr.names):
        rc, n = ctx.contig_records(tid)
        if rc == -2:
            continue
        data, mem, _ = ctx.per_base_bgzf(tid, cname)
        total_members += len(mem)
        if tid < 3:
            check_contig(ctx, tid, cname)
    assert total_members > len(hdr.names)
    run_tool(CLI, ["--plain", "-q", "0:1:10:30:", "p", bam], tmp_path)
    run_tool(CLI, ["-q", "0:1:10:30:", "z", bam], tmp_path, env={"MOSDEPTH_B200_DEVICE_BGZF": "1", "MSD_BGZF_MIN_RUNS": "0"})
    run_tool(CLI, ["-q", "0:1:10:30:", "h", bam], tmp_path, env={"MOSDEPTH_B200_DEVICE_BGZF": "0"})
    plain = (tmp_path / "p.per-base.bed").read_bytes()
    dev = (tmp_path / "z.per-base.bed.gz").read_bytes()
    assert gzip.decompress(dev) == plain
    assert gzip.decompress((tmp_path / "h.per-base.bed.gz").read_bytes()) == plain
    assert dev != (tmp_path / "h.per-base.bed.gz").read_bytes()          # the device path really wrote it
    assert gzip.decompress((tmp_path / "z.per-base.bed.gz.csi").read_bytes())[:4] == b"CSI\x01"
    assert gzip.decompress((tmp_path / "z.quantized.bed.gz").read_bytes()) == (tmp_path / "p.quantized.bed").read_bytes()
    assert gzip.decompress((tmp_path / "h.quantized.bed.gz").read_bytes()) == (tmp_path / "p.quantized.bed").read_bytes()
    assert (tmp_path / "z.mosdepth.summary.txt").read_bytes() == (tmp_path / "p.mosdepth.summary.txt").read_bytes()
