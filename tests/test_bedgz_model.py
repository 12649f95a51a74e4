"""CPU replay of K10 (mosdepth_b200/csrc/bedgz_core.cuh, the per-line arithmetic of the device's per-base.bed.gz encoder)
against gzip: tools/bedgz_model.cpp drives the same MSD_HD functions the kernels of bedgz.cuh call, lines emitted in a
scrambled order.  The members must be valid BGZF, decompress to the printf text of the same runs, never split a line, and
compress about as well as zlib level 1 (what the reference's writer uses for per-base.bed.gz, mosdepth.nim:650-651)."""
import gzip
import struct
import subprocess
import zlib

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    d = tmp_path_factory.mktemp("bedgz")
    exe = d / "bedgz_model"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-I/usr/local/cuda/include", "-o", str(exe), str(ROOT / "tools" / "bedgz_model.cpp"), "-lz"], check=True)
    return exe, d


def members(data):
    pos, out = 0, []
    while pos < len(data):
        assert data[pos:pos + 4] == b"\x1f\x8b\x08\x04" and data[pos + 10:pos + 12] == b"\x06\x00"
        assert data[pos + 12:pos + 14] == b"BC" and struct.unpack_from("<H", data, pos + 14)[0] == 2
        bsize = struct.unpack_from("<H", data, pos + 16)[0] + 1
        crc, isize = struct.unpack_from("<II", data, pos + bsize - 8)
        out.append((pos, bsize, isize, crc))
        pos += bsize
    assert pos == len(data)
    return out


CASES = [
    ("chr20", 300000, "wgs", 1, 0.95), ("chr1", 120000, "sparse", 2, 0.9), ("chrUn_KI270442v1", 90000, "wide", 3, 0.9), ("1", 50000, "tiny", 4, 0.9),
    ("X", 3000, "tiny", 5, 0.5), ("n" * 200, 20000, "wgs", 6, 0.8), ("HLA-DRB1*15:01:01:01", 100000, "sparse", 7, 0.9), ("MT", 1, "wgs", 8, 0.0), ("MT", 2, "wide", 9, 0.0),
    ("ab", 70000, "wgs", 10, 0.9),
    # label mode (quantized.bed): gaps between runs, labels of 3 to 48 bytes copied from the nearest line with the same label
    ("chr5", 100000, "quant", 11, 0.7), ("1", 2000, "quant", 12, 0.0), ("chrUn_X", 3, "quant", 13, 0.0),
]


@pytest.mark.parametrize("chrom,n,pattern,seed,min_vs_zlib", CASES, ids=[f"{c[0][:12]}-{c[1]}-{c[2]}" for c in CASES])
def test_replay_against_gzip(model, chrom, n, pattern, seed, min_vs_zlib):
    exe, d = model
    prefix = d / f"m{seed}"
    r = subprocess.run([str(exe), str(prefix), chrom, str(n), pattern, str(seed)], check=True, capture_output=True, text=True)
    n_lines, text_bytes, gz_bytes, n_members, hdr_bits = (int(x) for x in r.stdout.split())
    gz = (d / f"m{seed}.gz").read_bytes()
    text = (d / f"m{seed}.txt").read_bytes()
    assert (d / f"m{seed}.replay.txt").read_bytes() == text          # the bytes the CRC is taken over
    assert len(text) == text_bytes and len(gz) == gz_bytes and n_lines == n
    assert gzip.decompress(gz) == text
    ms = members(gz)
    assert len(ms) == n_members + 1 and ms[-1][1] == 28 and ms[-1][2] == 0
    at = 0
    for pos, bsize, isize, crc in ms[:-1]:
        assert 0 < isize <= 32768 and bsize <= 65536
        piece = zlib.decompress(gz[pos + 18:pos + bsize - 8], -15)          # every member stands alone (BFINAL = 1)
        assert piece == text[at:at + isize] and piece.endswith(b"\n") and zlib.crc32(piece) == crc
        at += isize
    assert at == len(text)
    if min_vs_zlib:
        assert len(text) / len(gz) >= min_vs_zlib * len(text) / len(zlib.compress(text, 1))


def test_name_longer_than_the_device_path_accepts_is_refused(model):
    exe, d = model
    r = subprocess.run([str(exe), str(d / "long"), "n" * 201, "5", "wgs"], capture_output=True, text=True)
    assert r.returncode == 3 and "refused" in r.stderr


def test_k1_model_inflates_what_k10_deflates(model, tmp_path):
    """Round trip inside the product's own arithmetic: the members K10 writes (dynamic code with all 316 symbols, repeat code 16,
    32 KB members) go through the host replay of the inflate kernels (tools/inflate_model, K1a + K1b + K1c against zlib)."""
    import json
    exe, d = model
    inf = tmp_path / "inflate_model"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", "-o", str(inf), str(ROOT / "tools" / "inflate_model.cpp"), "-lz"], check=True)
    for chrom, n, pattern, seed in [("chr7", 150000, "wgs", 21), ("HLA-DRB1*15:01:01:01", 60000, "wide", 22), ("1", 40000, "sparse", 23)]:
        subprocess.run([str(exe), str(tmp_path / "rt"), chrom, str(n), pattern, str(seed)], check=True, capture_output=True)
        r = subprocess.run([str(inf), str(tmp_path / "rt.gz")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = json.loads(r.stdout)
        assert out["bad"] == 0 and out["blocks"] >= 2 and out["bytes"] >= (tmp_path / "rt.txt").stat().st_size      # every 8th block is replayed at a second alignment
