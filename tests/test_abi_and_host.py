"""CPU-only checks: the C-ABI library loads and exports every symbol include/mosdepth_b200.h declares, fails loudly
without a GPU (no CPU fallback), and the host-side logic (header/index parsing, window arithmetic, sharding)."""
import ctypes
import json
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, run_tool


@pytest.fixture(scope="module")
def built():
    lib = ROOT / "mosdepth_b200" / "libmosdepth_b200.so"
    if not lib.exists() or not (ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200").exists() or not (ROOT / "tools" / "gen_bam").exists():
        import __graft_entry__ as ge
        ge.build()
    return lib


def test_library_exports_every_declared_symbol(built):
    hdr = (ROOT / "include" / "mosdepth_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(msd_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = ctypes.CDLL(str(built))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    import mosdepth_b200 as m
    assert set(m.EXPORTS) == declared
    assert m.load_library().msd_abi_version() == 2


def test_no_cpu_fallback(built):
    """Without an sm_100 device msd_create must fail with MSD_ENODEV and say so."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mosdepth_b200 as m
    with pytest.raises(m.MsdError) as ei:
        m.Context(0)
    assert ei.value.code == -10
    r = run_tool(ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200", ["t", ROOT / "tests/golden/fixtures/ovl.bam"], ROOT / "tests", check=False)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def test_product_never_references_oracle():
    for p in list((ROOT / "mosdepth_b200").rglob("*")) + [ROOT / "Makefile"]:
        if p.is_file() and p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "") and p.name != "Makefile":
            try:
                txt = p.read_text()
            except UnicodeDecodeError:
                continue
            assert "mosdepth_oracle" not in txt and "oracle/" not in txt, p


def test_n_windows_matches_window_gen(built):
    import mosdepth_b200 as m
    lib = m.load_library()

    def window_gen(window, tlen):   # mosdepth.nim:382-388 in uint32 arithmetic
        out, start = 0, 0
        while ((start + window) & 0xFFFFFFFF) < tlen:
            out += 1; start = (start + window) & 0xFFFFFFFF
        if start != tlen:
            out += 1
        return out
    for tlen, w in [(0, 5), (1, 1), (1, 500), (499, 500), (500, 500), (501, 500), (16569, 100), (600000045, 500), (1000, 7)]:
        assert lib.msd_n_windows(tlen, w) == window_gen(w, tlen), (tlen, w)


def test_header_and_index_parsing(fixtures):
    from mosdepth_b200 import hostio
    h = hostio.read_bam_header(fixtures / "ovl.bam")
    assert len(h.names) == 87 and h.names[24] == "MT" and h.lengths[24] == 16569
    ix = hostio.read_index(fixtures / "ovl.bam")
    assert [i for i, c in enumerate(ix) if c.has_data] == [24] and ix[24].n_mapped == 2 and ix[24].ref_beg == h.first_record_voff
    h = hostio.read_bam_header(fixtures / "big.bam")
    assert h.names == ["ref"] and h.lengths == [600000045]
    ix = hostio.read_index(fixtures / "big.bam")   # BGZF-compressed CSI, depth 6
    assert ix[0].has_data and ix[0].ref_beg == (69 << 16) and ix[0].ref_end == (166 << 16) and ix[0].n_mapped == 1
    ix = hostio.read_index(fixtures / "empty-tids.bam")
    assert sum(c.has_data for c in ix) == 13
    with pytest.raises(FileNotFoundError):
        hostio.read_index(fixtures / "track.bed")


def test_cli_argument_errors_need_no_gpu(built, fixtures, tmp_path):
    cli = ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200"
    r = run_tool(cli, ["-c", "nonexistent", "--by", "20000", "t", fixtures / "ovl.bam"], tmp_path, check=False)
    assert r.returncode == 1 and "[mosdepth] chromosome nonexistent not found" in r.stderr
    r = run_tool(cli, ["t", fixtures / "ovl.bam", "--min-frag-len", "10", "--max-frag-len", "9"], tmp_path, check=False)
    assert r.returncode == 2 and "--max-frag-len was lower than --min-frag-len." in r.stderr
    r = run_tool(cli, ["xx", "tests/tt.cram"], tmp_path, check=False)
    assert r.returncode == 1
    r = run_tool(cli, ["--by", fixtures / "bad.bed", "t", fixtures / "ovl.bam"], tmp_path, check=False)
    assert r.returncode == 1 and "skipping bad bed line:MT\t2" in r.stderr and "invalid integer: asdf" in r.stderr
    r = run_tool(cli, ["-t4prefix", "sample.bam"], tmp_path, check=False)
    assert r.returncode == 1 and "error parsing arguments" in r.stderr


def test_lpt_shards_cover_and_balance():
    from mosdepth_b200.shard import lpt_shards
    grch38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622,
              133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
    for n in (1, 2, 4, 8):
        sh = lpt_shards(grch38, n)
        assert sorted(t for s in sh for t in s) == list(range(25))
        loads = [sum(grch38[t] for t in s) for s in sh]
        assert max(loads) <= 1.05 * sum(grch38) / n
    assert lpt_shards([0, 5, 0, 3], 2) == [[1], [3]]


def test_generator_is_thread_count_independent(built, tmp_path):
    gen = ROOT / "tools" / "gen_bam"
    for t in (1, 5):
        subprocess.run([str(gen), "--config", "c3", "--genome-fraction", "0.0003", "--threads", str(t), "-o", str(tmp_path / f"t{t}.bam")], check=True, capture_output=True)
    from mosdepth_b200 import hostio
    a = hostio.inflate_all((tmp_path / "t1.bam").read_bytes()); b = hostio.inflate_all((tmp_path / "t5.bam").read_bytes())
    assert a == b


def test_split_contig_spans_hold_every_owned_record(tmp_path):
    """Intra-contig sharding, host side (no GPU): cuts are aligned to the index windows and the --by window, the parts
    partition [0, tlen), and the span of each part contains every record it owns plus one at or beyond its end."""
    import struct
    from mosdepth_b200 import hostio
    from mosdepth_b200.shard import split_contig
    bam = tmp_path / "c2.bam"
    subprocess.run([str(ROOT / "tools" / "gen_bam"), "--config", "c2", "--genome-fraction", "0.008", "--threads", "2", "-o", str(bam)], check=True, capture_output=True)
    hdr = hostio.read_bam_header(bam); index = hostio.read_index(bam)
    data = bam.read_bytes()
    tid, tlen = 0, hdr.lengths[0]

    def positions(vb, ve):
        pos, stream, first = vb >> 16, b"", True
        end_c, end_u = ve >> 16, ve & 0xFFFF
        for p, bl, xl, isz in hostio.iter_bgzf_blocks(data, pos):
            if p > end_c or (p == end_c and end_u == 0):
                break
            blk = hostio.inflate_block(data, p, bl, xl)
            if p == end_c:
                blk = blk[:end_u]
            if first:
                blk = blk[vb & 0xFFFF:]; first = False
            stream += blk
        out, o = [], 0
        while o + 12 <= len(stream):
            bs, t, ps = struct.unpack_from("<iii", stream, o)
            if o + 4 + bs > len(stream):
                break
            if t == tid:
                out.append(ps)
            o += 4 + bs
        return out

    every = positions(index[tid].ref_beg, index[tid].ref_end)
    for n_parts in (2, 4):
        parts = split_contig(index[tid], tlen, n_parts, window=500)
        assert len(parts) == n_parts
        assert parts[0][0] == 0 and parts[-1][1] == tlen and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        owned_total = 0
        for lo, hi, vb, ve in parts:
            assert lo % 32 == 0 and lo % 500 == 0
            got = positions(vb, ve)
            owned = [p for p in got if lo <= p < hi]
            assert len(owned) == sum(1 for p in every if lo <= p < hi)
            assert hi == tlen or any(p >= hi for p in got)
            owned_total += len(owned)
        assert owned_total == len(every)


def test_ingest_parallel_copy_and_block_chase(tmp_path):
    """mosdepth_b200/csrc/ingest.hpp on the CPU: the multi-threaded copies of the pinned-ring pump equal memcpy, and the parallel BSIZE
    chase (slices found by block signature, stitched) returns exactly the blocks of the serial walk -- also when block signatures are
    planted inside payloads.  (MSD_CHASE_SLICE_KB shrinks the slices so that a 100 MB test file is cut into many.)"""
    exe = tmp_path / "ingest_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-I/usr/local/cuda/include", "-o", str(exe), str(ROOT / "tools" / "ingest_test.cpp")], check=True)
    bam = tmp_path / "c3.bam"
    subprocess.run([str(ROOT / "tools" / "gen_bam"), "--config", "c3", "--genome-fraction", "0.002", "--threads", "4", "-o", str(bam)], check=True, capture_output=True)
    r = subprocess.run([str(exe), str(bam)], capture_output=True, text=True, env={**os.environ, "MSD_CHASE_SLICE_KB": "512"})
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["bad"] == 0 and out["chase_cases"] == 72


def test_cli_shard_plan_equals_python_plan(tmp_path):
    """The host CLI's multi-GPU plan (`--print-shards`, no GPU touched) is the plan of mosdepth_b200.shard: offset ranges cut inside
    contigs for window mode, whole contigs (LPT) for the modes whose queries need a contig on one GPU."""
    import json
    from mosdepth_b200 import hostio
    from mosdepth_b200.shard import contig_weights, lpt_shards, offset_shards
    bam = tmp_path / "s.bam"
    subprocess.run([str(ROOT / "tools" / "gen_bam"), "--config", "c3", "--genome-fraction", "0.003", "--threads", "4", "-o", str(bam)], check=True, capture_output=True)
    hdr, idx = hostio.read_bam_header(bam), hostio.read_index(bam)
    cli = ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200"
    for n in (2, 3, 8):
        got = json.loads(subprocess.run([str(cli), "--print-shards", "--gpus", str(n), "-n", "-x", "--by", "500", "x", str(bam)], check=True, capture_output=True, text=True).stdout)
        assert got["sharding"] == "offsets"
        want = offset_shards(idx, hdr.lengths, n, window=500, with_cover=True)
        assert [[tuple(p) for p in r] for r in got["ranks"]] == [[tuple(int(x) for x in p) for p in r] for r in want]
        for r in got["ranks"]:      # a cut piece's cover never lies beyond its margin
            for t, lo, hi, vb, ve, cover in r:
                assert cover <= max(lo - 65536, 0) and vb < ve
        # per-base output needs whole contigs: LPT
        got = json.loads(subprocess.run([str(cli), "--print-shards", "--gpus", str(n), "--by", "500", "x", str(bam)], check=True, capture_output=True, text=True).stdout)
        assert got["sharding"] == "contigs"
        assert [[p[0] for p in r] for r in got["ranks"]] == lpt_shards(contig_weights(idx), n)


def test_nim_shim_declares_every_export():
    """mosdepth_b200/nim/mosdepth_b200.nim (source only: no Nim compiler here) lists exactly the functions include/mosdepth_b200.h
    declares (the shim claimed to mirror the header and omitted five exports)."""
    hdr = (ROOT / "include" / "mosdepth_b200.h").read_text()
    nim = (ROOT / "mosdepth_b200" / "nim" / "mosdepth_b200.nim").read_text()
    declared = set(re.findall(r"\b(msd_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    shim = set(re.findall(r"^proc (msd_[a-z0-9_]+)\*", nim, flags=re.M))
    assert declared == shim, (sorted(declared - shim), sorted(shim - declared))
    import mosdepth_b200 as m
    assert declared == set(m.EXPORTS)
