"""CPU test of the HOST side of the CLI stand-in (mosdepth_b200/host/main.cpp + bgzf_writer.hpp + lines.hpp): flags, header and
index parsing, spans, BED tables, the per-contig loop and its emission order, row formatting, the BGZF/CSI writer.  The CLI is
linked, in a temporary directory, against tests/mock/msd_mock.c -- a TEST DOUBLE of the C ABI that answers from the CPU oracle's
own functions -- and every output file must equal the oracle's byte for byte on the reference's fixtures and flag combinations
(the same cases tests/test_gpu_cli_parity.py runs against the real library on a GPU).  This says nothing about the kernels; it pins
everything around them without a GPU."""
import gzip
import os
import subprocess

import pytest

from conftest import ROOT, run_tool
from test_gpu_cli_parity import FIX_CASES, SUFFIXES


@pytest.fixture(scope="module")
def mock_cli(tmp_path_factory):
    d = tmp_path_factory.mktemp("mockcli")
    subprocess.run(["gcc", "-O2", "-fPIC", "-w", "-c", "-o", str(d / "msd_mock.o"), str(ROOT / "tests" / "mock" / "msd_mock.c")], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-w", "-I/usr/local/cuda/include", "-c", "-o", str(d / "msd_mock_bgzf.o"), str(ROOT / "tests" / "mock" / "msd_mock_bgzf.cpp")], check=True)
    subprocess.run(["g++", "-shared", "-o", str(d / "libmosdepth_b200.so"), str(d / "msd_mock.o"), str(d / "msd_mock_bgzf.o"), "-lz", "-lpthread", "-lm"], check=True)
    (d / "lib_path.txt").write_text(str(d / "libmosdepth_b200.so"))
    exe = d / "mosdepth-b200-mock"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-o", str(exe), str(ROOT / "mosdepth_b200" / "host" / "main.cpp"), f"-I{ROOT / 'include'}",
                    f"-L{d}", "-lmosdepth_b200", "-lz", f"-Wl,-rpath,{d}"], check=True)
    return exe


def compare(oracle_bin, cli, tmp_path, bam, flags, env=None, expect_rc=0, gz=False):
    ro = run_tool(oracle_bin, flags + ["o", bam], tmp_path, env=env, check=False)
    rg = run_tool(cli, ([] if gz else ["--plain"]) + flags + ["g", bam], tmp_path, env=env, check=False)
    assert ro.returncode == expect_rc, ro.stderr
    assert rg.returncode == expect_rc, rg.stderr
    n = 0
    for s in SUFFIXES:
        po = tmp_path / ("o" + s)
        pg = tmp_path / ("g" + s + (".gz" if gz and s.endswith(".bed") else ""))
        assert po.exists() == pg.exists(), s
        if po.exists():
            a, b = po.read_bytes(), pg.read_bytes()
            if gz and s.endswith(".bed"):
                b = gzip.decompress(b)
                assert gzip.decompress((tmp_path / ("g" + s + ".gz.csi")).read_bytes())[:4] == b"CSI\x01"
            if a != b:
                la, lb = a.splitlines(), b.splitlines()
                for i, (x, y) in enumerate(zip(la, lb)):
                    assert x == y, f"{s} line {i + 1}: oracle {x!r} != cli {y!r}"
                assert len(la) == len(lb), f"{s}: {len(la)} vs {len(lb)} lines"
            n += 1
    assert n >= 2


HEAVY = {"big.bam"}          # a 600 Mb contig: 2.4 GB depth array in the oracle and again in the mock; three cases are enough
SLOW = {("ovl.bam", ("-m", "--by", "37"))}      # 6.7 M rows of a contig without reads: 50 s here; replaced by the same flags on -c MT
CASES = [(n, f) for n, f in FIX_CASES if n not in HEAVY and (n, tuple(f)) not in SLOW] + [("big.bam", ["--by", "500", "-n", "-x"]), ("ovl.bam", ["-m", "--by", "37", "-c", "MT"])]


@pytest.mark.parametrize("name,flags", CASES, ids=[f"{n}:{' '.join(f)}" for n, f in CASES])
def test_fixture_cli_parity_on_mock(oracle_bin, mock_cli, fixtures, tmp_path, name, flags):
    flags = [f.replace("FIX", str(fixtures)) for f in flags]
    compare(oracle_bin, mock_cli, tmp_path, fixtures / name, flags)


GZ_CASES = [("ovl.bam", ["--by", "100", "-T", "0,1,2,3,4,5", "-c", "MT", "-q", "0:1:1000"]), ("empty-tids.bam", ["-b", "100", "-q", "0:1:5:150"]),
            ("nanopore.bam", ["--by", "250", "-T", "50,100"]), ("empty-tids.bam", ["-n", "--thresholds", "1,5", "--by", "FIX/empty-tids.bed"])]


@pytest.mark.parametrize("name,flags", GZ_CASES, ids=[f"gz:{n}:{' '.join(f)}" for n, f in GZ_CASES])
@pytest.mark.parametrize("threads", ["1", "5"])
def test_gz_outputs_on_mock(oracle_bin, mock_cli, fixtures, tmp_path, name, flags, threads):
    flags = [f.replace("FIX", str(fixtures)) for f in flags]
    compare(oracle_bin, mock_cli, tmp_path, fixtures / name, flags, env={"MOSDEPTH_B200_WRITER_THREADS": threads}, gz=True)


def test_env_precision_and_labels_on_mock(oracle_bin, mock_cli, fixtures, tmp_path):
    compare(oracle_bin, mock_cli, tmp_path, fixtures / "ovl.bam", ["-q", "0:1:1000", "--by", "77", "-c", "MT"], env={"MOSDEPTH_Q0": "AAA", "MOSDEPTH_Q1": "BBB", "MOSDEPTH_PRECISION": "7"})


def test_error_paths_match_on_mock(oracle_bin, mock_cli, fixtures, tmp_path):
    for flags, rc in ((["-c", "nonexistent", "--by", "20000"], 1), (["--min-frag-len", "10", "--max-frag-len", "9"], 2), (["-x", "-a"], 2), (["-T", "1"], 2)):
        ro = run_tool(oracle_bin, flags + ["o", fixtures / "ovl.bam"], tmp_path, check=False)
        rg = run_tool(mock_cli, flags + ["g", fixtures / "ovl.bam"], tmp_path, check=False)
        assert ro.returncode == rc and rg.returncode == rc
        assert ro.stderr.strip() == rg.stderr.strip()


def test_synthetic_c3_on_mock(oracle_bin, mock_cli, tmp_path):
    """25 contigs, 500-base windows: the north-star flags through the host loop (merged spans, prefetch hint, parallel row writer)."""
    bam = tmp_path / "c3.bam"
    subprocess.run([str(ROOT / "tools" / "gen_bam"), "--config", "c3", "--genome-fraction", "0.0005", "--threads", "4", "-o", str(bam)], check=True, capture_output=True)
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    compare(oracle_bin, mock_cli, tmp_path / "a", bam, ["--by", "500", "-n", "-x"])
    compare(oracle_bin, mock_cli, tmp_path / "b", bam, ["--by", "500"], gz=True)


def test_device_bgzf_branch_on_mock(oracle_bin, mock_cli, fixtures, tmp_path):
    """MOSDEPTH_B200_DEVICE_BGZF=1 (the default for .gz output): the CLI takes ready BGZF members from msd_per_base_bgzf (here: the host replay of K10's arithmetic)
    and only appends and indexes them; the per-base file must still decompress to the oracle's, and differ from the host writer's bytes."""
    for k, (name, flags) in enumerate((("nanopore.bam", ["--by", "250", "-q", "0:1:50:100:"]), ("empty-tids.bam", ["-x", "-q", "0:1:5:150"]), ("ovl.bam", ["-q", "0:1:1000"]))):
        d = tmp_path / str(k); d.mkdir()
        compare(oracle_bin, mock_cli, d, fixtures / name, flags, env={"MOSDEPTH_B200_DEVICE_BGZF": "1"}, gz=True)
        dev = (d / "g.per-base.bed.gz").read_bytes()
        compare(oracle_bin, mock_cli, d, fixtures / name, flags, env={"MOSDEPTH_B200_DEVICE_BGZF": "0"}, gz=True)      # the host writer (the device path is the default since r2)
        assert dev != (d / "g.per-base.bed.gz").read_bytes()
    # labels from the environment (mosdepth.nim:111-122), one of them longer than the device path takes: the CLI falls back to its own writer
    d = tmp_path / "env"; d.mkdir()
    compare(oracle_bin, mock_cli, d, fixtures / "nanopore.bam", ["-q", "0:1:50:100:"], env={"MOSDEPTH_B200_DEVICE_BGZF": "1", "MOSDEPTH_Q0": "NO_COVERAGE", "MOSDEPTH_Q1": "LOW", "MOSDEPTH_Q2": "CALLABLE"}, gz=True)
    compare(oracle_bin, mock_cli, d, fixtures / "nanopore.bam", ["-q", "0:1:50:100:"], env={"MOSDEPTH_B200_DEVICE_BGZF": "1", "MOSDEPTH_Q0": "N" * 60}, gz=True)


def test_exome_bed_median_thresholds_quantize_on_mock(oracle_bin, mock_cli, tmp_path):
    """BASELINE config C4 at small scale: unsorted overlapping BED, -T, -m, -q through the host loop."""
    bam, bed = tmp_path / "c4.bam", tmp_path / "c4.bed"
    subprocess.run([str(ROOT / "tools" / "gen_bam"), "--config", "c4", "--genome-fraction", "0.0008", "--bed", str(bed), "--threads", "4", "-o", str(bam)], check=True, capture_output=True)
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    compare(oracle_bin, mock_cli, tmp_path / "a", bam, ["--by", str(bed), "-T", "1,10,20,30", "-m", "-n", "-q", "0:1:10:30:"])
    compare(oracle_bin, mock_cli, tmp_path / "b", bam, ["--by", str(bed), "-n"], gz=True)


def test_python_binding_of_per_base_bgzf_on_mock(mock_cli, fixtures):
    """The ctypes binding (mosdepth_b200.Context.per_base_bgzf: pointer-to-pointer outputs, the msd_bgzf_member layout) against the
    double, in a process of its own whose LIB_PATH points at it: members parse, gunzip to the text of the runs, offsets chain."""
    lib = mock_cli.parent / "libmosdepth_b200.so"
    for name in ("nanopore.bam", "empty-tids.bam"):
        r = subprocess.run([os.sys.executable, str(ROOT / "tests" / "mock" / "binding_check.py"), str(lib), str(fixtures / name)], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.startswith("ok"), r.stderr


def test_bench_bedgz_script_on_mock(mock_cli, tmp_path):
    """tools/bench_bedgz.py (the K10 side measurement bench.py runs in a process of its own) end to end against the double:
    its JSON line parses and every contig's members were verified against the runs."""
    import json
    lib = mock_cli.parent / "libmosdepth_b200.so"
    script = ROOT / "tools" / "bench_bedgz.py"
    code = ("import sys; sys.path.insert(0, %r); import mosdepth_b200 as m; from pathlib import Path; m.LIB_PATH = Path(%r); m._lib = None; "
            "sys.argv = [%r, '0.0003', '1', '--json']; __file__ = %r; exec(compile(open(%r).read(), %r, 'exec'))") % (str(ROOT), str(lib), str(script), str(script), str(script), str(script))
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env={**os.environ, "MSD_BENCH_DIR": str(tmp_path)})
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["contigs"] >= 20 and out["verified_contigs"] == out["contigs"] and out["lines"] > 10000 and 0 < out["dev_bytes"] < out["text"]


def test_truncated_or_corrupt_header_and_index_end_with_a_message(mock_cli, fixtures, tmp_path):
    """The host's BAM-header and BAI/CSI parsers are bounds-checked: a cut or damaged file ends the run with `[mosdepth] ...` and
    exit code 1 (found with an ASan build of this CLI: the unchecked parser read past a truncated .bai), never with a signal."""
    import random
    import shutil
    rng = random.Random(3)
    for name, ext in (("ovl.bam", ".bai"), ("big.bam", ".csi")):
        bam = (fixtures / name).read_bytes(); idx = (fixtures / (name + ext)).read_bytes()
        raw_idx = gzip.decompress(idx) if ext == ".csi" else idx
        t = tmp_path / "t.bam"
        for f in (tmp_path / "t.bam.bai", tmp_path / "t.bam.csi"):
            f.unlink(missing_ok=True)
        t.write_bytes(bam)
        for cut in (0, 3, 4, 8, 12, 16, 20, 40, 100, len(raw_idx) - 20):      # (the last 8 bytes, n_no_coor, are optional)
            (tmp_path / ("t.bam" + ext)).write_bytes(raw_idx[:cut])
            r = run_tool(mock_cli, ["--plain", "-n", "--by", "100000", "x", t], tmp_path, check=False)
            assert r.returncode == 1 and "[mosdepth]" in r.stderr, (name, cut, r.returncode, r.stderr[-300:])
        for k in range(25 if name == "ovl.bam" else 5):      # (big.bam: a 600 Mb contig, every successful run fills a 2.4 GB array)
            b = bytearray(raw_idx)
            for _ in range(3):
                b[rng.randrange(len(b))] = rng.randrange(256)
            (tmp_path / ("t.bam" + ext)).write_bytes(bytes(b))
            r = run_tool(mock_cli, ["--plain", "-n", "--by", "100000", "x", t], tmp_path, check=False)
            assert r.returncode in (0, 1), (name, k, r.returncode, r.stderr[-300:])
        (tmp_path / ("t.bam" + ext)).write_bytes(idx)
        for cut in (0, 10, 17, 18, 30, 100, 500):
            t.write_bytes(bam[:cut])
            r = run_tool(mock_cli, ["--plain", "-n", "--by", "100000", "x", t], tmp_path, check=False)
            assert r.returncode == 1 and r.stderr.strip() if cut < 100 else r.returncode in (0, 1), (name, cut, r.returncode)
        for k in range(20 if name == "ovl.bam" else 4):
            b = bytearray(bam)
            for _ in range(2):
                b[rng.randrange(min(len(b), 400))] = rng.randrange(256)
            t.write_bytes(bytes(b))
            r = run_tool(mock_cli, ["--plain", "-n", "--by", "100000", "x", t], tmp_path, check=False)
            assert r.returncode in (0, 1, 2), (name, k, r.returncode, r.stderr[-300:])


def test_bench_renders_the_reference_files_from_library_results(oracle_bin, mock_cli, tmp_path):
    """bench.py's parity gate renders regions.bed / summary / global.dist / region.dist from what the library returns per contig
    (bench.render_outputs) and compares them with the oracle's files.  Here the renderer itself is pinned on the CPU: the results come
    from the ABI test double (= the oracle's arithmetic), so any difference is the renderer's."""
    bam = tmp_path / "c3.bam"
    subprocess.run([str(ROOT / "tools" / "gen_bam"), "--config", "c3", "--genome-fraction", "0.0006", "--threads", "4", "-o", str(bam)], check=True, capture_output=True)
    run_tool(oracle_bin, ["-n", "-x", "--by", "500", "o", bam], tmp_path)
    lib = mock_cli.parent / "libmosdepth_b200.so"
    code = f"""
import sys; sys.path.insert(0, {str(ROOT)!r})
from pathlib import Path
import numpy as np
import mosdepth_b200 as m
m.LIB_PATH = Path({str(lib)!r}); m._lib = None
import bench
ctx = m.Context(0); hdr = ctx.open({str(bam)!r}); ctx.set_filters(mode=m.MODE_FAST); ctx.run()
per = {{}}
for t in range(len(hdr.names)):
    rc, n = ctx.contig_records(t)
    if rc < 0:
        continue
    st, d = ctx.contig_stats(t); v, rst, rd = ctx.windows(t, 500)
    per[t] = dict(values=v, stat=[st.cum_length, st.cum_depth, st.min_depth, st.max_depth], dist=d, rstat=[rst.cum_length, rst.cum_depth, rst.min_depth, rst.max_depth], rdist=rd)
texts, _tot = bench.render_outputs(hdr.names, hdr.lengths, 500, per)
for suf, txt in texts.items():
    want = (Path({str(tmp_path)!r}) / ("o" + suf)).read_text()
    assert txt == want, suf
print("ok", len(per))
"""
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stderr[-3000:]
