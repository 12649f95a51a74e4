"""The K8 (runs.cuh) and K10 (bedgz.cuh) KERNELS themselves, compiled by g++ and executed on host threads under a small CUDA
stand-in (tests/emu/cuda_runtime.h: CTAs one after the other, threads of a CTA as real threads, __syncthreads / warp shuffles as
barriers, atomicOr as a real atomic), launched in the order and geometry of msd_per_base_bgzf.  From a depth array to BGZF
members; gunzip must give the printf text of a plain host run-length pass.  This checks indexing, barriers, grid-stride loops and
the three-launch scans without a GPU; it is not a substitute for the GPU tests (tests/test_gpu_zzz_bedgz.py)."""
import gzip
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu")
    exe = d / "bedgz_emu"
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", f"-I{ROOT / 'tests' / 'emu'}", "-o", str(exe), str(ROOT / "tests" / "emu" / "bedgz_emu.cpp"), "-lz"], check=True)
    return exe, d


# (contig name, bases, CTA cap of the grid-stride kernels, seed): many scan tiles with few CTAs; the library's own cap; tiny inputs
# the last field: "" = per-base (K8 + K10, msd_per_base_bgzf), "quant" = quantized (K9 + host filter + K10 in label mode, msd_quantized_bgzf)
CASES = [("chr1", 400000, 7, 2, ""), ("HLA-DRB1*15:01:01:01", 150000, 1184, 3, ""), ("X", 5000, 3, 4, ""), ("MT", 17, 1184, 5, ""), ("c", 1, 1, 6, ""),
         ("chr1", 600000, 9, 7, "quant"), ("MT", 17, 1184, 8, "quant"), ("chrUn_KI270442v1", 2, 1, 9, "quant")]


@pytest.mark.parametrize("chrom,n_bases,max_ctas,seed,mode", CASES, ids=[f"{c[0][:8]}-{c[1]}-{c[2]}{c[4]}" for c in CASES])
def test_kernels_under_emulation(emu, chrom, n_bases, max_ctas, seed, mode):
    exe, d = emu
    prefix = d / f"e{seed}"
    r = subprocess.run([str(exe), str(prefix), chrom, str(n_bases), str(max_ctas), str(seed)] + ([mode] if mode else []), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    n_dev, n_host, text_bytes, gz_bytes, n_members = (int(v) for v in r.stdout.split())
    text = (d / f"e{seed}.txt").read_bytes()
    gz = (d / f"e{seed}.gz").read_bytes()
    assert n_dev == n_host and len(text) == text_bytes and len(gz) == gz_bytes and (n_members >= 1 or n_dev == 0)
    assert gzip.decompress(gz) == text


@pytest.fixture(scope="module")
def reduce_emu(tmp_path_factory):
    d = tmp_path_factory.mktemp("remu")
    exe = d / "reduce_emu"
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", f"-I{ROOT / 'tests' / 'emu'}", "-o", str(exe), str(ROOT / "tests" / "emu" / "reduce_emu.cpp")], check=True)
    return exe


# (tlen, window, CTA cap, seed): many scan tiles on few persistent CTAs (decoupled look-back), window 1, a window longer than the contig,
# windows wider than the quotient table's 1024 depths need, a contig of one tile + one cell
# (sizes: the warp-per-region kernel and the warp-aggregated histogram make several barrier exchanges per 32 cells under the stand-in;
# these sizes keep a case at 5-40 s)
REDUCE_CASES = [(60001, 500, 7, 4), (20000, 1, 16, 5), (16385, 100000000, 1184, 6), (33457, 3333, 5, 7), (4097, 4096, 2, 8), (40000, 33, 1184, 9), (16569, 100, 1184, 2),
                (777, 37, 3, 3)]


@pytest.mark.parametrize("tlen,window,max_ctas,seed", REDUCE_CASES, ids=[f"{c[0]}-{c[1]}-{c[2]}" for c in REDUCE_CASES])
def test_scan_window_region_kernels_under_emulation(reduce_emu, tlen, window, max_ctas, seed):
    """K6 scan_segments_kernel / totals_add_kernel, K9's keep_* compaction and K7 windows_mean_kernel / windows_kernel / regions_kernel under the CUDA stand-in against a plain
    restatement of the reference's arithmetic: prefix sum with int32 wrap, dist bins, depth stats, the sequential float64 mean bit for
    bit (Q1), the CountStat median (Q7), window bins (Q2), thresholds, intervals that reach past the contig end (Q5)."""
    r = subprocess.run([str(reduce_emu), str(tlen), str(window), str(max_ctas), str(seed)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


@pytest.fixture(scope="module")
def records_emu(tmp_path_factory):
    d = tmp_path_factory.mktemp("recemu")
    exe = d / "records_emu"
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", f"-I{ROOT / 'tests' / 'emu'}", "-o", str(exe), str(ROOT / "tests" / "emu" / "records_emu.cpp"), "-lz"], check=True)
    return exe


# (fixture, blocks per chunk): one block per chunk makes records and overlapping mates straddle chunk boundaries (carry buffer, live list)
RECORD_CASES = [("empty-tids.bam", 1), ("nanopore.bam", 1), ("overlapping-pairs.bam", 1), ("full-fragment-pairs.bam", 2)]
MODES = [(0, []), (1, ["-x"]), (2, ["-a"])]


@pytest.mark.parametrize("name,per_chunk", RECORD_CASES, ids=[f"{c[0]}-{c[1]}" for c in RECORD_CASES])
@pytest.mark.parametrize("mode,flags", MODES, ids=["default", "fast", "fragment"])
def test_framing_records_and_mate_join_under_emulation(records_emu, oracle_bin, tmp_path, name, per_chunk, mode, flags):
    """K2 (framing), K3/K5 (filters + CIGAR events) and K4 (mate join with its live list) under the CUDA stand-in, chunk by chunk in
    msd_run's launch order, blocks inflated by zlib: the per-base output must equal the oracle's byte for byte in all three modes."""
    from conftest import FIX, run_tool
    run_tool(oracle_bin, flags + ["o", FIX / name], tmp_path)
    r = subprocess.run([str(records_emu), str(FIX / name), str(tmp_path / "e.bed"), str(mode), str(per_chunk), "5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
    assert (tmp_path / "e.bed").read_bytes() == (tmp_path / "o.per-base.bed").read_bytes()


@pytest.mark.parametrize("mode,flags", [(0, []), (1, ["-x"])], ids=["default", "fast"])
def test_long_read_cigars_walked_by_the_warp_under_emulation(records_emu, oracle_bin, tmp_path, mode, flags):
    """K5's warp-cooperative CIGAR walk (records.cuh: coop_cigar, CIGARs of 48 operations or more) on synthetic nanopore reads (10-100 kb, an operation
    every ~10 bases, so thousands of operations per read and blocks that abut across the 32-operation steps): per-base output equal to the oracle's."""
    from conftest import run_tool
    gen = ROOT / "tools" / "gen_bam"
    if not gen.exists():
        subprocess.run(["make", "-C", str(ROOT), "tools/gen_bam"], check=True, capture_output=True)
    bam = tmp_path / "lr.bam"
    subprocess.run([str(gen), "--config", "c5", "--genome-fraction", "0.00004", "--threads", "2", "-o", str(bam)], check=True, capture_output=True)
    run_tool(oracle_bin, flags + ["o", bam], tmp_path)
    r = subprocess.run([str(records_emu), str(bam), str(tmp_path / "e.bed"), str(mode), "3", "5"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
    assert (tmp_path / "e.bed").read_bytes() == (tmp_path / "o.per-base.bed").read_bytes()
    assert (tmp_path / "e.bed").stat().st_size > 10000
