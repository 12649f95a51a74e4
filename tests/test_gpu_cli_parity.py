"""End-to-end parity: the host CLI stand-in over the CUDA library (`mosdepth-b200 --plain`) must write the same
bytes as the CPU oracle for every output file, across the reference's fixtures and flag combinations, and on
synthetic BAMs of the BASELINE.json shapes (small scale)."""
import subprocess
from pathlib import Path

import pytest

from conftest import ROOT, run_tool

pytestmark = pytest.mark.gpu

CLI = ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200"
GEN = ROOT / "tools" / "gen_bam"
SUFFIXES = [".mosdepth.summary.txt", ".mosdepth.global.dist.txt", ".mosdepth.region.dist.txt", ".regions.bed", ".per-base.bed",
            ".quantized.bed", ".thresholds.bed"]


def compare(oracle_bin, tmp_path, bam, flags, env=None, expect_rc=0):
    ro = run_tool(oracle_bin, flags + ["o", bam], tmp_path, env=env, check=False)
    rg = run_tool(CLI, ["--plain"] + flags + ["g", bam], tmp_path, env=env, check=False)
    assert ro.returncode == expect_rc, ro.stderr
    assert rg.returncode == expect_rc, rg.stderr
    n = 0
    for s in SUFFIXES:
        po, pg = tmp_path / ("o" + s), tmp_path / ("g" + s)
        assert po.exists() == pg.exists(), s
        if po.exists():
            a, b = po.read_bytes(), pg.read_bytes()
            if a != b:
                la, lb = a.splitlines(), b.splitlines()
                for i, (x, y) in enumerate(zip(la, lb)):
                    assert x == y, f"{s} line {i + 1}: oracle {x!r} != gpu {y!r}"
                assert len(la) == len(lb), f"{s}: {len(la)} vs {len(lb)} lines"
            n += 1
    assert n >= 2
    return ro, rg


FIX_CASES = [
    ("ovl.bam", []), ("ovl.bam", ["-x"]), ("ovl.bam", ["--by", "100000000"]), ("ovl.bam", ["-l", "80", "-u", "80"]), ("ovl.bam", ["-l", "81"]),
    ("ovl.bam", ["-u", "79"]), ("ovl.bam", ["-q", "0:1:1000"]), ("ovl.bam", ["--by", "100", "-T", "0,1,2,3,4,5", "-c", "MT"]),
    ("ovl.bam", ["--by", "FIX/track.bed", "-T", "0,1,2", "-c", "MT"]), ("ovl.bam", ["--by", "FIX/track.bed"]), ("ovl.bam", ["--by", "FIX/unordered.bed"]),
    ("ovl.bam", ["--by", "FIX/missing.bed"]), ("ovl.bam", ["-n", "-R", "GT04008021_119"]), ("ovl.bam", ["-n", "-R", "MISSING"]), ("ovl.bam", ["-Q", "30"]),
    ("ovl.bam", ["-F", "0", "-i", "16"]), ("ovl.bam", ["-m", "--by", "37"]), ("ovl.bam", ["-c", "MT:1-100", "--by", "50"]),
    ("full-fragment-pairs.bam", ["-a"]), ("full-fragment-pairs.bam", []), ("full-fragment-pairs.bam", ["-x", "--by", "1000", "-q", "0:1:2:"]),
    ("big.bam", []), ("big.bam", ["--by", "500", "-n", "-x"]), ("big.bam", ["--by", "1000000", "-m", "-T", "1"]),
    ("overlapping-pairs.bam", []), ("overlapping-pairs.bam", ["-x", "-q", "1:2"]),
    ("nanopore.bam", []), ("nanopore.bam", ["-x"]), ("nanopore.bam", ["-q", "60"]), ("nanopore.bam", ["--by", "100", "-m", "-T", "50,100,120"]),
    ("nanopore.bam", ["-F", "2048", "--by", "250"]),
    ("empty-tids.bam", []), ("empty-tids.bam", ["-n", "-b", "100"]), ("empty-tids.bam", ["-n", "--thresholds", "1,5", "--by", "FIX/empty-tids.bed"]),
    ("empty-tids.bam", ["-x", "-b", "FIX/empty-tids.bed", "-m"]), ("empty-tids.bam", ["-a", "-q", "0:1:5:150"]), ("empty-tids.bam", ["-c", "HPV18", "-b", "33", "-m"]),
    ("empty-tids.bam", ["-n", "-b", "FIX/empty-tids.bed"]), ("empty-tids.bam", ["-Q", "20", "-F", "1024", "-b", "1000"]),
    ("chrom-end/minimal.bam", ["-n", "--by", "FIX/chrom-end/region-not-ok.bed"]), ("chrom-end/minimal.bam", ["--by", "10000000"]),
]


@pytest.mark.parametrize("name,flags", FIX_CASES, ids=[f"{n}:{' '.join(f)}" for n, f in FIX_CASES])
def test_fixture_cli_parity(oracle_bin, fixtures, tmp_path, name, flags):
    flags = [f.replace("FIX", str(fixtures)) for f in flags]
    compare(oracle_bin, tmp_path, fixtures / name, flags)


def test_env_precision_and_labels(oracle_bin, fixtures, tmp_path):
    compare(oracle_bin, tmp_path, fixtures / "ovl.bam", ["-q", "0:1:1000", "--by", "77"], env={"MOSDEPTH_Q0": "AAA", "MOSDEPTH_Q1": "BBB", "MOSDEPTH_PRECISION": "7"})


def test_error_paths_match(oracle_bin, fixtures, tmp_path):
    for flags, rc in ((["-c", "nonexistent", "--by", "20000"], 1), (["--min-frag-len", "10", "--max-frag-len", "9"], 2), (["-x", "-a"], 2), (["-T", "1"], 2)):
        ro = run_tool(oracle_bin, flags + ["o", fixtures / "ovl.bam"], tmp_path, check=False)
        rg = run_tool(CLI, flags + ["g", fixtures / "ovl.bam"], tmp_path, check=False)
        assert ro.returncode == rc and rg.returncode == rc
        assert ro.stderr.strip() == rg.stderr.strip()


@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    d = tmp_path_factory.mktemp("synth")
    out = {}
    for key, args in {"c3": ["--config", "c3", "--genome-fraction", "0.001"], "c2": ["--config", "c2", "--genome-fraction", "0.02"],
                      "c2_straddle": ["--config", "c2", "--genome-fraction", "0.01", "--straddle", "--level", "1"],
                      "c5": ["--config", "c5", "--genome-fraction", "0.0015", "--level", "1"],
                      "c4": ["--config", "c4", "--genome-fraction", "0.002", "--bed", str(d / "c4.bed")],
                      # CIGARs kept in CG:B,I tags behind the <l_seq>S<reflen>N placeholder: every second multi-op
                      # record of a short-read BAM, one-op records too (htslib leaves those alone), and long reads with
                      # more than 65,535 operations
                      "c2_cg": ["--config", "c2", "--genome-fraction", "0.008", "--cg-every", "2", "--cg-short"],
                      "c5_cg": ["--config", "c5", "--genome-fraction", "0.0008", "--level", "1", "--dense", "--cg-every", "2"]}.items():
        subprocess.run([str(GEN)] + args + ["--threads", "8", "-o", str(d / f"{key}.bam")], check=True, capture_output=True)
        out[key] = d / f"{key}.bam"
    out["c4_bed"] = d / "c4.bed"
    return out


SYN_CASES = [
    ("c3", ["--by", "500", "-n", "-x"]), ("c3", ["--by", "500"]), ("c3", ["--by", "333", "-m", "-n"]), ("c3", ["-a", "-n", "--by", "10000"]),
    ("c2", []), ("c2", ["-x"]), ("c2", ["-F", "0", "-Q", "1"]), ("c2", ["-q", "0:1:10:30:", "-n"]), ("c2", ["-R", "grp1", "-n", "--by", "1000"]), ("c2", ["-R", "nope", "-n"]),
    ("c2_straddle", []), ("c2_straddle", ["-x", "--by", "500", "-n"]),
    ("c5", []), ("c5", ["-x", "-n", "--by", "500"]), ("c5", ["-a", "-n", "--by", "2000"]),
    ("c2_cg", []), ("c2_cg", ["-x", "--by", "500"]), ("c5_cg", []), ("c5_cg", ["-x", "-n", "--by", "500"]), ("c5_cg", ["-a", "-n", "--by", "2000"]),
]


@pytest.mark.parametrize("key,flags", SYN_CASES, ids=[f"{k}:{' '.join(f)}" for k, f in SYN_CASES])
def test_synthetic_cli_parity(oracle_bin, synth, tmp_path, key, flags):
    compare(oracle_bin, tmp_path, synth[key], flags)


def test_exome_bed_median_thresholds_quantize(oracle_bin, synth, tmp_path):   # BASELINE config C4 at small scale
    compare(oracle_bin, tmp_path, synth["c4"], ["--by", str(synth["c4_bed"]), "-T", "1,10,20,30", "-m", "-n", "-q", "0:1:10:30:"])
    compare(oracle_bin, tmp_path, synth["c4"], ["--by", str(synth["c4_bed"]), "-n"])


def test_small_chunks_exercise_carry_and_live_mates(oracle_bin, synth, tmp_path):
    """Force many tiny chunks so records and overlapping mates straddle chunk boundaries."""
    env = {"MSD_CHUNK_BLOCKS": "3", "MSD_CHUNK_INFL_MB": "1"}
    compare(oracle_bin, tmp_path, synth["c2_straddle"], [], env=env)
    compare(oracle_bin, tmp_path, synth["c5"], ["--by", "1000"], env=env)
    compare(oracle_bin, tmp_path, synth["c2"], ["--by", "500"], env=env)
    compare(oracle_bin, tmp_path, synth["c2_cg"], [], env=env)
    compare(oracle_bin, tmp_path, synth["c5_cg"], ["--by", "1000"], env=env)
