"""Pins the CPU oracle against every hot-path known answer of the reference's own test-suite
(/root/reference/functional-tests.sh, line numbers cited per test; depthstat.nim:59-79), run on the
reference's fixture BAMs (tests/golden/fixtures, copied data).  SURVEY.md section 8c lists the goldens.
The oracle writes *.bed where the reference writes *.bed.gz (parity is on decompressed content)."""
from pathlib import Path

import pytest

from conftest import run_tool


def lines(p):
    return Path(p).read_text().splitlines()


def grep(p, pre):
    return [l for l in lines(p) if l.startswith(pre)]


def test_overlapM(oracle_bin, fixtures, tmp_path):  # functional-tests.sh:26-32
    run_tool(oracle_bin, ["t", fixtures / "ovl.bam"], tmp_path)
    assert grep(tmp_path / "t.per-base.bed", "MT") == ["MT\t0\t80\t1", "MT\t80\t16569\t0"]
    assert grep(tmp_path / "t.per-base.bed", "1\t") == ["1\t0\t249250621\t0"]
    s = lines(tmp_path / "t.mosdepth.summary.txt")
    assert [l for l in s if "MT" in l] == ["MT\t16569\t80\t0.00\t0\t1"]
    assert [l for l in s if "total" in l] == ["total\t16569\t80\t0.00\t0\t1"]
    # 4 dist lines
    assert lines(tmp_path / "t.mosdepth.global.dist.txt") == ["MT\t1\t0.00", "MT\t0\t1.00", "total\t1\t0.00", "total\t0\t1.00"]


def test_overlapFastMode(oracle_bin, fixtures, tmp_path):  # :34-41
    run_tool(oracle_bin, ["t", "--fast-mode", fixtures / "ovl.bam"], tmp_path)
    assert grep(tmp_path / "t.per-base.bed", "MT") == ["MT\t0\t6\t1", "MT\t6\t42\t2", "MT\t42\t80\t1", "MT\t80\t16569\t0"]
    s = lines(tmp_path / "t.mosdepth.summary.txt")[1:]
    assert {l.split("\t")[1] for l in s} == {"16569"}
    assert {l.split("\t")[3] for l in s} == {"0.01"}


def test_missing_chrom(oracle_bin, fixtures, tmp_path):  # :44-46
    r = run_tool(oracle_bin, ["-c", "nonexistent", "--by", "20000", "t", fixtures / "ovl.bam"], tmp_path, check=False)
    assert r.returncode == 1 and "[mosdepth] chromosome nonexistent not found" in r.stderr


def test_unordered_bed(oracle_bin, fixtures, tmp_path):  # :48-50
    run_tool(oracle_bin, ["--by", fixtures / "unordered.bed", "t", fixtures / "ovl.bam"], tmp_path)
    assert len(lines(tmp_path / "t.regions.bed")) == 2


def test_missing_bed_chrom(oracle_bin, fixtures, tmp_path):  # :53-54
    run_tool(oracle_bin, ["--by", fixtures / "missing.bed", "t", fixtures / "ovl.bam"], tmp_path)


def test_bed_start_past_chrom_end(oracle_bin, fixtures, tmp_path):  # :56-57
    run_tool(oracle_bin, ["-n", "--by", fixtures / "chrom-end" / "region-not-ok.bed", "t", fixtures / "chrom-end" / "minimal.bam"], tmp_path)


def test_big_window(oracle_bin, fixtures, tmp_path):  # :59-62
    run_tool(oracle_bin, ["t", fixtures / "ovl.bam", "--by", "100000000"], tmp_path)
    assert len(grep(tmp_path / "t.per-base.bed", "MT")) == 2
    assert grep(tmp_path / "t.regions.bed", "MT") == ["MT\t0\t16569\t0.00"]


def test_length_filter(oracle_bin, fixtures, tmp_path):  # :66-81
    run_tool(oracle_bin, ["t", fixtures / "ovl.bam", "--min-frag-len", "80", "--max-frag-len", "80"], tmp_path)
    assert grep(tmp_path / "t.per-base.bed", "MT") == ["MT\t0\t80\t1", "MT\t80\t16569\t0"]
    run_tool(oracle_bin, ["t", fixtures / "ovl.bam", "--min-frag-len", "81"], tmp_path)
    assert grep(tmp_path / "t.per-base.bed", "MT") == ["MT\t0\t16569\t0"]
    run_tool(oracle_bin, ["t", fixtures / "ovl.bam", "--max-frag-len", "79"], tmp_path)
    assert grep(tmp_path / "t.per-base.bed", "MT") == ["MT\t0\t16569\t0"]
    r = run_tool(oracle_bin, ["t", fixtures / "ovl.bam", "--min-frag-len", "10", "--max-frag-len", "9"], tmp_path, check=False)
    assert r.returncode == 2 and "--max-frag-len was lower than --min-frag-len." in r.stderr


def test_fragment_mode(oracle_bin, fixtures, tmp_path):  # :85-95
    run_tool(oracle_bin, ["t", "--fragment-mode", fixtures / "full-fragment-pairs.bam"], tmp_path)
    c = "chr22:20000000-23000000"
    exp = [(0, 17318, 0), (17318, 17320, 1), (17320, 17420, 2), (17420, 17756, 1), (17756, 52130, 0),
           (52130, 52135, 1), (52135, 52235, 2), (52235, 52546, 1), (52546, 3000001, 0)]
    assert lines(tmp_path / "t.per-base.bed") == [f"{c}\t{a}\t{b}\t{d}" for a, b, d in exp]


def test_quantize(oracle_bin, fixtures, tmp_path):  # :102-106
    run_tool(oracle_bin, ["-q", "0:1:1000", "t", fixtures / "ovl.bam"], tmp_path)
    assert grep(tmp_path / "t.quantized.bed", "MT") == ["MT\t0\t80\t1:1000", "MT\t80\t16569\t0:1"]
    assert grep(tmp_path / "t.quantized.bed", "1\t") == ["1\t0\t249250621\t0:1"]


def test_single_quant_nanopore(oracle_bin, fixtures, tmp_path):  # :108-109 (exit code only in the reference)
    run_tool(oracle_bin, ["-q", "60", "t", fixtures / "nanopore.bam"], tmp_path)
    q = grep(tmp_path / "t.quantized.bed", "LXWQ01001294.1\t")
    assert len(q) == 25 and q[0].endswith("60:inf")
    assert lines(tmp_path / "t.quantized.bed")[0] == "LXWQ01000001.1\t0\t1443\t0:60"   # -2 contig, quantize[0]==0 (:795-799)


def test_thresholds_window(oracle_bin, fixtures, tmp_path):  # :113-116
    run_tool(oracle_bin, ["--by", "100", "-T", "0,1,2,3,4,5", "-c", "MT", "t", fixtures / "ovl.bam"], tmp_path)
    t = lines(tmp_path / "t.thresholds.bed")
    assert t[0] == "#chrom\tstart\tend\tregion\t0X\t1X\t2X\t3X\t4X\t5X"
    assert t[1] == "MT\t0\t100\tunknown\t100\t80\t0\t0\t0\t0"
    assert {l.split("\t")[6] for l in t[1:]} == {"0"}


def test_thresholds_bed(oracle_bin, fixtures, tmp_path):  # :119-121
    run_tool(oracle_bin, ["--by", fixtures / "track.bed", "-T", "0,1,2", "-c", "MT", "t", fixtures / "ovl.bam"], tmp_path)
    assert lines(tmp_path / "t.thresholds.bed")[1:] == ["MT\t2\t80\taregion\t78\t78\t0"]


def test_named_quant_and_precision(oracle_bin, fixtures, tmp_path):  # :123-132
    run_tool(oracle_bin, ["-q", "0:1:1000", "t", fixtures / "ovl.bam"], tmp_path,
             env={"MOSDEPTH_Q0": "AAA", "MOSDEPTH_Q1": "BBB", "MOSDEPTH_PRECISION": "7"})
    assert grep(tmp_path / "t.quantized.bed", "MT\t") == ["MT\t0\t80\tBBB", "MT\t80\t16569\tAAA"]
    assert lines(tmp_path / "t.mosdepth.global.dist.txt")[0] == "MT\t1\t0.0048283"
    s = lines(tmp_path / "t.mosdepth.summary.txt")
    assert s[1] == "MT\t16569\t80\t0.0048283\t0\t1"
    assert s[-1] == "total\t16569\t80\t0.0048283\t0\t1"


def test_track_header(oracle_bin, fixtures, tmp_path):  # :135-137
    run_tool(oracle_bin, ["--by", fixtures / "track.bed", "t", fixtures / "ovl.bam"], tmp_path)
    assert lines(tmp_path / "t.regions.bed") == ["MT\t2\t80\taregion\t1.00"]


def test_bad_bed(oracle_bin, fixtures, tmp_path):  # :139-142
    r = run_tool(oracle_bin, ["--by", fixtures / "bad.bed", "t", fixtures / "ovl.bam"], tmp_path, check=False)
    assert r.returncode == 1
    assert "skipping bad bed line:MT\t2" in r.stderr
    assert "invalid integer: asdf" in r.stderr


def test_read_groups(oracle_bin, fixtures, tmp_path):  # :145-153
    run_tool(oracle_bin, ["-n", "t", fixtures / "ovl.bam"], tmp_path)
    run_tool(oracle_bin, ["-n", "tt", fixtures / "ovl.bam", "-R", "GT04008021_119"], tmp_path)
    a = lines(tmp_path / "tt.mosdepth.global.dist.txt")
    assert len(a) == 4 and a == lines(tmp_path / "t.mosdepth.global.dist.txt")
    run_tool(oracle_bin, ["-n", "tt", fixtures / "ovl.bam", "-R", "MISSING"], tmp_path)
    assert lines(tmp_path / "tt.mosdepth.global.dist.txt") == ["MT\t0\t1.00", "total\t0\t1.00"]


def test_big_chrom(oracle_bin, fixtures, tmp_path):  # :155-160
    run_tool(oracle_bin, ["t", fixtures / "big.bam"], tmp_path)
    s = lines(tmp_path / "t.mosdepth.summary.txt")
    assert len(s) == 3
    assert {l.split("\t")[2] for l in s[1:]} == {"15"}
    assert {l.split("\t")[4] for l in s[1:]} == {"0"}
    assert {l.split("\t")[5] for l in s[1:]} == {"1"}
    # SURVEY appendix A derived runs
    assert lines(tmp_path / "t.per-base.bed") == [
        "ref\t0\t600000006\t0", "ref\t600000006\t600000018\t1", "ref\t600000018\t600000019\t0",
        "ref\t600000019\t600000022\t1", "ref\t600000022\t600000045\t0"]


def test_empty_tids(oracle_bin, fixtures, tmp_path):  # :165-169
    run_tool(oracle_bin, ["t", "-n", "--thresholds", "1,5", "--by", fixtures / "empty-tids.bed", fixtures / "empty-tids.bam"], tmp_path)
    assert not [l for l in lines(tmp_path / "t.mosdepth.region.dist.txt") if l.split("\t")[0] == "HPV26"]


def test_regions_vs_global_1436(oracle_bin, fixtures, tmp_path):  # :172-174
    import difflib
    run_tool(oracle_bin, ["t_regions", "-n", "-b", "100", fixtures / "empty-tids.bam"], tmp_path)
    import subprocess
    r = subprocess.run(["diff", "-u", "t_regions.mosdepth.region.dist.txt", "t_regions.mosdepth.global.dist.txt"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert len(r.stdout.splitlines()) == 1436


def test_overlapping_pairs(oracle_bin, fixtures, tmp_path):  # :176-181
    run_tool(oracle_bin, ["t", fixtures / "overlapping-pairs.bam"], tmp_path)
    assert lines(tmp_path / "t.per-base.bed") == ["1\t0\t565173\t0", "1\t565173\t565253\t1", "1\t565253\t249250621\t0"]


def test_c1_big_bam_windows_derived(oracle_bin, fixtures, tmp_path):
    """BASELINE config C1 (tests/big.bam --by 500 -n -x): SURVEY 8c derived golden (not asserted by the reference)."""
    run_tool(oracle_bin, ["t", "--by", "500", "-n", "-x", fixtures / "big.bam"], tmp_path)
    assert lines(tmp_path / "t.mosdepth.summary.txt")[1:] == [
        "ref\t600000045\t16\t0.00\t0\t1", "ref_region\t600000045\t16\t0.00\t0\t1",
        "total\t600000045\t16\t0.00\t0\t1", "total_region\t600000045\t16\t0.00\t0\t1"]
    assert lines(tmp_path / "t.mosdepth.global.dist.txt") == ["ref\t0\t1.00", "total\t0\t1.00"]
    assert lines(tmp_path / "t.mosdepth.region.dist.txt") == ["ref\t0\t1.00", "total\t0\t1.00"]
    reg = lines(tmp_path / "t.regions.bed")
    assert len(reg) == 1200001
    assert reg[-1] == "ref\t600000000\t600000045\t0.36"
    assert all(l.endswith("\t0.00") for l in reg[:-1])


def test_nanopore_derived(oracle_bin, fixtures, tmp_path):
    """SURVEY appendix A (derived; the reference only checks the exit code for nanopore.bam)."""
    run_tool(oracle_bin, ["t", fixtures / "nanopore.bam"], tmp_path)
    s = [l for l in lines(tmp_path / "t.mosdepth.summary.txt") if l.startswith("LXWQ01001294.1")]
    assert s == ["LXWQ01001294.1\t1706\t194207\t113.84\t31\t131"]
    assert len(grep(tmp_path / "t.per-base.bed", "LXWQ01001294.1")) == 1481
    run_tool(oracle_bin, ["t", "-x", fixtures / "nanopore.bam"], tmp_path)
    s = [l for l in lines(tmp_path / "t.mosdepth.summary.txt") if l.startswith("LXWQ01001294.1")]
    assert s == ["LXWQ01001294.1\t1706\t207828\t121.82\t33\t133"]
    assert len(grep(tmp_path / "t.per-base.bed", "LXWQ01001294.1")) == 170


def test_empty_tids_summary_derived(oracle_bin, fixtures, tmp_path):
    run_tool(oracle_bin, ["t", fixtures / "empty-tids.bam"], tmp_path)
    s = {l.split("\t")[0]: l for l in lines(tmp_path / "t.mosdepth.summary.txt")}
    assert s["HPV18"] == "HPV18\t7857\t1774383\t225.83\t0\t646"
    assert s["HCV-1"] == "HCV-1\t9646\t1994\t0.21\t0\t37"
    assert s["total"].startswith("total\t105198\t1785795\t16.98")
    assert len(s) == 1 + 13 + 1


def test_threads_give_same_output(oracle_bin, fixtures, tmp_path):
    run_tool(oracle_bin, ["a", "-t", "0", fixtures / "empty-tids.bam"], tmp_path)
    run_tool(oracle_bin, ["b", "-t", "3", fixtures / "empty-tids.bam"], tmp_path)
    for suf in [".per-base.bed", ".mosdepth.summary.txt", ".mosdepth.global.dist.txt"]:
        assert (tmp_path / ("a" + suf)).read_bytes() == (tmp_path / ("b" + suf)).read_bytes()


def test_median_rule():
    """depthstat.nim:62-79 self test, restated on the rule the oracle and the kernel both implement."""
    def median(vals, size=32768):
        n = len(vals); stop_n = int(0.5 + n * 0.5); cum = 0
        counts = [0] * size
        for v in vals: counts[min(size - 1, v)] += 1
        for i, c in enumerate(counts):
            cum += c
            if cum >= stop_n: return i
        return -1
    v = list(range(11))
    assert median(v) == 5
    assert median(v + [6, 6, 6, 6]) == 6
    assert median([]) == 0
