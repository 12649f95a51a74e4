"""CIGARs with more than 65,535 operations: BAM keeps them in a CG:B,I tag behind the placeholder
<l_seq>S<reflen>N and htslib's bam_read1 swaps them back in before mosdepth sees the record.  No reference test covers
this, so the oracle's restatement of bam_tag2cigar is pinned by invariance: the same alignments written with and without
the CG encoding must give byte-identical outputs (CPU only; the GPU path is compared with the oracle in
test_gpu_cli_parity.py)."""
import subprocess

import pytest

from conftest import ROOT, run_tool

GEN = ROOT / "tools" / "gen_bam"


def _gen(path, args):
    subprocess.run(["make", "-s", "-C", str(ROOT), "tools/gen_bam"], check=True)
    subprocess.run([str(GEN)] + args + ["--threads", "4", "-o", str(path)], check=True, capture_output=True)


@pytest.mark.parametrize("cfg,flags", [
    (["--config", "c2", "--genome-fraction", "0.002"], ["--by", "500"]),
    (["--config", "c2", "--genome-fraction", "0.002"], ["-x", "-n", "--by", "500"]),
    (["--config", "c5", "--genome-fraction", "0.0004", "--level", "1", "--dense"], ["--by", "1000"]),
])
def test_oracle_is_invariant_under_cg_encoding(oracle_bin, tmp_path, cfg, flags):
    _gen(tmp_path / "plain.bam", cfg)
    _gen(tmp_path / "cg.bam", cfg + ["--cg-every", "2"])
    assert (tmp_path / "plain.bam").read_bytes() != (tmp_path / "cg.bam").read_bytes()
    run_tool(oracle_bin, flags + ["p", tmp_path / "plain.bam"], tmp_path)
    run_tool(oracle_bin, flags + ["c", tmp_path / "cg.bam"], tmp_path)
    n = 0
    for suf in (".mosdepth.summary.txt", ".mosdepth.global.dist.txt", ".mosdepth.region.dist.txt", ".regions.bed", ".per-base.bed"):
        a, b = tmp_path / ("p" + suf), tmp_path / ("c" + suf)
        assert a.exists() == b.exists()
        if a.exists():
            assert a.read_bytes() == b.read_bytes(), suf
            n += 1
    assert n >= 3


def test_placeholder_is_kept_when_the_tag_is_shorter(oracle_bin, tmp_path):
    """htslib does not move a CG tag that holds fewer operations than the core CIGAR (bam_tag2cigar: CG_len < n_cigar):
    a one-operation CIGAR behind the two-operation placeholder stays <l_seq>S<reflen>N, which covers nothing in default mode."""
    cfg = ["--config", "c2", "--genome-fraction", "0.002"]
    _gen(tmp_path / "plain.bam", cfg)
    _gen(tmp_path / "short.bam", cfg + ["--cg-every", "1", "--cg-short"])
    run_tool(oracle_bin, ["-n", "p", tmp_path / "plain.bam"], tmp_path)
    run_tool(oracle_bin, ["-n", "s", tmp_path / "short.bam"], tmp_path)
    bases = lambda f: int((tmp_path / f).read_text().splitlines()[1].split("\t")[2])
    assert 0 < bases("s.mosdepth.summary.txt") < 0.5 * bases("p.mosdepth.summary.txt")   # 85 % of the reads are 150M
