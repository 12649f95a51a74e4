"""The host CLI's default output path (`*.bed.gz` + `.csi`, written by bgzf_writer.hpp's thread pool) must decompress to
exactly the bytes the `--plain` path writes (which tests/test_gpu_cli_parity.py pins to the oracle), for every bed file
and for 1 and 5 writer threads.  The writer itself is covered on the CPU by tests/test_bgzf_writer.py."""
import gzip

import pytest

from conftest import ROOT, run_tool

pytestmark = pytest.mark.gpu

CLI = ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200"
BEDS = [".regions.bed", ".per-base.bed", ".quantized.bed", ".thresholds.bed"]
CASES = [("big.bam", ["--by", "500", "-T", "1,5", "-q", "0:1:4:"]), ("empty-tids.bam", ["-b", "100", "-q", "0:1:5:150"]), ("nanopore.bam", ["-x", "--by", "250"])]


@pytest.mark.parametrize("name,flags", CASES, ids=[c[0] for c in CASES])
def test_gz_outputs_decompress_to_the_plain_outputs(fixtures, tmp_path, name, flags):
    bam = fixtures / name
    run_tool(CLI, ["--plain"] + flags + ["p", bam], tmp_path)
    run_tool(CLI, flags + ["z1", bam], tmp_path, env={"MOSDEPTH_B200_WRITER_THREADS": "1"})
    run_tool(CLI, flags + ["z5", bam], tmp_path, env={"MOSDEPTH_B200_WRITER_THREADS": "5"})
    n = 0
    for s in BEDS:
        plain = tmp_path / ("p" + s)
        if not plain.exists():
            continue
        z1, z5 = (tmp_path / ("z1" + s + ".gz")).read_bytes(), (tmp_path / ("z5" + s + ".gz")).read_bytes()
        assert z1 == z5, s
        assert gzip.decompress(z1) == plain.read_bytes(), s
        assert gzip.decompress((tmp_path / ("z1" + s + ".gz.csi")).read_bytes())[:4] == b"CSI\x01", s
        n += 1
    assert n >= 2
    for s in [".mosdepth.summary.txt", ".mosdepth.global.dist.txt"]:
        assert (tmp_path / ("z1" + s)).read_bytes() == (tmp_path / ("p" + s)).read_bytes()
