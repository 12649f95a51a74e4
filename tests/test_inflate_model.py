"""CPU check of mosdepth_b200/csrc/inflate_core.cuh (table-entry encodings, canonical bookkeeping, token format) and of
the resolve kernel's window schedule: tools/inflate_model replays the kernels' arithmetic lane by lane and compares
with zlib.  Not the product path (the kernels are tested on the GPU in test_gpu_stages.py)."""
import subprocess
import json

import pytest

from conftest import ROOT, FIX
from bgzf_cases import basic_payloads, bgzf, tricky_members


@pytest.fixture(scope="module")
def model():
    exe = ROOT / "tools" / "inflate_model"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", "-o", str(exe), str(ROOT / "tools" / "inflate_model.cpp"), "-lz"], check=True)
    return exe


def _run(model, path):
    r = subprocess.run([str(model), str(path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout)


@pytest.mark.parametrize("name", ["ovl.bam", "empty-tids.bam", "nanopore.bam", "big.bam"])
def test_model_on_reference_fixtures(model, name):
    assert _run(model, FIX / name)["bad"] == 0


@pytest.mark.parametrize("level", [1, 9, "fixed", "stored", "rle", "huffman"])
def test_model_on_synthetic_blocks(model, tmp_path, level):
    f = tmp_path / "b.gz"
    f.write_bytes(bgzf(basic_payloads(n_blocks=60, max_len=65000), level))
    out = _run(model, f)
    assert out["bad"] == 0 and out["blocks"] == 60


def test_model_on_tricky_members(model, tmp_path):
    cases = tricky_members()
    f = tmp_path / "t.gz"
    f.write_bytes(b"".join(m for _, m in cases))
    out = _run(model, f)
    assert out["bad"] == 0 and out["blocks"] == len(cases)


def test_crc_model_every_size_class_and_alignment(model):
    """K1c (bgzf_crc_kernel): the row-strided CRC-32 arithmetic of inflate_core.cuh against zlib's crc32 for block sizes
    around every row / tail / lane boundary at all 16 alignments of the block inside the inflated buffer."""
    r = subprocess.run([str(model), "--crc-selftest"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["bad"] == 0 and out["crc_cases"] > 400
