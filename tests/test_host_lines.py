"""CPU test of the host CLI's row formatting (mosdepth_b200/host/lines.hpp) against Python's own %-formatting, which is the
C printf the reference's Nim `formatFloat(ffDecimal)` / `$int` output is compared with in tests/test_oracle_golden.py."""
import random
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("lines")
    e = d / "lines_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", str(e), str(ROOT / "tools" / "lines_test.cpp")], check=True)
    return e


def test_rows_match_printf(exe):
    rng = random.Random(11)
    cases, want = [], []
    for i in range(3000):
        chrom = rng.choice(["chr1", "MT", "HLA-DRB1*15:01:01:01", "1"])
        s = rng.choice([0, 1, 9, 10, 99, 100, 16569, 2**31 - 2, 2**32 - 2, rng.randrange(0, 2**32 - 1)])
        e = rng.choice([s + 1, min(2**32 - 1, s + 500), 2**32 - 1])
        name = rng.choice([".", "-", "aregion", "exon_17|ENST0001"])
        v = rng.choice([0.0, 0.005, 0.015, 0.35555555555555546, 30.125, 1e6 / 3, rng.random() * rng.choice([1, 100, 1e5])])
        p = rng.choice([2, 2, 2, 0, 1, 7, 12])
        shown = "" if name in (".", "-") else name + "\t"
        if i % 2 == 0:
            cases.append(f"R {chrom} {s} {e} {name} {v!r} {p}")
            want.append(f"{chrom}\t{s}\t{e}\t{shown}{v:.{p}f}\n")
        else:
            n = rng.choice([-1, 0, 1, 4])
            cnt = [rng.choice([0, 1, 500, 2**40]) for _ in range(max(n, 0))]
            cases.append(f"T {chrom} {s} {e} {name} {n} " + " ".join(map(str, cnt)))
            tail = "\t0\t0\t0" if n < 0 else "".join(f"\t{c}" for c in cnt)
            want.append(f"{chrom}\t{s}\t{e}\t{'unknown' if name in ('.', '-') else name}{tail}\n")
    r = subprocess.run([str(exe)], input="\n".join(cases) + "\n", capture_output=True, text=True, check=True)
    assert r.stdout == "".join(want)


def test_window_intervals_follow_window_gen(exe):
    """mosdepth.nim:382-388: [k*w, (k+1)*w) while it fits, the last window ends at tlen; 64-bit products (k*w can pass 2^32)."""
    cases, want = [], []
    for tlen, w in [(16569, 500), (1000, 500), (1001, 500), (2**32 - 1, 2**31), (600000045, 500), (5, 100000000)]:
        n = (tlen + w - 1) // w
        for k in sorted({0, 1, n // 2, n - 2, n - 1} & set(range(n))):
            cases.append(f"W {k} {w} {tlen}")
            want.append(f"{k * w} {min((k + 1) * w, tlen)}\n")
    r = subprocess.run([str(exe)], input="\n".join(cases) + "\n", capture_output=True, text=True, check=True)
    assert r.stdout == "".join(want)
