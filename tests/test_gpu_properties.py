"""Bench-size checks (BASELINE config C3 at the scale bench.py runs, 18.5 M reads / 88 k BGZF blocks / four full chunks in
flight): byte-for-byte parity of `--by 500 -n -x` with the oracle, and size-independent properties of the results that need
no oracle at all -- conservation of bases between the three views of the depth (contig stats, dist histogram, window means),
run-to-run determinism under atomics, fast-mode depth dominating default-mode depth base by base."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_gpu_cli_parity import compare

pytestmark = pytest.mark.gpu

GEN = ROOT / "tools" / "gen_bam"
FRACTION = os.environ.get("MSD_TEST_FRACTION", "0.03")


@pytest.fixture(scope="module")
def big(tmp_path_factory):
    shm = "/dev/shm"      # the 1.7 GB file goes to memory when it can
    d = shm if os.path.isdir(shm) and os.access(shm, os.W_OK) else str(tmp_path_factory.mktemp("benchsize"))
    bam = os.path.join(d, f"msd_test_c3_{os.getpid()}.bam")
    subprocess.run([str(GEN), "--config", "c3", "--genome-fraction", FRACTION, "--threads", str(os.cpu_count() or 8), "-o", bam], check=True, capture_output=True)
    yield bam
    for f in (bam, bam + ".bai"):
        if os.path.exists(f):
            os.remove(f)


def test_bench_size_cli_parity(oracle_bin, big, tmp_path):
    compare(oracle_bin, tmp_path, big, ["-t", str(max(0, (os.cpu_count() or 2) - 1)), "--by", "500", "-n", "-x"])


def _run(bam, mode, window=500):
    import mosdepth_b200 as m
    from mosdepth_b200 import hostio
    hdr = hostio.read_bam_header(bam)
    ctx = m.Context(0)
    ctx.open(bam, hdr)
    ctx.set_filters(mode=mode)
    info = ctx.run()
    out = {"info": info.as_dict(), "contigs": {}}
    for t in range(len(hdr.names)):
        rc, n = ctx.contig_records(t)
        if rc < 0:
            continue
        st, dist = ctx.contig_stats(t)
        vals, rst, rdist = ctx.windows(t, window)
        out["contigs"][t] = dict(n=n, tlen=int(hdr.lengths[t]), stat=(st.cum_length, st.cum_depth, st.min_depth, st.max_depth), dist=dist,
                                 vals=vals, rstat=(rst.cum_length, rst.cum_depth, rst.min_depth, rst.max_depth), rdist=rdist)
    return ctx, out


def test_conservation_and_determinism(big):
    import mosdepth_b200 as m
    ctx, a = _run(big, m.MODE_FAST)
    assert len(a["contigs"]) == 25
    for t, c in a["contigs"].items():
        L, bases, lo, hi = c["stat"]
        tlen = c["tlen"]
        assert L == tlen
        # the dist histogram is the same array seen as counts: sum = tlen, first moment = bases, support = [min, max]
        d = c["dist"]
        assert int(d.sum()) == tlen
        assert int((d * np.arange(d.size, dtype=np.int64)).sum()) == bases
        nz = np.nonzero(d)[0]
        assert nz[0] == lo and nz[-1] == hi
        # windows: one value per 500 bases, the last one shorter; sum(mean * length) = bases up to float64 rounding
        nw = (tlen + 499) // 500
        assert c["vals"].size == nw
        lens = np.full(nw, 500.0); lens[-1] = tlen - 500 * (nw - 1)
        assert abs(float((c["vals"] * lens).sum()) - bases) <= 1e-9 * max(bases, 1) + 1e-3
        assert c["rstat"][0] == tlen and c["rstat"][1] == bases and c["rstat"][2] == lo and c["rstat"][3] == hi
        assert int(c["rdist"].sum()) == nw
        assert float(c["vals"].max()) <= hi + 1e-9 and float(c["vals"].min()) >= lo - 1e-9
    # fast mode: every read that passes the filters adds its reference span once
    assert a["info"]["n_events"] % 2 == 0 and a["info"]["n_events"] > 0
    ctx.close()
    # atomics commute: a second context gives bit-identical results (every integer and every float64)
    ctx2, b = _run(big, m.MODE_FAST)
    ctx2.close()
    assert a["info"]["n_events"] == b["info"]["n_events"] and a["info"]["n_records"] == b["info"]["n_records"]
    for t, c in a["contigs"].items():
        o = b["contigs"][t]
        assert c["stat"] == o["stat"] and c["rstat"] == o["rstat"] and c["n"] == o["n"]
        assert np.array_equal(c["dist"], o["dist"]) and np.array_equal(c["rdist"], o["rdist"])
        assert np.array_equal(c["vals"].view(np.uint64), o["vals"].view(np.uint64))


def test_fast_mode_dominates_default_mode(big):
    """Default mode removes CIGAR gaps and the overlap of mates from what fast mode counts: its depth is <= base by base."""
    import mosdepth_b200 as m
    cf, f = _run(big, m.MODE_FAST)
    cd, d = _run(big, m.MODE_DEFAULT)
    assert f["info"]["n_records"] == d["info"]["n_records"]
    less = 0
    for t in f["contigs"]:
        assert d["contigs"][t]["n"] == f["contigs"][t]["n"]
        assert d["contigs"][t]["stat"][1] <= f["contigs"][t]["stat"][1]
        if t in (0, 20, 24):        # chr1, chr21, chrM: every base
            a, b = cf.debug_depth(t), cd.debug_depth(t)
            assert a.min() >= 0 and b.min() >= 0
            assert bool((b <= a).all())
            less += int((b < a).sum())
    assert less > 0
    cf.close(); cd.close()
