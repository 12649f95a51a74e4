"""Intra-contig sharding on one GPU: the shards of a contig are run one after another by separate contexts
(the way separate ranks would), the depth each shard leaves beyond its end is handed to the next one, and the shard results
-- depth arrays, contig stats and dist, window values -- put together must equal the unsharded run bit for bit."""
import numpy as np
import pytest

from conftest import ROOT, run_tool

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bam(tmp_path_factory):
    d = tmp_path_factory.mktemp("intra")
    p = d / "c2.bam"
    run_tool(ROOT / "tools" / "gen_bam", ["--config", "c2", "--genome-fraction", "0.05", "--threads", "4", "-o", p], d)
    return p


@pytest.mark.parametrize("mode_name", ["fast", "default"])
@pytest.mark.parametrize("n_parts", [2, 3])
def test_sharded_contig_equals_unsharded(bam, mode_name, n_parts):
    import mosdepth_b200 as m
    from mosdepth_b200 import hostio
    from mosdepth_b200.shard import split_contig
    mode = m.MODE_FAST if mode_name == "fast" else m.MODE_DEFAULT
    window = 500
    hdr = hostio.read_bam_header(bam); index = hostio.read_index(bam)
    tid = max(range(len(hdr.names)), key=lambda t: index[t].n_mapped)
    tlen = hdr.lengths[tid]
    data = np.frombuffer(bam.read_bytes(), dtype=np.uint8)

    ref = m.Context(0)
    ref.open_mem(data, hdr); ref.set_filters(mode=mode); ref.run()
    want_depth = ref.debug_depth(tid)[:tlen]
    want_stat, want_dist = ref.contig_stats(tid)
    want_vals, want_rst, want_rdist = ref.windows(tid, window)
    ref.close()

    parts = split_contig(index[tid], tlen, n_parts, window=window)
    assert len(parts) == n_parts and parts[0][0] == 0 and parts[-1][1] == tlen
    ctxs, spills = [], []
    for lo, hi, vb, ve in parts:
        assert lo % 32 == 0 and lo % window == 0
        c = m.Context(0)
        c.open_mem(data, hdr)
        want = np.zeros(len(hdr.names), dtype=np.uint8); want[tid] = 1
        c.set_targets(want); c.set_spans([vb], [ve]); c.set_filters(mode=mode)
        c.set_contig_range(tid, lo, hi)
        c.run()
        start, n = c.contig_spill(tid)
        assert start == hi and (n > 0 or hi == tlen)
        spills.append(c.depth_copy(tid, start, n) if n else np.zeros(0, dtype=np.int32))
        ctxs.append(c)
    vals, dist, cum_len, cum_depth, mn, mx, n_rec = [], np.zeros(0, dtype=np.int64), 0, 0, 2**32, 0, 0
    rdist = np.zeros(512, dtype=np.int64)
    for i, (c, (lo, hi, _vb, _ve)) in enumerate(zip(ctxs, parts)):
        if i > 0 and spills[i - 1].size:
            c.depth_add(tid, lo, spills[i - 1])
        c.contig_finalize(tid)
        assert np.array_equal(c.debug_depth(tid)[lo:hi], want_depth[lo:hi]), f"depth of shard {i} differs"
        st, d = c.contig_stats(tid)
        cum_len += st.cum_length; cum_depth += st.cum_depth; mn = min(mn, st.min_depth); mx = max(mx, st.max_depth)
        dd = np.zeros(max(dist.size, d.size), dtype=np.int64); dd[:dist.size] += dist; dd[:d.size] += d; dist = dd
        v, rst, rd = c.windows(tid, window)
        vals.append(v); rdist += rd
        n_rec += max(c.contig_records(tid)[1], 0)
        c.close()
    assert (cum_len, cum_depth, mn, mx) == (want_stat.cum_length, want_stat.cum_depth, want_stat.min_depth, want_stat.max_depth)
    assert np.array_equal(dist, want_dist)
    assert np.array_equal(np.concatenate(vals).view(np.uint64), want_vals.view(np.uint64))   # the same float64 bits
    assert np.array_equal(rdist, want_rdist)
    assert n_rec == index[tid].n_mapped + index[tid].n_unmapped


def test_shard_span_too_short_is_refused(bam):
    import mosdepth_b200 as m
    from mosdepth_b200 import hostio
    from mosdepth_b200.shard import split_contig
    hdr = hostio.read_bam_header(bam); index = hostio.read_index(bam)
    tid = max(range(len(hdr.names)), key=lambda t: index[t].n_mapped)
    tlen = hdr.lengths[tid]
    lo, hi, vb, _ve = split_contig(index[tid], tlen, 2, window=500)[0]
    ve = index[tid].linear[max(1, hi // 16384 - 4)]   # ends before the cut: not every owned record is inside
    c = m.Context(0)
    c.open_mem(np.frombuffer(bam.read_bytes(), dtype=np.uint8), hdr)
    want = np.zeros(len(hdr.names), dtype=np.uint8); want[tid] = 1
    c.set_targets(want); c.set_spans([vb], [ve]); c.set_filters(mode=m.MODE_FAST); c.set_contig_range(tid, lo, hi)
    with pytest.raises(m.MsdError) as e:
        c.run()
    assert e.value.code == -11
    c.close()
