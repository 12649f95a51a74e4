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

    parts5 = split_contig(index[tid], tlen, n_parts, window=window, with_cover=True)
    parts = [p[:4] for p in parts5]
    assert len(parts) == n_parts and parts[0][0] == 0 and parts[-1][1] == tlen
    ctxs, spills = [], []
    for lo, hi, vb, ve, cover in parts5:
        assert lo % 32 == 0 and lo % window == 0
        c = m.Context(0)
        c.open_mem(data, hdr)
        want = np.zeros(len(hdr.names), dtype=np.uint8); want[tid] = 1
        c.set_targets(want); c.set_spans([vb], [ve]); c.set_filters(mode=mode)
        c.set_contig_range(tid, lo, hi)
        c.set_contig_cover(tid, cover)
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


def test_undeclared_span_start_is_refused_in_default_mode(bam):
    """Default mode on a shard that starts inside its contig: unless the caller declares how far back its span reaches
    (msd_set_contig_cover), a read that finds no mate although the mate lies before the owned range fails the run (MSD_ESPAN)
    instead of silently skipping a possible overlap correction."""
    import mosdepth_b200 as m
    from mosdepth_b200 import hostio
    from mosdepth_b200.shard import split_contig
    hdr = hostio.read_bam_header(bam); index = hostio.read_index(bam)
    tid = max(range(len(hdr.names)), key=lambda t: index[t].n_mapped)
    lo, hi, vb, ve, cover = split_contig(index[tid], hdr.lengths[tid], 2, window=500, with_cover=True)[1]
    c = m.Context(0)
    c.open_mem(np.frombuffer(bam.read_bytes(), dtype=np.uint8), hdr)
    want = np.zeros(len(hdr.names), dtype=np.uint8); want[tid] = 1
    c.set_targets(want); c.set_spans([vb], [ve]); c.set_filters(mode=m.MODE_DEFAULT); c.set_contig_range(tid, lo, hi)
    with pytest.raises(m.MsdError) as e:
        c.run()
    assert e.value.code == -17
    need, covered = c.contig_span_need(tid)
    assert need < lo and covered == lo and need >= cover       # the mates are within the margin: declaring it makes the run pass
    c.set_contig_cover(tid, cover)
    c.run()
    c.close()


def test_mate_in_front_of_the_span_is_detected(tmp_path):
    """a proper pair whose first mate starts more than the margin before the cut (a spliced read, 150 kb N gap)
    and reaches across it.  A shard whose span does not hold that mate must not pass; with the span extended the sharded result equals
    the unsharded one (the overlap of the two mates is counted once)."""
    import mosdepth_b200 as m
    from mosdepth_b200 import hostio
    from bam_writer import record, write_bam
    tlen, cut = 400000, 200000
    recs = []
    for k in range(40):     # ordinary pairs on both sides of the cut, some overlapping
        p = 5000 + k * 9000
        recs.append((p, record(0, p, "100M", flag=99, mtid=0, mpos=p + 60, tlen=160, qname=f"p{k}")))
        recs.append((p + 60, record(0, p + 60, "100M", flag=147, mtid=0, mpos=p, tlen=-160, qname=f"p{k}")))
    recs.append((100000, record(0, 100000, "50M150000N50M", flag=99, mtid=0, mpos=250020, tlen=150120, qname="spliced")))
    recs.append((250020, record(0, 250020, "100M", flag=147, mtid=0, mpos=100000, tlen=-150120, qname="spliced")))
    recs.sort(key=lambda r: r[0])
    bam = tmp_path / "n.bam"
    first, voffs, end = write_bam(bam, [("c1", tlen)], [r for _, r in recs])
    hdr = hostio.read_bam_header(bam)
    data = np.frombuffer(bam.read_bytes(), dtype=np.uint8)
    ref = m.Context(0); ref.open_mem(data, hdr); ref.set_filters(mode=m.MODE_DEFAULT); ref.run()
    want_depth = ref.debug_depth(0)[:tlen]; ref.close()
    # mate 1 covers [100000, 100050) and [250050, 250100), mate 2 [250020, 250120): the overlap [250050, 250100) counts once
    assert [int(want_depth[p]) for p in (100010, 250010, 250030, 250060, 250110, 250130)] == [1, 0, 1, 1, 1, 0]

    def first_at_or_after(pos):
        return next(v for (p, _), v in zip(recs, voffs) if p >= pos)

    def shard2(span_from):
        c = m.Context(0); c.open_mem(data, hdr); c.set_filters(mode=m.MODE_DEFAULT)
        c.set_spans([first_at_or_after(span_from)], [0]); c.set_contig_range(0, cut, tlen); c.set_contig_cover(0, span_from)
        return c
    c = shard2(cut - 65536)                       # the usual margin: the spliced mate at 100,000 is not in the span
    with pytest.raises(m.MsdError) as e:
        c.run()
    assert e.value.code == -17 and c.contig_span_need(0)[0] == 100000
    c.close()
    # span extended as msd_contig_span_need() says: shard 1 + shard 2 == unsharded
    c1 = m.Context(0); c1.open_mem(data, hdr); c1.set_filters(mode=m.MODE_DEFAULT)
    c1.set_spans([first], [first_at_or_after(cut + 60000)]); c1.set_contig_range(0, 0, cut); c1.run()
    start, n = c1.contig_spill(0)
    spill = c1.depth_copy(0, start, n)
    c1.contig_finalize(0)
    c2 = shard2(100000); c2.run()
    c2.depth_add(0, cut, spill); c2.contig_finalize(0)
    got = np.concatenate([c1.debug_depth(0)[:cut], c2.debug_depth(0)[cut:tlen]])
    assert np.array_equal(got, want_depth)
    c1.close(); c2.close()
