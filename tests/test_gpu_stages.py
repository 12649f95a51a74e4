"""GPU parity of single stages through the C ABI: inflate vs zlib, scan vs numpy, depth arrays vs the oracle."""
import zlib
from pathlib import Path

import numpy as np
import pytest

from conftest import run_tool

pytestmark = pytest.mark.gpu

FIXTURE_BAMS = ["ovl.bam", "big.bam", "empty-tids.bam", "nanopore.bam", "overlapping-pairs.bam", "full-fragment-pairs.bam",
                "chrom-end/minimal.bam"]


@pytest.fixture(scope="module")
def ctx():
    import mosdepth_b200 as m
    c = m.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", FIXTURE_BAMS)
def test_inflate_matches_zlib(ctx, fixtures, name):
    from mosdepth_b200 import hostio
    data = (fixtures / name).read_bytes()
    want = hostio.inflate_all(data)
    got = ctx.debug_inflate(data, len(want))
    assert got == want


from bgzf_cases import basic_payloads, bgzf as _bgzf, tricky_members


@pytest.mark.parametrize("level", [1, 6, 9, "fixed", "stored", "rle", "huffman"])
def test_inflate_synthetic_blocks(ctx, level):
    payloads = basic_payloads(max_len=65000 if level == "stored" else 65280)
    data = _bgzf(payloads, level)
    want = b"".join(payloads)
    assert ctx.debug_inflate(data, len(want)) == want


def test_inflate_tricky_members(ctx):
    """15-bit codes, several DEFLATE blocks + empty stored blocks per member, 258-byte matches at short periods,
    32 KB distances, ISIZE 65536, literal runs > 510 (skip tokens), 1-byte members and the empty EOF member."""
    cases = tricky_members()
    for i, (p, m) in enumerate(cases):
        assert ctx.debug_inflate(m, max(len(p), 1)) == p, f"case {i} alone"
    data = b"".join(m for _, m in cases)
    want = b"".join(p for p, _ in cases)
    assert ctx.debug_inflate(data, len(want)) == want


def test_inflate_many_blocks_lock_step(ctx):
    """More blocks than decoder lanes on the device with very uneven sizes: lanes finish at different times, take new
    tickets and rebuild tables while their neighbours keep decoding."""
    rng = np.random.default_rng(3)
    payloads = []
    for i in range(3000):
        n = int(rng.integers(1, 400)) if i % 3 else int(rng.integers(1, 65000))
        words = [rng.integers(65, 91, int(rng.integers(2, 9)), dtype=np.uint8).tobytes() for _ in range(50)]
        payloads.append(b"".join(words[int(j)] for j in rng.integers(0, 50, n // 4 + 1))[:n])
    data = _bgzf(payloads, 6)
    want = b"".join(payloads)
    assert ctx.debug_inflate(data, len(want)) == want


def test_inflate_checks_crc32(ctx):
    """A stored block whose payload byte is flipped still inflates to ISIZE bytes: only the CRC32 of the gzip trailer
    (which htslib's reader verifies) can tell."""
    import mosdepth_b200 as m
    from bgzf_cases import member, deflate
    p = bytes(range(256)) * 16
    good = member(p, deflate(p, "stored"))
    assert ctx.debug_inflate(good, len(p)) == p
    bad = bytearray(good); bad[18 + 5 + 100] ^= 0x01          # header 18 B, stored-block header 5 B, then the payload
    with pytest.raises(m.MsdError) as e:
        ctx.debug_inflate(bytes(bad), len(p))
    assert e.value.code == -14 and "status 10" in str(e.value)


@pytest.mark.parametrize("block_bytes", [1, 3, 4, 15, 16, 17, 496, 511, 512, 513, 1023, 4099, 65279, 65280, 65519, 65536])
def test_crc_kernel_matches_zlib(ctx, block_bytes):
    """K1c alone (msd_debug_crc32): every block size class around the row / lane / tail boundaries of the row-strided
    formulation, at several alignments of the first block (the following blocks then take every alignment by themselves)."""
    rng = np.random.default_rng(block_bytes)
    n_blocks = 3000 if block_bytes < 600 else 300
    data = rng.integers(0, 256, size=block_bytes * n_blocks - (block_bytes // 3), dtype=np.uint8)   # last block shorter
    for first_offset in (0, 1, 7, 13, 16, 31):
        got, _ = ctx.debug_crc32(data, block_bytes, first_offset)
        raw = data.tobytes()
        want = np.array([zlib.crc32(raw[i:i + block_bytes]) for i in range(0, len(raw), block_bytes)], dtype=np.uint32)
        assert got.size == want.size
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, f"first_offset {first_offset}: block {bad[0]} of {want.size}: {got[bad[0]]:08x} != {want[bad[0]]:08x}"


def test_crc_kernel_extremes(ctx):
    for fill in (0, 255):
        data = np.full(65536 * 3 + 5, fill, dtype=np.uint8)
        got, _ = ctx.debug_crc32(data, 65536, 9)
        raw = data.tobytes()
        assert [int(x) for x in got] == [zlib.crc32(raw[i:i + 65536]) for i in range(0, len(raw), 65536)]


def test_inflate_rejects_corrupt_block(ctx):
    import mosdepth_b200 as m
    p = bytes(range(256)) * 40
    data = bytearray(_bgzf([p], 6))
    data[40] ^= 0xFF; data[41] ^= 0x55; data[60] ^= 0xAA
    try:
        got = ctx.debug_inflate(bytes(data), len(p))
    except m.MsdError as e:
        assert e.code == -14
    else:
        assert got != p   # a corrupt stream may still decode to the right length; it must not be accepted as equal


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 4095, 4096, 4097, 1 << 20, (1 << 22) + 13])
def test_cumsum_matches_numpy(ctx, n):
    rng = np.random.default_rng(n)
    a = rng.integers(-3, 4, n, dtype=np.int32)
    got = ctx.debug_cumsum(a)
    assert np.array_equal(got, np.cumsum(a, dtype=np.int64).astype(np.int32))


def test_cumsum_wraps_like_int32(ctx):
    a = np.full(5000, 2**30, dtype=np.int32)
    want = (np.cumsum(a.astype(np.int64)) & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
    assert np.array_equal(ctx.debug_cumsum(a), want)


def _oracle_runs(oracle_bin, tmp_path, bam, args):
    run_tool(oracle_bin, ["o"] + args + [bam], tmp_path)
    per = {}
    for line in (tmp_path / "o.per-base.bed").read_text().splitlines():
        c, s, e, d = line.split("\t")
        per.setdefault(c, []).append((int(s), int(e), int(d)))
    return per


MODES = [("default", [], 0), ("fast", ["-x"], 1), ("fragment", ["-a"], 2)]


@pytest.mark.parametrize("name", FIXTURE_BAMS)
@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
def test_per_base_runs_match_oracle(ctx, oracle_bin, fixtures, tmp_path, name, mode):
    import mosdepth_b200 as m
    bam = fixtures / name
    per = _oracle_runs(oracle_bin, tmp_path, bam, mode[1])
    hdr = ctx.open(bam)
    ctx.set_filters(mode=mode[2])
    ctx.run()
    n_with = 0
    for tid, cname in enumerate(hdr.names):
        rc, n = ctx.contig_records(tid)
        want = per[cname]
        if rc == -2:
            assert want == [(0, hdr.lengths[tid], 0)], cname
            continue
        n_with += 1
        s, e, d = ctx.per_base_runs(tid)
        got = list(zip(s.tolist(), e.tolist(), d.tolist()))
        assert got == want, cname
    assert n_with >= 1


def test_sharded_compact_buffers_match_whole_file(ctx, tmp_path):
    """Emulates the N-rank path on one GPU: every 'rank' loads only its contigs' byte ranges (rebased offsets,
    merged neighbours) and must reproduce the whole-file per-contig results."""
    import subprocess
    import mosdepth_b200 as m
    from conftest import ROOT
    from mosdepth_b200 import hostio
    from mosdepth_b200.shard import contig_weights, load_compact, lpt_shards, merge_adjacent_spans
    bam = tmp_path / "s.bam"
    subprocess.run([str(ROOT / "tools" / "gen_bam"), "--config", "c3", "--genome-fraction", "0.001", "--threads", "8", "-o", str(bam)], check=True, capture_output=True)
    hdr = ctx.open(bam)
    ctx.set_filters(mode=m.MODE_FAST)
    ctx.run()
    whole = {}
    for t in range(len(hdr.names)):
        rc, n = ctx.contig_records(t)
        st, dist = ctx.contig_stats(t)
        vals, rst, rd = ctx.windows(t, 500)
        whole[t] = (n, st.cum_depth, st.max_depth, dist.tolist(), vals.tobytes(), rd.tolist())
    index = hostio.read_index(bam)
    for world in (2, 3):
        seen = set()
        for rank, mine in enumerate(lpt_shards(contig_weights(index), world)):
            begs, ends = merge_adjacent_spans([index[t].ref_beg for t in mine], [index[t].ref_end for t in mine])
            buf, nb, ne, used = load_compact(bam, begs, ends)
            c2 = m.Context(0)
            c2.open_mem(buf, hdr)
            want = np.zeros(len(hdr.names), dtype=np.uint8); want[mine] = 1
            c2.set_targets(want); c2.set_spans(nb, ne); c2.set_filters(mode=m.MODE_FAST)
            c2.run()
            for t in mine:
                rc, n = c2.contig_records(t)
                st, dist = c2.contig_stats(t)
                vals, rst, rd = c2.windows(t, 500)
                assert (n, st.cum_depth, st.max_depth, dist.tolist(), vals.tobytes(), rd.tolist()) == whole[t], (world, rank, t)
                seen.add(t)
            c2.close()
        assert seen == set(range(len(hdr.names)))
