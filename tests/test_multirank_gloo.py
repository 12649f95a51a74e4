"""World-size-2 `gloo` test (CPU) of the N>1 host path: contig sharding, reading only a rank's byte ranges with
rebased virtual offsets, and the cross-rank reduction of per-contig results into the `total` rows.  The records of
each rank's compact buffer are decoded with zlib here (no GPU), which checks that every span begins on a record
boundary and that the shards partition the file's records exactly."""
import os
import struct
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _records_per_contig(buf, beg, end, n_targets):
    """Decode records from virtual offset beg to end in a (compact) BGZF buffer; returns counts per tid."""
    from mosdepth_b200 import hostio
    data = bytes(buf)
    pos, stream, first = beg >> 16, b"", True
    end_c, end_u = end >> 16, end & 0xFFFF
    for p, bl, xl, isz in hostio.iter_bgzf_blocks(data, pos):
        if p > end_c or (p == end_c and end_u == 0):
            break
        blk = hostio.inflate_block(data, p, bl, xl)
        if p == end_c:
            blk = blk[:end_u]
        if first:
            blk = blk[beg & 0xFFFF:]; first = False
        stream += blk
        if p == end_c:
            break
    counts = np.zeros(n_targets, dtype=np.int64)
    o = 0
    while o < len(stream):
        bs, tid = struct.unpack_from("<ii", stream, o)
        assert bs >= 32
        counts[tid] += 1
        o += 4 + bs
    assert o == len(stream)
    return counts


def _worker(rank, world, bam, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT))
    from mosdepth_b200 import hostio
    from mosdepth_b200.shard import contig_weights, load_compact, lpt_shards
    hdr = hostio.read_bam_header(bam); index = hostio.read_index(bam)
    shards = lpt_shards(contig_weights(index), world)
    mine = shards[rank]
    begs = [index[t].ref_beg for t in mine]; ends = [index[t].ref_end for t in mine]
    buf, nb, ne, used = load_compact(bam, begs, ends)
    counts = np.zeros(len(hdr.names), dtype=np.int64)
    for t, b, e in zip(mine, nb, ne):
        c = _records_per_contig(buf[:used], b, e, len(hdr.names))
        assert c.sum() == c[t] == index[t].n_mapped + index[t].n_unmapped   # the span holds exactly this contig's records
        counts += c
    # "totals": sum of per-contig counts, min/max of a per-rank extreme (what the dist/stat all-reduce does)
    tot = torch.from_numpy(counts.copy()); dist.all_reduce(tot)
    mn = torch.tensor([min(mine) if mine else 10**9]); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    mx = torch.tensor([max(mine) if mine else -1]); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((tot.numpy().tolist(), int(mn), int(mx), [len(s) for s in shards], int(used)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction(tmp_path):
    gen = ROOT / "tools" / "gen_bam"
    if not gen.exists():
        import __graft_entry__ as ge
        ge.build()
    bam = tmp_path / "g.bam"
    subprocess.run([str(gen), "--config", "c3", "--genome-fraction", "0.0005", "--threads", "4", "-o", str(bam)], check=True, capture_output=True)
    sys.path.insert(0, str(ROOT))
    from mosdepth_b200 import hostio
    index = hostio.read_index(bam)
    want = [c.n_mapped + c.n_unmapped for c in index]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, str(bam), port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot, mn, mx, sizes, used = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tot == want
    assert mn == 0 and mx == len(index) - 1
    assert sum(sizes) == sum(1 for w in want if w)
    assert used < os.path.getsize(bam)   # each rank reads only its own byte ranges


class _FakeShard:
    """numpy stand-in for a Context that holds one shard of a contig after msd_run(): `own` is the depth of the shard's own
    reads over the whole contig (zero before lo).  Implements the calls exchange_spill() makes."""

    def __init__(self, own, lo, hi):
        self.arr = own.astype(np.int32).copy(); self.lo, self.hi = lo, hi; self.final = False

    def contig_spill(self, tid):
        nz = np.nonzero(self.arr[self.hi:])[0]
        return self.hi, (int(nz[-1]) + 1 if nz.size else 0)

    def depth_copy(self, tid, pos, n):
        return self.arr[pos:pos + n].copy()

    def depth_add(self, tid, pos, cells):
        self.arr[pos:pos + cells.size] += cells

    def contig_finalize(self, tid):
        self.final = True


def _spill_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT))
    from mosdepth_b200.shard import exchange_spill
    tlen, read_len = 4000, 150
    rng = np.random.default_rng(5)
    starts = np.sort(rng.integers(0, tlen - read_len, 900))
    cuts = [0, 1024, 2048 + 512, tlen][: world + 1]; cuts[-1] = tlen
    lo, hi = cuts[rank], cuts[rank + 1]
    own = np.zeros(tlen + 1, dtype=np.int64)
    for s in starts[(starts >= lo) & (starts < hi)]:
        own[s] += 1; own[s + read_len] -= 1
    shard = _FakeShard(np.cumsum(own), lo, hi)

    def send(dst, a):
        dist.send(torch.tensor([a.size], dtype=torch.int64), dst)
        if a.size:
            dist.send(torch.from_numpy(a.copy()), dst)

    def recv(src):
        n = torch.zeros(1, dtype=torch.int64); dist.recv(n, src)
        t = torch.zeros(int(n), dtype=torch.int32)
        if int(n):
            dist.recv(t, src)
        return t.numpy()

    exchange_spill(shard, 0, rank - 1 if rank > 0 else None, rank + 1 if rank + 1 < world else None, send, recv, lo, hi)
    full = np.zeros(tlen + 1, dtype=np.int64)
    for s in starts:
        full[s] += 1; full[s + read_len] -= 1
    want = np.cumsum(full)[lo:hi]
    q.put((rank, bool(shard.final and np.array_equal(shard.arr[lo:hi], want))))
    dist.destroy_process_group()


def test_spill_exchange_between_ranks():
    """The depth carry across shard boundaries (shard.exchange_spill) over gloo send/recv, world size 3."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_spill_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == {0: True, 1: True, 2: True}
