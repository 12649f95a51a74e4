#!/usr/bin/env python
"""K10 A/B on one GPU (for the next GPU session; nothing here has been measured yet): per-base output of a synthetic
30X BAM in default mode (BASELINE configs[1] shape), per contig:
  runs      msd_per_base_runs                       (K8 + 12 B/run D2H; the host still has to format and deflate)
  bgzf      msd_per_base_bgzf                       (K8 + K10 + ~6 B/line D2H; the host appends the bytes)
  zlib1     zlib level 1 over the same text on ONE host thread (what the reference's writer does, mosdepth.nim:650-651)
and checks that the device members decompress to the text of the runs.
usage: python tools/bench_bedgz.py [genome_fraction=0.02] [reps=3]"""
import gzip, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mosdepth_b200 as m
from mosdepth_b200 import hostio
import bench

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bam, n_reads, n_bytes = bench.make_bam("c3", frac, 20261022, os.cpu_count() or 8)
ctx = m.Context(0)
hdr = ctx.open(bam)
ctx.set_filters(mode=m.MODE_DEFAULT)
ctx.run()
EOF_BLOCK = bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
tot = dict(runs=0.0, bgzf=0.0, zlib1=0.0, lines=0, text=0, dev_bytes=0, zlib_bytes=0)
for tid, name in enumerate(hdr.names):
    rc, n = ctx.contig_records(tid)
    if rc == -2:
        continue
    t_runs = t_bgzf = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); s, e, d = ctx.per_base_runs(tid); t_runs = min(t_runs, time.perf_counter() - t0)
        t0 = time.perf_counter(); data, mem, _ = ctx.per_base_bgzf(tid, name); t_bgzf = min(t_bgzf, time.perf_counter() - t0)
    text = "".join(f"{name}\t{a}\t{b}\t{c}\n" for a, b, c in zip(s.tolist(), e.tolist(), d.tolist())).encode() if len(s) <= 3_000_000 else None
    if text is not None:
        assert gzip.decompress(data + EOF_BLOCK) == text, name
        t0 = time.perf_counter(); z = zlib.compress(text, 1); t_z = time.perf_counter() - t0
        tot["zlib1"] += t_z; tot["zlib_bytes"] += len(z); tot["text"] += len(text)
    tot["runs"] += t_runs; tot["bgzf"] += t_bgzf; tot["lines"] += len(s); tot["dev_bytes"] += len(data)
    print(f"{name:>6} {len(s):>10} lines  runs {t_runs * 1e3:8.2f} ms  bgzf {t_bgzf * 1e3:8.2f} ms  {len(data) / max(len(s), 1):5.2f} B/line in {len(mem)} members", flush=True)
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in tot.items()})
print("note: both timings include the Python copy of the result arrays; compare them with each other, not with bench.py")
