#!/usr/bin/env python
"""Where the time of the per-contig queries goes on C2 (30X chr20, default mode): wall-clock per call with a device synchronize on both
sides, and the library's own breakdown of the run calls (MSD_TRACE).  usage: python tools/time_queries.py [genome_fraction]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mosdepth_b200 as m
from mosdepth_b200 import hostio
import bench

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
bam, n_reads, n_bytes = bench.make_bam("c2", frac, 20261022, os.cpu_count() or 8)
hdr = hostio.read_bam_header(bam)
raw = np.fromfile(bam, dtype=np.uint8)
host = torch.empty(raw.size, dtype=torch.uint8, pin_memory=True); host.numpy()[:] = raw
ctx = m.Context(0); ctx.open_mem(host, hdr); ctx.set_filters(mode=m.MODE_DEFAULT); ctx.stage()


def wall(label, fn, reps=4):
    for i in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        print(f"{label} call {i}: {dt:8.3f} ms", flush=True)
    return r


for rep in range(2):
    ctx.totals_reset(); info = ctx.run()
    print(f"run {rep}: total {info.ms_total:.2f} ms, scan {info.ms_scan:.3f} ms, {info.n_depth_cells} cells", flush=True)
t = [i for i in range(len(hdr.names)) if ctx.contig_records(i)[0] >= 0][0]
wall("contig_stats", lambda: ctx.contig_stats(t))
os.environ["MSD_TRACE"] = "1"
a, b, c = wall("per_base_runs(copy=False)", lambda: ctx.per_base_runs(t, copy=False))
print("runs:", len(a), flush=True)
wall("quantized_runs(copy=False)", lambda: ctx.quantized_runs(t, [0, 1, 10, 30, 2**63 - 1], copy=False))
os.environ.pop("MSD_TRACE")
os.environ["MSD_BGZF_MIN_RUNS"] = "0"
wall("per_base_bgzf", lambda: ctx.per_base_bgzf(t, hdr.names[t]))
wall("windows(500)", lambda: ctx.windows(t, 500))
ctx.close()
