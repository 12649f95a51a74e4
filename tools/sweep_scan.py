#!/usr/bin/env python
"""K6 A/B in one process: the variants of the scan launch (MSD_SCAN, see run_segments in msd_api.cu), interleaved.
Prints the device time between the events around run_segments (ms_scan) and a checksum of the window means (identical for every variant).
usage: python tools/sweep_scan.py [genome_fraction]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mosdepth_b200 as m
from mosdepth_b200 import hostio
import bench

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.12
bam, n_reads, n_bytes = bench.make_bam("c3", frac, 20261022, os.cpu_count() or 8)
hdr = hostio.read_bam_header(bam)
raw = np.fromfile(bam, dtype=np.uint8)
host = torch.empty(raw.size, dtype=torch.uint8, pin_memory=True); host.numpy()[:] = raw
ctx = m.Context(0); ctx.open_mem(host, hdr); ctx.set_filters(mode=m.MODE_FAST); ctx.prefetch_windows(500); ctx.stage()
for rep in range(3):
    for variant in (0, 6):
        os.environ["MSD_SCAN"] = str(variant)
        ms = []
        for _ in range(3):
            ctx.totals_reset(); info = ctx.run(); ms.append(info.ms_scan)
        cells = info.n_depth_cells
        chk = 0.0
        for t in range(len(hdr.names)):
            if ctx.contig_records(t)[0] >= 0:
                v, st, d = ctx.windows(t, 500); chk += float(v.sum())
        st, dist = ctx.contig_stats(0)
        print(f"rep {rep} MSD_SCAN={variant}: scan " + " ".join(f"{x:.3f}" for x in ms) + f" ms  -> {8 * cells / 1e9 / (min(ms) / 1e3):7.0f} GB/s best ({8 * cells / 1e9 / (min(ms) / 1e3) / 6540.8:.3f} of 6540.8)  checksum {chk:.6f} dist0 {int(dist[:8].sum())}", flush=True)
ctx.close()
