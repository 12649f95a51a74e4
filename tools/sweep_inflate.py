#!/usr/bin/env python
"""Sweep the inflate pipeline's knobs on one GPU in one process (resident input, device-timed msd_run only):
chunk size (blocks per launch), decode warps per SM, resolve CTAs per SM, CRC on/off.  Prints one line per setting.
usage: python tools/sweep_inflate.py [genome_fraction] [knobs|pipe|repeat|one|ramp] [resident|e2e  (trace one run)]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mosdepth_b200 as m
from mosdepth_b200 import hostio
import bench

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.03
bam, n_reads, n_bytes = bench.make_bam("c3", frac, 20261022, os.cpu_count() or 8)
hdr = hostio.read_bam_header(bam)
raw = np.fromfile(bam, dtype=np.uint8)
host = torch.empty(raw.size, dtype=torch.uint8, pin_memory=True); host.numpy()[:] = raw


def timed(ctx, steps=4, warm=2):
    for _ in range(warm):
        ctx.run()
    torch.cuda.synchronize()
    ctx.mark(0)
    infos = [ctx.run() for _ in range(steps)]
    ctx.mark(1)
    torch.cuda.synchronize()
    return ctx.elapsed_ms(0, 1) / steps, sum(i.ms_inflate_span for i in infos) / steps


def with_ctx(env, fn, resident=True):
    for k, v in env.items():
        os.environ[k] = str(v)
    ctx = m.Context(0); ctx.open_mem(host, hdr); ctx.set_filters(mode=m.MODE_FAST); ctx.prefetch_windows(500)
    if resident:
        ctx.stage()
    try:
        return fn(ctx)
    finally:
        ctx.close()
        for k in env:
            os.environ.pop(k, None)


def sweep(ctx, label):
    for dc in (9, 8, 7, 6):
        for rc in (16, 12, 8, 6, 4):
            os.environ["MSD_DEC_CTAS"] = str(dc); os.environ["MSD_RESOLVE_CTAS"] = str(rc)
            ms, span = timed(ctx)
            print(f"{label} dec_ctas {dc} resolve_ctas {rc}: {ms:7.2f} ms/step  inflate span {span:7.2f}  {n_reads / ms / 1e3:7.1f} M reads/s", flush=True)
    os.environ.pop("MSD_DEC_CTAS", None); os.environ.pop("MSD_RESOLVE_CTAS", None)


mode = sys.argv[2] if len(sys.argv) > 2 else "knobs"
if mode == "knobs":
    for blocks in (24000, 12000, 16000, 32000, 44000):
        with_ctx({"MSD_CHUNK_BLOCKS": blocks, "MSD_CHUNK_INFL_MB": 3200, "MSD_CHUNK_COMP_MB": 1200}, lambda c: sweep(c, f"chunk {blocks}") if blocks == 24000 else print(f"chunk {blocks}: %.2f ms/step span %.2f" % timed(c), flush=True))
    with_ctx({"MSD_NO_CRC": 1}, lambda c: print("no crc: %.2f ms/step span %.2f" % timed(c), flush=True))
elif mode == "pipe":   # chunk slots x stream priority x chunk plan, input resident in HBM and crossing PCIe
    def show(label, resident):
        def f(c):
            ms, span = timed(c)
            print(f"{label} {'resident' if resident else 'e2e     '}: {ms:7.2f} ms/step  inflate span {span:7.2f}  {n_reads / ms / 1e3:7.1f} M reads/s", flush=True)
        return f
    for slots in (4, 8):
        for prio in (0, 1):
            for extra in ({}, {"MSD_CHUNK_BLOCKS": 32000, "MSD_CHUNK_INFL_MB": 2200, "MSD_CHUNK_COMP_MB": 800}):
                env = {"MSD_SLOTS": slots, "MSD_PRIO": prio, **extra}
                label = f"slots {slots} prio {prio} blocks {extra.get('MSD_CHUNK_BLOCKS', 24000)}"
                with_ctx(env, show(label, True), True)
                with_ctx(env, show(label, False), False)
    for r1 in (12000, 24000):
        env = {"MSD_SLOTS": 8, "MSD_PRIO": 1, "MSD_COPY_BLOCKS": r1}
        with_ctx(env, show(f"slots 8 prio 1 copy_blocks {r1}", False), False)


if mode == "repeat":    # the candidates for the defaults, interleaved three times (run-to-run noise is a few per cent)
    cands = [(4, 0, 24000), (4, 1, 24000), (8, 1, 24000), (4, 1, 32000), (8, 1, 32000), (8, 1, 44000)]
    for rep in range(3):
        for slots, prio, blocks in cands:
            env = {"MSD_SLOTS": slots, "MSD_PRIO": prio, "MSD_CHUNK_BLOCKS": blocks, "MSD_CHUNK_INFL_MB": 3000, "MSD_CHUNK_COMP_MB": 1100}
            def f(c, resident):
                ms, span = timed(c, 6, 2)
                print(f"rep {rep} slots {slots} prio {prio} blocks {blocks} {'resident' if resident else 'e2e     '}: {ms:7.2f} ms/step  span {span:7.2f}  {n_reads / ms / 1e3:7.1f} M reads/s", flush=True)
            with_ctx(env, lambda c: f(c, True), True)
            with_ctx(env, lambda c: f(c, False), False)


if mode == "one":       # the default configuration, three times (A/B of library builds)
    for rep in range(3):
        for resident in (True, False):
            def f(c):
                ms, span = timed(c, 6, 2)
                print(f"rep {rep} {'resident' if resident else 'e2e     '}: {ms:7.2f} ms/step  span {span:7.2f}  {n_reads / ms / 1e3:7.1f} M reads/s", flush=True)
            with_ctx({}, f, resident)


if mode == "ramp":      # how the first chunks of a copied input are sized (end-to-end only)
    for rep in range(2):
        for dbl, r0, cb in ((0, 6000, 16000), (1, 6000, 16000), (1, 3000, 16000), (1, 1500, 16000), (1, 3000, 24000), (0, 3000, 16000), (1, 3000, 12000)):
            def f(c):
                ms, span = timed(c, 6, 2)
                print(f"rep {rep} double {dbl} ramp0 {r0} copy_blocks {cb} e2e: {ms:7.2f} ms/step  span {span:7.2f}  {n_reads / ms / 1e3:7.1f} M reads/s", flush=True)
            with_ctx({"MSD_RAMP_DOUBLE": dbl, "MSD_RAMP0_BLOCKS": r0, "MSD_COPY_BLOCKS": cb}, f, False)


def tr(c):
    timed(c, 1, 2)
    os.environ["MSD_TRACE"] = "1"; c.run(); os.environ.pop("MSD_TRACE")


if len(sys.argv) > 3:
    with_ctx({"MSD_SLOTS": 8, "MSD_PRIO": 1} if mode == "pipe" else {}, tr, resident=sys.argv[3] == "resident")
