#!/usr/bin/env python
"""K10 A/B on one GPU (for the next GPU session; nothing here has been measured yet): per-base output of a synthetic
30X BAM in default mode (BASELINE configs[1] shape), per contig:
  runs      msd_per_base_runs                       (K8 + 12 B/run D2H; the host still has to format and deflate)
  bgzf      msd_per_base_bgzf                       (K8 + K10 + ~6 B/line D2H; the host appends the bytes)
  zlib1     zlib level 1 over the same text on ONE host thread (what the reference's writer does, mosdepth.nim:650-651)
and checks that the device members decompress to the text of the runs.
usage: python tools/bench_bedgz.py [genome_fraction=0.02] [reps=3] [--json]      (--json: one JSON line at the end, for bench.py)"""
import gzip, json, os, sys, time, zlib
os.environ.setdefault("MSD_BGZF_MIN_RUNS", "0")      # time the device path on every contig, however short
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mosdepth_b200 as m
from mosdepth_b200 import hostio
import bench

as_json = "--json" in sys.argv
argv = [a for a in sys.argv[1:] if a != "--json"]
frac = float(argv[0]) if len(argv) > 0 else 0.02
reps = int(argv[1]) if len(argv) > 1 else 3
bam, n_reads, n_bytes = bench.make_bam("c3", frac, 20261022, os.cpu_count() or 8)
ctx = m.Context(0)
hdr = ctx.open(bam)
ctx.set_filters(mode=m.MODE_DEFAULT)
ctx.run()
EOF_BLOCK = bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
tot = dict(runs=0.0, bgzf=0.0, runs_dev_ms=0.0, bgzf_dev_ms=0.0, zlib1=0.0, lines=0, text=0, dev_bytes=0, zlib_bytes=0, members=0, contigs=0, verified_contigs=0)
for tid, name in enumerate(hdr.names):
    rc, n = ctx.contig_records(tid)
    if rc == -2:
        continue
    t_runs = t_bgzf = d_runs = d_bgzf = 1e9
    for _ in range(reps):      # wall = the call as a host sees it; dev = CUDA events on the library's compute stream around it
        ctx.mark(0); t0 = time.perf_counter(); s, e, d = ctx.per_base_runs(tid); t_runs = min(t_runs, time.perf_counter() - t0); ctx.mark(1)
        t0 = time.perf_counter(); data, mem, _ = ctx.per_base_bgzf(tid, name); t_bgzf = min(t_bgzf, time.perf_counter() - t0); ctx.mark(2)
        d_runs = min(d_runs, ctx.elapsed_ms(0, 1)); d_bgzf = min(d_bgzf, ctx.elapsed_ms(1, 2))
    text = "".join(f"{name}\t{a}\t{b}\t{c}\n" for a, b, c in zip(s.tolist(), e.tolist(), d.tolist())).encode() if len(s) <= 3_000_000 else None
    if text is not None:
        assert gzip.decompress(data + EOF_BLOCK) == text, name
        t0 = time.perf_counter(); z = zlib.compress(text, 1); t_z = time.perf_counter() - t0
        tot["zlib1"] += t_z; tot["zlib_bytes"] += len(z); tot["text"] += len(text); tot["verified_contigs"] += 1
    tot["runs"] += t_runs; tot["bgzf"] += t_bgzf; tot["runs_dev_ms"] += d_runs; tot["bgzf_dev_ms"] += d_bgzf
    tot["lines"] += len(s); tot["dev_bytes"] += len(data); tot["members"] += len(mem); tot["contigs"] += 1
    print(f"{name:>6} {len(s):>10} lines  runs {t_runs * 1e3:8.2f} ms  bgzf {t_bgzf * 1e3:8.2f} ms  {len(data) / max(len(s), 1):5.2f} B/line in {len(mem)} members", file=sys.stderr, flush=True)
out = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in tot.items()}
out.update({"workload": f"C3 synthetic 30X WGS x {frac:g}, default mode, per-base output of every contig", "reads": n_reads,
            "what": "runs = msd_per_base_runs (K8, 12 B/run to the host, text + deflate left to the host); bgzf = msd_per_base_bgzf (K8 + K10: the same runs plus ready BGZF members); "
                    "zlib1 = zlib level 1 of the same text on one host thread (the reference's writer); seconds are host wall per call summed over contigs (they include the Python copy of "
                    "the results), *_dev_ms are CUDA events around the calls; every verified contig's members gunzip to the text of its runs"})
if as_json:
    print(json.dumps(out))
else:
    print(out)
