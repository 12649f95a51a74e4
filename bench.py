#!/usr/bin/env python
"""bench.py -- headline benchmark of the mosdepth depth hot path on B200.

Metric (BASELINE.json): reads/s (and wall seconds) for a synthetic 30X whole-genome BAM with `--by 500 -n -x`
(config C3: fast mode, 500-bp windows, no per-base output), contigs sharded across 1/2/4/8 GPUs.

A "step" is one full pass of the hot path over the synthetic BAM: BGZF inflate -> record framing -> filters ->
+1/-1 depth events -> per-contig prefix sum + contig stats + global dist -> window means / window stats /
window dist for every contig (what the reference's main loop computes for `--by 500 -n -x`, mosdepth.nim:691-764),
plus the NCCL all-reduce of the `total` histograms/stats when N > 1.

  value : compressed BAM bytes already staged in HBM (msd_stage) when the timed region starts.
  e2e   : the same pass from a pinned HOST buffer holding the BAM file (host->device copies of the compressed
          bytes and device->host copies of every result inside the timed region), through the C ABI.
  --impl reference : the CPU oracle (oracle/mosdepth_oracle, a restatement of mosdepth; the real binary cannot be
          built here) on the same BAM with all host threads, `-t nproc-1 --by 500 -n -x`.

The genome is scaled by --genome-fraction (25 GRCh38 contigs, 30X, 150-bp pairs are kept) because a full 30X
WGS BAM (~60 GB) cannot be generated on the box inside the bench budget; reads/s is what is reported.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def sh(cmd, **kw):
    return subprocess.run(cmd, check=True, capture_output=True, text=True, **kw)


def ensure_built():
    if not (ROOT / "mosdepth_b200" / "libmosdepth_b200.so").exists() or not (ROOT / "tools" / "gen_bam").exists() \
            or not (ROOT / "oracle" / "mosdepth_oracle").exists():
        import __graft_entry__ as ge
        ge.build()


def _candidates():
    for d in (os.environ.get("MSD_BENCH_DIR"), "/dev/shm", "/tmp"):
        if d and os.path.isdir(d) and os.access(d, os.W_OK):
            p = Path(d) / "msd_bench"
            p.mkdir(exist_ok=True)
            yield p


def data_dir(need_bytes=0, keep_tag=None):
    """A writable directory (memory-backed when possible) with room for `need_bytes`; BAMs cached for other sizes are
    dropped before a slower directory is tried."""
    cands = list(_candidates())
    if not cands:
        raise RuntimeError("no writable data dir")
    free = lambda p: os.statvfs(p).f_bavail * os.statvfs(p).f_frsize
    for p in cands:
        if free(p) >= need_bytes:
            return p
    for p in cands:
        for f in p.iterdir():
            if f.is_file() and (keep_tag is None or not f.name.startswith(keep_tag)):
                f.unlink()
        if free(p) >= need_bytes:
            return p
    raise RuntimeError(f"no data dir with {need_bytes / 1e9:.1f} GB free")


def make_bam(config, fraction, seed, threads, level=6):
    """Generate (or reuse) the synthetic BAM. Returns (path, n_reads, n_bytes)."""
    tag = f"{config}_f{fraction:g}_s{seed}_l{level}"
    for d in _candidates():      # generated earlier (by this rank or by rank 0)
        bam, meta = d / f"{tag}.bam", d / f"{tag}.json"
        if bam.exists() and meta.exists() and Path(str(bam) + ".bai").exists():
            info = json.loads(meta.read_text())
            return bam, info["reads"], info["bytes"]
    d = data_dir(int(fraction * 60e9 * 1.1) + (64 << 20), keep_tag=tag)     # a 30X genome is ~56 GB of BAM at level 6
    bam, meta = d / f"{tag}.bam", d / f"{tag}.json"
    tmp = d / f"{tag}.tmp.{os.getpid()}.bam"
    r = sh([str(ROOT / "tools" / "gen_bam"), "--config", config, "--genome-fraction", str(fraction), "--seed", str(seed),
            "--threads", str(threads), "--level", str(level), "-o", str(tmp)])
    info = json.loads(r.stdout.strip().splitlines()[-1])
    os.replace(str(tmp) + ".bai", str(bam) + ".bai")
    os.replace(tmp, bam)
    meta.write_text(json.dumps(info))
    return bam, info["reads"], info["bytes"]


class ClockSampler(threading.Thread):
    """Samples SM clocks + throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def run_reference(args):
    """CPU arm: the oracle restatement with all host threads on the same BAM."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ensure_built()
    ncores = os.cpu_count() or 1
    threads = max(0, min(ncores - 1, 64))
    # same workload as our arm at this N (weak scaling: genome_fraction x N), bounded so that W+K steps end in minutes
    frac = args.cpu_fraction or min(args.genome_fraction * max(args.gpus, 1), 0.06)   # N = 1: the same BAM as our arm
    bam, n_reads, n_bytes = make_bam(args.config, frac, args.seed, ncores)
    outp = data_dir() / "ref_out"
    cmd = [str(ROOT / "oracle" / "mosdepth_oracle"), "-t", str(threads), "-n", "-x", "--by", "500", str(outp), str(bam)]
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        subprocess.run(cmd, check=True, capture_output=True)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    val = n_reads / sec
    sample = f"{args.config} synthetic 30X WGS, genome_fraction={frac:g}, {n_reads} reads, {n_bytes} BAM bytes, whole file per step"
    line = {"metric": "reads_per_s", "value": val, "unit": "reads/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"C3 synthetic 30X WGS (25 GRCh38 contigs x {frac:g}, 150-bp pairs) --by 500 -n -x", "genome_fraction": frac, "reads": n_reads, "bam_bytes": n_bytes,
                       "flags": "--by 500 -n -x", "note": "CPU oracle restatement of mosdepth (zlib inflate threads + single record thread); the mosdepth binary cannot be built in this image"},
            "cpu_baseline": {"value": val, "unit": "reads/s", "cores": threads + 1, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3")
    ap.add_argument("--genome-fraction", type=float, default=float(os.environ.get("MSD_BENCH_FRACTION", "0.06")))
    ap.add_argument("--cpu-fraction", type=float, default=0.0, help="genome fraction for the CPU arm (default: same BAM)")
    ap.add_argument("--seed", type=int, default=20261022)
    ap.add_argument("--window", type=int, default=500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the side measurement of K10 (per-base.bed.gz on the device), which runs in a process of its own after the timed region")
    ap.add_argument("--sharding", default="offsets", choices=["contigs", "offsets", "intra"],
                    help="N > 1: whole contigs per rank (LPT, no data-path exchange); N equal BGZF offset ranges of the file, cut inside contigs (the depth that reads leave beyond a cut goes to the next rank over NCCL); or every contig cut into N ranges")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line (rank 0): everything else that native code prints to fd 1 while we run (NCCL's
    # version banner under NCCL_DEBUG=VERSION/WARN, for one) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ensure_built()
    if world > 1:
        dist.barrier()
    import mosdepth_b200 as m
    from mosdepth_b200 import hostio
    from mosdepth_b200.shard import contig_weights, load_compact, lpt_shards, merge_adjacent_spans, offset_shards, split_contig

    # ---- data: weak scaling = every rank count sees `fraction * world` of the genome, i.e. fixed work per GPU --------
    fraction = args.genome_fraction * world
    ncores = os.cpu_count() or 1
    if rank == 0:
        bam, n_reads, n_bytes = make_bam(args.config, fraction, args.seed, ncores)
    if world > 1:
        dist.barrier()
    bam, n_reads, n_bytes = make_bam(args.config, fraction, args.seed, ncores)
    hdr = hostio.read_bam_header(bam)
    index = hostio.read_index(bam)
    nt = len(hdr.names)
    weights = contig_weights(index)
    intra = args.sharding == "intra" and world > 1
    parts = {}      # tid -> (lo, hi) owned by this rank (intra-contig sharding)
    if intra:
        # every contig is cut into `world` position ranges at index windows; rank r owns range r of every contig
        mine, sb, se = [], [], []
        for t in range(nt):
            if weights[t] <= 0:
                continue
            pr = split_contig(index[t], hdr.lengths[t], world, window=args.window)
            if len(pr) != world:            # too short to cut (chrM): rank 0 takes it whole
                if rank == 0:
                    mine.append(t); sb.append(index[t].ref_beg); se.append(index[t].ref_end)
                continue
            lo, hi, vb, ve = pr[rank]
            mine.append(t); parts[t] = (lo, hi); sb.append(vb); se.append(ve)
        begs, ends = merge_adjacent_spans(sb, se)
    elif args.sharding == "offsets" and world > 1:
        intra = True
        mine, sb, se = [], [], []
        for t, lo, hi, vb, ve in offset_shards(index, hdr.lengths, world, window=args.window)[rank]:
            mine.append(t); sb.append(vb); se.append(ve)
            if lo > 0 or hi < hdr.lengths[t]:
                parts[t] = (lo, hi)
        begs, ends = merge_adjacent_spans(sb, se)
    else:
        shards = lpt_shards(weights, world)
        mine = shards[rank]
        begs, ends = merge_adjacent_spans([index[t].ref_beg for t in mine], [index[t].ref_end for t in mine])
    want = np.zeros(nt, dtype=np.uint8); want[mine] = 1

    # pinned host image of this rank's byte ranges of the BAM file (the e2e input), virtual offsets rebased onto it
    keep = {}

    def alloc(n):
        keep["t"] = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        return keep["t"].numpy()

    _buf, begs, ends, used = load_compact(bam, begs, ends, alloc)
    host = keep["t"]

    ctx = m.Context(local)
    ctx.open_mem(host, hdr)
    ctx.set_targets(want)
    ctx.set_spans(begs, ends)
    ctx.set_filters(mode=m.MODE_FAST)   # -x; defaults -F 1796 -Q 0
    ctx.prefetch_windows(args.window)   # --by 500: the host will ask for every contig's windows
    for t, (lo, hi) in parts.items():
        ctx.set_contig_range(t, lo, hi)
    g_ptr, r_ptr, n_bins, s_ptr = ctx.totals_device()

    class DevView:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}

    tot_g = torch.as_tensor(DevView(g_ptr, n_bins), device=f"cuda:{local}")
    tot_r = torch.as_tensor(DevView(r_ptr, n_bins), device=f"cuda:{local}")
    tot_s = torch.as_tensor(DevView(s_ptr, 8), device=f"cuda:{local}")

    d2h = {"bytes": 0}
    check = {}

    def exchange_spills():
        """Sharding inside contigs: the depth that this rank's reads leave beyond the end of a range it owns goes to the rank
        that owns the next range of that contig (always rank + 1): one NCCL send/recv pair per neighbour, straight out of /
        into GPU memory; then every shard is finalised."""
        dev = f"cuda:{local}"
        out_t = [t for t in sorted(parts) if parts[t][1] < hdr.lengths[t]]     # ranges that end inside their contig
        in_t = [t for t in sorted(parts) if parts[t][0] > 0]                   # ranges that start inside their contig
        n_out = torch.zeros(nt, dtype=torch.int64, device=dev)
        spill = {t: ctx.contig_spill(t) for t in out_t}
        for t in out_t:
            n_out[t] = spill[t][1]
        all_n = [torch.zeros_like(n_out) for _ in range(world)]
        dist.all_gather(all_n, n_out)
        total_out = int(n_out.sum())
        send_buf = torch.empty(max(total_out, 1), dtype=torch.int32, device=dev)
        off = 0
        for t in out_t:
            start, n = spill[t]
            if n:
                ctx.depth_copy(t, start, n, device_ptr=send_buf.data_ptr() + 4 * off)
            off += n
        n_in = [int(all_n[rank - 1][t]) for t in in_t] if rank > 0 else []
        recv_buf = torch.empty(max(sum(n_in), 1), dtype=torch.int32, device=dev)
        ops = []
        if rank + 1 < world and total_out:
            ops.append(dist.P2POp(dist.isend, send_buf[:total_out], rank + 1))
        if rank > 0 and sum(n_in):
            ops.append(dist.P2POp(dist.irecv, recv_buf[:sum(n_in)], rank - 1))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        torch.cuda.synchronize()
        off = 0
        for t, n in zip(in_t, n_in):
            if n:
                ctx.depth_add(t, parts[t][0], device_ptr=recv_buf.data_ptr() + 4 * off, n=n)
            off += n
        for t in sorted(parts):
            ctx.contig_finalize(t)
        return 4 * total_out

    def step():
        """One pass of the hot path for this rank's contigs + the cross-rank reduction of the totals."""
        ctx.totals_reset()
        info = ctx.run()
        nbytes = 0
        if parts:
            nbytes += exchange_spills()
        for t in mine:
            if t not in parts:
                rc, _n = ctx.contig_records(t)
                if rc < 0:
                    continue
            st, distv = ctx.contig_stats(t)
            vals, rst, rdist = ctx.windows(t, args.window)
            nbytes += vals.nbytes + rdist.nbytes + distv.nbytes + 48
        if world > 1:
            dist.all_reduce(tot_g[:1024]); dist.all_reduce(tot_r[:512])
            mm = torch.stack([tot_s[2], tot_s[6]]); dist.all_reduce(mm, op=dist.ReduceOp.MIN)
            mx = torch.stack([tot_s[3], tot_s[7]]); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sums = torch.stack([tot_s[0], tot_s[1], tot_s[4], tot_s[5]]); dist.all_reduce(sums)
            check["totals"] = [int(tot_g[:1024].sum()), int((tot_g[:1024] * torch.arange(1024, device=tot_g.device)).sum()), int(tot_r[:512].sum()), int(sums[0]), int(sums[1]), int(mm[0]), int(mx[0])]
        else:
            check["totals"] = [int(tot_g[:1024].sum()), int((tot_g[:1024] * torch.arange(1024, device=tot_g.device)).sum()), int(tot_r[:512].sum()), int(tot_s[0]), int(tot_s[1]), int(tot_s[2]), int(tot_s[3])]
        d2h["bytes"] = nbytes
        return info

    def timed(resident):
        if resident:
            ctx.stage()
        else:
            ctx.set_spans(begs, ends)   # drops the staged copy
        for _ in range(args.warmup):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local); sampler.start()
        infos = []
        lc0 = ctx.launch_count()
        t0 = time.perf_counter()
        ctx.mark(0)
        for _ in range(args.steps):
            infos.append(step())
        ctx.mark(1)
        torch.cuda.synchronize()
        ms_dev = ctx.elapsed_ms(0, 1)
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        sampler.stop_flag.set(); sampler.join()
        # the step makes blocking host calls (results come back per contig), so the device-side span between
        # the two marks is the honest step time; the host wall clock is reported beside it
        tt = torch.tensor([ms_dev, wall * 1e3], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt[0].item() / args.steps, tt[1].item() / args.steps, infos, ctx.launch_count() - lc0, sampler.summary()

    total_reads = n_reads
    ms_res, wall_res, infos_res, launches, clocks = timed(True)
    ms_e2e, wall_e2e, infos_e2e, _, clocks2 = timed(False)
    info = infos_res[-1]
    value = total_reads / (ms_res / 1e3)
    e2e_val = total_reads / (ms_e2e / 1e3)

    # roofline of the dominant kernel pair K1 = huffman_decode_kernel + lz77_resolve_kernel (BGZF inflate, > 80 % of the
    # step): algorithmic bytes per launch = compressed bytes of the chunk, read once
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    inf_ms = sum(i.ms_inflate_span for i in infos_res) / len(infos_res)   # launches of consecutive chunks overlap: use the span
    inf_launches = info.n_inflate_launches
    alg_bytes_per_launch = info.bytes_compressed / max(inf_launches, 1)
    achieved = (info.bytes_compressed / 1e9) / (inf_ms / 1e3) if inf_ms > 0 else 0.0
    # DRAM bytes per BGZF block from the ncu --set full capture profiles/r1j_inflate_full_summary.txt (85,062 blocks:
    # decode 5.30 GB read + 5.87 GB written, resolve 16.78 GB read + 5.45 GB written), scaled to this run's blocks per launch
    traffic_per_block = (5.298858e9 + 5.873518e9 + 16.775090e9 + 5.448415e9) / 85062.0
    roofline = {"kernel": "huffman_decode_kernel + lz77_resolve_kernel (K1, BGZF inflate)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic_per_block * info.n_blocks / max(inf_launches, 1),
                "traffic_source": "ncu --set full, profiles/r1j_inflate_full_summary.txt, per block x blocks per launch",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "launches_per_step": inf_launches, "ms_per_launch": inf_ms / max(inf_launches, 1),
                "inflated_gbs": (info.bytes_inflated / 1e9) / (inf_ms / 1e3) if inf_ms > 0 else 0.0,
                "stage_ms_per_step": {"inflate_span": inf_ms, "inflate_sum": sum(i.ms_inflate for i in infos_res) / len(infos_res), "frame": sum(i.ms_frame for i in infos_res) / len(infos_res),
                                      "records": sum(i.ms_records for i in infos_res) / len(infos_res), "scan": sum(i.ms_scan for i in infos_res) / len(infos_res),
                                      "zero": sum(i.ms_zero for i in infos_res) / len(infos_res), "run_total": sum(i.ms_total for i in infos_res) / len(infos_res)},
                "note": "inflate is bound by the serial Huffman chains (latency / issue), not by HBM; compressed GB/s and inflated GB/s are quoted against the HBM peak as north_star asks"}

    # K1c alone, timed live: the one kernel of K1 that is HBM bound (it reads the inflated bytes once: algorithmic bytes = U)
    try:
        if rank != 0:
            raise RuntimeError("rank 0 only")
        blk = 65280
        sample = np.random.default_rng(1).integers(0, 256, size=blk * 16384, dtype=np.uint8)      # 1.07 GB: far larger than L2
        _crcs, crc_ms = ctx.debug_crc32(sample, blk, 0, reps=10)
        roofline["k1c_crc"] = {"kernel": "bgzf_crc_kernel (K1c), timed alone over %d blocks of %d bytes" % (sample.size // blk, blk), "bound": "hbm",
                               "achieved": sample.size / 1e9 / (crc_ms / 1e3), "peak": peak, "unit": "GB/s", "frac": sample.size / 1e9 / (crc_ms / 1e3) / peak,
                               "ms_per_launch": crc_ms, "algorithmic_bytes_per_launch": int(sample.size),
                               "traffic": 5.504065e9 / 85062 * (sample.size // blk), "traffic_source": "ncu --set full, profiles/r1m_crc_full_summary.txt, per block x blocks"}
        del sample
    except Exception as e:      # never let the side measurement take the bench line down
        roofline["k1c_crc"] = {"error": str(e)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = max(0, min(ncores - 1, 64))
        outp = data_dir() / "cpu_out"
        cmd = [str(ROOT / "oracle" / "mosdepth_oracle"), "-t", str(threads), "-n", "-x", "--by", str(args.window), str(outp), str(bam)]
        subprocess.run(cmd, check=True, capture_output=True)   # warm page cache
        t0 = time.perf_counter(); subprocess.run(cmd, check=True, capture_output=True); dt = time.perf_counter() - t0
        cpu_baseline = {"value": n_reads / dt, "unit": "reads/s", "cores": threads + 1, "kind": "port", "seconds": dt,
                        "sample": f"whole bench BAM ({n_reads} reads, genome_fraction={fraction:g}), oracle -t {threads} --by {args.window} -n -x"}

    # side measurement, outside every timed region and in a process of its own (a failure there cannot touch this line):
    # K10 (msd_per_base_bgzf) against msd_per_base_runs + zlib on a small default-mode BAM, members verified against the runs
    aux = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_aux:
        try:
            r = subprocess.run([sys.executable, str(ROOT / "tools" / "bench_bedgz.py"), "0.004", "2", "--json"], capture_output=True, text=True, timeout=150)
            aux = {"k10_per_base_bgzf": json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 and r.stdout.strip() else {"error": (r.stderr or r.stdout)[-600:]}}
        except Exception as e:
            aux = {"k10_per_base_bgzf": {"error": str(e)}}

    if rank == 0:
        line = {"metric": "reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": f"C3 synthetic 30X WGS (25 GRCh38 contigs x {fraction:g}, 150-bp pairs) --by {args.window} -n -x", "genome_fraction": fraction,
                           "reads": n_reads, "bam_bytes": n_bytes, "flags": f"--by {args.window} -n -x", "sharding": (f"{world} equal BGZF offset ranges cut inside contigs, depth spill to the next rank over NCCL send/recv" if (args.sharding == "offsets" and world > 1) else f"every contig cut into {world} position ranges, depth spill over NCCL send/recv" if intra else f"contigs LPT over {world} GPU(s)"),
                           "totals_check": check.get("totals"),
                           "l2": "inputs (compressed BAM, depth arrays) are far larger than the 126 MB L2; no flush needed",
                           "wall_s_per_step": wall_res / 1e3, "extrapolated_full_genome_wall_s": (ms_res / 1e3) / fraction * world},
                "e2e": {"value": e2e_val, "unit": "reads/s", "ms_per_step": ms_e2e, "wall_ms_per_step": wall_e2e,
                        "h2d_bytes_per_step": int(infos_e2e[-1].bytes_compressed), "d2h_bytes_per_step": int(d2h["bytes"]),
                        "stage_ms_per_step": {"h2d": infos_e2e[-1].ms_h2d, "inflate": infos_e2e[-1].ms_inflate, "frame": infos_e2e[-1].ms_frame,
                                              "records": infos_e2e[-1].ms_records, "scan": infos_e2e[-1].ms_scan, "run_total": infos_e2e[-1].ms_total}},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "aux": aux}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
