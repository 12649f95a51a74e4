#!/usr/bin/env python
"""bench.py -- headline benchmark of the mosdepth depth hot path on B200.

Metric (BASELINE.json): reads/s (and wall seconds) for a synthetic 30X whole-genome BAM with `--by 500 -n -x`
(config C3: fast mode, 500-bp windows, no per-base output) on 1/2/4/8 GPUs of one node.

A "step" is one full pass of the hot path over the synthetic BAM: BGZF inflate -> record framing -> filters ->
+1/-1 depth events -> prefix sum + contig stats + global dist of every contig (one launch) -> window means / window stats /
window dist of every contig (what the reference's main loop computes for `--by 500 -n -x`, mosdepth.nim:691-764); with
N > 1 GPUs the genome is cut into N equal BGZF offset ranges (cuts inside contigs), the depth that reads leave beyond a cut
goes to the next rank (ncclSend/ncclRecv inside the library, msd_exchange_spills) and the `total` rows are all-reduced
(msd_totals_allreduce).  The SAME BAM is used at every N ("scaling": "strong": the metric is one 30X genome on 1/2/4/8 GPUs).

  value : compressed BAM bytes already staged in HBM (msd_stage) when the timed region starts.
  e2e   : the same pass from a pinned HOST buffer holding this rank's bytes of the BAM file (host->device copies of the
          compressed bytes and device->host copies of every result inside the timed region), through the C ABI.
          e2e.cli_wall_s is the wall clock of the `mosdepth-b200` PROCESS on the same BAM (file in the page cache ->
          regions.bed.gz + .csi closed), which is what BASELINE.json's "wall s" names.
  --impl reference : the CPU oracle (oracle/mosdepth_oracle, a restatement of mosdepth; the real binary cannot be
          built here) with the best of T in {0, 3, 7, 15, nproc-1} inflate threads, on a BOUNDED SAMPLE of the same workload.

Before anything is timed, a small BAM of the same shape goes through the very same code path (all N ranks) and every output
file is rendered and compared byte for byte with the oracle's: `config.parity_checked`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SAMPLE_FRACTION = 0.03      # the bounded CPU sample: 18.5 M reads, ~1.7 GB of BAM, 2-4 s of oracle time per pass
PARITY_FRACTION = 0.004     # the small BAM every output of which is compared with the oracle before the timed region


def sh(cmd, **kw):
    return subprocess.run(cmd, check=True, capture_output=True, text=True, **kw)


def ensure_built():
    need = ["mosdepth_b200/libmosdepth_b200.so", "mosdepth_b200/bin/mosdepth-b200", "tools/gen_bam", "oracle/mosdepth_oracle"]
    if not all((ROOT / n).exists() for n in need):
        import __graft_entry__ as ge
        ge.build()


def _candidates():
    for d in (os.environ.get("MSD_BENCH_DIR"), "/dev/shm", "/tmp"):
        if d and os.path.isdir(d) and os.access(d, os.W_OK):
            p = Path(d) / "msd_bench"
            p.mkdir(exist_ok=True)
            yield p


def data_dir(need_bytes=0, keep_tags=()):
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
            if f.is_file() and not any(f.name.startswith(t) for t in keep_tags):
                f.unlink()
        if free(p) >= need_bytes:
            return p
    raise RuntimeError(f"no data dir with {need_bytes / 1e9:.1f} GB free")


def bam_tag(config, fraction, seed, level=6):
    return f"{config}_f{fraction:g}_s{seed}_l{level}"


def make_bam(config, fraction, seed, threads, level=None, keep_tags=()):
    """Generate (or reuse) the synthetic BAM (+ BAI, + the exon BED for c4: <bam>.bed). Returns (path, n_reads, n_bytes)."""
    level = level if level is not None else (1 if config == "c5" else 6)      # nanopore BAMs: QUAL barely compresses, level 1 like the fixture
    tag = bam_tag(config, fraction, seed, level)
    for d in _candidates():      # generated earlier (by this rank, by rank 0, or by an earlier N of the same scaling run)
        bam, meta = d / f"{tag}.bam", d / f"{tag}.json"
        if bam.exists() and meta.exists() and Path(str(bam) + ".bai").exists():
            info = json.loads(meta.read_text())
            return bam, info["reads"], info["bytes"]
    d = data_dir(int(fraction * 60e9 * 1.1) + (64 << 20), keep_tags=(tag,) + tuple(keep_tags))     # a 30X genome is ~56 GB of BAM at level 6
    bam, meta = d / f"{tag}.bam", d / f"{tag}.json"
    tmp = d / f"{tag}.tmp.{os.getpid()}.bam"
    t0 = time.perf_counter()
    r = sh([str(ROOT / "tools" / "gen_bam"), "--config", config, "--genome-fraction", str(fraction), "--seed", str(seed),
            "--threads", str(threads), "--level", str(level), "-o", str(tmp)] + (["--bed", str(tmp) + ".bed"] if config == "c4" else []))
    info = json.loads(r.stdout.strip().splitlines()[-1])
    info["gen_s"] = time.perf_counter() - t0
    if config == "c4":
        os.replace(str(tmp) + ".bed", str(bam) + ".bed")
    os.replace(str(tmp) + ".bai", str(bam) + ".bai")
    os.replace(tmp, bam)
    meta.write_text(json.dumps(info))
    return bam, info["reads"], info["bytes"]


def workload_config(args, world, parity):
    """The `config` object: the same keys and values in both arms (the reference arm runs a bounded sample OF THIS workload)."""
    f = args.genome_fraction
    return {"workload": f"C3 synthetic 30X WGS (25 GRCh38 contigs x {f:g} = {f * 3088286401 / 1e6:.0f} Mbp, 150-bp pairs, BGZF level 6) --by {args.window} -n -x",
            "genome_fraction": f, "flags": f"--by {args.window} -n -x", "window": args.window, "gpus": world,
            "l2": "inputs (compressed BAM, depth arrays) are far larger than the 126 MB L2; no flush needed",
            "parity_checked": bool(parity)}


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


# ---- the CPU arm ------------------------------------------------------------------------------------------------------------------
def clean_env():
    return {k: v for k, v in os.environ.items() if not k.startswith("MOSDEPTH_")}


def oracle_cmd(threads, window, outp, bam):
    return [str(ROOT / "oracle" / "mosdepth_oracle"), "-t", str(threads), "-n", "-x", "--by", str(window), str(outp), str(bam)]


def time_oracle(bam, window, t_list, repeats):
    """Process wall seconds of the oracle per thread count (best of `repeats`); the first run of all warms the page cache."""
    outp = data_dir() / f"cpu_out_{os.getpid()}"
    subprocess.run(oracle_cmd(t_list[0], window, outp, bam), check=True, capture_output=True, env=clean_env())
    best = {}
    for t in t_list:
        for _ in range(repeats):
            t0 = time.perf_counter()
            subprocess.run(oracle_cmd(t, window, outp, bam), check=True, capture_output=True, env=clean_env())
            dt = time.perf_counter() - t0
            best[t] = min(best.get(t, 1e30), dt)
    return best, outp


def thread_choices():
    n = os.cpu_count() or 1
    return sorted({t for t in (0, 3, 7, 15, n - 1) if 0 <= t <= max(n - 1, 0)})


def golden_check(tmp):
    """The CPU arm checks itself against the reference's own known answer before it is timed (functional-tests.sh:34-41,
    `ovl.bam --fast-mode`)."""
    fix = ROOT / "tests" / "golden" / "fixtures" / "ovl.bam"
    subprocess.run([str(ROOT / "oracle" / "mosdepth_oracle"), "-x", str(Path(tmp) / "g"), str(fix)], check=True, capture_output=True, env=clean_env())
    rows = [l.split("\t") for l in (Path(tmp) / "g.per-base.bed").read_text().splitlines() if l.startswith("MT\t")]
    return rows == [["MT", "0", "6", "1"], ["MT", "6", "42", "2"], ["MT", "42", "80", "1"], ["MT", "80", "16569", "0"]]


def run_reference(args):
    """CPU arm: the oracle restatement on a bounded sample of the workload, best thread count of {0, 3, 7, 15, nproc-1}."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ensure_built()
    ncores = os.cpu_count() or 1
    frac = min(args.cpu_fraction or SAMPLE_FRACTION, args.genome_fraction)
    bam, n_reads, n_bytes = make_bam(args.config, frac, args.seed, ncores, keep_tags=(bam_tag(args.config, args.genome_fraction, args.seed),))
    with tempfile.TemporaryDirectory() as tmp:
        parity = golden_check(tmp)
    # which -t is best on this box: one pass per candidate, best of 2 (these passes are also the warm-up)
    best, outp = time_oracle(bam, args.window, thread_choices(), 2)
    t_best = min(best, key=best.get)
    cmd = oracle_cmd(t_best, args.window, outp, bam)
    for _ in range(max(args.warmup - 2 * len(best), 0)):
        subprocess.run(cmd, check=True, capture_output=True, env=clean_env())
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        subprocess.run(cmd, check=True, capture_output=True, env=clean_env())
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    val = n_reads / sec
    sample = (f"bounded sample of the workload: the same generator at genome_fraction={frac:g} ({n_reads} reads, {n_bytes} BAM bytes), whole file per step, "
              f"process wall incl. writing regions.bed.gz; oracle -t {t_best} (best of -t {sorted(best)}: " + ", ".join(f"{t}: {n_reads / best[t] / 1e6:.2f} M reads/s" for t in sorted(best)) + ")")
    line = {"metric": "reads_per_s", "value": val, "unit": "reads/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args, max(args.gpus, 1), parity),
            "cpu_baseline": {"value": val, "unit": "reads/s", "cores": t_best + 1, "kind": "port", "sample": sample, "best_value": n_reads / min(times), "host_cores": ncores,
                             "note": "CPU oracle restatement of mosdepth (zlib inflate on -t worker threads + one record thread, the reference's own structure); the mosdepth binary (Nim + htslib/libdeflate) cannot be built in this image"},
            "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---- rendering the output files from the library's results (parity check; mirrors mosdepth_b200/host/main.cpp) --------------------------
def _dist_text(chrom, d, precision=2):
    total = int(sum(d))
    if total < 1:
        return ""
    out, cum, n = [], 0.0, len(d)
    for i in range(n):
        irev = n - 1 - i
        v = int(d[irev])
        if irev > 300 and v == 0:
            continue
        cum += v / total
        if cum < 8e-5:
            continue
        out.append(f"{chrom}\t{irev}\t{cum:.{precision}f}\n")
    return "".join(out)


def _summary_row(name, st):
    length, bases, mn, mx = st
    mean = bases / length if length > 0 else 0.0
    return f"{name}\t{length}\t{bases}\t{mean:.2f}\t{0 if mn == 0xFFFFFFFF else mn}\t{mx}\n"


def render_outputs(names, lengths, window, per_contig):
    """per_contig[tid] = None (no records: coverage() returned -2) or dict(values, stat, dist, rstat, rdist).  Returns the four texts of
    `--by <window> -n` exactly as the reference's main loop emits them (mosdepth.nim:691-813)."""
    import numpy as np
    regions, summary, gdist, rdist = [], ["chrom\tlength\tbases\tmean\tmin\tmax\n"], [], []
    tot_g, tot_r = np.zeros(512, dtype=np.int64), np.zeros(512, dtype=np.int64)
    gs, rs = [0, 0, 0xFFFFFFFF, 0], [0, 0, 0xFFFFFFFF, 0]
    add = lambda a, b: [a[0] + b[0], a[1] + b[1], min(a[2], b[2]), max(a[3], b[3])]

    def grow(a, n):
        if n > a.size:
            b = np.zeros(n, dtype=np.int64); b[:a.size] = a; return b
        return a
    for tid, (name, tlen) in enumerate(zip(names, lengths)):
        pc = per_contig.get(tid)
        nw = (tlen + window - 1) // window if tlen else 0
        starts = np.arange(nw, dtype=np.int64) * window
        stops = np.minimum(starts + window, tlen)
        vals = pc["values"] if pc else np.zeros(nw)
        assert len(vals) == nw, (name, len(vals), nw)
        regions.append("".join(f"{name}\t{s}\t{e}\t{v:.2f}\n" for s, e, v in zip(starts.tolist(), stops.tolist(), vals.tolist())))
        cg, cr = np.zeros(512, dtype=np.int64), np.zeros(512, dtype=np.int64)
        if pc:
            cg = grow(cg, len(pc["dist"])); cg[:len(pc["dist"])] += pc["dist"]
            cr[:512] += pc["rdist"]
            summary.append(_summary_row(name, pc["stat"])); summary.append(_summary_row(name + "_region", pc["rstat"]))
            gs, rs = add(gs, pc["stat"]), add(rs, pc["rstat"])
        gdist.append(_dist_text(name, cg)); rdist.append(_dist_text(name, cr))
        if pc:
            tot_g = grow(tot_g, cg.size); tot_g[:cg.size] += cg; tot_r += cr
    summary.append(_summary_row("total", gs)); summary.append(_summary_row("total_region", rs))
    gdist.append(_dist_text("total", tot_g)); rdist.append(_dist_text("total", tot_r))
    return {".regions.bed": "".join(regions), ".mosdepth.summary.txt": "".join(summary), ".mosdepth.global.dist.txt": "".join(gdist),
            ".mosdepth.region.dist.txt": "".join(rdist)}, (gs, rs, tot_g, tot_r)


# ---- the other BASELINE.json configs (C2, C4, C5): one GPU, same contract (value resident / e2e from pinned host memory / parity / CPU sample) ----
OTHER = {
    "c2": {"fraction": 1.0, "parity_fraction": 0.02, "sample_fraction": 0.25, "flags": lambda bam: [],
           "workload": "C2 synthetic 30X chr20 (64,444,167 bp x {f:g}, 150-bp pairs, BGZF level 6), default mode (CIGAR + mate-overlap correction), per-base output"},
    "c4": {"fraction": 1.0, "parity_fraction": 0.01, "sample_fraction": 0.1, "flags": lambda bam: ["-n", "--by", str(bam) + ".bed", "-T", "1,10,20,30", "-m", "-q", "0:1:10:30:"],
           "workload": "C4 synthetic 100X exome (25 GRCh38 contigs x {f:g}) + {n_bed} -interval BED (unsorted, overlapping), -n --by BED -T 1,10,20,30 -m -q 0:1:10:30:"},
    "c5": {"fraction": 0.1, "parity_fraction": 0.0015, "sample_fraction": 0.01, "flags": lambda bam: ["-a", "-n", "--by", "1000"],
           "workload": "C5 synthetic nanopore 30X WGS (25 GRCh38 contigs x {f:g}, reads 10-100 kb, an operation every ~10 bp, + 2 % short proper pairs), -a -n --by 1000 (fragment mode on the paired subset); default (CIGAR) mode reported beside it"},
}
SUFFIXES = [".mosdepth.summary.txt", ".mosdepth.global.dist.txt", ".mosdepth.region.dist.txt", ".regions.bed", ".per-base.bed", ".quantized.bed", ".thresholds.bed"]


def other_config(args, frac):
    spec = OTHER[args.config]
    return {"workload": spec["workload"].format(f=frac, n_bed=int(round(1195764 * frac))), "genome_fraction": frac, "flags": " ".join(spec["flags"]("BAM")).replace("BAM.bed", "exons.bed"), "gpus": 1,
            "l2": "inputs (compressed BAM, depth arrays) are far larger than the 126 MB L2; no flush needed"}


def cli_vs_oracle(bam, flags, threads):
    """Every output file of the mosdepth-b200 process against the oracle's, byte for byte.  Returns (ok, detail, oracle_wall_s, cli_wall_s)."""
    with tempfile.TemporaryDirectory(dir=str(data_dir())) as tmp:
        t0 = time.perf_counter()
        subprocess.run([str(ROOT / "oracle" / "mosdepth_oracle"), "-t", str(threads)] + flags + [str(Path(tmp) / "o"), str(bam)], check=True, capture_output=True, env=clean_env())
        t1 = time.perf_counter()
        subprocess.run([str(ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200"), "--plain"] + flags + [str(Path(tmp) / "g"), str(bam)], check=True, capture_output=True, env=clean_env())
        t2 = time.perf_counter()
        bad, n = [], 0
        for s in SUFFIXES:
            po, pg = Path(tmp) / ("o" + s), Path(tmp) / ("g" + s)
            if po.exists() != pg.exists():
                bad.append(s + " (missing)")
            elif po.exists():
                n += 1
                if po.read_bytes() != pg.read_bytes():
                    bad.append(s)
        return (not bad and n >= 3), (f"{n} files byte-identical" if not bad else f"DIFFER: {bad}"), t1 - t0, t2 - t1


def load_bed(path, names):
    """bed_to_table (mosdepth.nim:362-380): per contig, stable sort by start."""
    import numpy as np
    by = {}
    with open(path) as f:
        for line in f:
            c = line.rstrip("\n").split("\t")
            by.setdefault(c[0], []).append((int(c[1]), int(c[2])))
    out = {}
    for t, name in enumerate(names):
        if name in by:
            a = np.array(by[name], dtype=np.int64)
            a = a[np.argsort(a[:, 0], kind="stable")]
            out[t] = (a[:, 0].astype(np.uint32), a[:, 1].astype(np.uint32))
    return out


def run_other(args):
    """`--config c2|c4|c5` on one GPU."""
    spec = OTHER[args.config]
    frac = args.genome_fraction if args.fraction_given else spec["fraction"]
    ncores = os.cpu_count() or 1
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    ensure_built()
    if args.impl == "reference":
        sfrac = min(args.cpu_fraction or spec["sample_fraction"], frac)
        bam, n_reads, n_bytes = make_bam(args.config, sfrac, args.seed, ncores)
        flags = spec["flags"](bam)
        with tempfile.TemporaryDirectory() as tmp:
            parity = golden_check(tmp)
        outp = data_dir() / f"cpu_out_{os.getpid()}"
        best = {}
        for t in [x for x in thread_choices() if x >= 3] or thread_choices():
            cmd = [str(ROOT / "oracle" / "mosdepth_oracle"), "-t", str(t)] + flags + [str(outp), str(bam)]
            for _ in range(2):
                t0 = time.perf_counter(); subprocess.run(cmd, check=True, capture_output=True, env=clean_env()); best[t] = min(best.get(t, 1e30), time.perf_counter() - t0)
        t_best = min(best, key=best.get)
        cmd = [str(ROOT / "oracle" / "mosdepth_oracle"), "-t", str(t_best)] + flags + [str(outp), str(bam)]
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter(); subprocess.run(cmd, check=True, capture_output=True, env=clean_env()); times.append(time.perf_counter() - t0)
        for f in data_dir().glob(f"cpu_out_{os.getpid()}*"):
            f.unlink()
        sec = sum(times) / len(times); val = n_reads / sec
        print(json.dumps({"metric": "reads_per_s", "value": val, "unit": "reads/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": dict(other_config(args, frac), parity_checked=parity),
                          "cpu_baseline": {"value": val, "unit": "reads/s", "cores": t_best + 1, "kind": "port", "host_cores": ncores,
                                           "sample": f"bounded sample: the same generator at genome_fraction={sfrac:g} ({n_reads} reads, {n_bytes} BAM bytes), oracle process wall (writes every output file), best of -t {sorted(best)}"},
                          "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import mosdepth_b200 as m
    from mosdepth_b200 import hostio
    torch.cuda.set_device(0)
    # parity first (small), optionally at bench size as well
    pbam, p_reads, _ = make_bam(args.config, spec["parity_fraction"], args.seed, ncores)
    ok, detail, _ow, _cw = cli_vs_oracle(pbam, spec["flags"](pbam), min(ncores - 1, 7))
    parity_note = f"{p_reads} reads (genome_fraction {spec['parity_fraction']:g}): mosdepth-b200 vs oracle, {detail}"
    bam, n_reads, n_bytes = make_bam(args.config, frac, args.seed, ncores, keep_tags=(bam_tag(args.config, spec["parity_fraction"], args.seed, 1 if args.config == "c5" else 6),))
    full = None
    if args.full_parity:
        fok, fdetail, ow, cw = cli_vs_oracle(bam, spec["flags"](bam), min(ncores - 1, 15))
        full = {"ok": fok, "detail": fdetail, "oracle_wall_s": ow, "cli_plain_wall_s": cw}
        ok = ok and fok
        parity_note += f"; bench size ({n_reads} reads): {fdetail} (oracle {ow:.1f} s, mosdepth-b200 --plain {cw:.1f} s)"
    hdr = hostio.read_bam_header(bam)
    raw = np.fromfile(bam, dtype=np.uint8)
    host = torch.empty(raw.size, dtype=torch.uint8, pin_memory=True); host.numpy()[:] = raw; del raw
    ctx = m.Context(0)
    ctx.open_mem(host, hdr)
    nt = len(hdr.names)
    quants = [0, 1, 10, 30, 2**63 - 1]
    bed = load_bed(str(bam) + ".bed", hdr.names) if args.config == "c4" else None
    mode = {"c2": m.MODE_DEFAULT, "c4": m.MODE_DEFAULT, "c5": m.MODE_FRAGMENT}[args.config]
    ctx.set_filters(mode=mode)
    if args.config == "c5":
        ctx.prefetch_windows(1000)
    state = {"d2h": 0, "query_ms": 0.0, "units": 0}

    def step(time_queries=False):
        ctx.totals_reset()
        info = ctx.run()
        nbytes = 0
        if time_queries:
            ctx.mark(2)
        for t in range(nt):
            if ctx.contig_records(t)[0] < 0:
                continue
            st, dist = ctx.contig_stats(t); nbytes += dist.nbytes + 24
            if args.config == "c2":          # gen_depths (:58-96) for the per-base file
                a, b, c = ctx.per_base_runs(t, copy=False); nbytes += 12 * len(a); state["units"] = len(a)      # the runs stay in the library's pinned buffers, as for a C host
            elif args.config == "c4":        # BED loop :720-743 with median + thresholds, then gen_quantized :124-141
                if t in bed:
                    v, rst, rd, cnt = ctx.regions(t, bed[t][0], bed[t][1], use_median=True, thresholds=(1, 10, 20, 30)); nbytes += v.nbytes + rd.nbytes + cnt.nbytes
                a, b, c = ctx.quantized_runs(t, quants, copy=False); nbytes += 12 * len(a)
            else:                            # window loop :720-741
                v, rst, rd = ctx.windows(t, 1000); nbytes += v.nbytes + rd.nbytes
        if time_queries:
            ctx.mark(3); state["query_ms"] = ctx.elapsed_ms(2, 3)
        state["d2h"] = nbytes
        return info

    def timed(resident):
        if resident:
            ctx.stage()
        else:
            ctx.set_spans([], [])
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        sampler = ClockSampler(0); sampler.start()
        lc0 = ctx.launch_count()
        ctx.mark(0)
        infos = [step() for _ in range(args.steps)]
        ctx.mark(1)
        torch.cuda.synchronize()
        ms = ctx.elapsed_ms(0, 1) / args.steps
        sampler.stop_flag.set(); sampler.join()
        return ms, infos, ctx.launch_count() - lc0, sampler.summary()

    ms_res, infos_res, launches, clocks = timed(True)
    step(time_queries=True)
    query_ms = state["query_ms"]
    ms_e2e, infos_e2e, _l, _c = timed(False)
    info = infos_res[-1]
    extra = {}
    if args.config == "c5":      # the CIGAR-heavy path of the same BAM: default mode (every long read's blocks, no pairs to join)
        ctx.set_filters(mode=m.MODE_DEFAULT); ctx.stage()
        for _ in range(2):
            step()
        ctx.mark(4); i2 = [step() for _ in range(3)]; ctx.mark(5)
        ms_def = ctx.elapsed_ms(4, 5) / 3
        extra = {"default_mode_ms_per_step": ms_def, "default_mode_reads_per_s": n_reads / (ms_def / 1e3), "default_mode_records_ms": i2[-1].ms_records, "default_mode_events": i2[-1].n_events,
                 "default_mode_events_per_s": i2[-1].n_events / (i2[-1].ms_records / 1e3) if i2[-1].ms_records > 0 else None}
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    avg = lambda f: sum(f(i) for i in infos_res) / len(infos_res)
    inf_ms = avg(lambda i: i.ms_inflate_span)
    achieved = (info.bytes_compressed / 1e9) / (inf_ms / 1e3) if inf_ms > 0 else 0.0
    traffic_per_block = (5.321760e9 + 5.912786e9 + 16.776851e9 + 5.449551e9) / 85062.0
    cells = info.n_depth_cells
    # algorithmic bytes of the query stage: C2 runs = 8 B per base (two passes) + 12 B per run; C4 regions = 4 B per covered base
    # (+ the bisection passes of the median) + quantized runs 8 B per base; C5 windows = 4 B per base + 8 B per window
    q_bytes = {"c2": 8.0 * cells + 12.0 * state["units"], "c4": 4.0 * (sum(int((e.astype(np.int64) - s.astype(np.int64)).sum()) for s, e in bed.values()) if bed else 0) + 8.0 * cells, "c5": 4.0 * cells}[args.config]
    roofline = {"kernel": "huffman_decode_kernel + lz77_resolve_kernel (K1, BGZF inflate): the largest share of the step in this config too", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic_per_block * info.n_blocks / max(info.n_inflate_launches, 1), "traffic_source": "ncu --set full, profiles/r2a_fast_full_summary.txt, DRAM bytes per block x blocks per launch",
                "algorithmic_bytes_per_launch": info.bytes_compressed / max(info.n_inflate_launches, 1), "launches_per_step": info.n_inflate_launches, "ms_per_launch": inf_ms / max(info.n_inflate_launches, 1),
                "inflate_span_ms": inf_ms, "frame_ms": avg(lambda i: i.ms_frame), "records_ms": avg(lambda i: i.ms_records), "scan_ms": avg(lambda i: i.ms_scan), "run_total_ms": avg(lambda i: i.ms_total),
                "records_events_per_s": info.n_events / (avg(lambda i: i.ms_records) / 1e3) if avg(lambda i: i.ms_records) > 0 else None, "events": info.n_events,
                "query_ms": query_ms, "query_algorithmic_bytes": q_bytes, "query_gbs": q_bytes / 1e9 / (query_ms / 1e3) if query_ms > 0 else None,
                "query_frac": (q_bytes / 1e9 / (query_ms / 1e3) / peak) if query_ms > 0 else None,
                "query_kernels": {"c2": "K8 runs_count/scan/write/stops (gen_depths)", "c4": "K7 regions_kernel (median + thresholds) + K9 quantized runs", "c5": "K7 windows_all_kernel (served from msd_run)"}[args.config]}
    roofline.update(extra)
    ctx.close()
    cpu_baseline = None
    if not args.no_cpu_baseline:
        sfrac = min(args.cpu_fraction or spec["sample_fraction"], frac)
        sbam, s_reads, s_bytes = make_bam(args.config, sfrac, args.seed, ncores, keep_tags=(bam_tag(args.config, frac, args.seed, 1 if args.config == "c5" else 6),))
        flags = spec["flags"](sbam); outp = data_dir() / f"cpu_out_{os.getpid()}"; best = {}
        for t in [x for x in thread_choices() if x >= 7] or thread_choices():
            cmd = [str(ROOT / "oracle" / "mosdepth_oracle"), "-t", str(t)] + flags + [str(outp), str(sbam)]
            for _ in range(2):
                t0 = time.perf_counter(); subprocess.run(cmd, check=True, capture_output=True, env=clean_env()); best[t] = min(best.get(t, 1e30), time.perf_counter() - t0)
        t_best = min(best, key=best.get)
        t0 = time.perf_counter()
        subprocess.run([str(ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200")] + flags + [str(outp), str(sbam)], check=True, capture_output=True, env=clean_env())
        cw = time.perf_counter() - t0
        for f in data_dir().glob(f"cpu_out_{os.getpid()}*"):
            f.unlink()
        cpu_baseline = {"value": s_reads / best[t_best], "unit": "reads/s", "cores": t_best + 1, "kind": "port", "seconds": best[t_best], "host_cores": ncores,
                        "gpu_cli_wall_s_same_sample": cw, "gpu_cli_reads_per_s_same_sample": s_reads / cw,
                        "sample": f"bounded sample: the same generator at genome_fraction={sfrac:g} ({s_reads} reads, {s_bytes} BAM bytes), oracle process wall with every output file written (.gz), best of -t {sorted(best)} x 2"}
    line = {"metric": "reads_per_s", "value": n_reads / (ms_res / 1e3), "unit": "reads/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": dict(other_config(args, frac), parity_checked=bool(ok)),
            "e2e": {"value": n_reads / (ms_e2e / 1e3), "unit": "reads/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(infos_e2e[-1].bytes_compressed), "d2h_bytes_per_step": int(state["d2h"])},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "detail": {"reads": n_reads, "bam_bytes": n_bytes, "parity": parity_note, "full_parity": full, "depth_cells": cells}}
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c2", "c4", "c5"], help="BASELINE.json configs: c3 = the headline (30X WGS --by 500 -n -x, 1..8 GPUs); c2 / c4 / c5 run on one GPU")
    ap.add_argument("--full-parity", action="store_true", help="c2 / c4 / c5: also diff every output file of mosdepth-b200 against the oracle at BENCH size (minutes of CPU time)")
    ap.add_argument("--genome-fraction", type=float, default=float(os.environ.get("MSD_BENCH_FRACTION", "0.5")),
                    help="scale of the 30X genome (every contig length x f).  0.5 = 1.54 Gbp, 309 M reads, 28 GB of BAM: the largest the 1-GPU box generates (16 cores, ~5 min of zlib level 6) and holds twice (page cache + pinned image) inside the bench budget; 1.0 needs ~11 min of generation and 112 GB of host memory")
    ap.add_argument("--cpu-fraction", type=float, default=0.0, help="genome fraction of the CPU arm's bounded sample (default 0.03)")
    ap.add_argument("--seed", type=int, default=20261022)
    ap.add_argument("--window", type=int, default=500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the wall-clock run of the mosdepth-b200 process on the bench BAM")
    ap.add_argument("--sharding", default="offsets", choices=["contigs", "offsets"],
                    help="N > 1: N equal BGZF offset ranges of the file, cut inside contigs (the depth that reads leave beyond a cut goes to the next rank over NCCL), or whole contigs per rank (LPT, no data-path exchange)")
    args = ap.parse_args()
    args.fraction_given = any(a.startswith("--genome-fraction") for a in sys.argv[1:]) or "MSD_BENCH_FRACTION" in os.environ
    if args.config != "c3":
        return run_other(args)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 1)

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
    from mosdepth_b200.shard import contig_weights, load_compact, lpt_shards, merge_adjacent_spans, offset_shards

    ncores = os.cpu_count() or 1
    t_begin = time.perf_counter()

    def barrier():
        if world > 1:
            dist.barrier()

    def get_bam(fraction, keep=()):
        if rank == 0:
            make_bam(args.config, fraction, args.seed, ncores, keep_tags=keep)
        barrier()
        return make_bam(args.config, fraction, args.seed, ncores, keep_tags=keep)

    def fresh_comm_id():
        """The library's own communicator: a new NCCL id from rank 0 for every communicator (an id is good for one ncclCommInitRank round);
        torch.distributed is only the messenger."""
        box = [m.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    class Job:
        """One BAM through this rank's context: the shard plan, the pinned host image of this rank's bytes, and the step."""

        def __init__(self, bam, pin=True):
            self.bam = bam
            self.hdr = hostio.read_bam_header(bam)
            index = hostio.read_index(bam)
            self.nt = len(self.hdr.names)
            L = self.hdr.lengths
            if world > 1 and args.sharding == "offsets":
                plan = offset_shards(index, L, world, window=args.window, with_cover=True)
            else:
                plan = [[(t, 0, L[t], index[t].ref_beg, index[t].ref_end, 0) for t in r] for r in lpt_shards(contig_weights(index), world)]
            self.plan = plan
            self.pieces = plan[rank]
            self.cut = [(t, lo, hi, cover) for t, lo, hi, _vb, _ve, cover in self.pieces if lo > 0 or hi < L[t]]
            self.any_cut = any(lo > 0 or hi < L[t] for r in plan for t, lo, hi, *_ in r)
            begs, ends = merge_adjacent_spans([p[3] for p in self.pieces], [p[4] for p in self.pieces]) if self.pieces else ([1 << 16], [1 << 16])
            keep = {}

            def alloc(n):
                keep["t"] = torch.empty(n, dtype=torch.uint8, pin_memory=pin)
                return keep["t"].numpy()
            if self.pieces:
                _buf, self.begs, self.ends, self.used = load_compact(bam, begs, ends, alloc)
            else:   # nothing for this rank: an empty image with one span that reads nothing
                alloc(1 << 16); self.begs, self.ends, self.used = [1 << 16], [1 << 16], 0
            self.host = keep["t"]
            want = np.zeros(self.nt, dtype=np.uint8)
            for p in self.pieces:
                want[p[0]] = 1
            self.ctx = ctx = m.Context(local)
            ctx.open_mem(self.host, self.hdr)
            ctx.set_targets(want)
            ctx.set_spans(self.begs, self.ends)
            ctx.set_filters(mode=m.MODE_FAST)       # -x; defaults -F 1796 -Q 0
            ctx.prefetch_windows(args.window)       # --by 500: the host will ask for every contig's windows
            for t, lo, hi, cover in self.cut:
                ctx.set_contig_range(t, lo, hi); ctx.set_contig_cover(t, cover)
            if world > 1:
                ctx.comm_init(rank, world, fresh_comm_id())
            self.d2h = 0
            self.results = None

        def step(self, keep_results=False):
            """One pass of the hot path for this rank's pieces + the two exchanges."""
            ctx = self.ctx
            ctx.totals_reset()
            info = ctx.run()
            if self.any_cut:
                ctx.exchange_spills()
            nbytes, res = 0, []
            for t, lo, hi, *_ in self.pieces:
                is_cut = lo > 0 or hi < self.hdr.lengths[t]
                rc, n = ctx.contig_records(t)
                if rc < 0 and not is_cut:
                    continue
                st, distv = ctx.contig_stats(t)
                vals, rst, rdist = ctx.windows(t, args.window)
                nbytes += vals.nbytes + rdist.nbytes + distv.nbytes + 48
                if keep_results:
                    res.append((t, lo, max(n, 0), (st.cum_length, st.cum_depth, st.min_depth, st.max_depth), distv.copy(), vals.copy(),
                                (rst.cum_length, rst.cum_depth, rst.min_depth, rst.max_depth), rdist.copy()))
            if world > 1:
                ctx.totals_allreduce()
            g, r, gd, rd = ctx.totals()
            nbytes += gd.nbytes + rd.nbytes + 64
            self.d2h = nbytes
            if keep_results:
                self.results = (res, ((g.cum_length, g.cum_depth, g.min_depth, g.max_depth), (r.cum_length, r.cum_depth, r.min_depth, r.max_depth), gd.copy(), rd.copy()))
            return info

        def close(self):
            self.ctx.close()

    # ---- parity first: a small BAM of the same shape through the same path, every output file against the oracle ------------------
    keep_main = (bam_tag(args.config, args.genome_fraction, args.seed), bam_tag(args.config, SAMPLE_FRACTION, args.seed))
    pbam, p_reads, _ = get_bam(PARITY_FRACTION, keep_main)
    pj = Job(pbam, pin=True)
    pj.ctx.stage()
    pj.step(keep_results=True)
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, pj.results)
    else:
        gathered = [pj.results]
    parity, parity_note = False, ""
    if rank == 0:
        per = {}
        for res, _tot in gathered:          # ranks are in genome order: the shards of a contig arrive in position order
            for t, lo, nrec, st, dv, vals, rst, rdv in res:
                e = per.setdefault(t, {"n": 0, "values": [], "stat": [0, 0, 0xFFFFFFFF, 0], "rstat": [0, 0, 0xFFFFFFFF, 0], "dist": np.zeros(1, dtype=np.int64), "rdist": np.zeros(512, dtype=np.int64), "at": 0})
                assert lo == e["at"] * args.window, "shards of a contig do not meet at a window boundary"
                e["n"] += nrec; e["values"].append(vals); e["at"] += len(vals)
                e["stat"] = [e["stat"][0] + st[0], e["stat"][1] + st[1], min(e["stat"][2], st[2]), max(e["stat"][3], st[3])]
                e["rstat"] = [e["rstat"][0] + rst[0], e["rstat"][1] + rst[1], min(e["rstat"][2], rst[2]), max(e["rstat"][3], rst[3])]
                d = np.zeros(max(e["dist"].size, dv.size), dtype=np.int64); d[:e["dist"].size] += e["dist"]; d[:dv.size] += dv; e["dist"] = d
                e["rdist"] += rdv
        per = {t: dict(e, values=np.concatenate(e["values"])) if e["n"] > 0 else None for t, e in per.items()}
        texts, (gs, rs, tot_g, tot_r) = render_outputs(pj.hdr.names, pj.hdr.lengths, args.window, per)
        with tempfile.TemporaryDirectory() as tmp:
            subprocess.run(oracle_cmd(min(ncores - 1, 7), args.window, Path(tmp) / "o", pbam), check=True, capture_output=True, env=clean_env())
            bad = [s for s, txt in texts.items() if (Path(tmp) / ("o" + s)).read_text() != txt]
            # the same BAM through the mosdepth-b200 process (N GPUs: host threads + the library's NCCL exchange): files against the oracle's
            cli_bad = []
            if world == 1:
                subprocess.run([str(ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200"), "--plain", "--device", str(local), "-n", "-x", "--by", str(args.window), str(Path(tmp) / "g"), str(pbam)],
                               check=True, capture_output=True, env=clean_env())
                cli_bad = [s for s in texts if (Path(tmp) / ("o" + s)).read_bytes() != (Path(tmp) / ("g" + s)).read_bytes()]
        # the all-reduced `total` rows every rank holds must be the sums of the per-contig results
        _, (tg, tr, tgd, trd) = gathered[0]
        tot_ok = list(tg) == gs and list(tr) == rs and np.array_equal(np.trim_zeros(tgd, "b"), np.trim_zeros(tot_g, "b")) and np.array_equal(np.trim_zeros(trd, "b"), np.trim_zeros(tot_r, "b"))
        parity = not bad and not cli_bad and tot_ok
        parity_note = (f"{p_reads} reads (genome_fraction {PARITY_FRACTION:g}) through the same {world}-rank path: regions.bed, summary, global.dist, region.dist byte-identical to the oracle"
                       if parity else f"PARITY FAILED: files {bad} cli {cli_bad} totals_ok {tot_ok}")
        if not parity:
            print("[bench] " + parity_note, file=sys.stderr)
    pj.close(); del pj
    barrier()

    # ---- the workload ---------------------------------------------------------------------------------------------------------------
    bam, n_reads, n_bytes = get_bam(args.genome_fraction, keep_main)
    t_load = time.perf_counter()
    job = Job(bam)
    t_loaded = time.perf_counter()
    ctx = job.ctx

    def timed(resident):
        if resident:
            ctx.stage()
        else:
            ctx.set_spans(job.begs, job.ends)   # drops the staged copy
        for _ in range(args.warmup):
            job.step()
        barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local); sampler.start()
        infos = []
        lc0 = ctx.launch_count()
        t0 = time.perf_counter()
        ctx.mark(0)
        for _ in range(args.steps):
            infos.append(job.step())
        ctx.mark(1)
        torch.cuda.synchronize()
        ms_dev = ctx.elapsed_ms(0, 1)
        wall = time.perf_counter() - t0
        barrier()
        sampler.stop_flag.set(); sampler.join()
        # the step makes blocking host calls (results come back per contig), so the device-side span between
        # the two marks is the honest step time; the host wall clock is reported beside it
        tt = torch.tensor([ms_dev, wall * 1e3], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt[0].item() / args.steps, tt[1].item() / args.steps, infos, ctx.launch_count() - lc0, sampler.summary()

    ms_res, wall_res, infos_res, launches, clocks = timed(True)
    ms_e2e, wall_e2e, infos_e2e, _, _clocks2 = timed(False)
    info = infos_res[-1]
    value = n_reads / (ms_res / 1e3)
    e2e_val = n_reads / (ms_e2e / 1e3)
    h2d = torch.tensor([float(infos_e2e[-1].bytes_compressed), float(job.d2h)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(h2d)

    # roofline of the dominant kernel pair K1 = huffman_decode_kernel + lz77_resolve_kernel (BGZF inflate, > 80 % of the
    # step): algorithmic bytes per launch = compressed bytes of the chunk, read once
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    avg = lambda f: sum(f(i) for i in infos_res) / len(infos_res)
    inf_ms = avg(lambda i: i.ms_inflate_span)      # launches of consecutive chunks overlap: use the span
    inf_launches = info.n_inflate_launches
    achieved = (info.bytes_compressed / 1e9) / (inf_ms / 1e3) if inf_ms > 0 else 0.0
    # DRAM bytes per BGZF block from the ncu --set full capture of the shipped kernels (profiles/r2a_fast_full_summary.txt, 85,062 blocks in one launch each:
    # decode 5.32 GB read + 5.91 GB written, resolve 16.78 GB read + 5.45 GB written), scaled to this run's blocks per launch
    traffic_per_block = (5.321760e9 + 5.912786e9 + 16.776851e9 + 5.449551e9) / 85062.0
    scan_ms = avg(lambda i: i.ms_scan)
    scan_bytes = 8.0 * info.n_depth_cells
    roofline = {"kernel": "huffman_decode_kernel + lz77_resolve_kernel (K1, BGZF inflate)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic_per_block * info.n_blocks / max(inf_launches, 1),
                "traffic_source": "ncu --set full, profiles/r2a_fast_full_summary.txt, DRAM bytes per block x blocks per launch",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": info.bytes_compressed / max(inf_launches, 1), "launches_per_step": inf_launches, "ms_per_launch": inf_ms / max(inf_launches, 1),
                "inflated_gbs": (info.bytes_inflated / 1e9) / (inf_ms / 1e3) if inf_ms > 0 else 0.0,
                "inflate_span_ms": inf_ms, "inflate_sum_ms": avg(lambda i: i.ms_inflate), "frame_ms": avg(lambda i: i.ms_frame), "records_ms": avg(lambda i: i.ms_records),
                "scan_ms": scan_ms, "zero_ms": avg(lambda i: i.ms_zero), "run_total_ms": avg(lambda i: i.ms_total),
                # K6 (one segmented launch: prefix sum + stats + dist of every contig), live: 4 B read + 4 B written per depth cell
                "k6_scan_gbs": scan_bytes / 1e9 / (scan_ms / 1e3) if scan_ms > 0 else 0.0, "k6_scan_frac": (scan_bytes / 1e9 / (scan_ms / 1e3) / peak) if scan_ms > 0 else 0.0,
                "k6_scan_bytes": scan_bytes,
                "note": "inflate is bound by the serial Huffman chains (latency / issue), not by HBM; compressed GB/s and inflated GB/s are quoted against the HBM peak as north_star asks"}

    # K1c alone, timed live: the one kernel of K1 that is HBM bound (it reads the inflated bytes once: algorithmic bytes = U)
    try:
        if rank != 0:
            raise RuntimeError("rank 0 only")
        blk = 65280
        sample = np.random.default_rng(1).integers(0, 256, size=blk * 16384, dtype=np.uint8)      # 1.07 GB: far larger than L2
        _crcs, crc_ms = ctx.debug_crc32(sample, blk, 0, reps=10)
        roofline["k1c_crc_gbs"] = sample.size / 1e9 / (crc_ms / 1e3)
        roofline["k1c_crc_frac"] = roofline["k1c_crc_gbs"] / peak
        roofline["k1c_crc_ms_per_launch"] = crc_ms
        del sample
    except Exception as e:      # never let the side measurement take the bench line down
        roofline["k1c_crc_error"] = str(e)

    job.close()
    del job
    barrier()

    cpu_baseline, cli = None, {}
    if rank == 0:
        env = dict(clean_env(), MOSDEPTH_B200_STATS="1")
        exe = str(ROOT / "mosdepth_b200" / "bin" / "mosdepth-b200")
        outp = data_dir() / f"cli_out_{os.getpid()}"

        def cli_wall(path, reps=2):
            best, stats = 1e30, None
            for _ in range(reps):
                t0 = time.perf_counter()
                r = subprocess.run([exe, "--gpus", str(world), "-n", "-x", "--by", str(args.window), str(outp), str(path)], capture_output=True, text=True, env=env)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    return None, r.stderr[-400:]
                if dt < best:
                    best = dt
                    for l in r.stderr.splitlines():
                        if l.startswith("[mosdepth-b200] {"):
                            stats = json.loads(l[len("[mosdepth-b200] "):])
            return best, stats
        if not args.no_cli:
            try:
                w, st = cli_wall(bam)
                cli = {"cli_wall_s": w, "cli_reads_per_s": (n_reads / w) if w else None, "cli_stages": st}
            except Exception as e:
                cli = {"cli_error": str(e)}
        if world == 1 and not args.no_cpu_baseline:
            frac = min(args.cpu_fraction or SAMPLE_FRACTION, args.genome_fraction)
            sbam, s_reads, s_bytes = make_bam(args.config, frac, args.seed, ncores, keep_tags=keep_main)
            t_list = [t for t in thread_choices() if t >= 3] or thread_choices()
            best, _o = time_oracle(sbam, args.window, t_list, 2)
            t_best = min(best, key=best.get)
            cpu_baseline = {"value": s_reads / best[t_best], "unit": "reads/s", "cores": t_best + 1, "kind": "port", "seconds": best[t_best], "host_cores": ncores,
                            "sample": f"bounded sample: the same generator at genome_fraction={frac:g} ({s_reads} reads, {s_bytes} BAM bytes), whole file, oracle process wall, best of -t {sorted(best)} x 2 runs: "
                                      + ", ".join(f"-t {t}: {s_reads / best[t] / 1e6:.2f} M reads/s" for t in sorted(best))}
            if not args.no_cli:
                try:
                    w, _st = cli_wall(sbam)
                    cpu_baseline["gpu_cli_wall_s_same_sample"] = w      # the mosdepth-b200 process on the very file the oracle was timed on
                    cpu_baseline["gpu_cli_reads_per_s_same_sample"] = (s_reads / w) if w else None
                except Exception:
                    pass
        for f in data_dir().glob(f"cli_out_{os.getpid()}*"):
            f.unlink()
        for f in data_dir().glob(f"cpu_out_{os.getpid()}*"):
            f.unlink()

    if rank == 0:
        cfg = workload_config(args, world, parity)
        line = {"metric": "reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_res, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": cfg,
                "e2e": dict({"value": e2e_val, "unit": "reads/s", "ms_per_step": ms_e2e, "wall_ms_per_step": wall_e2e,
                             "h2d_bytes_per_step": int(h2d[0].item()), "d2h_bytes_per_step": int(h2d[1].item()),
                             "h2d_ms": infos_e2e[-1].ms_h2d, "inflate_ms": infos_e2e[-1].ms_inflate, "run_total_ms": infos_e2e[-1].ms_total},
                            **{k: v for k, v in cli.items() if not isinstance(v, dict)}),
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "detail": {"reads": n_reads, "bam_bytes": n_bytes, "parity": parity_note, "wall_s_per_step": wall_res / 1e3, "genome_wall_s_at_f1": (ms_res / 1e3) / args.genome_fraction,
                           "sharding": ("one GPU" if world == 1 else f"{world} equal BGZF offset ranges cut inside contigs; depth spill to the next rank and `total` rows over NCCL inside the library (msd_exchange_spills, msd_totals_allreduce)" if args.sharding == "offsets" else f"whole contigs, LPT over {world} GPUs"),
                           "load_pin_s": t_loaded - t_load, "setup_s": t_load - t_begin, "cli_stages": cli.get("cli_stages")}}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
