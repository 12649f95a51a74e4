#!/bin/bash
# Round 2: K6 with a 128-tile look-back, next-ticket hand-over and L2 prefetch; K1c with a rolling 16-row register pipeline; where the time of
# the per-contig queries and of the CLI's first run goes (MSD_TRACE host laps).
tag=${1:-r2j}
out=gpurun_out
mkdir -p $out
( timeout 400 python -m pytest tests/test_gpu_stages.py -x -q > $out/${tag}_pytest_stages.log 2>&1; echo "pytest stages rc=$?"; tail -2 $out/${tag}_pytest_stages.log )
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log )
( timeout 600 python bench.py --genome-fraction 0.12 --no-cpu-baseline --no-cli > $out/${tag}_bench_n1_f0.12.json 2> $out/${tag}_bench_n1_f0.12.err; echo "bench f=0.12 rc=$?"; python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_n1_f0.12.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f M reads/s, e2e %.1f M; K6 %.0f GB/s (%.3f of peak), scan %.3f ms; K1c %.0f GB/s (%.3f), parity %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, r.get("k6_scan_gbs",0), r.get("k6_scan_frac",0), r.get("scan_ms",0), r.get("k1c_crc_gbs",0), r.get("k1c_crc_frac",0), d["config"].get("parity_checked")))
PY
)
( timeout 400 python tools/time_queries.py 1.0 > $out/${tag}_time_queries_c2.log 2>&1; echo "time_queries rc=$?"; grep -v "^\[gen" $out/${tag}_time_queries_c2.log | tail -40 )
TIMEFORMAT='wall %R s user %U s sys %S s'
B=/dev/shm/msd_r2j; mkdir -p $B
tools/gen_bam --config c2 --genome-fraction 1.0 --threads $(nproc) -o $B/c2.bam > /dev/null 2>&1
for i in 1 2; do { time MSD_TRACE=1 MOSDEPTH_B200_STATS=1 mosdepth_b200/bin/mosdepth-b200 $B/o $B/c2.bam > $out/${tag}_cli_c2_trace_$i.log 2>&1 ; } 2> $out/${tag}_cli_c2_$i.time; echo "c2 cli run $i: $(cat $out/${tag}_cli_c2_$i.time)"; grep -E "host:|wall_s|runs:" $out/${tag}_cli_c2_trace_$i.log | head -40; done
rm -rf $B
ls -la $out | grep ${tag}
