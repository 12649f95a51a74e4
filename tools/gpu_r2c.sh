#!/bin/bash
# Round 2, third GPU session (one GPU): the whole gpu suite on the new ingest / scan / regions code, the headline at the round-1 size (A/B against
# r2a_bench_base.json) and at its new default size, the CLI's wall clock, and the first bench lines of C2 / C4 / C5.
# usage (from the repo root on the box): bash tools/gpu_r2c.sh r2c
tag=${1:-r2c}
out=gpurun_out
mkdir -p $out
TIMEFORMAT='wall %R s user %U s sys %S s'
( timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log )
( timeout 600 python bench.py --steps 5 --warmup 3 --genome-fraction 0.06 > $out/${tag}_bench_f006.json 2> $out/${tag}_bench_f006.err; echo "bench f0.06 rc=$?"; tail -3 $out/${tag}_bench_f006.err )
B=/dev/shm/msd_r2c; mkdir -p $B
{ time tools/gen_bam --config c3 --genome-fraction 0.12 --seed 20261022 --threads $(nproc) --level 6 -o $B/c3.bam > /dev/null 2>&1 ; } 2> $out/${tag}_gen_f012.log; echo "gen f0.12: $(cat $out/${tag}_gen_f012.log)"
for i in 1 2 3; do { time MSD_TRACE=1 MOSDEPTH_B200_STATS=1 mosdepth_b200/bin/mosdepth-b200 -n -x --by 500 $B/o_gz $B/c3.bam > $out/${tag}_cli_gz_$i.log 2>&1 ; } 2> $out/${tag}_cli_gz_$i.time; echo "cli gz $i: $(cat $out/${tag}_cli_gz_$i.time) $(grep wall_s $out/${tag}_cli_gz_$i.log)"; done
{ time oracle/mosdepth_oracle -t 15 -n -x --by 500 $B/o_or $B/c3.bam > /dev/null 2>&1 ; } 2> $out/${tag}_oracle_t15.time; echo "oracle -t 15: $(cat $out/${tag}_oracle_t15.time)"
( zcat $B/o_gz.regions.bed.gz | cmp - $B/o_or.regions.bed && cmp $B/o_gz.mosdepth.summary.txt $B/o_or.mosdepth.summary.txt && echo "cli f=0.12 parity ok" ) 2>&1 | tail -1
rm -rf $B
( timeout 900 python bench.py --config c2 --steps 5 --warmup 3 --full-parity > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; echo "bench c2 rc=$?"; tail -3 $out/${tag}_bench_c2.err )
( timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --genome-fraction 0.5 > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err; echo "bench c4 rc=$?"; tail -3 $out/${tag}_bench_c4.err )
( timeout 900 python bench.py --config c5 --steps 5 --warmup 3 --genome-fraction 0.05 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; echo "bench c5 rc=$?"; tail -3 $out/${tag}_bench_c5.err )
rm -f /dev/shm/msd_bench/c2_* /dev/shm/msd_bench/c4_* /dev/shm/msd_bench/c5_*
( timeout 1500 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_f05.json 2> $out/${tag}_bench_f05.err; echo "bench f0.5 rc=$?"; tail -3 $out/${tag}_bench_f05.err )
for f in $out/${tag}_bench_f006.json $out/${tag}_bench_f05.json $out/${tag}_bench_c2.json $out/${tag}_bench_c4.json $out/${tag}_bench_c5.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = j.get("roofline") or {}
    print(sys.argv[1], "value %.1f M  e2e %.1f M  ms %.2f  inflate_span %s scan_ms %s k6 %s query_ms %s q_frac %s parity %s cli %s cpu %s" % (j["value"] / 1e6, j["e2e"]["value"] / 1e6, j["ms_per_step"], r.get("inflate_span_ms"), r.get("scan_ms"), r.get("k6_scan_frac"), r.get("query_ms"), r.get("query_frac"), j["config"].get("parity_checked"), j["e2e"].get("cli_wall_s"), (j.get("cpu_baseline") or {}).get("value")))
    print("   ", (j.get("detail") or {}).get("parity"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
ls -la $out | grep ${tag}
