#!/bin/bash
# Round 2, final one-GPU validation of the shipped build: the whole gpu suite, smoke(), the default bench (f = 0.5) and its reference arm, the ncu launch
# list of the bench command, and the bench lines of C2 / C4 / C5.
# usage (from the repo root on the box): bash tools/gpu_r2o.sh r2o
tag=${1:-r2o}
out=gpurun_out
mkdir -p $out
( timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log )
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log )
( timeout 1500 python bench.py > $out/${tag}_bench_n1_f0.5.json 2> $out/${tag}_bench_n1_f0.5.err; echo "bench default rc=$?"; tail -3 $out/${tag}_bench_n1_f0.5.err )
( timeout 900 python bench.py --impl reference > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; echo "bench reference arm rc=$?"; tail -2 $out/${tag}_bench_reference_arm.err )
rm -f /dev/shm/msd_bench/c3_*0.5* 2>/dev/null
( timeout 600 python bench.py --steps 2 --warmup 1 --genome-fraction 0.06 --no-cpu-baseline --no-cli > $out/${tag}_bench_plain_f0.06.json 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_bench_c3_f0.06.csv python bench.py --steps 2 --warmup 1 --genome-fraction 0.06 --no-cpu-baseline --no-cli > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launch list rc=$?" )
( timeout 900 python bench.py --config c2 --steps 5 --warmup 3 --full-parity > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; echo "bench c2 rc=$?"; tail -2 $out/${tag}_bench_c2.err )
( timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --genome-fraction 0.5 > $out/${tag}_bench_c4_f0.5.json 2> $out/${tag}_bench_c4.err; echo "bench c4 rc=$?"; tail -2 $out/${tag}_bench_c4.err )
( timeout 900 python bench.py --config c5 --steps 5 --warmup 3 --genome-fraction 0.05 > $out/${tag}_bench_c5_f0.05.json 2> $out/${tag}_bench_c5.err; echo "bench c5 rc=$?"; tail -2 $out/${tag}_bench_c5.err )
for f in $out/${tag}_bench_n1_f0.5.json $out/${tag}_bench_reference_arm.json $out/${tag}_bench_c2.json $out/${tag}_bench_c4_f0.5.json $out/${tag}_bench_c5_f0.05.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = j.get("roofline") or {}
    print(sys.argv[1], "value %.1f M  e2e %.1f M  ms %.2f  inflate_span %s scan_ms %s k6 %s k1c %s query_ms %s parity %s cli %s cpu %s" % (j["value"] / 1e6, j["e2e"]["value"] / 1e6, j["ms_per_step"], r.get("inflate_span_ms"), r.get("scan_ms"), r.get("k6_scan_frac"), r.get("k1c_crc_frac"), r.get("query_ms"), j["config"].get("parity_checked"), j["e2e"].get("cli_wall_s"), (j.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
ls -la $out | grep ${tag}
