#!/bin/bash
# Round 2, last GPU minutes: K1c with 8 rows in 64 registers (default) against 16 rows in 128 (libmosdepth_b200_crc16.so) inside the pipeline, interleaved;
# the stage tests on the shipped build (K1c depth 8, warp-cooperative CIGAR walk); C5 with the cooperative walk.
tag=${1:-r2p}
out=gpurun_out
mkdir -p $out
( timeout 100 python -m pytest tests/test_gpu_stages.py -x -q > $out/${tag}_pytest_stages.log 2>&1; echo "pytest stages rc=$?"; tail -2 $out/${tag}_pytest_stages.log )
for rep in 1 2; do
  ( timeout 60 python tools/sweep_inflate.py 0.06 one > $out/${tag}_ab_crc8_$rep.log 2>&1; echo "A/B crc depth 8 (default) $rep rc=$?"; grep -E "^rep . resident" $out/${tag}_ab_crc8_$rep.log )
  ( MSD_LIB_PATH=$PWD/mosdepth_b200/libmosdepth_b200_crc16.so timeout 60 python tools/sweep_inflate.py 0.06 one > $out/${tag}_ab_crc16_$rep.log 2>&1; echo "A/B crc depth 16 $rep rc=$?"; grep -E "^rep . resident" $out/${tag}_ab_crc16_$rep.log )
done
( timeout 70 python bench.py --genome-fraction 0.06 --steps 3 --warmup 3 --no-cpu-baseline --no-cli > $out/${tag}_bench_n1_f0.06.json 2> $out/${tag}_bench.err; echo "bench f=0.06 rc=$?"; python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_n1_f0.06.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("value %.1f M, e2e %.1f M, K6 %.3f, K1c %.0f GB/s (%.3f), parity %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, r.get("k6_scan_frac",0), r.get("k1c_crc_gbs",0), r.get("k1c_crc_frac",0), d["config"].get("parity_checked")))
PY
)
( timeout 80 python bench.py --config c5 --genome-fraction 0.02 --steps 3 --warmup 3 > $out/${tag}_bench_c5_f0.02.json 2> $out/${tag}_bench_c5.err; echo "bench c5 rc=$?"; python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_c5_f0.02.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("c5: value %.2f M ms %.1f records_ms %.2f default-mode ms %.1f records %.2f events/s %.3g parity %s" % (d["value"]/1e6, d["ms_per_step"], r.get("records_ms",0), r.get("default_mode_ms_per_step",0), r.get("default_mode_records_ms",0), r.get("default_mode_events_per_s",0) or 0, d["config"].get("parity_checked")))
PY
)
ls -la $out | grep ${tag}
