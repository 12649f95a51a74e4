#!/bin/bash
# One GPU-box session: bench A/B, parity tests, launch list, ncu capture of the CRC kernel.  Outputs under gpurun_out/<tag>_*.
# usage (from the repo root on the box): bash tools/gpu_check.sh r1m
tag=${1:-run}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $out/${tag}_gpu.txt 2>&1
( timeout 300 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_f003.json 2> $out/${tag}_bench_f003.err; echo "bench f0.03 rc=$?" )
( MSD_NO_CRC=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_f003_nocrc.json 2> $out/${tag}_bench_f003_nocrc.err; echo "bench nocrc rc=$?" )
( timeout 400 python bench.py --steps 5 --warmup 3 --genome-fraction 0.06 > $out/${tag}_bench_f006.json 2> $out/${tag}_bench_f006.err; echo "bench f0.06 rc=$?" )
( timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log )
( timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?" )
tools/gen_bam --config c3 --genome-fraction 0.0288 --threads 16 -o /dev/shm/p.bam > /dev/null 2>&1
( MSD_CHUNK_BLOCKS=200000 MSD_CHUNK_INFL_MB=8000 MSD_CHUNK_COMP_MB=3000 MSD_RAMP0_BLOCKS=200000 MSD_RAMP1_BLOCKS=200000 timeout 300 \
  ncu --set full --clock-control none --import-source on -k regex:bgzf_crc -c 1 -f -o $out/${tag}_crc \
  mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by 500 /tmp/p /dev/shm/p.bam > $out/${tag}_ncu_crc.log 2>&1; echo "ncu crc rc=$?" )
ncu -i $out/${tag}_crc.ncu-rep --page raw --csv > $out/${tag}_crc_raw.csv 2>/dev/null
for f in $out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f M  e2e %.1f M  ms %.2f  inflate_span %.2f" % (j["value"] / 1e6, j["e2e"]["value"] / 1e6, j["ms_per_step"], j["roofline"]["stage_ms_per_step"]["inflate_span"]))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
