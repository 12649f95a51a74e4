#!/bin/bash
# Round-end style check on one GPU box: smoke(), the gpu test suite, both bench arms, launch list.  usage: bash tools/gpu_final.sh r1n
tag=${1:-final}
out=gpurun_out
mkdir -p $out
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log )
( timeout 400 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?" )
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; echo "reference arm rc=$?" )
( timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log )
( timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?" )
for f in $out/${tag}_bench_n1.json $out/${tag}_bench_reference_arm.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f M  e2e %.1f M  ms %.2f" % (j["value"] / 1e6, j["e2e"]["value"] / 1e6, j["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
