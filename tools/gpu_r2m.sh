#!/bin/bash
# Round 2: K6 with hoisted loads and lean stats (sum / min / max off the bins, branch-free run atomics), four variants: registers (0), registers + wide
# look-back (2), shared-memory ring (1), ring + wide look-back (4).
tag=${1:-r2m}
out=gpurun_out
mkdir -p $out
for v in 0 2 1 4; do ( MSD_SCAN=$v timeout 200 python -m pytest tests/test_gpu_stages.py -x -q > $out/${tag}_pytest_stages_scan$v.log 2>&1; echo "pytest stages MSD_SCAN=$v rc=$?"; tail -2 $out/${tag}_pytest_stages_scan$v.log ); done
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log )
( timeout 500 python tools/sweep_scan.py 0.12 > $out/${tag}_sweep_scan_f0.12.log 2>&1; echo "sweep_scan rc=$?"; grep -E "^rep" $out/${tag}_sweep_scan_f0.12.log )
ls -la $out | grep ${tag}
