#!/bin/bash
# Round 2: K6 variants (classic / pipelined / pipelined + wide look-back): correctness of the stage tests under each, then the interleaved timing.
tag=${1:-r2k}
out=gpurun_out
mkdir -p $out
for v in 1 4 0; do ( MSD_SCAN=$v timeout 300 python -m pytest tests/test_gpu_stages.py -x -q > $out/${tag}_pytest_stages_scan$v.log 2>&1; echo "pytest stages MSD_SCAN=$v rc=$?"; tail -2 $out/${tag}_pytest_stages_scan$v.log ); done
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log )
( timeout 500 python tools/sweep_scan.py 0.12 > $out/${tag}_sweep_scan_f0.12.log 2>&1; echo "sweep_scan rc=$?"; grep -E "^rep" $out/${tag}_sweep_scan_f0.12.log )
ls -la $out | grep ${tag}
