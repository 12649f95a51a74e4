#!/bin/bash
# Round 2, the shipped build once more: stage tests + smoke().
out=gpurun_out; tag=${1:-r2q}; mkdir -p $out
( timeout 90 python -m pytest tests/test_gpu_stages.py -x -q > $out/${tag}_pytest_stages.log 2>&1; echo "pytest stages rc=$?"; tail -2 $out/${tag}_pytest_stages.log )
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log )
