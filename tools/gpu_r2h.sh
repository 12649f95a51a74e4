#!/bin/bash
# Round 2, fused K1 (inflate_fused_kernel, MSD_FUSED=1): correctness against zlib / the oracle, A/B against the two-kernel pipeline in one
# session (interleaved), ncu --set full of the fused kernel.
# usage (from the repo root on the box): bash tools/gpu_r2h.sh r2h
tag=${1:-r2h}
out=gpurun_out
mkdir -p $out
( MSD_FUSED=1 timeout 300 python -m pytest tests/test_gpu_stages.py -x -q > $out/${tag}_pytest_stages_fused.log 2>&1; echo "pytest stages (fused) rc=$?"; tail -4 $out/${tag}_pytest_stages_fused.log )
( MSD_FUSED=1 timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke_fused.log 2>&1; echo "smoke (fused) rc=$?"; tail -1 $out/${tag}_smoke_fused.log )
( timeout 300 python -m pytest tests/test_gpu_stages.py -x -q > $out/${tag}_pytest_stages.log 2>&1; echo "pytest stages (two kernels, refactored) rc=$?"; tail -2 $out/${tag}_pytest_stages.log )
for rep in 1 2; do
  ( MSD_FUSED=0 timeout 300 python tools/sweep_inflate.py 0.06 one > $out/${tag}_ab_two_$rep.log 2>&1; echo "A/B two kernels $rep rc=$?"; grep -E "^rep" $out/${tag}_ab_two_$rep.log | tail -6 )
  ( MSD_FUSED=1 timeout 300 python tools/sweep_inflate.py 0.06 one > $out/${tag}_ab_fused_$rep.log 2>&1; echo "A/B fused $rep rc=$?"; grep -E "^rep" $out/${tag}_ab_fused_$rep.log | tail -6 )
done
( MSD_FUSED=1 timeout 600 python -m pytest tests/test_gpu_cli_parity.py -x -q > $out/${tag}_pytest_cli_fused.log 2>&1; echo "pytest cli parity (fused) rc=$?"; tail -3 $out/${tag}_pytest_cli_fused.log )
B=/dev/shm/msd_r2h; mkdir -p $B
tools/gen_bam --config c3 --genome-fraction 0.0288 --threads $(nproc) -o $B/p.bam > /dev/null 2>&1
( MSD_FUSED=1 MSD_CHUNK_BLOCKS=200000 MSD_CHUNK_INFL_MB=8000 MSD_CHUNK_COMP_MB=3000 MSD_RAMP0_BLOCKS=200000 timeout 600 \
  ncu --set full --clock-control none -k regex:"inflate_fused" -c 1 -f -o $out/${tag}_fused_full \
  mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by 500 $B/pf $B/p.bam > $out/${tag}_ncu_fused.log 2>&1; echo "ncu fused rc=$?" )
ncu -i $out/${tag}_fused_full.ncu-rep --page raw --csv > $out/${tag}_fused_full_raw.csv 2>/dev/null
rm -rf $B
ls -la $out | grep ${tag}
