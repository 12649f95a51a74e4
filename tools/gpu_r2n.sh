#!/bin/bash
# Round 2: K6 tile size and L2 prefetch-ahead variants (MSD_SCAN = 0 / 3 / 5 / 6), then ncu --set full of 0 and 6 at 370 M cells.
tag=${1:-r2n}
out=gpurun_out
mkdir -p $out
for v in 3 5 6; do ( MSD_SCAN=$v timeout 200 python -m pytest tests/test_gpu_stages.py -x -q > $out/${tag}_pytest_stages_scan$v.log 2>&1; echo "pytest stages MSD_SCAN=$v rc=$?"; tail -2 $out/${tag}_pytest_stages_scan$v.log ); done
( timeout 500 python tools/sweep_scan.py 0.12 > $out/${tag}_sweep_scan_f0.12.log 2>&1; echo "sweep_scan rc=$?"; grep -E "^rep" $out/${tag}_sweep_scan_f0.12.log )
B=/dev/shm/msd_r2n; mkdir -p $B
tools/gen_bam --config c3 --genome-fraction 0.12 --threads $(nproc) -o $B/p.bam > /dev/null 2>&1
for v in 0 6; do
  ( MSD_SCAN=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:"scan_segments" -c 1 -f -o $out/${tag}_scan${v}_full \
    mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by 500 $B/pf $B/p.bam > $out/${tag}_ncu_scan$v.log 2>&1; echo "ncu scan variant $v rc=$?" )
  ncu -i $out/${tag}_scan${v}_full.ncu-rep --page raw --csv > $out/${tag}_scan${v}_full_raw.csv 2>/dev/null
done
rm -rf $B
ls -la $out | grep ${tag}
