#!/bin/bash
# First GPU session for K10 (msd_per_base_bgzf) and the host writer paths, none of which has run on a GPU yet:
# the new gpu tests, the A/B against msd_per_base_runs + zlib, a launch list and one full ncu capture of the emit kernel.
# Outputs under gpurun_out/<tag>_*.   usage (repo root on the box): bash tools/gpu_k10.sh r2a
tag=${1:-k10}
out=gpurun_out
mkdir -p $out
( timeout 600 python -m pytest tests/test_gpu_zzz_bedgz.py tests/test_gpu_zz_cli_gz.py -q > $out/${tag}_pytest_k10.log 2>&1; echo "pytest k10 rc=$?"; tail -5 $out/${tag}_pytest_k10.log )
( timeout 600 python tools/bench_bedgz.py 0.02 3 > $out/${tag}_bench_bedgz.log 2>&1; echo "bench_bedgz rc=$?"; tail -3 $out/${tag}_bench_bedgz.log )
tools/gen_bam --config c3 --genome-fraction 0.02 --threads 16 -o /dev/shm/k10.bam > /dev/null 2>&1
for v in 0 1; do
  t0=$(date +%s.%N)
  MOSDEPTH_B200_DEVICE_BGZF=$v timeout 300 mosdepth_b200/bin/mosdepth-b200 -q 0:1:10:30: /tmp/k10_$v /dev/shm/k10.bam > /dev/null 2> $out/${tag}_cli_bgzf$v.err
  echo "cli per-base + quantized, DEVICE_BGZF=$v: rc=$? $(python -c "import sys; print(round(float(sys.argv[1]) - float(sys.argv[2]), 2))" $(date +%s.%N) $t0) s wall" | tee -a $out/${tag}_cli_cmp.log
done
( cmp <(gzip -dc /tmp/k10_0.per-base.bed.gz) <(gzip -dc /tmp/k10_1.per-base.bed.gz) && echo "per-base content identical"; ls -l /tmp/k10_0.per-base.bed.gz /tmp/k10_1.per-base.bed.gz ) >> $out/${tag}_cli_cmp.log 2>&1
( MOSDEPTH_B200_DEVICE_BGZF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches_k10.csv mosdepth_b200/bin/mosdepth-b200 /tmp/k10_n /dev/shm/k10.bam > $out/${tag}_ncu_launches_k10.log 2>&1; echo "launch list rc=$?" )
( MOSDEPTH_B200_DEVICE_BGZF=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:bz_emit -s 1 -c 1 -f -o $out/${tag}_bz_emit mosdepth_b200/bin/mosdepth-b200 /tmp/k10_f /dev/shm/k10.bam > $out/${tag}_ncu_bz_emit.log 2>&1; echo "ncu bz_emit rc=$?" )
ncu -i $out/${tag}_bz_emit.ncu-rep --page raw --csv > $out/${tag}_bz_emit_raw.csv 2>/dev/null
