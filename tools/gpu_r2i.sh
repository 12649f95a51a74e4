#!/bin/bash
# Round 2: second fused-K1 variant (4 decode warps at 96 registers + 20 resolve warps per CTA, mosdepth_b200/libmosdepth_b200_fused20.so) against the
# two-kernel pipeline, interleaved; then C2 with K10 as the CLI's default for .gz output.
tag=${1:-r2i}
out=gpurun_out
mkdir -p $out
( MSD_FUSED=1 MSD_LIB_PATH=$PWD/mosdepth_b200/libmosdepth_b200_fused20.so timeout 300 python -m pytest tests/test_gpu_stages.py -x -q -k "inflate or crc" > $out/${tag}_pytest_stages_fused20.log 2>&1; echo "pytest stages (fused20) rc=$?"; tail -2 $out/${tag}_pytest_stages_fused20.log )
for rep in 1 2; do
  ( MSD_FUSED=0 timeout 300 python tools/sweep_inflate.py 0.06 one > $out/${tag}_ab_two_$rep.log 2>&1; echo "A/B two kernels $rep rc=$?"; grep -E "^rep" $out/${tag}_ab_two_$rep.log | tail -6 )
  ( MSD_FUSED=1 MSD_LIB_PATH=$PWD/mosdepth_b200/libmosdepth_b200_fused20.so timeout 300 python tools/sweep_inflate.py 0.06 one > $out/${tag}_ab_fused20_$rep.log 2>&1; echo "A/B fused20 $rep rc=$?"; grep -E "^rep" $out/${tag}_ab_fused20_$rep.log | tail -6 )
done
TIMEFORMAT='wall %R s user %U s sys %S s'
B=/dev/shm/msd_r2i; mkdir -p $B
tools/gen_bam --config c2 --genome-fraction 1.0 --threads $(nproc) -o $B/c2.bam > /dev/null 2>&1
for v in 1 0; do for i in 1 2; do { time MOSDEPTH_B200_DEVICE_BGZF=$v MOSDEPTH_B200_STATS=1 mosdepth_b200/bin/mosdepth-b200 $B/o_$v $B/c2.bam > $out/${tag}_cli_c2_dev$v.log 2>&1 ; } 2> $out/${tag}_cli_c2_dev$v.time; echo "c2 cli .gz device_bgzf=$v run $i: $(cat $out/${tag}_cli_c2_dev$v.time) $(grep wall_s $out/${tag}_cli_c2_dev$v.log)"; done; done
{ time oracle/mosdepth_oracle -t 15 $B/o_or $B/c2.bam > /dev/null 2>&1 ; } 2> $out/${tag}_oracle_c2.time; echo "oracle c2 -t 15: $(cat $out/${tag}_oracle_c2.time)"
( zcat $B/o_1.per-base.bed.gz | cmp - $B/o_or.per-base.bed && zcat $B/o_0.per-base.bed.gz | cmp - $B/o_or.per-base.bed && echo "c2 per-base parity ok (device and host writers)" ) 2>&1 | tail -1
ls -la $B | head; rm -rf $B
ls -la $out | grep ${tag}
