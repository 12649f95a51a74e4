#!/bin/bash
# Round 2: ncu --set full of K6 (classic and pipelined) at 370 M cells, of the new K1c, and the per-kernel times of K10 on C2; host laps of the CLI.
tag=${1:-r2l}
out=gpurun_out
mkdir -p $out
B=/dev/shm/msd_r2l; mkdir -p $B
tools/gen_bam --config c3 --genome-fraction 0.12 --threads $(nproc) -o $B/p.bam > /dev/null 2>&1
for v in 0 1; do
  ( MSD_SCAN=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:"scan_segments" -c 1 -f -o $out/${tag}_scan${v}_full \
    mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by 500 $B/pf $B/p.bam > $out/${tag}_ncu_scan$v.log 2>&1; echo "ncu scan variant $v rc=$?" )
done
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:"bgzf_crc" -s 1 -c 1 -f -o $out/${tag}_crc_full \
  mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by 500 $B/pf $B/p.bam > $out/${tag}_ncu_crc.log 2>&1; echo "ncu crc rc=$?" )
rm -f $B/p.bam*
tools/gen_bam --config c2 --genome-fraction 1.0 --threads $(nproc) -o $B/c2.bam > /dev/null 2>&1
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bz_|runs_|bgzf_crc" --csv --log-file $out/${tag}_launches_c2_k10.csv \
  mosdepth_b200/bin/mosdepth-b200 $B/o $B/c2.bam > $out/${tag}_ncu_k10.log 2>&1; echo "ncu k10 launch list rc=$?" )
TIMEFORMAT='wall %R s user %U s sys %S s'
for i in 1 2; do { time MSD_TRACE=1 MOSDEPTH_B200_STATS=1 mosdepth_b200/bin/mosdepth-b200 $B/o $B/c2.bam > $out/${tag}_cli_c2_trace_$i.log 2>&1 ; } 2> $out/${tag}_cli_c2_$i.time; echo "c2 cli run $i: $(cat $out/${tag}_cli_c2_$i.time)"; grep -E "create:|bgzf:|wall_s" $out/${tag}_cli_c2_trace_$i.log | head -40; done
rm -rf $B
for f in scan0 scan1 crc; do ncu -i $out/${tag}_${f}_full.ncu-rep --page raw --csv > $out/${tag}_${f}_full_raw.csv 2>/dev/null; done
ls -la $out | grep ${tag}
