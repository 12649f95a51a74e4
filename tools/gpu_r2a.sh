#!/bin/bash
# Round 2, first GPU session: what the box can do (H2D probe), the round-1 build on this box (baseline), the CLI's wall clock with its
# stages, and ncu --set full captures of every kernel of the path (K1a/K1b/K1c re-profiled at the shipped build, K2..K9 for the first time).
# usage (from the repo root on the box): bash tools/gpu_r2a.sh r2a
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader; nproc; free -g | head -2; df -h /dev/shm /tmp | cat; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; } > $out/${tag}_box.txt 2>&1
( timeout 120 tools/h2d_probe 1 4 256 8 > $out/${tag}_h2d_probe_n1.json 2> $out/${tag}_h2d_probe_n1.err; echo "h2d probe rc=$?"; cat $out/${tag}_h2d_probe_n1.json )
( timeout 120 tools/h2d_probe 1 4 64 16 > $out/${tag}_h2d_probe_n1_p64.json 2>> $out/${tag}_h2d_probe_n1.err; cat $out/${tag}_h2d_probe_n1_p64.json )
( timeout 400 python bench.py --steps 5 --warmup 3 --no-aux > $out/${tag}_bench_base.json 2> $out/${tag}_bench_base.err; echo "bench base rc=$?" )
# ---- CLI wall clock, C3 x 0.12 (74 M reads, 6.7 GB BAM) -------------------------------------------------------------------
B=/dev/shm/msd_r2a; mkdir -p $B
( /usr/bin/time -f "gen_bam f=0.12 wall %e s user %U s" tools/gen_bam --config c3 --genome-fraction 0.12 --seed 20261022 --threads $(nproc) --level 6 -o $B/c3.bam > $out/${tag}_gen.log 2>&1; tail -2 $out/${tag}_gen.log )
for i in 1 2; do ( MSD_TRACE=1 MOSDEPTH_B200_STATS=1 /usr/bin/time -f "cli gz run $i wall %e s user %U s sys %S s maxrss %M KB" mosdepth_b200/bin/mosdepth-b200 -n -x --by 500 $B/o_gz $B/c3.bam > $out/${tag}_cli_gz_$i.log 2>&1; tail -1 $out/${tag}_cli_gz_$i.log ); done
( MOSDEPTH_B200_STATS=1 /usr/bin/time -f "cli plain wall %e s user %U s sys %S s" mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by 500 $B/o_pl $B/c3.bam > $out/${tag}_cli_plain.log 2>&1; tail -1 $out/${tag}_cli_plain.log )
for t in 3 15; do ( /usr/bin/time -f "oracle -t $t wall %e s user %U s" oracle/mosdepth_oracle -t $t -n -x --by 500 $B/o_or $B/c3.bam > $out/${tag}_oracle_t$t.log 2>&1; tail -1 $out/${tag}_oracle_t$t.log ); done
( zcat $B/o_gz.regions.bed.gz | cmp - $B/o_or.regions.bed && cmp $B/o_gz.mosdepth.summary.txt $B/o_or.mosdepth.summary.txt && echo "cli f=0.12 parity ok" ) 2>&1 | tail -1
rm -f $B/c3.bam $B/c3.bam.bai $B/o_*
# ---- ncu --set full: fast mode, one full-size chunk (85 k blocks) ------------------------------------------------------------
tools/gen_bam --config c3 --genome-fraction 0.0288 --threads $(nproc) -o $B/p.bam > /dev/null 2>&1
( MSD_CHUNK_BLOCKS=200000 MSD_CHUNK_INFL_MB=8000 MSD_CHUNK_COMP_MB=3000 MSD_RAMP0_BLOCKS=200000 timeout 600 \
  ncu --set full --clock-control none -c 90 -f -o $out/${tag}_fast_full \
  mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by 500 $B/pf $B/p.bam > $out/${tag}_ncu_fast.log 2>&1; echo "ncu fast rc=$?" )
ncu -i $out/${tag}_fast_full.ncu-rep --page raw --csv > $out/${tag}_fast_full_raw.csv 2>/dev/null
# ---- ncu --set full: default mode + per-base runs on a chr20-like contig (C2 x 0.25) -----------------------------------------
tools/gen_bam --config c2 --genome-fraction 0.25 --threads $(nproc) -o $B/c2.bam > /dev/null 2>&1
( timeout 600 ncu --set full --clock-control none -k regex:"mate_|records_kernel|runs_|scan_fused|frame_" -c 24 -f -o $out/${tag}_default_full \
  mosdepth_b200/bin/mosdepth-b200 --plain $B/pd $B/c2.bam > $out/${tag}_ncu_default.log 2>&1; echo "ncu default rc=$?" )
ncu -i $out/${tag}_default_full.ncu-rep --page raw --csv > $out/${tag}_default_full_raw.csv 2>/dev/null
# ---- ncu --set full: BED regions, mean and median + thresholds (C4 x 0.05) ---------------------------------------------------
tools/gen_bam --config c4 --genome-fraction 0.05 --threads $(nproc) --bed $B/c4.bed -o $B/c4.bam > /dev/null 2>&1
( timeout 600 ncu --set full --clock-control none -k regex:"regions_kernel|records_kernel" -c 6 --launch-skip 0 -f -o $out/${tag}_regions_full \
  mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by $B/c4.bed -T 1,10,20,30 -m -c chr1 $B/pr $B/c4.bam > $out/${tag}_ncu_regions.log 2>&1; echo "ncu regions rc=$?" )
ncu -i $out/${tag}_regions_full.ncu-rep --page raw --csv > $out/${tag}_regions_full_raw.csv 2>/dev/null
( /usr/bin/time -f "cli c4 -m -T wall %e s" mosdepth_b200/bin/mosdepth-b200 --plain -n --by $B/c4.bed -T 1,10,20,30 -m -q 0:1:10:30: $B/pr2 $B/c4.bam 2>&1 | tail -1 )
( /usr/bin/time -f "oracle c4 -m -T wall %e s" oracle/mosdepth_oracle -t 3 -n --by $B/c4.bed -T 1,10,20,30 -m -q 0:1:10:30: $B/pr3 $B/c4.bam 2>&1 | tail -1 )
rm -rf $B
python - "$out/${tag}_bench_base.json" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f M  e2e %.1f M  ms %.2f  inflate_span %.2f cpu %.2f M (%s cores)" % (j["value"] / 1e6, j["e2e"]["value"] / 1e6, j["ms_per_step"], j["roofline"]["stage_ms_per_step"]["inflate_span"], j["cpu_baseline"]["value"] / 1e6, j["cpu_baseline"]["cores"]))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
ls -la $out | grep ${tag}
