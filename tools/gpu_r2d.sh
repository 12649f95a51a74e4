#!/bin/bash
# Round 2, fourth GPU session (one GPU): literal packing in K1 (decode writes literals packed, resolve writes whole windows) -- correctness,
# A/B against the previous build (mosdepth_b200/libmosdepth_b200_prev.so, same box, interleaved), then ncu --set full of every kernel of the
# shipped build and the launch list of the bench.
# usage (from the repo root on the box): bash tools/gpu_r2d.sh r2d
tag=${1:-r2d}
out=gpurun_out
mkdir -p $out
( timeout 600 python -m pytest tests/test_gpu_stages.py -x -q > $out/${tag}_pytest_stages.log 2>&1; echo "pytest stages rc=$?"; tail -3 $out/${tag}_pytest_stages.log )
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log )
for rep in 1 2; do
  ( MSD_LIB_PATH=$PWD/mosdepth_b200/libmosdepth_b200_prev.so timeout 300 python tools/sweep_inflate.py 0.06 one > $out/${tag}_ab_prev_$rep.log 2>&1; echo "A/B prev $rep rc=$?"; grep -E "^rep" $out/${tag}_ab_prev_$rep.log | tail -6 )
  ( timeout 300 python tools/sweep_inflate.py 0.06 one > $out/${tag}_ab_new_$rep.log 2>&1; echo "A/B new $rep rc=$?"; grep -E "^rep" $out/${tag}_ab_new_$rep.log | tail -6 )
done
B=/dev/shm/msd_r2d; mkdir -p $B
tools/gen_bam --config c3 --genome-fraction 0.0288 --threads $(nproc) -o $B/p.bam > /dev/null 2>&1
( MSD_CHUNK_BLOCKS=200000 MSD_CHUNK_INFL_MB=8000 MSD_CHUNK_COMP_MB=3000 MSD_RAMP0_BLOCKS=200000 timeout 900 \
  ncu --set full --clock-control none -k regex:"huffman|lz77|bgzf_crc|frame_|records_kernel|scan_segments|totals_add|windows_all" -c 14 -f -o $out/${tag}_fast_full \
  mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by 500 $B/pf $B/p.bam > $out/${tag}_ncu_fast.log 2>&1; echo "ncu fast rc=$?" )
ncu -i $out/${tag}_fast_full.ncu-rep --page raw --csv > $out/${tag}_fast_full_raw.csv 2>/dev/null
tools/gen_bam --config c4 --genome-fraction 0.05 --threads $(nproc) --bed $B/c4.bed -o $B/c4.bam > /dev/null 2>&1
( timeout 600 ncu --set full --clock-control none -k regex:"regions_kernel|keep_|runs_" -c 12 -f -o $out/${tag}_regions_full \
  mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by $B/c4.bed -T 1,10,20,30 -m -q 0:1:10:30: -c chr1 $B/pr $B/c4.bam > $out/${tag}_ncu_regions.log 2>&1; echo "ncu regions (median) rc=$?" )
ncu -i $out/${tag}_regions_full.ncu-rep --page raw --csv > $out/${tag}_regions_full_raw.csv 2>/dev/null
( timeout 600 ncu --set full --clock-control none -k regex:"regions_kernel" -c 1 -f -o $out/${tag}_regions_mean_full \
  mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by $B/c4.bed -c chr1 $B/pr2 $B/c4.bam > $out/${tag}_ncu_regions_mean.log 2>&1; echo "ncu regions (mean) rc=$?" )
ncu -i $out/${tag}_regions_mean_full.ncu-rep --page raw --csv > $out/${tag}_regions_mean_full_raw.csv 2>/dev/null
rm -rf $B
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches_bench_c3_f0.06.csv python bench.py --steps 1 --warmup 1 --genome-fraction 0.06 --no-cpu-baseline --no-cli > $out/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?" )
rm -f $out/*.ncu-rep.tmp
ls -la $out | grep ${tag}
