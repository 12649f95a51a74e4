#!/bin/bash
# Round 2, second GPU session: the new code on one GPU -- new tests first (fail fast), the whole gpu suite, the bench at the round-1 size
# (A/B against r2a_bench_base.json of the same day), the CLI's wall clock with its stages, then the bench at its new default size.
# usage (from the repo root on the box): bash tools/gpu_r2b.sh r2b
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
( timeout 600 python -m pytest tests/test_gpu_intra_contig.py tests/test_gpu_multi.py "tests/test_gpu_stages.py" -x -q > $out/${tag}_pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -5 $out/${tag}_pytest_new.log )
( timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log )
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log )
( timeout 600 python bench.py --steps 5 --warmup 3 --genome-fraction 0.06 > $out/${tag}_bench_f006.json 2> $out/${tag}_bench_f006.err; echo "bench f0.06 rc=$?"; tail -3 $out/${tag}_bench_f006.err )
# ---- CLI wall clock, C3 x 0.12 (74 M reads, 6.7 GB BAM) -------------------------------------------------------------------
B=/dev/shm/msd_r2b; mkdir -p $B
TIMEFORMAT='wall %R s user %U s sys %S s'
{ time tools/gen_bam --config c3 --genome-fraction 0.12 --seed 20261022 --threads $(nproc) --level 6 -o $B/c3.bam > /dev/null 2>&1 ; } 2> $out/${tag}_gen_f012.log; echo "gen f0.12: $(cat $out/${tag}_gen_f012.log)"
for i in 1 2; do { time MSD_TRACE=1 MOSDEPTH_B200_STATS=1 mosdepth_b200/bin/mosdepth-b200 -n -x --by 500 $B/o_gz $B/c3.bam > $out/${tag}_cli_gz_$i.log 2>&1 ; } 2> $out/${tag}_cli_gz_$i.time; echo "cli gz $i: $(cat $out/${tag}_cli_gz_$i.time) $(grep wall_s $out/${tag}_cli_gz_$i.log)"; done
{ time MOSDEPTH_B200_STATS=1 mosdepth_b200/bin/mosdepth-b200 --plain -n -x --by 500 $B/o_pl $B/c3.bam > $out/${tag}_cli_plain.log 2>&1 ; } 2> $out/${tag}_cli_plain.time; echo "cli plain: $(cat $out/${tag}_cli_plain.time) $(grep wall_s $out/${tag}_cli_plain.log)"
for t in 7 15; do { time oracle/mosdepth_oracle -t $t -n -x --by 500 $B/o_or $B/c3.bam > /dev/null 2>&1 ; } 2> $out/${tag}_oracle_t$t.time; echo "oracle -t $t: $(cat $out/${tag}_oracle_t$t.time)"; done
( zcat $B/o_gz.regions.bed.gz | cmp - $B/o_or.regions.bed && cmp $B/o_gz.mosdepth.summary.txt $B/o_or.mosdepth.summary.txt && cmp $B/o_pl.regions.bed $B/o_or.regions.bed && echo "cli f=0.12 parity ok" ) 2>&1 | tail -1
rm -rf $B
( timeout 1500 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_f05.json 2> $out/${tag}_bench_f05.err; echo "bench f0.5 rc=$?"; tail -3 $out/${tag}_bench_f05.err )
( timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "bench ref rc=$?" )
for f in $out/${tag}_bench_f006.json $out/${tag}_bench_f05.json $out/${tag}_bench_ref.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = j.get("roofline") or {}
    print(sys.argv[1], "value %.1f M  e2e %.1f M  ms %.2f  inflate_span %s scan_ms %s k6 %s parity %s cli %s cpu %s" % (j["value"] / 1e6, j["e2e"]["value"] / 1e6, j["ms_per_step"], r.get("inflate_span_ms"), r.get("scan_ms"), r.get("k6_scan_frac"), j["config"].get("parity_checked"), j["e2e"].get("cli_wall_s"), (j.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
ls -la $out | grep ${tag}
