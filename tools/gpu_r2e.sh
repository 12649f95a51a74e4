#!/bin/bash
# Round 2, multi-GPU session (N GPUs of one box, N = $2): the multi-GPU product path on hardware -- `mosdepth-b200 --gpus N` against the oracle
# in nine modes, the bench under torchrun (library-side NCCL spill exchange + totals all-reduce, parity gate across N ranks), the CLI's wall clock
# at N GPUs, and the host->device ceiling of the box at 1..N GPUs.
# usage (from the repo root on the box): bash tools/gpu_r2e.sh r2e 2 [genome_fraction]
tag=${1:-r2e}; N=${2:-2}; F=${3:-0.12}
out=gpurun_out
mkdir -p $out
TIMEFORMAT='wall %R s user %U s sys %S s'
{ nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader; nproc; free -g | head -2; nvidia-smi topo -m; } > $out/${tag}_box.txt 2>&1
for n in ${PROBE_NS:-1 2 4 8}; do [ $n -le $N ] && ( timeout 180 tools/h2d_probe $n 4 256 8 > $out/${tag}_h2d_probe_n$n.json 2> $out/${tag}_h2d_probe_n$n.err; echo "h2d probe n=$n: $(cut -c1-170 $out/${tag}_h2d_probe_n$n.json)" ); done
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_intra_contig.py -q ${PYTEST_K:+-k "$PYTEST_K"} > $out/${tag}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $out/${tag}_pytest_multi.log )
if [ "$K10" = 1 ]; then ( timeout 600 python -m pytest tests/test_gpu_zzz_bedgz.py tests/test_gpu_zz_cli_gz.py tests/test_gpu_stages.py -x -q > $out/${tag}_pytest_k10.log 2>&1; echo "pytest k10/stages rc=$?"; tail -3 $out/${tag}_pytest_k10.log ); fi
if [ "$C5" = 1 ]; then ( timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --genome-fraction 0.05 --no-cpu-baseline > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; echo "bench c5 rc=$?"; python -c "
import json; j = json.loads(open('$out/${tag}_bench_c5.json').read().strip().splitlines()[-1]); r = j['roofline']
print('c5 value %.2f M reads/s ms %.1f inflate_span %.1f frame_ms %.1f default_mode_ms %.1f parity %s' % (j['value'] / 1e6, j['ms_per_step'], r['inflate_span_ms'], r['frame_ms'], r.get('default_mode_ms_per_step', 0), j['config']['parity_checked']))" ); fi
( timeout 900 python bench.py --steps 5 --warmup 3 --genome-fraction $F --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench N=1 rc=$?"; tail -2 $out/${tag}_bench_n1.err )
for n in ${BENCH_NS:-2 4 8}; do [ $n -le $N ] && ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510 + n)) bench.py --gpus $n --steps 5 --warmup 3 --genome-fraction $F > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err; echo "bench N=$n rc=$?"; tail -2 $out/${tag}_bench_n$n.err ); done
for f in $out/${tag}_bench_n*.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = j.get("roofline") or {}
    print(sys.argv[1], "N %d value %.1f M  e2e %.1f M  ms %.2f / %.2f  inflate_span %s scan_ms %s k6 %s parity %s cli %s" % (j["n_gpus"], j["value"] / 1e6, j["e2e"]["value"] / 1e6, j["ms_per_step"], j["e2e"]["ms_per_step"], r.get("inflate_span_ms"), r.get("scan_ms"), r.get("k6_scan_frac"), j["config"].get("parity_checked"), j["e2e"].get("cli_wall_s")))
    print("   ", (j.get("detail") or {}).get("parity"), (j.get("detail") or {}).get("cli_stages"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
ls -la $out | grep ${tag}
