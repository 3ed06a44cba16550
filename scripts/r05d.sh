# usage: bash scripts/r05d.sh N   (inside gpurun --gpus N): cfg2 (default) and cfg4 at N ranks
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
n=$1
run() { # tag, args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline $2 > $O/r05d_bench_$1_n${n}.json 2> $O/r05d_bench_$1_n${n}.err
  echo "n=$n $1 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('$O/r05d_bench_$1_n${n}.json').read().strip().splitlines()[-1])
    print('$1 n$n', d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('forward') or {}).get('value'))
except Exception as e:
    print('no line', e)
PY
  grep -v "^\*\|OMP_NUM\|^$" $O/r05d_bench_$1_n${n}.err | tail -3
}
run cfg2 ""
