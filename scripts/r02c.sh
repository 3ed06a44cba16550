# usage: bash scripts/r02c.sh N [tag]   (inside gpurun --gpus N)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
n=$1
T=${2:-r02c}
run() { # tag, extra args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline $2 > $O/${T}_bench_cfg2_n${n}_$1.json 2> $O/${T}_bench_n${n}_$1.err
  echo "n=$n $1 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('$O/${T}_bench_cfg2_n${n}_$1.json').read().strip().splitlines()[-1])
    print('cfg2 n$n $1', d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('forward') or {}).get('value'))
except Exception as e:
    print('no line', e)
PY
  grep -v "^\*\|OMP_NUM\|^$" $O/${T}_bench_n${n}_$1.err | tail -4
}
run split ""
if [ "$3" = "both" ]; then run graphcoll "--graph-collectives --no-forward-record"; fi
