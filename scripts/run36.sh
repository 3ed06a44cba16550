cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
n=4
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > $O/r01t_bench_cfg2_n$n.json 2> $O/bench_n$n.err; echo rc=$?; python -c "import json;d=json.loads(open('$O/r01t_bench_cfg2_n$n.json').read().strip().splitlines()[-1]);print('cfg2 n$n',d['value'],d['ms_per_step'],d['e2e']['value'])"; grep -v "^\*\|OMP_NUM\|^$" $O/bench_n$n.err | tail -4
