cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "edge_attention and f16x3" > $O/r01i_pytest_edge.log 2>&1; tail -2 $O/r01i_pytest_edge.log
n=2
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $O/r01i_bench_cfg2_n$n.json 2> $O/bench_n$n.err; echo rc=$?; tail -c 1500 $O/r01i_bench_cfg2_n$n.json; grep -v "^\*\|OMP_NUM\|^$" $O/bench_n$n.err | tail -8
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r01i_bench_cfg2_n1.json 2> $O/bench_n1.err; python -c "import json;d=json.loads(open('$O/r01i_bench_cfg2_n1.json').read().strip().splitlines()[-1]);print('cfg2 n1',d['value'],d['ms_per_step'],d['roofline']['own_kernel_shares'])"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg3_infer > $O/r01i_bench_cfg3_n1.json 2>> $O/bench_n1.err; python -c "import json;d=json.loads(open('$O/r01i_bench_cfg3_n1.json').read().strip().splitlines()[-1]);print('cfg3 n1',d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['own_kernel_shares'])"
