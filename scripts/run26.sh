cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "edge_attention_fused_backward" > $O/r01k_pytest_edge.log 2>&1; tail -2 $O/r01k_pytest_edge.log
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu > $O/r01k_pytest_model.log 2>&1; tail -2 $O/r01k_pytest_model.log
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'launches',d['gpu_launches'],d['roofline']['kernel'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['roofline']['own_kernels_ms_per_step'],d['roofline']['own_kernel_shares'])" || tail -5 $O/bench.err; }
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r01k_bench_cfg2.json 2>> $O/bench.err; show $O/r01k_bench_cfg2.json cfg2
