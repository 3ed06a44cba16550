cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/r01e_pytest_gpu.log 2>&1; tail -8 $O/r01e_pytest_gpu.log
grep -E "Error|error|assert |mismatch" $O/r01e_pytest_gpu.log | head -20
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'launches',d['gpu_launches'],d['roofline']['kernel'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['roofline']['own_kernels_ms_per_step'],d['roofline']['own_kernel_shares'])" || tail -5 $O/bench.err; }
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r01e_bench_cfg2_graph.json 2>> $O/bench.err; show $O/r01e_bench_cfg2_graph.json cfg2-graph
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > $O/r01e_bench_cfg2_nograph.json 2>> $O/bench.err; show $O/r01e_bench_cfg2_nograph.json cfg2-nograph
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline > $O/r01e_bench_cfg3_graph.json 2>> $O/bench.err; show $O/r01e_bench_cfg3_graph.json cfg3-graph
timeout 300 python scripts/profile_step.py cfg2_train > $O/r01e_cfg2_train_torch_profiler.txt 2>&1
tail -5 $O/bench.err
