cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/r01s_pytest_gpu.log 2>&1; tail -6 $O/r01s_pytest_gpu.log
grep -E "^FAILED|^ERROR" $O/r01s_pytest_gpu.log | head
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'launches',d['gpu_launches'],d['roofline']['kernel'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['roofline']['own_kernels_ms_per_step'],d['roofline']['own_kernel_shares'])" || tail -5 $O/bench.err; }
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r01s_bench_cfg2.json 2>> $O/bench.err; show $O/r01s_bench_cfg2.json cfg2
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline > $O/r01s_bench_cfg3.json 2>> $O/bench.err; show $O/r01s_bench_cfg3.json cfg3
tail -3 $O/bench.err
