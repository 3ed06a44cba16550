cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "hyper_linear" > $O/r01d_pytest_hyper.log 2>&1; tail -3 $O/r01d_pytest_hyper.log
grep -E "Error|error|assert|mismatch" $O/r01d_pytest_hyper.log | head -10
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "graphed or padded" > $O/r01d_pytest_graph.log 2>&1; tail -5 $O/r01d_pytest_graph.log
grep -E "Error|error|assert|mismatch" $O/r01d_pytest_graph.log | head -30
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'launches',d['gpu_launches'],d['roofline']['kernel'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['roofline']['own_kernels_ms_per_step'],d['roofline']['own_kernel_shares'])" || tail -5 $O/bench.err; }
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r01d_bench_cfg2_graph.json 2>> $O/bench.err; show $O/r01d_bench_cfg2_graph.json cfg2-graph
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > $O/r01d_bench_cfg2_nograph.json 2>> $O/bench.err; show $O/r01d_bench_cfg2_nograph.json cfg2-nograph
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline > $O/r01d_bench_cfg3_graph.json 2>> $O/bench.err; show $O/r01d_bench_cfg3_graph.json cfg3-graph
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline --no-graph > $O/r01d_bench_cfg3_nograph.json 2>> $O/bench.err; show $O/r01d_bench_cfg3_nograph.json cfg3-nograph
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:hyper_rowdot_f16' -c 4 -o /tmp/r01d_hyper python scripts/ncu_layer.py cfg2_train 1 > $O/ncu_h.log 2>&1
ncu -i /tmp/r01d_hyper.ncu-rep --page raw --csv > $O/r01d_hyper_raw.csv 2>/dev/null
python scripts/ncu_metrics.py $O/r01d_hyper_raw.csv $O/r01d_hyper_metrics.json
tail -20 $O/bench.err
