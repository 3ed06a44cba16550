cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "edge" ) > $O/r02n_pytest_edge.log 2>&1; echo pytest edge rc=$?
tail -30 $O/r02n_pytest_edge.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg4_wide --no-cpu-baseline --no-forward-record > $O/r02n_bench_cfg4.json 2> $O/r02n_bench_cfg4.err; echo bench cfg4 rc=$?
python - <<PY
import json
try:
    d=json.loads(open('$O/r02n_bench_cfg4.json').read().strip().splitlines()[-1])
    print('cfg4', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['own_kernels_ms_per_step'])
    for k in d['roofline']['per_kernel']: print('   ',k['kernel'],k['achieved'],k['frac'],k['share'])
except Exception as e: print('no line', e)
PY
tail -3 $O/r02n_bench_cfg4.err
