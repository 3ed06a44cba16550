cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "wide_f256 or edge_hypernet" ) > $O/r02r_pytest_wide.log 2>&1; echo pytest rc=$?
grep -E "strict|passed|failed|FAILED|Error" $O/r02r_pytest_wide.log | tail -8 | cut -c1-400
timeout 900 python bench.py --steps 5 --warmup 3 --workload cfg4_wide --no-cpu-baseline --no-forward-record > $O/r02r_bench_cfg4.json 2> $O/r02r_bench_cfg4.err; echo bench cfg4 rc=$?
python - <<PY
import json
try:
    d=json.loads(open('$O/r02r_bench_cfg4.json').read().strip().splitlines()[-1])
    print('cfg4', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['own_kernels_ms_per_step'])
    for k in d['roofline']['per_kernel']: print('   ',k['kernel'],k['achieved'],k['frac'],k['share'])
except Exception as e: print('no line', e)
PY
timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg5_large --no-cpu-baseline --no-forward-record > $O/r02r_bench_cfg5.json 2> $O/r02r_bench_cfg5.err; echo bench cfg5 rc=$?
python - <<PY
import json
try:
    d=json.loads(open('$O/r02r_bench_cfg5.json').read().strip().splitlines()[-1])
    print('cfg5', d['value'], d['ms_per_step'], d['e2e']['value'])
except Exception as e: print('no line', e)
PY
