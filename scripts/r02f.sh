cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "hyper" ) > $O/r02f_pytest_hyper.log 2>&1; echo pytest hyper rc=$?
tail -15 $O/r02f_pytest_hyper.log | cut -c1-300
( timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "gradients" ) > $O/r02f_pytest_model.log 2>&1; echo pytest model rc=$?
grep -E "strict|passed|failed|FAILED" $O/r02f_pytest_model.log | cut -c1-400 | tail -12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-forward-record > $O/r02f_bench_cfg2.json 2> $O/r02f_bench.err; echo bench rc=$?
python - <<PY
import json
for f in ('r02f_bench_cfg2.json',):
    try:
        d=json.loads(open('$O/'+f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['own_kernels_ms_per_step'])
        for k in d['roofline']['per_kernel']: print('   ',k['kernel'],k['achieved'],k['frac'],k['share'])
    except Exception as e: print(f, 'no line', e)
PY
tail -3 $O/r02f_bench.err
