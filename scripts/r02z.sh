cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r02z}
# first on the debug build: a protocol bug traps after 4 s instead of hanging the GPU
( CGAT_B200_LIB=trap timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "hyper_linear" ) > $O/${T}_pytest_hyper_trap.log 2>&1; echo pytest hyper trap rc=$?
tail -5 $O/${T}_pytest_hyper_trap.log | cut -c1-300
if grep -q "passed" $O/${T}_pytest_hyper_trap.log && ! grep -q "failed" $O/${T}_pytest_hyper_trap.log; then
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${T}_bench_cfg2.json 2> $O/${T}_bench.err; echo bench rc=$?
  python - <<PY
import json
for f in ('${T}_bench_cfg2.json',):
    try:
        d=json.loads(open('$O/'+f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['own_kernels_ms_per_step'], (d.get('forward') or {}).get('value'))
        for k in d['roofline']['per_kernel'][:9]: print('   ',k['kernel'],k['achieved'],k['frac'],k['share'])
    except Exception as e: print(f, 'no line', e)
PY
  tail -3 $O/${T}_bench.err
fi
