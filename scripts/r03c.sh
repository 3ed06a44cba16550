cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
for v in "" cs1; do
  CGAT_B200_LIB=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r03c_bench_cfg2_$v.json 2> $O/r03c_bench_$v.err; echo "bench [$v] rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('$O/r03c_bench_cfg2_$v.json').read().strip().splitlines()[-1])
    print('[$v]', d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('forward') or {}).get('value'))
    for k in d['roofline']['per_kernel'][:9]:
        if 'hyper' in k['kernel']: print('   ',k['kernel'],k['achieved'],k['frac'],k['share'])
except Exception as e: print('no line', e)
PY
done
CGAT_B200_LIB=cs1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "hyper_linear" 2>&1 | tail -2
