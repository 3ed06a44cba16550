cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -s ) > $O/r02e_pytest_gpu.log 2>&1; echo pytest rc=$? | tee -a $O/r02e_pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $O/r02e_pytest_gpu.log | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02e_bench_cfg2.json 2> $O/r02e_bench.err; echo bench rc=$?
CGAT_B200_LINEAR3X=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-forward-record > $O/r02e_bench_cfg2_linear3x.json 2> $O/r02e_bench_l3x.err; echo bench l3x rc=$?
python - <<PY
import json
for f in ('r02e_bench_cfg2.json','r02e_bench_cfg2_linear3x.json'):
    try:
        d=json.loads(open('$O/'+f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['own_kernels_ms_per_step'], d.get('forward'))
        for k in d['roofline']['per_kernel']: print('   ',k['kernel'],k['achieved'],k['frac'],k['share'])
    except Exception as e: print(f, 'no line', e)
PY
tail -3 $O/r02e_bench.err
