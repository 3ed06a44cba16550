cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
CGAT_B200_LIB=trap timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -4
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r05b_bench_cfg2.json 2> $O/r05b_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$O/r05b_bench_cfg2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('forward') or {}).get('value'))
for k in d['roofline']['per_kernel'][:14]: print('   ',k['kernel'],k['achieved'],k['frac'],k['share'])
PY
