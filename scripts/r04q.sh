cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
CGAT_B200_LIB=trap timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 600 python scripts/profile_graph_step.py cfg2_train 4 > $O/r04q_graph_step_kernels.txt 2>/dev/null; head -14 $O/r04q_graph_step_kernels.txt | cut -c1-120
bash scripts/r04p.sh r04p > $O/r04p.log 2>&1; tail -3 $O/r04p.log
cp $O/r04p_cfg2_train_kernel_metrics.json profiles/ 2>/dev/null
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r04q_bench_cfg2.json 2> $O/r04q_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$O/r04q_bench_cfg2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('forward') or {}).get('value'), d['cpu_baseline'])
print({k:d['roofline'][k] for k in ('kernel','achieved','peak','frac','traffic','traffic_source')})
for k in d['roofline']['per_kernel'][:14]: print('   ',k['kernel'],k['achieved'],k['frac'],k['share'])
PY
for w in cfg4_wide cfg5_large; do
timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-forward-record > $O/r04q_bench_$w.json 2> $O/r04q_$w.err
python - <<PY
import json
d=json.loads(open('$O/r04q_bench_$w.json').read().strip().splitlines()[-1])
print('$w', d['value'], d['ms_per_step'])
PY
done
