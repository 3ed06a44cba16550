cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
for w in cfg4_wide cfg5_large cfg3_infer; do
timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-forward-record > $O/r04e_bench_$w.json 2> $O/r04e_$w.err; echo "$w rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('$O/r04e_bench_$w.json').read().strip().splitlines()[-1])
    print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
    for k in d['roofline']['per_kernel'][:6]: print('   ',k['kernel'],k['achieved'],k['frac'],k['share'])
except Exception as e: print('no line', e)
PY
done
