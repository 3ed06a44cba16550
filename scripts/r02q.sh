cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q ) > $O/r02q_pytest_gpu.log 2>&1; echo pytest rc=$?
grep -E "passed|failed|FAILED|Error" $O/r02q_pytest_gpu.log | tail -8 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-forward-record > $O/r02q_bench_cfg2.json 2> $O/r02q_bench.err; echo bench rc=$?
python - <<PY
import json
for f in ('r02q_bench_cfg2.json',):
    try:
        d=json.loads(open('$O/'+f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['own_kernels_ms_per_step'])
    except Exception as e: print(f, 'no line', e)
PY
