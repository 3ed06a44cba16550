cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | grep -E "AssertionError|passed|failed|Error|error" | tail -10
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s 2>&1 | grep -E "outliers|passed|failed|Error" | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline'])"
