cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "edge_attention" 2>&1 | grep -E "AssertionError|passed|failed|Error|error" | tail -6
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s 2>&1 | grep -E "outliers|passed|failed|Error" | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline'])"
timeout 300 python scripts/profile_step.py cfg2_train 2>&1 | grep -E "cgat::|Self CUDA time" | cut -c1-75,150-230 | head -16
timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline'])"
