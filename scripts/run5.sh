cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | grep -E "AssertionError|passed|failed|Error|error" | tail -20
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu 2>&1 | grep -E "AssertionError|passed|failed|Error" | tail -8
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s 2>&1 | grep -E "outliers|passed|failed|Error" | tail -25
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline 2>&1 | tail -2
