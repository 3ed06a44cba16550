set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -s 2>&1 | grep -E "gemm3x|passed|failed|Error" | tail -30
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s 2>&1 | grep -E "outliers|passed|failed|Error|assert" | tail -25
timeout 300 python scripts/gemm_bench.py 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline 2>&1 | tail -3
