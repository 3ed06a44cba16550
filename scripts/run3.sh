cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -s 2>&1 | grep -E "^gemm3x|AssertionError|passed|failed" | tail -30
