cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -s -k tn 2>&1 | grep -E "^gemm3x|passed|failed|AssertionError|rror" | tail -14
