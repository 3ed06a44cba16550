cd $GRAFT_REPO_ROOT
echo "== fused (hyper+gemm with separate correction accumulators)"; timeout 300 python scripts/grad_diag.py 2>&1 | tail -8
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -s 2>&1 | grep -E "^gemm3x|passed|failed" | tail -12
timeout 300 python scripts/gemm_bench.py 2>&1 | tail -5
