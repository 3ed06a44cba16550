cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "edge_attention" 2>&1 | grep -E "AssertionError|passed|failed|Error|error" | tail -10
