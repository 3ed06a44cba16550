cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "edge_attention" 2>&1 | grep -E "AssertionError|passed|failed|Error|error" | tail -6
timeout 300 python scripts/profile_step.py cfg2_train 2>&1 | grep -E "cgat::|Self CUDA time" | cut -c1-75,150-230 | head -16
