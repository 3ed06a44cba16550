cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CGAT_B200_LIB=trap timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "hyper" 2>&1 | tail -3
timeout 600 python scripts/profile_graph_step.py cfg2_train 4 2>/dev/null | grep "cfg2_train\|kernels:\|hyper_wgrad\|sum_parts"
