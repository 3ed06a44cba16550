cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/edge_time.py 2>&1 | tail -11 | tee gpurun_out/r03u_edge_time.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "edge" 2>&1 | tail -3
