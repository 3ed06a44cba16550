cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CGAT_B200_LIB=trap timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "edge" 2>&1 | tail -3
for v in "" e1 e4 e8 e15; do CGAT_B200_LIB=$v timeout 300 python scripts/edge_time.py 2>&1 | grep "dgrad"; done | tee gpurun_out/r03r_dgrad_ablation.txt
