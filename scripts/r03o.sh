cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" w1 w2 w4 w7; do CGAT_B200_LIB=$v timeout 300 python scripts/edge_time.py 2>&1 | grep "wgrad"; done | tee gpurun_out/r04s_ewgrad_ablation.txt
