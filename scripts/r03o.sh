cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" f1 f2 f4 f8 f15; do CGAT_B200_LIB=$v timeout 300 python scripts/edge_time.py 2>&1 | grep "edge_attn_fwd\|bwd_prep"; done | tee gpurun_out/r03s_efwd_ablation.txt
