cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" b12 b16; do CGAT_B200_LIB=$v timeout 300 python scripts/edge_time.py 2>&1 | grep "reduce"; done | tee gpurun_out/r05e_reduce_batch.txt
