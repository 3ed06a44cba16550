cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" sr8 sr6 sr4; do CGAT_B200_LIB=$v timeout 300 python scripts/edge_time.py 2>&1 | grep "reduce\|sum_parts"; done | tee gpurun_out/r05a_reduce_chunks_grouped.txt
