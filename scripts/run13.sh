cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"edge_dgrad|edge_attn_kernel" -s 12 -c 4 -o gpurun_out/prof_edge python scripts/one_step.py cfg2_train 2 > gpurun_out/ncu_edge.log 2>&1
tail -3 gpurun_out/ncu_edge.log
ls -la gpurun_out
