cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/profile_graph_step.py cfg2_train 4 > gpurun_out/r03m_graph_step_kernels.txt 2>gpurun_out/r03m.err; tail -3 gpurun_out/r03m.err; cat gpurun_out/r03m_graph_step_kernels.txt
