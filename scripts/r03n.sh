cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CGAT_B200_LIB=trap timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python scripts/profile_graph_step.py cfg2_train 4 > gpurun_out/r03n_graph_step_kernels.txt 2>gpurun_out/r03n.err; tail -3 gpurun_out/r03n.err; head -24 gpurun_out/r03n_graph_step_kernels.txt
