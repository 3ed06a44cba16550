cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CGAT_B200_LIB=trap timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "hyper" 2>&1 | tail -4
for n in 5632; do CGAT_B200_LIB=dbg64 timeout 120 python scripts/hyper_timeline.py $n 2>&1 | tail -32; done | tee gpurun_out/r03k_hyper_timeline.txt
for n in 2944 5632 11776 23552; do timeout 120 python scripts/hyper_time.py $n 2>&1 | tail -2; done | tee gpurun_out/r03k_hyper_time.txt
