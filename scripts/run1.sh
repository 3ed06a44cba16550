set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -c "import os; print('cores', os.cpu_count())"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -25
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu 2>&1 | tail -30
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
