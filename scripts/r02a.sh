cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r02a_smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q -s ) > $O/r02a_pytest_gpu.log 2>&1; echo pytest rc=$? | tee -a $O/r02a_pytest_gpu.log
CGAT_B200_LINEAR3X=1 timeout 300 python -m pytest tests/test_gpu_gemm.py -q -k linear3x > $O/r02a_pytest_linear3x.log 2>&1; echo linear3x rc=$?
for c in default_k12 default_k24; do
  timeout 300 python scripts/grad_diag.py $c > $O/r02a_graddiag_${c}_fused.txt 2>&1
  CGAT_B200_FUSED=0 timeout 300 python scripts/grad_diag.py $c > $O/r02a_graddiag_${c}_lib.txt 2>&1
  CGAT_B200_F16X3=0 CGAT_B200_F16X3_EDGE=0 timeout 300 python scripts/grad_diag.py $c > $O/r02a_graddiag_${c}_tf32.txt 2>&1
done
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02a_bench_cfg2.json 2> $O/r02a_bench.err; echo bench rc=$?
tail -c 600 $O/r02a_bench_cfg2.json
tail -3 $O/r02a_pytest_gpu.log
