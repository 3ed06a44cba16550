cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -s ) > $O/r02b_pytest_gpu.log 2>&1; echo pytest rc=$? | tee -a $O/r02b_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02b_bench_cfg2.json 2> $O/r02b_bench.err; echo bench rc=$?
CGAT_B200_LINEAR3X=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-forward-record > $O/r02b_bench_cfg2_linear3x.json 2> $O/r02b_bench_l3x.err; echo bench l3x rc=$?
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r02b_bench_reference.json 2> $O/r02b_ref.err; echo ref rc=$?
tail -c 1500 $O/r02b_bench_cfg2.json; echo; tail -c 300 $O/r02b_bench_cfg2_linear3x.json; echo; cat $O/r02b_bench_reference.json | cut -c1-400; tail -3 $O/r02b_ref.err
grep -E "passed|failed|error" $O/r02b_pytest_gpu.log | tail -3
