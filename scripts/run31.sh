cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/r01p_pytest_gpu.log 2>&1; tail -5 $O/r01p_pytest_gpu.log
timeout 400 python bench.py > $O/r01p_bench_default.json 2> $O/bench.err; tail -c 600 $O/r01p_bench_default.json; tail -3 $O/bench.err
