cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python bench.py --steps 3 --warmup 3 --workload cfg4_wide --no-cpu-baseline > $O/r01q_bench_cfg4_wide.json 2> $O/bench.err; tail -c 1200 $O/r01q_bench_cfg4_wide.json; tail -5 $O/bench.err
