cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/r01b_pytest_gpu.log 2>&1; tail -3 $O/r01b_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r01b_bench_cfg2_train.json 2> $O/bench_cfg2.err; tail -c 1500 $O/r01b_bench_cfg2_train.json
timeout 400 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline > $O/r01b_bench_cfg3_infer.json 2> $O/bench_cfg3.err; tail -c 600 $O/r01b_bench_cfg3_infer.json
timeout 200 python scripts/cpu_profile.py cfg2_train > $O/r01b_cpu_profile.txt 2>&1; head -3 $O/r01b_cpu_profile.txt
timeout 300 python scripts/profile_step.py cfg2_train > $O/r01b_cfg2_train_torch_profiler.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/r01b_launches.csv python scripts/one_step.py cfg2_train 2 > $O/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:edge_|hyper_|gemm3x' -c 75 -o $O/r01b_full python scripts/one_step.py cfg2_train 1 > $O/ncu_f.log 2>&1
ncu -i $O/r01b_full.ncu-rep --page raw --csv > $O/r01b_full_raw.csv 2>/dev/null
ls -la $O
