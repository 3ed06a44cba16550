cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
R=r01n
timeout 400 python bench.py --steps 10 --warmup 3 > $O/${R}_bench_cfg2_train.json 2> $O/bench.err; tail -c 900 $O/${R}_bench_cfg2_train.json
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline > $O/${R}_bench_cfg3_infer.json 2>> $O/bench.err; tail -c 300 $O/${R}_bench_cfg3_infer.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference_arm.json 2>> $O/bench.err; tail -c 300 $O/${R}_bench_reference_arm.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file /tmp/launches.csv python scripts/one_step.py cfg2_train 2 > $O/ncu_l.log 2>&1
python scripts/summarize_launches.py /tmp/launches.csv > $O/${R}_cfg2_train_launches_summary.txt 2>&1
python scripts/summarize_launches.py /tmp/launches.csv --slim $O/${R}_cfg2_train_launches.csv 2>&1 | tail -2; gzip -f $O/${R}_cfg2_train_launches.csv
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:edge_attn|edge_dgrad|edge_wgrad|edge_reduce|hyper_rowdot|hyper_wgrad|hyper_trunk|gemm3x_nt_res' -c 26 -o /tmp/${R}_full python scripts/ncu_layer.py cfg2_train 1 > $O/ncu_f.log 2>&1
ncu -i /tmp/${R}_full.ncu-rep --page raw --csv > $O/${R}_cfg2_train_full_raw.csv 2>/dev/null
python scripts/ncu_metrics.py $O/${R}_cfg2_train_full_raw.csv $O/${R}_cfg2_train_kernel_metrics.json
gzip -f $O/${R}_cfg2_train_full_raw.csv
timeout 200 python scripts/profile_step.py cfg2_train > $O/${R}_cfg2_train_torch_profiler.txt 2>&1
du -sh $O
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:edge_wgrad' -c 1 -o /tmp/${R}_wgrad python scripts/ncu_layer.py cfg2_train 1 > $O/ncu_w.log 2>&1
ncu -i /tmp/${R}_wgrad.ncu-rep --page source --csv --print-source sass > /tmp/wgrad_src.csv 2>/dev/null
python scripts/sass_hot.py /tmp/wgrad_src.csv 0 45 > $O/${R}_edge_wgrad_sass_hot.txt 2>&1; head -2 $O/${R}_edge_wgrad_sass_hot.txt | cut -c1-120
du -sh $O
