cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
export CGAT_B200_F16X3=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "hyper_linear" > $O/r01c_pytest_hyper.log 2>&1; rc=$?
tail -5 $O/r01c_pytest_hyper.log
if [ $rc -ne 0 ]; then echo "F16X3 FAILED -> falling back to tf32 for the rest"; grep -E "Error|error|assert|mismatch" $O/r01c_pytest_hyper.log | head -20; export CGAT_B200_F16X3=0; fi
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/r01c_pytest_gpu.log 2>&1; tail -6 $O/r01c_pytest_gpu.log
for f in $CGAT_B200_F16X3; do
  CGAT_B200_F16X3=$f timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/r01c_bench_cfg2_f16_$f.json 2>> $O/bench.err
  python -c "import json;d=json.load(open('$O/r01c_bench_cfg2_f16_$f.json'));print('cfg2 f16=$f',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel'],d['roofline']['frac'],d['roofline']['own_kernels_ms_per_step'],d['roofline']['own_kernel_shares'])"
  CGAT_B200_F16X3=$f timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline > $O/r01c_bench_cfg3_f16_$f.json 2>> $O/bench.err
  python -c "import json;d=json.load(open('$O/r01c_bench_cfg3_f16_$f.json'));print('cfg3 f16=$f',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel'],d['roofline']['frac'],d['roofline']['own_kernels_ms_per_step'],d['roofline']['own_kernel_shares'])"
done
timeout 300 python scripts/profile_step.py cfg2_train > $O/r01c_cfg2_train_torch_profiler.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file /tmp/launches.csv python scripts/one_step.py cfg2_train 2 > $O/ncu_l.log 2>&1
python scripts/summarize_launches.py /tmp/launches.csv > $O/r01c_cfg2_train_launches_summary.txt 2>&1
python scripts/summarize_launches.py /tmp/launches.csv --slim $O/r01c_cfg2_train_launches.csv 2>&1 | tail -2; gzip -f $O/r01c_cfg2_train_launches.csv
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:edge_attn|edge_dgrad|edge_wgrad|edge_reduce|hyper_rowdot|hyper_wgrad|hyper_trunk|gemm3x_nt_res' -c 26 -o /tmp/r01c_full python scripts/ncu_layer.py cfg2_train 1 > $O/ncu_f.log 2>&1
ncu -i /tmp/r01c_full.ncu-rep --page raw --csv > $O/r01c_cfg2_train_full_raw.csv 2>/dev/null
python scripts/ncu_metrics.py $O/r01c_cfg2_train_full_raw.csv $O/r01c_cfg2_train_kernel_metrics.json
gzip -f $O/r01c_cfg2_train_full_raw.csv
ls -la $O; du -sh $O
