cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
R=${1:-r04p}
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file /tmp/launches.csv python scripts/one_step.py cfg2_train 2 > $O/ncu_l.log 2>&1
python scripts/summarize_launches.py /tmp/launches.csv > $O/${R}_cfg2_train_launches_summary.txt 2>&1
python scripts/summarize_launches.py /tmp/launches.csv --slim $O/${R}_cfg2_train_launches.csv 2>&1 | tail -2; gzip -f $O/${R}_cfg2_train_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:edge_attn|edge_dgrad|edge_wgrad|edge_reduce|hyper_rowdot|hyper_wgrad|hyper_trunk|gemm3x|sum_parts|adamw' -c 110 -o /tmp/${R}_full python scripts/ncu_layer.py cfg2_train 1 > $O/ncu_f.log 2>&1
ncu -i /tmp/${R}_full.ncu-rep --page raw --csv > $O/${R}_cfg2_train_full_raw.csv 2>/dev/null
python scripts/ncu_metrics.py $O/${R}_cfg2_train_full_raw.csv $O/${R}_cfg2_train_kernel_metrics.json
gzip -f $O/${R}_cfg2_train_full_raw.csv
head -30 $O/${R}_cfg2_train_launches_summary.txt | cut -c1-150
for k in edge_dgrad_zr edge_attn_kernel hyper_rowdot_f16; do
  timeout 200 ncu --set full --clock-control none --import-source on -k "regex:$k" -c 1 -o /tmp/${R}_$k python scripts/ncu_layer.py cfg2_train 1 > $O/ncu_w.log 2>&1
  ncu -i /tmp/${R}_$k.ncu-rep --page source --csv --print-source sass > /tmp/${k}_src.csv 2>/dev/null
  python scripts/sass_hot.py /tmp/${k}_src.csv 0 40 > $O/${R}_${k}_sass_hot.txt 2>&1
done
du -sh $O
