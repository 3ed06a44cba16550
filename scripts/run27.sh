cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:edge_dgrad|edge_wgrad|edge_reduce' -c 3 -o /tmp/r01l_bwd python scripts/ncu_layer.py cfg2_train 1 > $O/ncu_b.log 2>&1
tail -2 $O/ncu_b.log
ncu -i /tmp/r01l_bwd.ncu-rep --page source --csv --print-source sass > /tmp/bwd_src.csv 2>/dev/null
grep -c "Kernel Name" /tmp/bwd_src.csv
for i in 0 1 2; do python scripts/sass_hot.py /tmp/bwd_src.csv $i 40 > $O/r01l_bwd_sass_hot_$i.txt 2>&1; head -1 $O/r01l_bwd_sass_hot_$i.txt | cut -c1-90; done
ncu -i /tmp/r01l_bwd.ncu-rep --page raw --csv > $O/r01l_bwd_raw.csv 2>/dev/null
python scripts/ncu_metrics.py $O/r01l_bwd_raw.csv $O/r01l_bwd_metrics.json
gzip -f $O/r01l_bwd_raw.csv
