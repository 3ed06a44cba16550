cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:edge_attn_kernel|edge_dgrad|edge_wgrad' -c 4 -o /tmp/r01h_edge python scripts/ncu_layer.py cfg2_train 1 > $O/ncu_e.log 2>&1
tail -3 $O/ncu_e.log
ncu -i /tmp/r01h_edge.ncu-rep --page source --csv --print-source sass > /tmp/edge_src.csv 2>/dev/null
ls -la /tmp/edge_src.csv
for i in 0 1 2 3; do python scripts/sass_hot.py /tmp/edge_src.csv $i 45 > $O/r01h_edge_sass_hot_$i.txt 2>&1; head -3 $O/r01h_edge_sass_hot_$i.txt; done
ncu -i /tmp/r01h_edge.ncu-rep --page raw --csv > $O/r01h_edge_raw.csv 2>/dev/null
python scripts/ncu_metrics.py $O/r01h_edge_raw.csv $O/r01h_edge_metrics.json
gzip -f $O/r01h_edge_raw.csv
