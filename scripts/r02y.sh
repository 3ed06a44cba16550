cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
R=r02y
for k in hyper_rowdot_f16; do
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 9 -c 1 -o /tmp/${R}_$k python scripts/ncu_layer.py cfg2_train 1 > $O/ncu_w.log 2>&1
  ncu -i /tmp/${R}_$k.ncu-rep --page source --csv --print-source sass > /tmp/${k}_src.csv 2>/dev/null
  python scripts/sass_hot.py /tmp/${k}_src.csv 0 60 > $O/${R}_${k}_sass_hot.txt 2>&1
  python scripts/sass_sync.py /tmp/${k}_src.csv 0 > $O/${R}_${k}_sass_sync.txt 2>&1
  ncu -i /tmp/${R}_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for i,n in enumerate(h):
    if any(t in n for t in ('Kernel Name','tensor','smsp__average_warp','issue_active','l1tex__data_pipe','shared','lts__t_sectors_op_read.sum','gpu__time_duration','warps_issue_stalled')): print(n, r[i])
" > $O/${R}_${k}_raw.txt 2>&1
done
cat $O/${R}_hyper_rowdot_f16_sass_sync.txt | cut -c1-200
