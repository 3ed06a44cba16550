cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
n=2
for m in A B; do
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n scripts/nccl_graph_probe.py $m > $O/r02d_probe_$m.log 2>&1
  echo "probe $m rc=$?"; grep -E "PROBE|err|captur|Error|error" $O/r02d_probe_$m.log | tail -8
done
