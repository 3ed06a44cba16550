cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
for n in 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $O/r01f_bench_cfg2_n$n.json 2> $O/bench_n$n.err; tail -c 1200 $O/r01f_bench_cfg2_n$n.json; tail -5 $O/bench_n$n.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline --workload cfg3_infer > $O/r01f_bench_cfg3_n$n.json 2> $O/bench3_n$n.err; tail -c 600 $O/r01f_bench_cfg3_n$n.json; tail -5 $O/bench3_n$n.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $n --steps 1 --warmup 0 > $O/r01f_ref_n$n.json 2> $O/ref_n$n.err; tail -c 600 $O/r01f_ref_n$n.json
done
