cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "edge_attention and f16x3" > $O/r01o_pytest_edge.log 2>&1; tail -2 $O/r01o_pytest_edge.log
grep -E "Error|error|assert |mismatch" $O/r01o_pytest_edge.log | head -5
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['roofline']['kernel'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['roofline']['own_kernels_ms_per_step'],d['roofline']['own_kernel_shares'])" || tail -5 $O/bench.err; }
for pf in 1 0; do
CGAT_B200_EDGE_PREFETCH=$pf timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg3_infer --no-cpu-baseline > $O/r01o_bench_cfg3_pf$pf.json 2>> $O/bench.err; show $O/r01o_bench_cfg3_pf$pf.json cfg3-pf$pf
done
CGAT_B200_EDGE_PREFETCH=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r01o_bench_cfg2_pf1.json 2>> $O/bench.err; show $O/r01o_bench_cfg2_pf1.json cfg2-pf1
