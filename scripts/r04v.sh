cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 10 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "(edge_attention_fused_backward_all_grads and f16x3 and not 700) or (hyper_linear_fused_fwd_bwd and f16x3 and not 5559) or hyper_trunks or seg_softmax" > gpurun_out/r04v_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r04v_memcheck.log; grep "ERROR SUMMARY\|passed\|failed\|Invalid" gpurun_out/r04v_memcheck.log | head -10
