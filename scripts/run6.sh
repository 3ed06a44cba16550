cd $GRAFT_REPO_ROOT
echo "== fused"; timeout 300 python scripts/grad_diag.py 2>&1 | tail -17
echo "== unfused"; CGAT_B200_FUSED=0 timeout 300 python scripts/grad_diag.py 2>&1 | tail -17
echo "== profile train"; timeout 300 python scripts/profile_step.py cfg2_train 2>&1 | tail -60
