cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/edge_wide_diag.py 256 8 40 12 2>&1 | tail -16
python scripts/edge_wide_diag.py 128 5 40 12 2>&1 | tail -16
python scripts/edge_wide_diag.py 256 3 30 12 2>&1 | tail -16
