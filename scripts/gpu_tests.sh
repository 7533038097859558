cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
