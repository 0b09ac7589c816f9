cd $GRAFT_REPO_ROOT
timeout 30 python tools/mini_parity.py 2>&1 | tail -2
