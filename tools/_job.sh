cd $GRAFT_REPO_ROOT
timeout 60 python tools/acq_bench.py 2>&1 | grep "path\|checksum\|acquired" | tail -3 | cut -c1-200
