cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" timeout 120 python tools/acq_bench.py 2>&1 | grep "path\|checksum" | tail -2 | cut -c1-150; }
run GC_DUMMY=1
run GC_ROWS_MINB5=1
