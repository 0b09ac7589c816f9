cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python tools/acq_bench.py 2>&1 | grep "path\|checksum" | tail -2; }
run GC_DUMMY=1
run GC_ACQ_LEGACY=1
for c in 8 4 2; do for p in 0 1 2; do run GC_ACQ_OVERLAP=1 GC_ACQ_CHUNK_PRNS=$c GC_COLS_PERSIST=$p; done; done
