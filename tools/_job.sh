cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "acquisition or acquire or golden or multi or sharded or wrappers" 2>&1 | tail -8
python tools/acq_bench.py 2>&1 | tail -12
GC_ACQ_NO_SHIFT=1 python tools/acq_bench.py 2>&1 | tail -4
