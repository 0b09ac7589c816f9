cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
big() { echo "== big $*"; env "$@" timeout 200 python tools/big_bench.py L2C B1C E1C20 E1C18 B1I 2>&1 | grep "fft\|rror" | cut -c1-120; }
big GC_COLS_BIG_PIPE=1
big GC_COLS_BIG_PIPE=0
timeout 500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 200 --timeout-method thread -k "b1c or l2c or e1c or varb or b1i or golden" > gpurun_out/s3_pytest_big.txt 2>&1; echo "pytest big rc $?"; tail -3 gpurun_out/s3_pytest_big.txt
