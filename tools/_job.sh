cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python tools/acq_bench.py 2>&1 | grep "path\|checksum" | tail -2 | cut -c1-120; }
run GC_X=1
big() { echo "== big $*"; env "$@" timeout 300 python tools/big_bench.py E5B L2C B1C E1C20 B1I 2>&1 | grep "fft" | cut -c1-120; }
big GC_X=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -m gpu --timeout 240 --timeout-method thread -k "acq or golden or one_call or device or graph or multi" > gpurun_out/s3_pytest_acq.txt 2>&1; echo "pytest acq rc $?"; tail -3 gpurun_out/s3_pytest_acq.txt
