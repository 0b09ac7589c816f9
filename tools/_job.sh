cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python tools/acq_bench.py 2>&1 | grep "path\|checksum" | tail -2 | cut -c1-150; }
run GC_X=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -m gpu --timeout 240 --timeout-method thread -k "acq or golden or one_call or device or graph or multi" > gpurun_out/s3_pytest_acq.txt 2>&1; echo "pytest acq rc $?"; tail -3 gpurun_out/s3_pytest_acq.txt
