cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GC_TRACK_DEBUG=1 timeout 200 python tools/prof_track.py 12 3000 1 2>&1 | grep -v "rank [1-6]" | tail -8
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 400 --timeout-method thread -k "track or golden or wrappers or full_size or one_call" > gpurun_out/s2_pytest_trk.txt 2>&1; echo "pytest trk rc $?"; tail -4 gpurun_out/s2_pytest_trk.txt
