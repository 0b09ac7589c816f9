cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GC_TRACK_DEBUG=1 timeout 200 python tools/prof_track.py 12 3000 2 2>&1 | grep -v "rank [1-6]" | tail -7
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 400 --timeout-method thread -k "test_tracking_vs_oracle or cluster_variants or glonass_tracking or e1c_tracking" > gpurun_out/s2_pytest_trk.txt 2>&1; echo "pytest trk rc $?"; tail -3 gpurun_out/s2_pytest_trk.txt
