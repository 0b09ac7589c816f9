cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 70 compute-sanitizer --tool racecheck --log-file gpurun_out/r02c_racecheck.log python tools/sanitize_mini.py 2>&1 | tail -2
tail -2 gpurun_out/r02c_racecheck.log
