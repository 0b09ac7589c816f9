cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_racecheck.log python tools/sanitize_run.py > gpurun_out/r02_racecheck.out 2>&1; echo "racecheck rc $?"; tail -3 gpurun_out/r02_racecheck.log; tail -2 gpurun_out/r02_racecheck.out
