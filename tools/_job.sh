cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/sanitize_run.py 2>&1 | tail -8
timeout 1700 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log python tools/sanitize_run.py > gpurun_out/r02_memcheck.out 2>&1; echo "memcheck rc $?"; tail -3 gpurun_out/r02_memcheck.log
