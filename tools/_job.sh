cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r02_gputests.log
tail -60 gpurun_out/r02_gputests.log
