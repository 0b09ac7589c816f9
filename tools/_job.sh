set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -s -k "fam5_tracking or b1i_tracking or l2c_tracking or l2c_cl_pilot or b1c_wb or b1c_nb or varb_acq or test_tracking_vs_oracle or full_size_tracking" 2>&1 | grep -v "^$" | tail -80 > gpurun_out/r02_parity1.log
cat gpurun_out/r02_parity1.log
