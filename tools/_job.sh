cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sig_power|fwd_cols|fwd_rows|inv_rows|inv_cols|peak_select|fine_|pack_results|finish_replica" -c 400 --csv --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-tracking --no-cpu-baseline > gpurun_out/b.log 2>&1
tail -3 gpurun_out/b.log | cut -c1-300
