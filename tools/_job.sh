cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:inv_rows|inv_cols_big' -c 4 -o gpurun_out/r02b_b1c -f python tools/big_bench.py B1C > gpurun_out/r02b_ncu_b1c.log 2>&1; echo "b1c capture rc $?"
ls -la gpurun_out/r02b_b1c.ncu-rep
