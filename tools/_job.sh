cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/fine_diag.py 2>&1 | grep variant
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 240 --timeout-method thread -k "acq or golden or one_call" > gpurun_out/s2_pytest_acq.txt 2>&1; echo "pytest acq rc $?"; tail -3 gpurun_out/s2_pytest_acq.txt
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu --timeout 150 --timeout-method thread > gpurun_out/s2_pytest_multi.txt 2>&1; echo "pytest multi rc $?"; tail -3 gpurun_out/s2_pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/s2_bench_n2.err > gpurun_out/s2_bench_n2.json; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_n2.json').read())
for k in ['value','ms_per_step','n_gpus','tracking_value','tracking_batch_value','gal_e1c_value','all_constellation_ms','replica_value','e2e_multi_abi']:
    print(k, d.get(k))
print(d['e2e'])
PY
