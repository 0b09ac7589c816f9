cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu --timeout 150 --timeout-method thread > gpurun_out/s2_pytest_multi8.txt 2>&1; echo "pytest multi rc $?"; tail -3 gpurun_out/s2_pytest_multi8.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/s2_bench_n8.err > gpurun_out/s2_bench_n8.json; echo "bench rc $?"
tail -3 gpurun_out/s2_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_n8.json').read())
for k in ['value','ms_per_step','n_gpus','tracking_value','tracking_batch_value','gal_e1c_value','all_constellation_ms','replica_value','e2e_multi_abi']:
    print(k, d.get(k))
print(d['e2e'])
a=d['widened']['all_constellation_acquisition']
print(a['device_ms_per_rank'], a['predicted_ms_per_rank'], a['ms'], a['device_ms_slowest_rank'])
PY
