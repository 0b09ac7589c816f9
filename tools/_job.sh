cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 400 --timeout-method thread > gpurun_out/s2_pytest_gpu.txt 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/s2_pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/s2_bench_n1.err > gpurun_out/s2_bench_n1.json; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_n1.json').read())
for k in ['value','ms_per_step','tracking_value','tracking_batch_value','gal_e1c_value','all_constellation_ms','gpu_launches']:
    print(k, d.get(k))
print(d['e2e'], d['e2e_cold'])
print(d['roofline'])
print(d['widened']['all_constellation_acquisition']['per_signal_ms_this_rank'])
PY
