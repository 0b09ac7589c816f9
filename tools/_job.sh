cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 400 --timeout-method thread > gpurun_out/r02b_pytest_gpu.txt 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02b_pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-400
timeout 600 python bench.py 2>gpurun_out/r02b_bench_n1.err > gpurun_out/r02b_bench_n1.json; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02b_bench_n1.json').read())
for k in ['value','ms_per_step','tracking_value','tracking_batch_value','gal_e1c_value','all_constellation_ms','gpu_launches']:
    print(k, d.get(k))
print(d['e2e'], d['e2e_cold']['ms'])
print(d['roofline']['frac'], d['roofline']['launch_ms'], d['cpu_baseline'])
print(d['parity'])
print(d['widened']['all_constellation_acquisition']['per_signal_ms_this_rank'])
PY
