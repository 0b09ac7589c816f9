cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 400 --timeout-method thread > gpurun_out/r02b_pytest_gpu.txt 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02b_pytest_gpu.txt
timeout 600 python bench.py 2>gpurun_out/r02b_bench_n1.err > gpurun_out/r02b_bench_n1.json; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02b_bench_n1.json').read())
for k in ['value','ms_per_step','tracking_value','tracking_batch_value','gal_e1c_value','all_constellation_ms','gpu_launches']:
    print(k, d.get(k))
print(d['e2e']['ms_per_step'], d['e2e_cold']['ms'], d['roofline']['frac'], d['clocks'])
PY
