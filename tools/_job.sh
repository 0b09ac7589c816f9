cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/acq_diag.py > gpurun_out/s2_diag.txt 2>&1; echo "diag rc $?"
tail -60 gpurun_out/s2_diag.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "acq" > gpurun_out/s2_pytest_acq.txt 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/s2_pytest_acq.txt
python bench.py --steps 10 --warmup 3 --no-tracking --no-cpu-baseline 2>gpurun_out/s2_bench.err | tee gpurun_out/s2_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'cold',d['e2e_cold']['ms'], 'launches', d['gpu_launches'], d['roofline']['launch_ms'], d['roofline']['kernel_share_of_step'])"
