cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -m gpu --timeout 240 --timeout-method thread -k "acq or golden or one_call or device or graph or multi" > gpurun_out/s2_pytest_acq.txt 2>&1; echo "pytest acq rc $?"; tail -3 gpurun_out/s2_pytest_acq.txt
timeout 300 python bench.py --no-tracking --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'cold',d['e2e_cold']['ms'], 'launches', d['gpu_launches'])"
