cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 --timeout-method thread > gpurun_out/s3_pytest_gpu.txt 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/s3_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/s3_bench_n1.json 2> gpurun_out/s3_bench_n1.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('gpurun_out/s3_bench_n1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'cold',d['e2e_cold']['ms'],'launches',d['gpu_launches'],'frac',d['roofline']['frac'],'rows ms',d['roofline']['launch_ms'])
print('trk',d['tracking_value'],'batch',d['tracking_batch_value'],'e1c',d['gal_e1c_value'],'allc',d['all_constellation_ms'], 'parity', d['parity']['acquisition'])
"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
