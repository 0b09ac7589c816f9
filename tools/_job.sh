cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 > gpurun_out/s3_bench_n8.json 2> gpurun_out/s3_bench_n8.err; echo "bench rc $?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/s3_bench_n8.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'multi_abi',d.get('e2e_multi_abi'))
print('trk',d['tracking_value'],'batch',d['tracking_batch_value'],'e1c',d['gal_e1c_value'],'allc',d['all_constellation_ms'])
"
