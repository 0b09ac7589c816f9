cd $GRAFT_REPO_ROOT
GC_TRACK_DEBUG=1 python tools/prof_track.py 12 3000 1 2>&1 | grep -v "rank [1-6]"
python tools/acq_bench.py 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-tracking --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'cold',d['e2e_cold']['ms'])"
