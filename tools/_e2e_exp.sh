python -m pytest tests -m gpu -x -q -k "packed or ingest or sweep" 2>&1 | tail -2
XS3D_HOST_THREADS=12 python tools/gpu_stream_trace.py 2>&1 | tail -14
for t in 10 12 14 15; do XS3D_HOST_THREADS=$t XS_BENCH_NO_SMI=1 python bench.py --steps 20 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('threads $t', e['value'], e['ms_per_step'], e['one_sweep_at_a_time']['ms_per_step'])"; done
