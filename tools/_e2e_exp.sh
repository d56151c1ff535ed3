python -m pytest tests -m gpu -x -q -k "ingest or packed or sweep or random_vs" 2>&1 | tail -2
python tools/gpu_ingest.py > gpurun_out/ingest_r1j.log 2>&1; cat gpurun_out/ingest_r1j.log
ncu --set full --clock-control none --import-source on -k regex:ingest -c 12 -o gpurun_out/prof_r1j_ingest -f python tools/gpu_ingest.py > gpurun_out/ncu_r1j_ingest.log 2>&1
