python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu10.log 2>&1; tail -2 gpurun_out/pytest_gpu10.log
python tools/gpu_ingest.py > gpurun_out/ingest_r1j.log 2>&1; cat gpurun_out/ingest_r1j.log
