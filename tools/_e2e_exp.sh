python -m pytest tests -m gpu -x -q -k "ingest or packed" 2>&1 | tail -3
python tools/gpu_misc.py 2>&1 | tail -7
