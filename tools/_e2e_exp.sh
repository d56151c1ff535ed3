python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu12.log 2>&1; tail -2 gpurun_out/pytest_gpu12.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1k_ref.log 2>&1
python bench.py --steps 20 > gpurun_out/bench_r1k.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/gpu_misc.py > gpurun_out/misc_r1k.log 2>&1
