lscpu | grep -i "model name\|^CPU(s)\|MHz" ; cat /proc/loadavg
python tools/gpu_pipe.py 2>&1 | grep "^pack\|^H2D\|^total"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1j_ref.log 2>&1
python bench.py --steps 20 > gpurun_out/bench_r1j.log 2>&1
XS_BENCH_TWO_HANDLES=1 python bench.py --steps 20 > gpurun_out/bench_r1j_twohandles.log 2>&1
cat /proc/loadavg
