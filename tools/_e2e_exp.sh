python -m pytest tests -m gpu -x -q -k "packed" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r1j_n2.log 2>&1; tail -c 800 gpurun_out/bench_r1j_n2.log
