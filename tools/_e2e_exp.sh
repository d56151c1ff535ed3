python bench.py --steps 20 > gpurun_out/bench_r1j.log 2>&1 || tail -20 gpurun_out/bench_r1j.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r1j_n2.log 2>&1 || tail -20 gpurun_out/bench_r1j_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py 2>&1 | tail -3
