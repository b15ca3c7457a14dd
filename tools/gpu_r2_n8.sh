#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,memory.total --format=csv | tail -8 | tr '\n' ';'; echo
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -8
PLB200_BENCH_CONFIG4=0 PLB200_BENCH_CHECKS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_N4.json 2>gpurun_out/bench_N4.err
tail -c 1800 gpurun_out/bench_N4.json; grep -v "^\*\|OMP_NUM" gpurun_out/bench_N4.err | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_N8.json 2>gpurun_out/bench_N8.err
tail -c 3500 gpurun_out/bench_N8.json; grep -v "^\*\|OMP_NUM" gpurun_out/bench_N8.err | tail -8
