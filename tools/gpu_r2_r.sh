#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -4
PLB200_BENCH_CHECKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_N2_full.json 2>gpurun_out/bench_N2_full.err
tail -c 2600 gpurun_out/bench_N2_full.json; grep -v "^\*\|OMP_NUM" gpurun_out/bench_N2_full.err | grep -i "error\|Traceback" | head -5
