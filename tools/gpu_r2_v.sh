#!/bin/bash
# 2 GPUs: sharded parity tests with the jit forms, then the N=2 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -5
PLB200_BENCH_CHECKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2v_bench_N2.json 2>gpurun_out/r2v_bench_N2.err
tail -c 2000 gpurun_out/r2v_bench_N2.json; grep -v "^\*\|OMP_NUM" gpurun_out/r2v_bench_N2.err | grep -i "error\|Traceback" | head -5
