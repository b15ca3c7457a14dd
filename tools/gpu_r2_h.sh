#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_dist_gpu.py" -m gpu -x -q -k "2-peer and not multi and not nofuse" 2>&1 | tail -60 > gpurun_out/dist_test.log
tail -40 gpurun_out/dist_test.log
export PLB200_BENCH_CONFIG4=0 PLB200_BENCH_CHECKS=0
PLB200_SWAP_FUSED=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2_fused.out 2>gpurun_out/bench2_fused.err
tail -c 1500 gpurun_out/bench2_fused.out; grep -v "^\*\|OMP_NUM" gpurun_out/bench2_fused.err | tail -30
