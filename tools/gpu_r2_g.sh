#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -6
export PLB200_BENCH_CONFIG4=0 PLB200_BENCH_CHECKS=0
for mode in "PLB200_SWAP_FUSED=1" "PLB200_SWAP_FUSED=0 PLB200_SWAP_MULTI=1" "PLB200_SWAP_FUSED=0 PLB200_SWAP_MULTI=0"; do
  echo "== $mode"
  env $mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench2.err | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.read()); print({k: l[k] for k in ('value','ms_per_step','gpu_launches')}, l['config']['index_bit_swaps_per_step'], l['e2e']['value'], l.get('jit'))"
  tail -3 gpurun_out/bench2.err
done
