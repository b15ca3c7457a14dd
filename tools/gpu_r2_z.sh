#!/bin/bash
# 4 GPUs: routed parity (k-bit exchanges, interleaved destinations) + the N=4 bench line with its sharded check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -k "matches_single and 8-peer and not multi and not nofuse" 2>&1 | tail -3
PLB200_BENCH_CONFIG4=0 PLB200_BENCH_CHECKS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2z_bench_N8.json 2>gpurun_out/r2z_bench_N8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench_N8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps')}, d['checks'], d['config']['index_bit_swaps_per_step'])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/r2z_bench_N8.err | grep -i "error\|Traceback" | head -5
