#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -k "2-" 2>&1 | tail -5
export PLB200_BENCH_CONFIG4=0 PLB200_BENCH_CHECKS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench2.err | tail -1 > gpurun_out/bench2_nosync.json
python - <<PY
import json
l = json.loads(open("gpurun_out/bench2_nosync.json").read())
print({k: l[k] for k in ("value", "ms_per_step", "gpu_launches")}, l["config"]["index_bit_swaps_per_step"], l["config"].get("nvlink_counters"), l["e2e"]["value"])
PY
grep -v "^\*\|OMP_NUM" gpurun_out/bench2.err | grep -i "error\|Traceback" | head -3
