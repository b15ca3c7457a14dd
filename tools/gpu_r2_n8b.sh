#!/bin/bash
mkdir -p gpurun_out
export PLB200_BENCH_CONFIG4=0 PLB200_BENCH_CHECKS=0
for bits in 1 3; do
  echo "== PLB200_SWAP_MAX_BITS=$bits"
  PLB200_SWAP_MAX_BITS=$bits timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$bits bench.py --gpus 8 --steps 4 --warmup 3 2>gpurun_out/bench8_$bits.err | tail -1 > gpurun_out/bench8_bits$bits.json
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/bench8_bits$bits.json").read())
    print({k: l[k] for k in ("value", "ms_per_step", "gpu_launches")}, "swaps/step", l["config"]["index_bit_swaps_per_step"], "bytes/swap", l["config"]["nvlink_bytes_per_swap_per_gpu"], "e2e", l["e2e"]["value"])
except Exception as e:
    print("parse error", e)
PY
  grep -v "^\*\|OMP_NUM" gpurun_out/bench8_$bits.err | grep -i "error\|Traceback" | head -3
done
