#!/bin/bash
mkdir -p gpurun_out
export PLB200_BENCH_CONFIG4=0 PLB200_BENCH_CHECKS=0 PLB200_BENCH_QUBITS=33
for multi in 1 0; do
  echo "== 33 local qubits, 4 GPUs, PLB200_SWAP_MULTI=$multi"
  PLB200_SWAP_MULTI=$multi timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2955$multi bench.py --gpus 4 --steps 1 --warmup 1 2>gpurun_out/bench4_33_$multi.err | tail -1 > gpurun_out/bench4_33_multi$multi.json
  echo "exit $?"
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/bench4_33_multi$multi.json").read())
    print({k: l[k] for k in ("value", "ms_per_step", "gpu_launches")}, "swaps/step", l["config"]["index_bit_swaps_per_step"], "bytes/swap", l["config"]["nvlink_bytes_per_swap_per_gpu"], "e2e ms", l["e2e"]["step_ms"])
except Exception as e:
    print("parse error", e)
PY
  grep -v "^\*\|OMP_NUM" gpurun_out/bench4_33_$multi.err | grep -i "error\|Traceback" | head -3
done
