#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jit.py tests/test_device_shim.py tests/test_parity_measure.py tests/test_fusion.py -m gpu -q 2>&1 | tail -8
for mb in 2 3; do
  echo "== adjoint MINB $mb"; PLB200_JIT_ADJ_MINB=$mb PLB200_JIT_DISK_CACHE=0 timeout 600 python tools/adjoint_bench.py 24 1000 2>&1 | tail -2 | cut -c1-250
done
PLB200_JIT=0 timeout 600 python tools/adjoint_bench.py 24 1000 2>&1 | tail -2 | cut -c1-250
