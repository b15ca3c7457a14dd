#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_jit.py tests/test_conformance.py -m gpu -x -q 2>&1 | tail -5
for tape in light light2; do
  for nd in 0 1; do
    if [ $nd = 1 ]; then export PLB200_JIT_NO_DIRECT=1; else unset PLB200_JIT_NO_DIRECT; fi
    PLB200_PROF_TAPE=$tape PLB200_JIT=sync PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c128 fuse 5 2>&1 | tail -1
  done
  PLB200_PROF_TAPE=$tape PLB200_JIT=0 timeout 300 python tools/fused_prof.py 30 c128 fuse 5 2>&1 | tail -1
done
unset PLB200_JIT_NO_DIRECT
for nd in 0 1; do
  if [ $nd = 1 ]; then export PLB200_JIT_NO_DIRECT=1; else unset PLB200_JIT_NO_DIRECT; fi
  PLB200_JIT=sync PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c128 fuse 3 2>&1 | tail -1
  PLB200_JIT=sync PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -1
done
unset PLB200_JIT_NO_DIRECT
PLB200_JIT=sync PLB200_JIT_MINB=3 PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c128 fuse 3 2>&1 | tail -1
PLB200_JIT=sync PLB200_JIT_MINB=3 PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -1
PLB200_FUSE_TRACE=1 PLB200_JIT=sync timeout 300 python tools/fused_prof.py 30 c128 fuse 1 2>&1 | grep -E "tile pass" | tail -20
