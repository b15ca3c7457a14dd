#!/bin/bash
export PLB200_LIB_PATH=$PWD/pennylane-lightning_b200/lib_r3/libplb200.so
for mb in 2 3 4; do
  PLB200_JIT_MINB=$mb PLB200_JIT=sync PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c128 fuse 3 2>&1 | tail -1 | cut -c1-200
done
for mb in 1 2; do
  PLB200_JIT_NO_FFMA2=1 PLB200_JIT_MINB=$mb PLB200_JIT=sync PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -1 | cut -c1-200
done
PLB200_JIT_MINB=3 PLB200_FUSE_TRACE=1 PLB200_JIT=sync timeout 300 python tools/fused_prof.py 30 c128 fuse 1 2>&1 | grep -E "tile pass" | tail -20 | sed 's/tile bits.*:/:/'
