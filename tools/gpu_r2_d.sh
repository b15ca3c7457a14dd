#!/bin/bash
for v in lib_m12 lib_f12 lib_f14; do
  export PLB200_LIB_PATH=$PWD/pennylane-lightning_b200/$v/libplb200.so
  echo "== $v"
  PLB200_JIT=sync PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c128 fuse 3 2>&1 | tail -1 | cut -c1-200
  PLB200_JIT=sync PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -1 | cut -c1-200
done
