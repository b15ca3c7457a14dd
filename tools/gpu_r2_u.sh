#!/bin/bash
# c64 with 2^12-amplitude tiles and 4 register bits (half the code per op): occupancy sweep
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -1; }
L=$PWD/pennylane-lightning_b200/lib_c64r4/libplb200.so
{
run PLB200_LIB_PATH=$L PLB200_JIT_MINB=2
run PLB200_LIB_PATH=$L PLB200_JIT_MINB=3
run PLB200_LIB_PATH=$L PLB200_JIT_MINB=4
} 2>&1 | tee gpurun_out/r2u_c64r4.log
