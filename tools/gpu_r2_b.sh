#!/bin/bash
mkdir -p gpurun_out
PLB200_JIT=sync timeout 900 ncu --set full --clock-control none --import-source on -k regex:plb_pass -s 60 -c 2 -f -o gpurun_out/r2_jit_pass_30q_c128_mb4 python tools/fused_prof.py 30 c128 fuse 1 > gpurun_out/ncu_jit128.log 2>&1
PLB200_JIT=sync timeout 900 ncu --set full --clock-control none --import-source on -k regex:plb_pass -s 51 -c 2 -f -o gpurun_out/r2_jit_pass_30q_c64_mb2 python tools/fused_prof.py 30 c64 fuse 1 > gpurun_out/ncu_jit64.log 2>&1
tail -n 2 gpurun_out/ncu_jit128.log gpurun_out/ncu_jit64.log
PLB200_FUSE_TRACE=1 PLB200_JIT=sync timeout 300 python tools/fused_prof.py 30 c128 fuse 1 2>&1 | grep -E "tile pass|n=30" | tail -22
