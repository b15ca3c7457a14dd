#!/bin/bash
timeout 600 python -m pytest tests/test_jit.py -m gpu -x -q 2>&1 | tail -3
PLB200_JIT=sync PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -1 | cut -c1-200
PLB200_JIT_NO_FFMA2=1 PLB200_JIT=sync PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -1 | cut -c1-200
PLB200_JIT_MINB=1 PLB200_JIT=sync PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -1 | cut -c1-200
PLB200_JIT=sync timeout 900 ncu --set full --clock-control none --import-source on -k regex:plb_pass -s 51 -c 2 -f -o gpurun_out/r2_jit_pass_30q_c64_ffma2 python tools/fused_prof.py 30 c64 fuse 1 > gpurun_out/ncu_jit64b.log 2>&1
PLB200_JIT=sync timeout 900 ncu --set full --clock-control none --import-source on -k regex:plb_pass -s 60 -c 2 -f -o gpurun_out/r2_jit_pass_30q_c128_direct python tools/fused_prof.py 30 c128 fuse 1 > gpurun_out/ncu_jit128b.log 2>&1
