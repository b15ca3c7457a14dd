#!/bin/bash
# round 2, GPU call A: JIT parity + first timings + ncu + bench with checks/extras
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tail -1
nproc; free -g | head -2
timeout 600 python -m pytest tests/test_jit.py -m gpu -x -q 2>&1 | tail -15
for mode in 0 async; do
  PLB200_JIT=$mode timeout 300 python tools/fused_prof.py 30 c128 fuse 3 2>&1 | tail -2
  PLB200_JIT=$mode timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -2
done
for mb in 4 5; do
  PLB200_JIT_MINB=$mb PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c128 fuse 3 2>&1 | tail -1
done
PLB200_JIT_MINB=2 PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -1
PLB200_JIT_MINB=4 PLB200_JIT_DISK_CACHE=0 timeout 300 python tools/fused_prof.py 30 c64 fuse 3 2>&1 | tail -1
# ncu: specialised kernels (3 launches from the steady state) and the interpreter kernel
PLB200_JIT=sync timeout 900 ncu --set full --clock-control none --import-source on -k regex:plb_pass -s 60 -c 3 -f -o gpurun_out/r2_jit_pass_30q_c128 python tools/fused_prof.py 30 c128 fuse 1 > gpurun_out/ncu_jit.log 2>&1
PLB200_JIT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 60 -c 3 -f -o gpurun_out/r2_tile_kernel_30q_c128 python tools/fused_prof.py 30 c128 fuse 1 > gpurun_out/ncu_interp.log 2>&1
tail -3 gpurun_out/ncu_jit.log gpurun_out/ncu_interp.log
timeout 900 python tools/adjoint_bench.py 24 1000 ref 2>&1 | tail -4
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 6000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
