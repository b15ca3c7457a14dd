#!/bin/bash
# round 2, jit forms: c64 variants, ncu captures of the new kernels, full bench line, launch list
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/fused_prof.py 30 ${PREC:-c128} fuse 3 2>&1 | tail -1; }
{
PREC=c64 run PLB200_JIT_FORMS=1
PREC=c64 run PLB200_JIT_FORMS=1 PLB200_JIT_FFMA2=1
PREC=c64 run PLB200_JIT_FORMS=1 PLB200_JIT_FFMA2=1 PLB200_JIT_MINB=1
run PLB200_JIT_FORMS=1
} 2>&1 | tee gpurun_out/r2t_c64_ab.log
PLB200_JIT=sync timeout 900 ncu --set full --clock-control none --import-source on -k regex:plb_pass -s 60 -c 2 -f -o gpurun_out/r2t_jit_pass_30q_c128_forms python tools/fused_prof.py 30 c128 fuse 1 > gpurun_out/r2t_ncu128.log 2>&1
PLB200_JIT=sync timeout 900 ncu --set full --clock-control none --import-source on -k regex:plb_pass -s 51 -c 2 -f -o gpurun_out/r2t_jit_pass_30q_c64_forms python tools/fused_prof.py 30 c64 fuse 1 > gpurun_out/r2t_ncu64.log 2>&1
timeout 1500 python bench.py > gpurun_out/r2t_bench_N1.json 2> gpurun_out/r2t_bench_N1.err
tail -c 1500 gpurun_out/r2t_bench_N1.json
PLB200_BENCH_CHECKS=0 PLB200_BENCH_EXTRA=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r2t_launches_bench.csv | cut -c1-300
