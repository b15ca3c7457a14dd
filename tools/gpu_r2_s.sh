#!/bin/bash
# jit forms (scaled rotations, in-stream ladders) on the device: parity tests, then A/B timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jit.py tests/test_fusion.py -m gpu -q -x 2>&1 | tail -5
run() { echo "== $*"; env "$@" timeout 300 python tools/fused_prof.py 30 ${PREC:-c128} fuse 3 2>&1 | tail -1; }
{
run PLB200_JIT_FORMS=1
run PLB200_JIT_FORMS=0
run PLB200_JIT_FORMS=1 PLB200_JIT_MINB=3
run PLB200_JIT_FORMS=1 PLB200_JIT_MINB=5
run PLB200_JIT_FORMS=1 PLB200_LIB_PATH=$PWD/pennylane-lightning_b200/lib_m12/libplb200.so
PREC=c64 run PLB200_JIT_FORMS=1
PREC=c64 run PLB200_JIT_FORMS=0
PREC=c64 run PLB200_JIT_FORMS=1 PLB200_JIT_MINB=3
} 2>&1 | tee gpurun_out/r2s_forms_ab.log
PLB200_FUSE_TRACE=1 timeout 300 python tools/fused_prof.py 30 c128 fuse 1 2>&1 | grep "trace" | tail -20 > gpurun_out/r2s_trace_c128.log
tail -20 gpurun_out/r2s_trace_c128.log
