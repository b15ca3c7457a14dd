#!/bin/bash
# L2 prefetch distance of the specialised passes
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/fused_prof.py 30 ${PREC:-c128} fuse 3 2>&1 | tail -1 | cut -c1-150; }
{
run PLB200_JIT_PREFETCH=1
run PLB200_JIT_PREFETCH=0
run PLB200_JIT_PREFETCH=2
PREC=c64 run PLB200_JIT_PREFETCH=0
PREC=c64 run PLB200_JIT_PREFETCH=2
} 2>&1 | tee gpurun_out/r2aa_prefetch.log
