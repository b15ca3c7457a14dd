#!/bin/bash
# last run of the round: the GPU tests touched since the full-suite run, then the bench line without the 3-minute
# lightning.qubit check (the full line with that check: profiles/r2_bench_N1.json)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_jit.py tests/test_fusion.py tests/test_parity_measure.py -m gpu -q -x 2>&1 | tail -4
PLB200_BENCH_CHECKS=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2final_bench_N1.json 2> gpurun_out/r2final_bench_N1.err
tail -c 2500 gpurun_out/r2final_bench_N1.json; tail -2 gpurun_out/r2final_bench_N1.err
