#!/bin/bash
# full GPU suite, then the bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 1500 python bench.py > gpurun_out/r2w_bench_N1.json 2> gpurun_out/r2w_bench_N1.err
tail -c 1200 gpurun_out/r2w_bench_N1.json; tail -3 gpurun_out/r2w_bench_N1.err
