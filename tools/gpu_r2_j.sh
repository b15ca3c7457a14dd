#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) 2>&1 | tail -20
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_N1.json 2> gpurun_out/bench_N1.err
tail -c 7000 gpurun_out/bench_N1.json; tail -5 gpurun_out/bench_N1.err
PLB200_BENCH_CHECKS=0 PLB200_BENCH_EXTRA=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/r2_launches_bench.csv | cut -c1-300
