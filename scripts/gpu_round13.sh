#!/bin/bash
# final evidence of the round: both bench arms as the driver runs them, the launch list of the same command
mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_err.log; tail -3 gpurun_out/bench_final_err.log
( time timeout 900 python bench.py ) > gpurun_out/bench_final.json 2>> gpurun_out/bench_final_err.log; tail -3 gpurun_out/bench_final_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-fast > gpurun_out/b_ncu.log 2>&1
cut -c1-200 gpurun_out/bench_final.json; cut -c1-200 gpurun_out/bench_final_reference.json
