#!/bin/bash
# Developer aid (run under gpurun): GPU tests, smoke, both bench arms, ncu launch list of the default bench command.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_r01d.log 2>&1; tail -8 gpurun_out/t_r01d.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_c5_reference_r01d.json 2> gpurun_out/bench_err.log
( time timeout 900 python bench.py ) > gpurun_out/bench_c5_r01d.json 2>> gpurun_out/bench_err.log; tail -c 900 gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-fast > gpurun_out/b_ncu.log 2>&1
cut -c1-700 gpurun_out/bench_c5_r01d.json
