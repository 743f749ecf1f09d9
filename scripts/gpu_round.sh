#!/bin/bash
# Developer aid (run under gpurun): GPU tests, smoke, bench lines, ncu launch list and one full capture
# of the dominant kernel; everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_c5_r01b.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c5_reference_r01b.json 2>> gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-fast --batch 148 > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_quad_mma -s 1 -c 1 -f -o gpurun_out/prof_mma_v15 \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-fast --batch 148 > gpurun_out/ncu_mma_v15.log 2>&1
cat gpurun_out/bench_c5_r01b.json | cut -c1-400
