#!/bin/bash
# round 2, one GPU: full GPU suite, bench (both arms), ncu --set full of k_fast_tiled (hoisted map mode)
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r02c_tests.log 2>&1; tail -15 gpurun_out/r02c_tests.log
( time timeout 900 python bench.py ) > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench_err.log; tail -5 gpurun_out/r02c_bench_err.log; cut -c1-300 gpurun_out/r02c_bench.json
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/r02c_bench_ref.json 2>> gpurun_out/r02c_bench_err.log; cut -c1-200 gpurun_out/r02c_bench_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fast_tiled -s 1 -c 1 -f -o gpurun_out/r02c_k_fast_tiled \
  python scripts/fast_probe.py > gpurun_out/r02c_ncu_fast.log 2>&1; tail -2 gpurun_out/r02c_ncu_fast.log
