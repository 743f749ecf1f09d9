#!/bin/bash
# final evidence after the packed remainder tiles: tests, bench, launch list, one full ncu capture of each k_quad_mma instantiation
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/t_r01m.log 2>&1; tail -4 gpurun_out/t_r01m.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-80
( time timeout 900 python bench.py ) > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2_err.log; tail -3 gpurun_out/bench_final2_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-fast > gpurun_out/b_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_quad_mma -s 2 -c 2 -f -o gpurun_out/prof_mma_final \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-fast > gpurun_out/ncu_mma_final.log 2>&1; tail -2 gpurun_out/ncu_mma_final.log
cut -c1-120 gpurun_out/bench_final2.json
