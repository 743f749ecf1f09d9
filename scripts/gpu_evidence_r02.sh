#!/bin/bash
# Round-2 evidence on ONE GPU (run under gpurun): GPU test suite, smoke, both bench arms, every reference configuration
# end to end, launch lists and `ncu --set full` captures of the kernels DESIGN.md discusses, compute-sanitizer.
# Outputs land in gpurun_out/; the summaries copied into profiles/ are listed in profiles/README.md.
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/ev_tests.log 2>&1; tail -10 gpurun_out/ev_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-100
( time timeout 900 python bench.py ) > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench_err.log; tail -3 gpurun_out/ev_bench_err.log; cut -c1-200 gpurun_out/ev_bench.json
( time timeout 900 python bench.py --impl reference ) > gpurun_out/ev_bench_ref.json 2>> gpurun_out/ev_bench_err.log; cut -c1-200 gpurun_out/ev_bench_ref.json
timeout 900 python scripts/full_configs.py --repeat 2 --out gpurun_out/ev_full_configs.jsonl 2>&1 | tail -14 | cut -c1-200
timeout 300 python scripts/lat_chain_probe.py 2>&1 | tail -5
timeout 300 python scripts/rel_time.py 2>&1 | tail -3
# launch lists (share of each kernel in a step) and full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-fast --no-map --no-extra > gpurun_out/ev_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ev_launches_fast.csv \
  python scripts/fast_probe.py > gpurun_out/ev_ncu_fast_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_quad_mma -s 2 -c 2 -f -o gpurun_out/ev_k_quad_mma \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-fast --no-map --no-extra > gpurun_out/ev_ncu_quad.log 2>&1; tail -1 gpurun_out/ev_ncu_quad.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fast_tiled -s 1 -c 1 -f -o gpurun_out/ev_k_fast_tiled \
  python scripts/fast_probe.py > gpurun_out/ev_ncu_fast.log 2>&1; tail -1 gpurun_out/ev_ncu_fast.log
for k in k_rel_pv k_rel_direct k_rel_tiled; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 1 -c 1 -f -o gpurun_out/ev_$k \
  python scripts/prof_rel.py 2048 > gpurun_out/ev_ncu_$k.log 2>&1; tail -1 gpurun_out/ev_ncu_$k.log
done
bash scripts/gpu_sanitizer.sh
