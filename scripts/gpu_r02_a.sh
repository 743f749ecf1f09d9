#!/bin/bash
# round 2, first GPU session: full GPU test suite (new operating-point tests), bench, A/B of the hoisted-path kernels,
# launch list + ncu --set full of k_fast (old) / k_fast_tiled (new)
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/r02a_tests.log 2>&1; tail -15 gpurun_out/r02a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-100
( time timeout 900 python bench.py ) > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench_err.log; tail -3 gpurun_out/r02a_bench_err.log; cut -c1-400 gpurun_out/r02a_bench.json
for v in 0 1; do echo "== FAST_VARIANT=$v"; ALPS_B200_FAST_VARIANT=$v timeout 300 python scripts/fast_probe.py 2>&1 | tail -8; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02a_launches_fast.csv \
  python scripts/fast_probe.py > gpurun_out/r02a_ncu_fast_list.log 2>&1
for v in 0 1; do
ALPS_B200_FAST_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fast -s 1 -c 1 -f -o gpurun_out/r02a_k_fast_v$v \
  python scripts/fast_probe.py > gpurun_out/r02a_ncu_fast_v$v.log 2>&1; tail -2 gpurun_out/r02a_ncu_fast_v$v.log
done
ls -la gpurun_out | tail -12
