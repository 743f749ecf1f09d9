#!/bin/bash
# round 2, one GPU: full GPU suite + smoke + bench both arms + every reference configuration end to end + latency probe
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r02h_tests.log 2>&1; tail -12 gpurun_out/r02h_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-100
( time timeout 900 python bench.py ) > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench_err.log; tail -3 gpurun_out/r02h_bench_err.log; cut -c1-200 gpurun_out/r02h_bench.json
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r02h_bench_ref.json 2>> gpurun_out/r02h_bench_err.log; cut -c1-200 gpurun_out/r02h_bench_ref.json
timeout 900 python scripts/full_configs.py --out gpurun_out/r02h_full_configs.jsonl 2>&1 | tail -14 | cut -c1-220
timeout 300 python scripts/latency_probe.py 2>&1 | tail -6
