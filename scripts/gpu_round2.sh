#!/bin/bash
# Developer aid (run under gpurun): GPU tests, smoke, A/B of the k_quad_mma block order, one full ncu capture.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t_r01c.log; tail -4 gpurun_out/t_r01c.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu --no-fast > gpurun_out/bench_tilemajor.json 2> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_err.log
ALPS_B200_OMEGA_MAJOR=1 timeout 600 python bench.py --no-cpu --no-fast > gpurun_out/bench_omegamajor.json 2>> gpurun_out/bench_err.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_quad_mma -s 1 -c 1 -f -o gpurun_out/prof_mma_tilemajor \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-fast > gpurun_out/ncu_mma_tm.log 2>&1
for f in gpurun_out/bench_tilemajor.json gpurun_out/bench_omegamajor.json; do python - "$f" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], j["value"], j["e2e"]["value"], j["roofline"]["frac"], j["roofline"]["kernel_ms_per_step"])
PY
done
