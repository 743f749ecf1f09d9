#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu --no-fast --steps 10 > gpurun_out/bench_clocks.json 2> gpurun_out/bench_clocks_err.log; tail -2 gpurun_out/bench_clocks_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_clocks.json').read().strip().splitlines()[-1])
print(j["value"], j["roofline"]["frac"], j["roofline"]["peak"], json.dumps(j["clocks"]))
PY
nvidia-smi --query-gpu=power.limit,power.max_limit,clocks.max.sm --format=csv,noheader
