#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_final3_reference.json 2> gpurun_out/bench_final3_err.log; tail -3 gpurun_out/bench_final3_err.log
( time timeout 900 python bench.py ) > gpurun_out/bench_final3.json 2>> gpurun_out/bench_final3_err.log; tail -3 gpurun_out/bench_final3_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_final3.json').read().strip().splitlines()[-1])
print(j["value"], j["e2e"]["value"], j["roofline"]["frac"], json.dumps(j["cpu_baseline"])[:260], json.dumps(j["clocks"]))
r=json.loads(open('gpurun_out/bench_final3_reference.json').read().strip().splitlines()[-1])
print(r["value"], r["ms_per_step"], r["cpu_baseline"]["sample"][:120])
PY
