#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/pack_probe.py 2>&1 | tail -7
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/t_r01l.log 2>&1; tail -4 gpurun_out/t_r01l.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_packed.json 2> gpurun_out/bench_packed_err.log; tail -2 gpurun_out/bench_packed_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_packed.json').read().strip().splitlines()[-1])
print(j["value"], j["e2e"]["value"], j["roofline"]["frac"], j["roofline"]["kernel_ms_per_step"], j["gpu_launches"], j["fast_path"]["value_per_gpu"])
PY
