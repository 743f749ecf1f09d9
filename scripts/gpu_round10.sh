#!/bin/bash
# Developer aid: GPU tests, full-size configurations with the disp() memo, latency probe.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/t_r01i.log 2>&1; tail -4 gpurun_out/t_r01i.log
timeout 900 python scripts/full_configs.py --repeat 2 --out gpurun_out/full_configs_memo.jsonl 2>&1 | cut -c1-220
timeout 300 python scripts/latency_probe.py 2>&1 | head -2
