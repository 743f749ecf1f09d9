#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/t_r01j.log 2>&1; tail -4 gpurun_out/t_r01j.log
timeout 900 python scripts/full_configs.py --repeat 2 --out gpurun_out/full_configs_graphbatch.jsonl 2>&1 | cut -c1-250
timeout 300 python scripts/latency_probe.py 2>&1 | head -5
