#!/bin/bash
# Developer aid (run under gpurun, 1 GPU): GPU tests, then every reference configuration at full size, timed end to end.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/t_r01g.log 2>&1; tail -6 gpurun_out/t_r01g.log
( time timeout 1200 python scripts/full_configs.py --out gpurun_out/full_configs.jsonl ) 2>&1 | tail -16
