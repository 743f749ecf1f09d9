#!/bin/bash
# Developer aid (run under gpurun, 1 GPU): GPU tests + the complete 512x512 C5 map (direct and fast path).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_r01e.log 2>&1; tail -8 gpurun_out/t_r01e.log
( time timeout 900 python scripts/full_map_c5.py --out gpurun_out/full_map_c5_n1.json ) 2>&1 | tail -6
