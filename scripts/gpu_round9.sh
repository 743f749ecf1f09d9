#!/bin/bash
# Developer aid: GPU tests, full-size configurations (best of 3), latency probe.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/t_r01h.log 2>&1; tail -4 gpurun_out/t_r01h.log
timeout 900 python scripts/full_configs.py --repeat 3 --out gpurun_out/full_configs.jsonl cfg_electron_mode cfg_ICW cfg_kperp cfg_bimax cfg_kpar_fast cfg_relativistic cfg_map cfg_cold_plasma cfg_chebyshev cfg_analytical 2>&1 | cut -c1-150
timeout 300 python scripts/latency_probe.py 2>&1 | head -2
