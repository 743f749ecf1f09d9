#!/bin/bash
# Developer aid (run under gpurun, 1 GPU): ncu captures of the four kernels of the single-omega chain (C1).
mkdir -p gpurun_out
ALPS_B200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"k_plan|k_quad_mma|k_resonant_lat|k_chi_assemble" -s 32 -c 4 -f -o gpurun_out/prof_lat \
  python scripts/prof_lat.py > gpurun_out/ncu_lat.log 2>&1
tail -3 gpurun_out/ncu_lat.log
echo "--- latency defaults"; timeout 300 python scripts/latency_probe.py 2>&1 | head -2
