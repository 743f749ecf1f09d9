#!/bin/bash
# Developer aid (run under gpurun, 1 GPU): GPU tests, latency A/B of the single-omega graph, the complete 512x512 C5 map.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_r01e.log 2>&1; tail -8 gpurun_out/t_r01e.log
for k in "0 0 0" "1 0 0" "1 1 0" "1 1 1"; do set -- $k
  echo "--- latency ZC=$1 FUSE=$2 PDL=$3"; ALPS_B200_ZC=$1 ALPS_B200_FUSE=$2 ALPS_B200_PDL=$3 timeout 300 python scripts/latency_probe.py 2>&1 | head -4
done
( time timeout 900 python scripts/full_map_c5.py --out gpurun_out/full_map_c5_n1.json ) 2>&1 | tail -6
