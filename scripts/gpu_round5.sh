#!/bin/bash
# Developer aid (run under gpurun, 1 GPU): GPU tests, latency probe with the new defaults and with the old chain.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/t_r01f.log 2>&1; tail -12 gpurun_out/t_r01f.log
echo "--- latency defaults"; timeout 300 python scripts/latency_probe.py 2>&1 | head -8
echo "--- latency ZC=0 FUSE=0"; ALPS_B200_ZC=0 ALPS_B200_FUSE=0 timeout 300 python scripts/latency_probe.py 2>&1 | head -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
