#!/bin/bash
# Developer aid (gpurun, 1 GPU): latency chain A/B (knobs), parity tests, device-side timeline of the chain (trace build).
mkdir -p gpurun_out
for kn in "ALPS_B200_FORK=0" "ALPS_B200_FORK=1"; do echo "--- $kn"; env $kn timeout 300 python scripts/lat_chain_probe.py 2>&1 | tail -5; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -5
if [ -f alps_b200/libalps_b200_trace.so ]; then
for cfg in c1 c2; do
ALPS_B200_LIB=alps_b200/libalps_b200_trace.so timeout 300 python scripts/lat_trace.py $cfg 2>&1 | tail -22
done
fi
