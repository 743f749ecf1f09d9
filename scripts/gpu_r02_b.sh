#!/bin/bash
# round 2, two GPUs: device group + library communicator tests, C4 / C5 partition timings
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --durations=5 ) > gpurun_out/r02b_tests.log 2>&1; tail -25 gpurun_out/r02b_tests.log
timeout 900 python scripts/group_probe.py --out gpurun_out/r02b_group_probe_n2.json 2>&1 | tail -4
