#!/bin/bash
# round 2, two GPUs: device group + library communicator tests, C4 partition timings
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --durations=5 ) > gpurun_out/r02b_tests.log 2>&1; tail -25 gpurun_out/r02b_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  scripts/multi_gpu_check.py gpurun_out/r02b_c4_partitions_n2.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8
