#!/bin/bash
# Round-2 evidence on N >= 2 GPUs (gpurun --gpus N): multi-GPU tests, the torchrun bench line (strong_map, harmonic_shard;
# add --full-map for the complete direct 512x512 map), the device-group probe and the complete C5 map on the device group.
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > gpurun_out/mg_tests.log 2>&1; tail -4 gpurun_out/mg_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 3 --warmup 3 --no-cpu ${FULLMAP:+--full-map} ) > gpurun_out/mg_bench_n$N.json 2> gpurun_out/mg_bench_err.log
tail -2 gpurun_out/mg_bench_err.log | cut -c1-200; tail -c 2500 gpurun_out/mg_bench_n$N.json
timeout 900 python scripts/group_probe.py --out gpurun_out/mg_group_probe_n$N.json 2>&1 | tail -2 | cut -c1-1500
timeout 600 python scripts/full_map_c5.py --ngpu $N $([ -n "$FULLMAP" ] || echo --skip-direct) --out gpurun_out/mg_full_map_c5_group$N.json 2>&1 | tail -1 | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  scripts/multi_gpu_check.py gpurun_out/mg_c4_partitions_n$N.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | cut -c1-600
