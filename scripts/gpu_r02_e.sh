#!/bin/bash
# round 2, eight GPUs: multi-GPU tests at N=8, torchrun bench, device-group probe, the complete C5 map on the device group
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > gpurun_out/r02e_tests.log 2>&1; tail -4 gpurun_out/r02e_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --full-map ) > gpurun_out/r02e_bench_n8.json 2> gpurun_out/r02e_bench_n8_err.log; tail -3 gpurun_out/r02e_bench_n8_err.log | cut -c1-300; tail -c 2600 gpurun_out/r02e_bench_n8.json
timeout 600 python scripts/group_probe.py --out gpurun_out/r02e_group_probe_n8.json 2>&1 | tail -3 | cut -c1-1600
timeout 600 python scripts/full_map_c5.py --ngpu 8 --out gpurun_out/r02e_full_map_c5_group8.json 2>&1 | tail -1 | cut -c1-1200
