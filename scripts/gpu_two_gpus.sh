#!/bin/bash
# Developer aid (run under gpurun --gpus 2): NCCL checks of both sharding modes incl. the sharded map_search, the twin main
# program under torchrun on a use_map input, and the 2-GPU bench line.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8
timeout 300 python -m alps_b200.run tests/inputs/test_map_small.in --dist tests/inputs/test_kpar_fast_dist.in --out gpurun_out/sol1 --nproc 4 2>&1 | tail -3
timeout 300 $TR -m alps_b200.run tests/inputs/test_map_small.in --dist tests/inputs/test_kpar_fast_dist.in --out gpurun_out/sol2 --emulate-nproc 4 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3
cmp gpurun_out/sol1/test_map_small.map gpurun_out/sol2/test_map_small.map && echo "MAP FILES IDENTICAL"
cmp gpurun_out/sol1/test_map_small.roots gpurun_out/sol2/test_map_small.roots && echo "ROOTS FILES IDENTICAL"
timeout 600 $TR scripts/full_map_c5.py --nr 128 --ni 128 --out gpurun_out/full_map_c5_128_n2.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -2
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c5_n2_r01.json 2> gpurun_out/bench_n2_err.log; cut -c1-300 gpurun_out/bench_c5_n2_r01.json
