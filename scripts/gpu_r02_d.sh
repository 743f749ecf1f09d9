#!/bin/bash
# round 2, two GPUs: multi-GPU tests, torchrun bench (strong_map / harmonic_shard legs), device-group probe
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_golden.py tests/test_gpu_ops.py -m gpu -q -x --durations=5 ) > gpurun_out/r02d_tests.log 2>&1; tail -12 gpurun_out/r02d_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2_err.log; tail -5 gpurun_out/r02d_bench_n2_err.log | cut -c1-300; tail -c 3000 gpurun_out/r02d_bench_n2.json
( time timeout 600 python bench.py --no-cpu --no-extra ) > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1_err.log; tail -3 gpurun_out/r02d_bench_n1_err.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench_n1.json').read().strip().splitlines()[-1])
print('N=1 value', d['value'], 'fast', d['fast_path']['value'], d['fast_path']['roofline']['frac'], d['fast_path']['ms_per_step'], 'strong', d['strong_map']['hoisted_full_map'], d['strong_map']['direct_map'])
PY
timeout 600 python scripts/group_probe.py --out gpurun_out/r02d_group_probe_n2.json 2>&1 | tail -3 | cut -c1-1500
