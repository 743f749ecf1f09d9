#!/bin/bash
# two GPUs: multi-GPU tests after the unsharded-small-call rule, sanitizer on the device-group kernels, C4 probe, comm check
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x --durations=4 ) > gpurun_out/r02k_tests.log 2>&1; tail -8 gpurun_out/r02k_tests.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_multi.py::test_device_group_harmonic_partition" tests/test_gpu_multi.py::test_device_group_omega_partition_is_bitwise_the_single_gpu_result -q -x > gpurun_out/sanitizer_memcheck_multi.log 2>&1
echo "memcheck (device group) rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_multi.log | tail -3
timeout 600 python scripts/group_probe.py --skip-c5 --out gpurun_out/r02k_group_probe_n2.json 2>&1 | tail -2 | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  scripts/multi_gpu_check.py gpurun_out/r02k_c4_partitions_n2.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -4 | cut -c1-900
