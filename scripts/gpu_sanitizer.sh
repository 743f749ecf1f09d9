#!/bin/bash
# compute-sanitizer on the tests that exercise round 2's kernels: k_fast_tiled / k_fast_tables (TMA ring + mbarriers),
# k_rel_plan / k_rel_pv / k_rel_direct / k_rel_tiled, the harmonic-shard buffer sizing, the class_n logic of chunked
# batches; with two GPUs visible also k_reduce_partials (peer memory) and the device-group threads
mkdir -p gpurun_out
T="tests/test_gpu_ops.py::test_mode1_parity_battery tests/test_gpu_ops.py::test_relativistic_throughput_class_against_oracle tests/test_gpu_ops.py::test_c2_bimax_150x300_every_class_against_oracle tests/test_gpu_parity.py::test_batch_chunking_and_api_errors tests/test_gpu_golden.py::test_map_search_finds_the_root_region"
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then T="$T tests/test_gpu_multi.py::test_device_group_harmonic_partition tests/test_gpu_multi.py::test_device_group_omega_partition_is_bitwise_the_single_gpu_result"; fi
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T -q -x -k "not c3" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fullsize.py::test_harmonic_shard_after_an_unsharded_small_batch -q -x > gpurun_out/sanitizer_memcheck_shard.log 2>&1
echo "memcheck (harmonic shard after small batch, C5) rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_shard.log | tail -2
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_ops.py::test_relativistic_throughput_class_against_oracle[parallel]" tests/test_gpu_ops.py::test_mode1_parity_battery -q -x > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck.log | tail -3
