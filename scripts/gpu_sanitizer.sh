#!/bin/bash
# compute-sanitizer on the tests that exercise this session's kernels (k_plan<true>, k_chi_assemble, zero-copy graph, small-batch
# graphs, k_tps_eval, memo/prefetch drivers)
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_single_omega_graph_knobs_are_bitwise_neutral tests/test_gpu_parity.py::test_quadrature_variants_and_batch_classes_agree tests/test_relativistic_setup.py::test_device_spline_evaluation_matches_the_host_statement tests/test_gpu_golden.py::test_disp_memo_is_transparent tests/test_gpu_parity.py::test_small_bimax_all_branches"
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T -q -x > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py::test_single_omega_graph_knobs_are_bitwise_neutral tests/test_relativistic_setup.py::test_device_spline_evaluation_matches_the_host_statement -q -x > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck.log | tail -3
