#!/bin/bash
# Evidence for DESIGN.md 4c (gpurun, 1 GPU): wall-clock per disp() with the chain's knobs, device-side timeline of the chain
# (trace build: make -C alps_b200/csrc trace), compute-sanitizer on the tests that drive the chain.
mkdir -p gpurun_out
{
for kn in "ALPS_B200_PDL=0 ALPS_B200_SPIN=0 ALPS_B200_FORK=0" "ALPS_B200_EARLY=0 ALPS_B200_SPIN=0 ALPS_B200_FORK=0" "ALPS_B200_SPIN=0" "ALPS_B200_SPIN=1"; do
  echo "--- $kn"; env $kn timeout 300 python scripts/lat_chain_probe.py 2>&1 | tail -4
done
if [ -f alps_b200/libalps_b200_trace.so ]; then
for cfg in c1 c2 c4; do ALPS_B200_LIB=alps_b200/libalps_b200_trace.so timeout 300 python scripts/lat_trace.py $cfg 2>&1 | tail -32; done
fi
} > gpurun_out/lat_chain_timeline.txt 2>&1
tail -5 gpurun_out/lat_chain_timeline.txt
T="tests/test_gpu_parity.py::test_single_omega_graph_knobs_are_bitwise_neutral tests/test_gpu_golden.py::test_kpar_fast_scan_eigen_heat_files tests/test_gpu_golden.py::test_root_batching_is_bit_identical_to_the_serial_order"
for tool in memcheck synccheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T -q -x > gpurun_out/sanitizer_lat_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_lat_$tool.log | tail -3
done
