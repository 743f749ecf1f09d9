#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -q -x -s -k "c5 or c4 or c2 or shard_after or drifting or map_search" ) > gpurun_out/r02j_tests.log 2>&1; grep -E "worst|passed|failed|Error" gpurun_out/r02j_tests.log | cut -c1-400
bash scripts/gpu_sanitizer.sh
