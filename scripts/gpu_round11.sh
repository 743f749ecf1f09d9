#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py -m gpu -q 2>&1 | tail -2
timeout 900 python scripts/full_configs.py --repeat 2 --out gpurun_out/full_configs_prefetch.jsonl 2>&1 | cut -c1-220
