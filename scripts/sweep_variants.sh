#!/bin/bash
# Developer aid: parity + C5 throughput for every k_quad tile variant (run under gpurun).
for v in ${VARIANTS:-0 5 8}; do
  echo "=== variant $v"
  ALPS_B200_QUAD_VARIANT=$v timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
  ALPS_B200_QUAD_VARIANT=$v timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('value %.1f D/s  kernel %.2f ms/step  executed %.2f TF/s  frac %.3f  clocks %s'%(d['value'],r['kernel_ms_per_step'],r['achieved'],r['frac'],d['clocks']))
"
done
