#!/bin/bash
# round 2, one GPU: the omega-tiled relativistic kernels -- parity, A/B timing, launch list, ncu --set full
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q -x -k "relativistic or batch_classes" --durations=5 ) > gpurun_out/r02f_tests.log 2>&1; tail -15 gpurun_out/r02f_tests.log
for v in 0 1; do echo "== REL_TILED=$v"; ALPS_B200_REL_TILED=$v timeout 300 python scripts/rel_time.py 2>&1 | tail -3; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02f_launches_rel.csv \
  python scripts/prof_rel.py 2048 > gpurun_out/r02f_ncu_rel_list.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02f_launches_rel.csv')))
hi = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn, mv = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r) > mv: agg.setdefault(r[kn].split('(')[0][:50], []).append(float(r[mv].replace(',','')))
for k,v in agg.items(): print("%-50s n=%3d total=%12.1f us  last=%10.1f us" % (k, len(v), sum(v)/1e3, v[-1]/1e3))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rel_tiled -s 1 -c 1 -f -o gpurun_out/r02f_k_rel_tiled \
  python scripts/prof_rel.py 2048 > gpurun_out/r02f_ncu_a.log 2>&1; tail -1 gpurun_out/r02f_ncu_a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rel<" -s 1 -c 1 -f -o gpurun_out/r02f_k_rel_pv \
  python scripts/prof_rel.py 2048 > gpurun_out/r02f_ncu_b.log 2>&1; tail -1 gpurun_out/r02f_ncu_b.log
