#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x -k "relativistic or batch_classes or knobs" --durations=3 ) > gpurun_out/r02g_tests.log 2>&1; tail -8 gpurun_out/r02g_tests.log
for v in 1 2; do echo "== PV_MINB=$v"; ALPS_B200_REL_PV_MINB=$v timeout 300 python scripts/rel_time.py 2>&1 | tail -3; done
echo "== REL_TILED=0"; ALPS_B200_REL_TILED=0 timeout 300 python scripts/rel_time.py 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02g_launches_rel.csv \
  python scripts/prof_rel.py 2048 > gpurun_out/r02g_ncu_rel_list.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02g_launches_rel.csv')))
hi = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn, mv = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r) > mv: agg.setdefault(r[kn].split('(')[0][:50], []).append(float(r[mv].replace(',','')))
for k,v in agg.items():
    if 'rel' in k: print("%-50s n=%3d last=%10.1f us" % (k, len(v), v[-1]/1e3))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_rel_pv' -s 1 -c 1 -f -o gpurun_out/r02g_k_rel_pv \
  python scripts/prof_rel.py 2048 > gpurun_out/r02g_ncu.log 2>&1; tail -1 gpurun_out/r02g_ncu.log
