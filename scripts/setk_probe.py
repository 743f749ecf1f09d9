"""Developer aid: wall-clock of alps_b200_set_k along a scan (same nmax / changing nmax) and of the first disp() after it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alps_b200 import tables
from alps_b200.solver import Solver
pl = tables.config_kpar_fast(); sol = Solver(pl, emulate_nproc=4)
om = 9.98811e-3 - 2.31322e-7j
for label, ks in (("k_par scan, nmax constant", [(1e-2, 1e-2 * (1 + 0.01 * i)) for i in range(40)]),
                  ("k_perp scan, nmax grows", [(0.1 * (1 + 0.1 * i), 1e-3) for i in range(40)])):
    ts, td, td2 = [], [], []
    for k in ks:
        t = time.perf_counter(); nm = sol.set_k(*k); ts.append(time.perf_counter() - t)
        t = time.perf_counter(); sol.disp(om); td.append(time.perf_counter() - t)
        t = time.perf_counter(); sol.disp(om * 1.0001); sol.disp(om * 1.0002); sol.disp(om * 1.0003); td2.append((time.perf_counter() - t) / 3)
    print("%-28s set_k median %.0f us (max %.0f), first disp %.0f us, later disp %.0f us, last nmax %s" %
          (label, np.median(ts) * 1e6, np.max(ts) * 1e6, np.median(td) * 1e6, np.median(td2) * 1e6, list(nm)))
sol.close()
