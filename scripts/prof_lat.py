"""Developer aid: a few single-omega disp() calls of C1 with plain launches, for ncu captures of the latency chain
(ALPS_B200_NO_GRAPH=1 python scripts/prof_lat.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alps_b200 import tables
from alps_b200.solver import Solver
pl = tables.config_kpar_fast(); sol = Solver(pl, emulate_nproc=4); sol.set_k(1e-2, 1e-2)
om = 9.98811e-3 - 2.31322e-7j
for i in range(12): sol.disp(om * (1 + 1e-6 * i))
sol.close()
