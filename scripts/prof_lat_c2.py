"""Developer aid: a few single-omega disp() calls of C2 (use_bM protons) with plain launches, for ncu launch lists
(ALPS_B200_NO_GRAPH=1 python scripts/prof_lat_c2.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alps_b200 import tables
from alps_b200.solver import Solver
pl = tables.config_bimax(); sol = Solver(pl, emulate_nproc=4); sol.set_k(1e-3, 1e-3)
om = 1.0e-3 - 1e-6j
for i in range(12): sol.disp(om * (1 + 1e-6 * i))
sol.close()
