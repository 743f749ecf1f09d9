"""Developer aid: one batch of the C3 configuration (relativistic pair plasma, 500x500 (Gamma, pbar_par) grid)
for ncu captures of k_rel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alps_b200 import tables
from alps_b200.solver import Solver
pl = tables.config_relativistic(rel_backend="device")
sol = Solver(pl); sol.set_k(1e-3, 1e-1)
rng = np.random.default_rng(5)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
oms = 6.2713e-2 * (1.0 + 0.05 * rng.uniform(-1, 1, n)) + 1j * 6.2713e-2 * 0.02 * rng.uniform(-1, 1, n)
for _ in range(2): sol.disp_batch(oms)
sol.close()
