"""Developer aid: throughput of the C3 configuration (relativistic pair plasma, 500x500 (Gamma, pbar_par) grid):
batched and one disp() at a time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alps_b200 import tables
from alps_b200.solver import Solver
pl = tables.config_relativistic(rel_backend="device")
sol = Solver(pl); sol.set_k(1e-3, 1e-1)
rng = np.random.default_rng(5)
for n in (256, 2048):
    oms = 6.2713e-2 * (1.0 + 0.05 * rng.uniform(-1, 1, n)) + 1j * 6.2713e-2 * 0.02 * rng.uniform(-1, 1, n)
    sol.disp_batch(oms)
    t = time.perf_counter()
    for _ in range(3): D = sol.disp_batch(oms)
    dt = (time.perf_counter() - t) / 3
    print("C3 batch %d: %.0f D/s  (checksum %.15e)" % (n, n / dt, float(np.sum(np.abs(D)))))
for _ in range(20): sol.disp(complex(oms[0]))
t = time.perf_counter()
for i in range(200): sol.disp(complex(oms[i]))
print("C3 single: %.0f D/s" % (200 / (time.perf_counter() - t)))
sol.close()
