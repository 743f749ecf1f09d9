"""Developer aid: element-wise chi / wave differences at the golden root (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alps_b200 import tables
from alps_b200.solver import Solver
from oracle.oracle import Oracle
np.set_printoptions(linewidth=200, precision=3)
pl = tables.config_kpar_fast()
for nproc in (4, 0):
    orc = Oracle(pl, nproc=nproc); sol = Solver(pl, emulate_nproc=nproc)
    orc.set_k(1e-2, 1e-2); sol.set_k(1e-2, 1e-2)
    for om in [9.98811e-3 - 2.31322e-7j, 0.3 + 0.01j]:
        Do, co, lo, wo = orc.disp(om, full=True)
        Dg, cg, lg, wg = sol.disp(om, full=True)
        print("nproc", nproc, "om", om, "D", Do, Dg)
        for s in range(2):
            print(" chi0 oracle s=%d\n" % s, co[s])
            print(" elementwise rel err\n", np.abs(cg[s] - co[s]) / np.maximum(np.abs(co[s]), 1e-300))
        print(" wave oracle\n", wo, "\n rel err\n", np.abs(wg - wo) / np.abs(wo))
    sol.close()
