"""Developer aid: verbose CUDA-vs-oracle comparison (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alps_b200 import tables
from alps_b200.solver import Solver
from oracle.oracle import Oracle
from tests.util import det_scale, tensor_err, omega_samples

def run(name, pl, kperp, kpar, oms, nproc=0):
    print("=== ", name, flush=True)
    orc = Oracle(pl, nproc=nproc); sol = Solver(pl, emulate_nproc=nproc)
    print("nmax oracle", orc.set_k(kperp, kpar), "gpu", sol.set_k(kperp, kpar), flush=True)
    t = time.time(); Db = sol.disp_batch(oms); print("batch time", time.time() - t, flush=True)
    for i, om in enumerate(oms):
        Do, co, lo, wo = orc.disp(complex(om), full=True)
        Dg, cg, lg, wg = sol.disp(complex(om), full=True)
        errs = [tensor_err(cg[s], co[s]) for s in range(pl.nspec)]
        lerr = [max(tensor_err(lg[s, :, :, m], lo[s, :, :, m]) if np.max(np.abs(lo[s, :, :, m])) > 0 else 0.0 for m in range(3)) for s in range(pl.nspec)]
        print("om=%s chi_err=%s low_err=%s wave_err=%.2e D_err=%.2e batch_err=%.2e" % (
            om, ["%.1e" % e for e in errs], ["%.1e" % e for e in lerr], tensor_err(wg, wo),
            abs(Dg - Do) / det_scale(wo), abs(Db[i] - Dg) / det_scale(wo)), flush=True)
        if max(errs) > 1e-9:
            for s in range(pl.nspec):
                print(" species", s, "\n gpu", cg[s], "\n ora", co[s])
    sol.close()

if __name__ == "__main__":
    pl = tables.config_small(24, 48, kind=1)
    oms = list(omega_samples(1, 6, (0.02, 1.5), (-0.05, 0.05))) + [0.3 + 0j, 0.011 - 1e-6j, 1.0 + 1e-5j]
    run("small bimax", pl, 0.3, 0.05, oms)
    pl = tables.config_small(28, 56, kind=2)
    run("small kappa", pl, 0.2, 0.08, list(omega_samples(2, 5, (0.02, 1.2), (-0.03, 0.03))))
    pl = tables.config_kpar_fast()
    run("C1", pl, 1e-2, 1e-2, [9.98811e-3 - 2.31322e-7j, 9.9e-3 - 5.5e-6j, 5e-2 - 3e-4j, 0.3 + 0.01j, 0.9 + 0j], nproc=4)
