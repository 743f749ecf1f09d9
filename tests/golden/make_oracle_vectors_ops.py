"""Writes tests/golden/oracle_vectors_ops.npz: D, chi0, chi0_low and wave of the CPU oracle (oracle/, the restatement
of the reference's disp()) at the OPERATING POINTS that bench.py measures and DESIGN.md quotes -- the sizes at which the
oracle is too slow to run inside a test:

  c5   BASELINE config 5: 3-species bi-kappa, 1024x2048 grid, nmax = 200 forced, k = (15.5, 1e-2); four omegas of the
       512x512 map: Im < 0 with resonant harmonics (protons n=1, electrons n=0, alphas n=2 at the grid edge),
       Im > 0 (protons n=2, alphas n=4 resonant), Im = 0 exactly (electrons n=0 resonant), and a weakly damped one
       whose |Im p_res| <= Tlim (the linearised near-pole branch, src/ALPS_fns.f90:1088-1165)
  c4   tests/test_kperp.in at its last scan point k_perp = 3, k_par = 1e-3 with mpirun -np 4 emulated
       (nmax 88 / 29, split_processes ranges past nmax, src/ALPS_fns.f90:4048-4064)
  c2   tests/test_bimax.in on its own 150x300 grid (protons use_bM: closed-form NHDS chi; electrons from the table)

About one minute of CPU per c5 omega on 8 cores.      python tests/golden/make_oracle_vectors_ops.py [case ...]"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from alps_b200 import tables          # noqa: E402
from oracle.oracle import Oracle      # noqa: E402

OUT = os.path.join(HERE, "oracle_vectors_ops.npz")
CASES = {
    # name: (plasma builder, emulated nproc, nmax_force, kperp, kpar, omegas)
    "c5": (lambda: tables.config_kappa3(1024, 2048), 0, 200, 15.5, 1.0e-2,
           [1.03 - 0.02j, 2.01 + 0.01j, 0.52 + 0j, 3.0002 - 5.0e-5j]),
    "c4": (tables.config_kpar_fast, 4, 0, 3.0, 1.0e-3,
           [2.5e-3 - 1.0e-4j, 2.5e-3 + 1.0e-4j, 2.5e-3 + 0j, 1.002 - 1.0e-4j, 5.5e-3 - 2.0e-5j, 0.4 + 0.01j]),
    "c2": (lambda: tables.config_bimax(150, 300), 0, 0, 1.0e-3, 3.0e-2,
           [3.0e-2 - 1.0e-5j, 4.5e-2 - 1.9e-2j, 3.0e-2 + 0j, 2.0e-2 + 3.0e-3j]),
}


def compute(names):
    out = {}
    for name in names:
        make, nproc, nmax_force, kperp, kpar, oms = CASES[name]
        pl = make()
        orc = Oracle(pl, nproc=nproc, nmax_force=nmax_force)
        out[name + "_nmax"] = np.asarray(orc.set_k(kperp, kpar), dtype=np.int64)
        out[name + "_k"] = np.array([kperp, kpar])
        oms = np.array(oms)
        res = []
        for o in oms:
            t0 = time.time()
            res.append(orc.disp(complex(o), full=True))
            print("%s om=%s D=%s  (%.1f s)" % (name, o, res[-1][0], time.time() - t0), flush=True)
        out[name + "_om"] = oms
        out[name + "_D"] = np.array([r[0] for r in res])
        out[name + "_chi0"] = np.array([r[1] for r in res])
        out[name + "_chi0_low"] = np.array([r[2] for r in res])
        out[name + "_wave"] = np.array([r[3] for r in res])
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    old = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    old.update(compute(names))
    np.savez_compressed(OUT, **old)
    print("wrote", OUT)
