"""Writes tests/golden/oracle_vectors.npz: D, chi0, chi0_low and wave of the CPU oracle (oracle/, the restatement of
the reference's disp()) for seeded inputs -- small bi-Maxwellian and bi-kappa plasmas (tables.config_small) and the
test_kpar_fast tables, omegas in both half planes, on the real axis and next to a resonance.  The fixture guards the
oracle against regressions (tests/test_oracle_golden.py) and gives the CUDA path committed vectors to match without
running the oracle (tests/test_gpu_parity.py).      python tests/golden/make_oracle_vectors.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from alps_b200 import tables          # noqa: E402
from oracle.oracle import Oracle      # noqa: E402

CASES = {
    "small_bimax": (lambda: tables.config_small(24, 48, kind=1), 0, 0.3, 0.05),
    "small_kappa": (lambda: tables.config_small(24, 48, kind=2), 0, 0.8, -0.07),
    "kpar_fast": (tables.config_kpar_fast, 4, 1.0e-2, 2.0e-2),
}
OMEGAS = {
    "small_bimax": [0.3 + 0.01j, 0.31 - 0.02j, 0.011 - 1e-6j, 0.7 + 0j, 1.3 - 0.004j],
    "small_kappa": [0.25 + 0.02j, 0.9 - 0.03j, 0.05 + 0j, 1.7 - 1e-4j],
    "kpar_fast": [1.99e-2 - 1.0e-5j, 5.0e-3 + 2.0e-4j, 2.0e-2 + 0j, 0.3 - 1.0e-3j],
}


def compute():
    out = {}
    for name, (make, nproc, kperp, kpar) in CASES.items():
        pl = make()
        orc = Oracle(pl, nproc=nproc)
        out[name + "_nmax"] = np.asarray(orc.set_k(kperp, kpar), dtype=np.int64)
        out[name + "_k"] = np.array([kperp, kpar])
        oms = np.array(OMEGAS[name])
        res = [orc.disp(complex(o), full=True) for o in oms]
        out[name + "_om"] = oms
        out[name + "_D"] = np.array([r[0] for r in res])
        out[name + "_chi0"] = np.array([r[1] for r in res])
        out[name + "_chi0_low"] = np.array([r[2] for r in res])
        out[name + "_wave"] = np.array([r[3] for r in res])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **compute())
    print("wrote", os.path.join(HERE, "oracle_vectors.npz"))
