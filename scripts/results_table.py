"""Measured D(omega,k) evaluations per second for the BASELINE.md configurations (run under gpurun).
GPU: batched (disp_batch, host buffers) and one-at-a-time (disp) through the C ABI; CPU: the restated
reference (oracle, OpenMP) on the same tables."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alps_b200 import tables
from alps_b200.solver import Solver
from oracle.oracle import Oracle

def measure(name, pl, kperp, kpar, om0, nbatch, nproc=0, ncpu=3, modes=(0, 1)):
    rng = np.random.default_rng(5)
    oms = om0 * (1.0 + 0.05 * rng.uniform(-1, 1, nbatch)) + 1j * abs(om0) * 0.02 * rng.uniform(-1, 1, nbatch)
    out = {"config": name, "nbatch": nbatch}
    sol = Solver(pl, emulate_nproc=nproc)
    out["nmax"] = [int(n) for n in sol.set_k(kperp, kpar)]
    for mode in modes:
        if mode == 1 and any(s.relativistic for s in pl.species):
            continue
        sol.set_mode(mode); sol.set_k(kperp, kpar)
        sol.disp_batch(oms)
        t = time.perf_counter(); sol.disp_batch(oms); dt = time.perf_counter() - t
        out["gpu_batch_Dps_mode%d" % mode] = nbatch / dt
    sol.set_mode(0); sol.set_k(kperp, kpar)
    for _ in range(20): sol.disp(complex(oms[0]))
    t = time.perf_counter()
    for i in range(200): sol.disp(complex(oms[i % nbatch]))
    out["gpu_single_Dps"] = 200 / (time.perf_counter() - t)
    sol.close()
    orc = Oracle(pl, nproc=nproc); orc.set_k(kperp, kpar)
    t = time.perf_counter()
    for i in range(ncpu): orc.disp(complex(oms[i]))
    out["cpu_oracle_Dps"] = ncpu / (time.perf_counter() - t)
    out["cpu_cores"] = os.cpu_count()
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    only = sys.argv[1:]
    _measure = measure
    def measure(name, *a, **k):
        if not only or any(name.startswith(o) for o in only):
            _measure(name, *a, **k)
    measure("C1 test_kpar_fast (120x240, nmax 21/13)", tables.config_kpar_fast(), 1e-2, 1e-2, 9.98811e-3, 4096, nproc=4)
    measure("test_map-like 50x50 map on the C1 tables", tables.config_kpar_fast(), 1e-2, 1e-2, 1.0e-2, 2500, nproc=4)
    measure("C2 test_bimax (150x300, protons NHDS, electrons table)", tables.config_bimax(), 1e-3, 0.03, 3.0e-2, 4096)
    measure("C3 test_relativistic (rel grid 500x500, nmax 14/14)", tables.config_relativistic(), 1e-3, 1e-1, 6.2713e-2, 256, ncpu=2)
    measure("C4 test_kperp at k_perp=3 (120x240)", tables.config_kpar_fast(), 3.0, 1e-3, 9.9e-4, 2048, nproc=4, ncpu=2)
