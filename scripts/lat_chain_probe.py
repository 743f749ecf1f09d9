"""Developer aid: wall-clock per alps_b200_disp call (and per small batch) on the shipped configurations -- the latency
chain of DESIGN.md 4c.  Knobs come from the environment (ALPS_B200_PDL, ALPS_B200_ZC, ...)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alps_b200 import tables
from alps_b200.solver import Solver

def probe(name, pl, k, om, **kw):
    sol = Solver(pl, **kw); sol.set_k(*k)
    for _ in range(50): sol.disp(om)
    out = []
    for nb in (1, 3, 8):
        best = 1e9
        for rep in range(5):
            N = 400
            oms = [om * (1 + 1e-6 * (rep * N + i + 1) * np.arange(1, nb + 1)) for i in range(N)]   # never the same omega twice (memo)
            t = time.perf_counter()
            if nb == 1:
                for i in range(N): sol.disp(complex(oms[i][0]))
            else:
                for i in range(N): sol.disp_batch(oms[i])
            best = min(best, (time.perf_counter() - t) / N * 1e6)
        out.append("n=%d %.1f us" % (nb, best))
    # the C ABI itself: alps_b200_disp with preallocated buffers (Solver.disp allocates two arrays and builds the pointers
    # on every call: ~5 us of Python)
    import ctypes
    from alps_b200 import _lib
    omv, D = np.zeros(2), np.zeros(2)
    p_om, p_D = omv.ctypes.data_as(ctypes.c_void_p), D.ctypes.data_as(ctypes.c_void_p)
    f = sol.L.alps_b200_disp
    best = 1e9
    for rep in range(5):
        N = 400
        t = time.perf_counter()
        for i in range(N):
            o = om * (1 + 1e-6 * (7777 + rep * N + i))
            omv[0] = o.real; omv[1] = o.imag
            f(p_om, p_D, None, None, None)
        best = min(best, (time.perf_counter() - t) / N * 1e6)
    out.append("C ABI n=1 %.1f us" % best)
    print("%-16s %s   D(om) = %r" % (name, "  ".join(out), sol.disp(om)), flush=True)
    sol.close()

probe("C1 kpar_fast", tables.config_kpar_fast(), (1e-2, 1e-2), 9.98811e-3 - 2.31322e-7j, emulate_nproc=4)
probe("C2 bimax", tables.config_bimax(), (1e-3, 1e-3), 1.0e-3 - 1e-6j, emulate_nproc=4)
probe("C4 kperp k=3", tables.config_kpar_fast(), (3.0, 1e-3), 9.9e-4 - 2e-6j, emulate_nproc=4)
try:
    probe("C3 relativistic", tables.config_relativistic(rel_backend="device"), (1e-3, 1e-3), 1.0e-3 - 1e-6j, emulate_nproc=4)
except Exception as e:
    print("C3 skipped:", e)
