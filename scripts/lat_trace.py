"""Developer aid: device-side timeline of the single-omega chain (k_plan -> k_quad_mma -> k_resonant_lat ->
k_chi_assemble) from %globaltimer stamps of the trace build:

    make -C alps_b200/csrc trace
    ALPS_B200_LIB=alps_b200/libalps_b200_trace.so [ALPS_B200_PDL=0|1] python scripts/lat_trace.py [c1|c2|c3|c4]

Prints the median over calls of every stamp relative to the start of k_plan, and the host-side time per call."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alps_b200 import tables, _lib
from alps_b200.solver import Solver

NAMES = {0: "plan start", 21: "plan: omega read from host", 1: "plan end (last item)",
         2: "quad first CTA start", 3: "quad contraction done (last CTA)", 5: "quad plan fetched (last CTA)",
         4: "quad first CTA end", 7: "quad last CTA end",
         8: "resonant first block start", 10: "resonant first block past the wait", 23: "resonant last block past the wait",
         9: "resonant Landau parts computed (last)", 11: "resonant near-pole parts computed (last)",
         13: "resonant last ticket", 15: "resonant combine done (last item)",
         16: "chi first block start", 18: "chi past its first wait (quad counter or predecessor)", 17: "chi harmonic sums done (last warp)",
         19: "chi determinant written", 25: "chi: species constants in shared memory", 35: "chi: first loads arrived",
         37: "resonant near-pole blocks past the wait for quad (last)", 39: "chi: past the wait for resonant", 41: "resonant near: first point interpolated (thread 0, last block)",
         43: "resonant: block sums in shared memory (last)", 47: "chi: resonant parts combined (last unit)",
         27: "chi: components in shared memory (phase 1)", 29: "chi: ordered sums done (phase 2)"}
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
cfg = {"c1": (tables.config_kpar_fast, (1e-2, 1e-2), 9.98811e-3 - 2.31322e-7j),
       "c2": (tables.config_bimax, (1e-3, 1e-3), 1.0e-3 - 1e-6j),
       "c4": (tables.config_kpar_fast, (3.0, 1e-3), 9.9e-4 - 2e-6j),
       "c3": (lambda: tables.config_relativistic(rel_backend="device"), (1e-3, 1e-3), 1.0e-3 - 1e-6j)}[which]
lib = _lib.lib()
trace = lib.alps_b200_debug_trace
trace.argtypes = [ctypes.c_void_p]
sol = Solver(cfg[0](), emulate_nproc=4); sol.set_k(*cfg[1])
om = cfg[2]
for i in range(50): sol.disp(om * (1 + 1e-7 * i))
buf = np.zeros(64, dtype=np.uint64)
trace(None)
rows, host = [], []
for i in range(300):
    o = om * (1 + 1e-6 * (i + 1))
    t = time.perf_counter()
    sol.disp(o)
    host.append((time.perf_counter() - t) * 1e6)
    trace(buf.ctypes.data)
    rows.append(buf.astype(np.int64).copy())
rows = np.array(rows)
t0 = rows[:, 0:1]
rel = (rows - t0) / 1e3
print("config %s  PDL=%s  host time per disp(): median %.1f us (min %.1f)" %
      (which, os.environ.get("ALPS_B200_PDL", "default"), np.median(host), np.min(host)))
for slot in sorted(NAMES, key=lambda s: np.median(rel[:, s])):
    v = rel[:, slot]
    ok = (rows[:, slot] != 0) & (rows[:, slot] != -1)
    if not ok.any():
        continue
    print("  %7.2f us  (p10 %6.2f, p90 %6.2f)  %s" % (np.median(v[ok]), np.percentile(v[ok], 10), np.percentile(v[ok], 90), NAMES[slot]))
sol.close()
