"""Developer aid: device time of the quadrature kernel on the C5 workload for the current
ALPS_B200_QUAD_VARIANT (no result checks -- also usable with the ALPS_QUAD_DEBUG ablation variants)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from alps_b200 import _lib
from alps_b200.solver import Solver

w = bench.WORKLOADS["c5"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
sol = Solver(bench.build_plasma(w), nmax_force=w["nmax_force"], batch_max=B)
sol.set_k(w["kperp"], w["kpar"])
om = bench.map_omegas(w, 0, 1, B)
ms = []
for _ in range(3):
    try:
        sol.disp_batch(om)
    except Exception as e:  # ablation variants produce garbage
        print("note:", str(e)[:80])
    ms.append(sol.info(_lib.INFO_LAST_KERNEL_MS))
print(json.dumps({"variant": os.environ.get("ALPS_B200_QUAD_VARIANT"), "batch": B, "kernel_ms": ms}))
