"""Developer aid: is the DMMA quadrature kernel's time data dependent?  Same launch (C5 grid, 208 harmonics = 13 full
tiles, 296 omegas) with the real bi-kappa tables, all-zero tables, and random O(1) tables."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from alps_b200 import _lib
from alps_b200.solver import Solver
w = bench.WORKLOADS["c5"]
B = 296
om = bench.map_omegas(w, 0, 1, B)
rng = np.random.default_rng(1)
for name in ("real bi-kappa tables", "all-zero tables", "random O(1) tables", "real tables clipped below at 1e-30"):
    pl = bench.build_plasma(w)
    if name.startswith("all-zero"):
        pl.f0[...] = 0.0
    elif name.startswith("random"):
        pl.f0[...] = rng.uniform(0.5, 1.5, pl.f0.shape)
    elif "clipped" in name:
        pl.f0[...] = np.maximum(pl.f0, 1e-30)
    print(name, "f0 range", float(pl.f0.min()), float(pl.f0.max()), flush=True)
    sol = Solver(pl, device=0, nmax_force=207, batch_max=B)
    sol.set_k(w["kperp"], w["kpar"])
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); sol.set_stream(st.cuda_stream)
    om_d = torch.from_numpy(om.view(np.float64).copy()).cuda(); D_d = torch.zeros(2 * B, dtype=torch.float64, device="cuda")
    for _ in range(2): sol.disp_batch_dev(B, om_d.data_ptr(), D_d.data_ptr())
    sol.sync(); ms = []
    for _ in range(3):
        sol.disp_batch_dev(B, om_d.data_ptr(), D_d.data_ptr()); sol.sync(); ms.append(sol.info(_lib.INFO_LAST_KERNEL_MS))
    print("   %s: %.2f ms per launch" % (name, min(ms)), flush=True)
    sol.close()
