"""Developer aid: k_quad_mma time per 296-omega launch on the C5 tables for harmonic counts around a tile boundary
(packed / empty remainder groups)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from alps_b200 import _lib
from alps_b200.solver import Solver
w = bench.WORKLOADS["c5"]
pl = bench.build_plasma(w)
B = 296
om = bench.map_omegas(w, 0, 1, B)
for nmax in (191, 192, 193, 199, 200, 207):
    sol = Solver(pl, device=0, nmax_force=nmax, batch_max=B)
    sol.set_k(w["kperp"], w["kpar"])
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); sol.set_stream(st.cuda_stream)
    om_d = torch.from_numpy(om.view(np.float64).copy()).cuda(); D_d = torch.zeros(2 * B, dtype=torch.float64, device="cuda")
    for _ in range(2): sol.disp_batch_dev(B, om_d.data_ptr(), D_d.data_ptr())
    sol.sync(); ms = []
    for _ in range(3):
        sol.disp_batch_dev(B, om_d.data_ptr(), D_d.data_ptr()); sol.sync(); ms.append(sol.info(_lib.INFO_LAST_KERNEL_MS))
    print("nmax %d: %d harmonics, %.2f ms per launch, %.4f ms per 16-harmonic tile-equivalent" % (nmax, nmax + 1, min(ms), min(ms) / ((nmax + 1) / 16.0)), flush=True)
    sol.close()
