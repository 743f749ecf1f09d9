"""Developer aid: where the time of the map fast path goes (C5)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alps_b200 import tables, _lib
from alps_b200.solver import Solver
import bench
w = bench.WORKLOADS["c5"]
pl = bench.build_plasma(w)
sol = Solver(pl, nmax_force=200)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sol.set_stream(st.cuda_stream)
sol.set_mode(1)
t = time.perf_counter(); sol.set_k(w["kperp"], w["kpar"]); torch.cuda.synchronize(); print("first set_k %.1f ms" % ((time.perf_counter() - t) * 1e3))
for _ in range(3):
    t = time.perf_counter(); sol.set_k(w["kperp"], w["kpar"]); torch.cuda.synchronize(); print("set_k (same k) %.1f ms" % ((time.perf_counter() - t) * 1e3))
t = time.perf_counter(); sol.set_k(w["kperp"] * 1.01, w["kpar"]); torch.cuda.synchronize(); print("set_k (new kperp) %.1f ms" % ((time.perf_counter() - t) * 1e3))
for n in (4736, 4736 * 4):
    om = bench.map_omegas(w, 0, 1, n)
    om_d = torch.from_numpy(om.view(np.float64).copy()).cuda(); D_d = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
    sol.disp_batch_dev(n, om_d.data_ptr(), D_d.data_ptr()); sol.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sol.disp_batch_dev(n, om_d.data_ptr(), D_d.data_ptr()); e1.record(); e1.synchronize(); sol.sync()
    ms = e0.elapsed_time(e1)
    print("n=%d: %.1f ms total, k_fast %.1f ms -> %.0f D/s" % (n, ms, sol.info(_lib.INFO_LAST_KERNEL_MS), n / ms * 1e3))
sol.close()
