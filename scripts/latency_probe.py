"""Developer aid: latency of one alps_b200_disp call (sequential root finding is latency bound)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alps_b200 import tables
from alps_b200.solver import Solver
pl = tables.config_kpar_fast(); sol = Solver(pl, emulate_nproc=4); sol.set_k(1e-2, 1e-2)
om = 9.98811e-3 - 2.31322e-7j
for _ in range(50): sol.disp(om)
N = 2000
t = time.perf_counter()
for i in range(N): sol.disp(om * (1 + 1e-6 * i))
dt = time.perf_counter() - t
print("disp: %.1f us per call (%.0f D/s)" % (dt / N * 1e6, N / dt))
t = time.perf_counter()
for i in range(200): sol.disp(om, full=True)
print("disp full: %.1f us per call" % ((time.perf_counter() - t) / 200 * 1e6))
for nb in (8, 64, 512, 4096):
    oms = om * (1 + 1e-5 * np.arange(nb))
    sol.disp_batch(oms)
    t = time.perf_counter()
    for _ in range(5): sol.disp_batch(oms)
    dt = (time.perf_counter() - t) / 5
    print("batch %5d: %.1f us per call, %.0f D/s" % (nb, dt * 1e6, nb / dt))
sol.close()

# device-side time of the 5-kernel chain for one omega (async launches, one sync at the end)
import torch
sol = Solver(pl, emulate_nproc=4); sol.set_k(1e-2, 1e-2)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sol.set_stream(st.cuda_stream)
om_d = torch.tensor([om.real, om.imag], dtype=torch.float64, device="cuda")
D_d = torch.zeros(2, dtype=torch.float64, device="cuda")
for _ in range(20): sol.disp_batch_dev(1, om_d.data_ptr(), D_d.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(500): sol.disp_batch_dev(1, om_d.data_ptr(), D_d.data_ptr())
e1.record(); e1.synchronize()
print("device chain for n=1: %.1f us per D" % (e0.elapsed_time(e1) / 500 * 1e3))
sol.close()
