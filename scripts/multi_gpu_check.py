"""torchrun check of both multi-GPU modes on real GPUs (NCCL):
   omega sharding (no data-path collective, one all_gather) and harmonic sharding (all_reduce of the
   chi partials).  torchrun --nproc-per-node N scripts/multi_gpu_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from alps_b200 import tables, sharding
from alps_b200.solver import Solver

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pl = tables.config_small(48, 96, kind=2)
sol = Solver(pl, device=lr)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sol.set_stream(st.cuda_stream)
kperp, kpar = 1.5, 0.05
sol.set_k(kperp, kpar)
rng = np.random.default_rng(7)
n = 128 * world     # every shard stays in the throughput batch class (> 64 omegas), where the summation order is
                    # independent of the batch size: the gathered map is bitwise the single-GPU map
oms = rng.uniform(0.05, 2.0, n) + 1j * rng.uniform(-0.03, 0.03, n)
D_full = sol.disp_batch(oms)          # every rank computes the reference answer on its own GPU
# (i) omega sharding
lo, hi = sharding.omega_shard(n, rank, world)
local = sol.disp_batch(oms[lo:hi])
def all_gather(pad):
    t = torch.from_numpy(pad.view(np.float64).copy()).cuda()
    outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [o.cpu().numpy().view(np.complex128) for o in outs]
D1 = sharding.gather_omega_shards(local, n, rank, world, all_gather)
ok1 = np.array_equal(D1, D_full)
# (ib) map_search sharded over the ranks (alps_b200_map_grid / disp_batch slices / all_gather / alps_b200_map_finish)
# against the single-process alps_b200_map_search on this rank's GPU: same D, same val, same minima
margs = (0.05, 2.0, -0.03, 0.03, 40, 32)
om_s, val_s, cal_s, roots_s = sol.map_search(*margs, shard=(rank, world, sharding.torch_all_gather()))
om_1, val_1, cal_1, roots_1 = sol.map_search(*margs)
ok_map = (np.array_equal(om_s, om_1) and np.array_equal(cal_s, cal_1) and np.array_equal(val_s, val_1)
          and roots_s == roots_1)
print("rank %d/%d sharded map_search identical=%s minima=%d" % (rank, world, ok_map, len(roots_1)), flush=True)
assert ok_map
# (ii) harmonic sharding + NCCL all_reduce of the partials
sol.set_harmonic_shard(rank, world); sol.set_k(kperp, kpar)
L = sol.chi_partial_len()
om_d = torch.from_numpy(oms.view(np.float64).copy()).cuda()
part = torch.zeros(n * L, dtype=torch.float64, device="cuda")
sol.chi_partial_dev(n, om_d.data_ptr(), part.data_ptr())
dist.all_reduce(part, op=dist.ReduceOp.SUM)
D_d = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
sol.assemble_dev(n, om_d.data_ptr(), part.data_ptr(), D_d.data_ptr())
torch.cuda.synchronize()
D2 = D_d.cpu().numpy().view(np.complex128)
err2 = float(np.max(np.abs(D2 - D_full) / np.abs(D_full)))
print("rank %d/%d omega-shard identical=%s harmonic-shard max rel err=%.2e nmax=%s" % (rank, world, ok1, err2, list(sol.nmax)), flush=True)
assert ok1 and err2 < 1e-10
sol.close()

# (iii) C4 (tests/test_kperp.in at k_perp = 3: nmax 88 / 29 on the 120x240 grid): time per call of the
# harmonic-sharded chain (chi partials -> NCCL all_reduce -> assemble) against the unsharded chain on one GPU
import json
pl = tables.config_kpar_fast()
sol = Solver(pl, device=lr, emulate_nproc=4)
sol.set_stream(st.cuda_stream)
kperp, kpar = 3.0, 1e-3
res = {"config": "C4 test_kperp at k_perp=3", "n_gpus": world}
for n in (1, 64, 1024):
    oms = 9.9e-4 * (1.0 + 0.05 * rng.uniform(-1, 1, n)) + 1j * 2e-5 * rng.uniform(-1, 1, n)
    om_d = torch.from_numpy(oms.view(np.float64).copy()).cuda()
    D_d = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
    sol.set_harmonic_shard(0, 1); nmax = sol.set_k(kperp, kpar)
    def timed(fn, reps=30):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3   # us per call, max over ranks
    t_one = timed(lambda: sol.disp_batch_dev(n, om_d.data_ptr(), D_d.data_ptr()))
    D_ref = D_d.cpu().numpy().view(np.complex128).copy()
    sol.set_harmonic_shard(rank, world); sol.set_k(kperp, kpar)
    L = sol.chi_partial_len()
    part = torch.zeros(n * L, dtype=torch.float64, device="cuda")
    def sharded():
        sol.chi_partial_dev(n, om_d.data_ptr(), part.data_ptr())
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        sol.assemble_dev(n, om_d.data_ptr(), part.data_ptr(), D_d.data_ptr())
    t_sh = timed(sharded)
    D_sh = D_d.cpu().numpy().view(np.complex128)
    res["n%d" % n] = {"unsharded_us_per_call": t_one, "harmonic_sharded_us_per_call": t_sh,
                      "allreduce_bytes": int(n * L * 8), "max_rel_diff": float(np.max(np.abs(D_sh - D_ref) / np.abs(D_ref)))}
res["nmax"] = [int(v) for v in nmax]
if rank == 0:
    print(json.dumps(res), flush=True)
sol.close(); dist.destroy_process_group()
