"""torchrun check of the one-process-per-GPU mode on real GPUs: the library-owned NCCL communicator
(alps_b200_comm_init) under both partitions (include/alps_b200.h):
   OMEGA     disp_batch / map_search are collective, every rank evaluates a slice, one ncclAllGather inside the
             library -- bitwise the single-GPU result on every rank;
   HARMONIC  every rank sums its block of harmonics, ncclAllReduce of the chi partials on the library's stream.
and the time per call of a C4 batch (tests/test_kperp.in at k_perp = 3) under each.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_check.py [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from alps_b200 import tables, _lib
from alps_b200.solver import Solver

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pl = tables.config_small(48, 96, kind=2)
sol = Solver(pl, device=lr)
kperp, kpar = 1.5, 0.05
sol.set_k(kperp, kpar)
rng = np.random.default_rng(7)
margs = (0.05, 2.0, -0.03, 0.03, 40, 32)
sizes = (9, 37, 64, 65, 100 * world + 3)
oms = {n: rng.uniform(0.05, 2.0, n) + 1j * rng.uniform(-0.03, 0.03, n) for n in sizes}
# single-GPU answers, before the communicator exists (every rank on its own GPU)
D_one = {n: sol.disp_batch(oms[n]) for n in sizes}
map_one = sol.map_search(*margs)
full_one = sol.disp_batch_full(oms[37])
d_single = sol.disp(complex(oms[9][0]))
sol.comm_init_torch()
assert int(sol.info(_lib.INFO_NGPU)) == world
# (i) OMEGA partition (default): collective disp_batch, bitwise the single-GPU result for every batch class
for n in sizes:
    D = sol.disp_batch(oms[n])
    assert np.array_equal(D.view(np.float64), D_one[n].view(np.float64)), ("omega partition differs", n)
om_s, val_s, cal_s, roots_s = sol.map_search(*margs)
ok_map = (np.array_equal(om_s, map_one[0]) and np.array_equal(cal_s, map_one[2]) and np.array_equal(val_s, map_one[1])
          and roots_s == map_one[3])
assert ok_map
assert sol.disp(complex(oms[9][0])) == d_single          # single omegas are not partitioned
print("rank %d/%d OMEGA partition: disp_batch %s and map_search bitwise the single-GPU results, %d minima"
      % (rank, world, list(sizes), len(roots_s)), flush=True)
# (ii) HARMONIC partition: ncclAllReduce inside the library
sol.set_partition(_lib.PARTITION_HARMONIC)
nmax = sol.set_k(kperp, kpar)
worst = 0.0
for n in sizes:
    D = sol.disp_batch(oms[n])
    worst = max(worst, float(np.max(np.abs(D - D_one[n]) / np.abs(D_one[n]))))
Dh, chi_h, low_h, wave_h = sol.disp_batch_full(oms[37])
worst = max(worst, float(np.max(np.abs(chi_h - full_one[1])) / np.max(np.abs(full_one[1]))),
            float(np.max(np.abs(wave_h - full_one[3])) / np.max(np.abs(full_one[3]))))
d1 = sol.disp(complex(oms[9][0]))
worst = max(worst, abs(d1 - d_single) / abs(d_single))
print("rank %d/%d HARMONIC partition: max rel diff to the single-GPU result %.2e, nmax=%s" % (rank, world, worst, list(nmax)),
      flush=True)
assert worst < 1e-10
sol.set_partition(_lib.PARTITION_OMEGA)
sol.close()

# (iii) C4 (tests/test_kperp.in at k_perp = 3: nmax 88 / 29 on the 120x240 grid): wall time per host-buffer call, max
# over ranks, single GPU vs the two partitions
pl = tables.config_kpar_fast()
sol = Solver(pl, device=lr, emulate_nproc=4)
sol.comm_init_torch()
kperp, kpar = 3.0, 1e-3
res = {"config": "C4 test_kperp at k_perp=3, k_par=1e-3 (nmax 88/29), host-buffer calls, us per call (max over ranks)",
       "n_gpus": world}


def timed(fn, reps):
    for _ in range(3):
        fn()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t = torch.tensor([(time.perf_counter() - t0) / reps * 1e6], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for n in (1, 64, 1024, 16384):
    om = 9.9e-4 * (1.0 + 0.05 * rng.uniform(-1, 1, n)) + 1j * 2e-5 * rng.uniform(-1, 1, n)
    om_d = torch.from_numpy(om.view(np.float64).copy()).cuda()
    D_d = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
    reps = 200 if n <= 64 else (30 if n <= 1024 else 5)
    sol.set_partition(_lib.PARTITION_OMEGA); nmax = sol.set_k(kperp, kpar)

    def one_gpu():      # device buffers are never partitioned: this rank's GPU alone
        sol.disp_batch_dev(n, om_d.data_ptr(), D_d.data_ptr()); sol.sync()
    t_one = timed(one_gpu, reps)
    D_ref = D_d.cpu().numpy().view(np.complex128).copy()
    t_om = timed(lambda: sol.disp_batch(om), reps)
    D_om = sol.disp_batch(om)
    sol.set_partition(_lib.PARTITION_HARMONIC); sol.set_k(kperp, kpar)
    t_h = timed(lambda: sol.disp_batch(om), reps)
    D_h = sol.disp_batch(om)
    res["n%d" % n] = {"one_gpu_device_buffers_us": t_one, "omega_partition_us": t_om, "harmonic_partition_us": t_h,
                      "omega_bitwise": bool(np.array_equal(D_om.view(np.float64), D_ref.view(np.float64))) if n > 8 else None,
                      "harmonic_max_rel_diff": float(np.max(np.abs(D_h - D_ref) / np.abs(D_ref)))}
res["nmax"] = [int(v) for v in nmax]
if rank == 0:
    print(json.dumps(res), flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)
sol.close()
dist.destroy_process_group()
