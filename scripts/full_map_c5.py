"""BASELINE.json configs[4] end to end: the complete nr x ni complex-omega map of the C5 workload (3-species
bi-kappa f0, 1024x2048 grid, nmax = 200) through alps_b200_map_search -- direct quadrature and the k-hoisted map mode --
on one GPU, on a device group (--ngpu N, one process) or collectively over one process per GPU (torchrun: the
library-owned NCCL communicator, OMEGA partition):

    python scripts/full_map_c5.py [--nr 512 --ni 512] [--ngpu N] [--out gpurun_out/full_map_c5.json]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/full_map_c5.py ...

Wall-clock per map includes the host<->device copies, the gather, the sentinels and find_minima (what a user waits for).
bench.py's strong_map leg reports the same numbers in the driver-run line (`--full-map` for the direct 512x512 map)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from alps_b200.solver import Solver
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--nr", type=int, default=512)
ap.add_argument("--ni", type=int, default=512)
ap.add_argument("--ngpu", type=int, default=1)
ap.add_argument("--workload", default="c5")
ap.add_argument("--skip-direct", action="store_true")
ap.add_argument("--out", default="gpurun_out/full_map_c5.json")
a = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
w = bench.WORKLOADS[a.workload]
sol = Solver(bench.build_plasma(w), device=lr, nmax_force=w["nmax_force"], ngpu=a.ngpu if world == 1 else 1)
if world > 1:
    sol.comm_init_torch()
sol.set_map_mode(0)      # the formulation is chosen explicitly below
margs = (w["omr"][0], w["omr"][1], w["omi"][0], w["omi"][1], a.nr, a.ni)
res = {"workload": w["desc"], "nr": a.nr, "ni": a.ni, "n_gpus": world * a.ngpu,
       "how": "device group of one process" if a.ngpu > 1 else ("one process per GPU, library communicator" if world > 1 else "one GPU")}
out = {}
for mode, name in ((0, "direct"), (1, "hoisted")):
    if mode == 0 and a.skip_direct:
        continue
    sol.set_mode(mode)
    nmax = sol.set_k(w["kperp"], w["kpar"])
    sol.map_search(*margs[:4], 16, 9 * max(world, a.ngpu))   # warm-up (allocations, communicator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    sol.set_k(w["kperp"], w["kpar"])              # mode 1: builds the k-hoisted tables (inside the timed region)
    om, val, cal, roots = sol.map_search(*margs, numroots=1000)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out[name] = (val, cal, roots)
    res[name] = {"seconds_per_map": t2 - t1, "D_per_s": a.nr * a.ni / (t2 - t1), "minima_found": len(roots),
                 "first_minima": [[r.real, r.imag] for r in roots[:8]], "finite": bool(np.all(np.isfinite(val)))}
res["nmax"] = [int(n) for n in nmax]
if "direct" in out and "hoisted" in out:
    d, f = out["direct"][1], out["hoisted"][1]
    res["hoisted_vs_direct_max_rel_diff_D"] = float(np.max(np.abs(d - f) / np.abs(d)))
    res["same_minima"] = out["direct"][2] == out["hoisted"][2]
if rank == 0:
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"))
    print(json.dumps(res), flush=True)
sol.close()
if world > 1:
    dist.destroy_process_group()
