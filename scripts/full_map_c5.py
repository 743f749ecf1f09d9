"""BASELINE.json configs[4] end to end: the complete nr x ni complex-omega map of the C5 workload (3-species
bi-kappa f0, 1024x2048 grid, nmax = 200) through map_search -- direct quadrature and, separately, the k-hoisted
map fast path -- on one GPU or sharded over the GPUs of a box:

    python scripts/full_map_c5.py [--nr 512 --ni 512] [--out gpurun_out/full_map_c5.json]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/full_map_c5.py ...

Wall-clock per map includes the host<->device copies, the all_gather and find_minima (what a user waits for)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from alps_b200 import sharding
from alps_b200.solver import Solver
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--nr", type=int, default=512)
ap.add_argument("--ni", type=int, default=512)
ap.add_argument("--workload", default="c5")
ap.add_argument("--skip-direct", action="store_true")
ap.add_argument("--out", default="gpurun_out/full_map_c5.json")
a = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
shard = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    shard = (rank, world, sharding.torch_all_gather())
w = bench.WORKLOADS[a.workload]
sol = Solver(bench.build_plasma(w), device=lr, nmax_force=w["nmax_force"])
margs = (w["omr"][0], w["omr"][1], w["omi"][0], w["omi"][1], a.nr, a.ni)
res = {"workload": w["desc"], "nr": a.nr, "ni": a.ni, "n_gpus": world}
out = {}
for mode, name in ((0, "direct"), (1, "fast_path")):
    if mode == 0 and a.skip_direct:
        continue
    sol.set_mode(mode)
    t0 = time.perf_counter()
    nmax = sol.set_k(w["kperp"], w["kpar"])       # mode 1: builds the k-hoisted tables (inside the timed region)
    sol.map_search(*margs[:4], 8, 8 * max(world, 1) * 9, shard=shard)   # warm-up (allocations, NCCL)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    sol.set_k(w["kperp"], w["kpar"])
    om, val, cal, roots = sol.map_search(*margs, shard=shard, numroots=1000)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out[name] = (val, cal, roots)
    res[name] = {"seconds_per_map": t2 - t1, "D_per_s": a.nr * a.ni / (t2 - t1), "minima_found": len(roots),
                 "first_minima": [[r.real, r.imag] for r in roots[:8]], "finite": bool(np.all(np.isfinite(val)))}
res["nmax"] = [int(n) for n in nmax]
if "direct" in out and "fast_path" in out:
    d, f = out["direct"][1], out["fast_path"][1]
    res["fast_vs_direct_max_rel_diff_D"] = float(np.max(np.abs(d - f) / np.abs(d)))
    res["same_minima"] = out["direct"][2] == out["fast_path"][2]
if rank == 0:
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"))
    print(json.dumps(res), flush=True)
sol.close()
if world > 1:
    dist.destroy_process_group()
