"""Device group (alps_b200_cfg.ngpu = N, ONE process) on a box with N GPUs: time per host-buffer call under the OMEGA and
HARMONIC partitions against the same call on one GPU, for C4 (tests/test_kperp.in at k_perp = 3, nmax 88/29 on 120x240: a
D is ~50 us of GPU work) and C5 (1024x2048, nmax 200: a D is ~0.45 ms of a whole GPU).
    python scripts/group_probe.py [--ngpu N] [--out gpurun_out/group_probe.json]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from alps_b200 import tables, _lib
from alps_b200.solver import Solver
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--ngpu", type=int, default=torch.cuda.device_count())
ap.add_argument("--out", default="gpurun_out/group_probe.json")
ap.add_argument("--skip-c5", action="store_true")
a = ap.parse_args()
rng = np.random.default_rng(5)


def timed(fn, reps):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


def probe(name, make, kw, kperp, kpar, om_of, sizes):
    out = {"config": name}
    ref = {}
    oms = {n: om_of(n) for n in sizes}      # the same omegas for every partition
    for label, ngpu, part, red in (("one_gpu", 1, _lib.PARTITION_OMEGA, None),
                                   ("omega", a.ngpu, _lib.PARTITION_OMEGA, None),
                                   ("harmonic_p2p", a.ngpu, _lib.PARTITION_HARMONIC, "p2p"),
                                   ("harmonic_nccl", a.ngpu, _lib.PARTITION_HARMONIC, "nccl")):
        if red:
            os.environ["ALPS_B200_REDUCE"] = red
        sol = Solver(make(), ngpu=ngpu, **kw)
        try:
            sol.set_partition(part)
            out["nmax"] = [int(v) for v in sol.set_k(kperp, kpar)]
            for n in sizes:
                om = oms[n]
                reps = 300 if n == 1 else max(3, min(100, 20000 // n))
                if n == 1:
                    k = [0]

                    def one():      # a fresh omega per call: no memo hits
                        k[0] += 1
                        sol.disp(complex(om[0]) * (1.0 + 1e-9 * k[0]))
                    t = timed(one, reps)
                    D = np.array([sol.disp(complex(om[0]))])
                else:
                    t = timed(lambda: sol.disp_batch(om), reps)
                    D = sol.disp_batch(om)
                r = out.setdefault("n%d" % n, {})
                r[label + "_us"] = t
                if label == "one_gpu":
                    ref[n] = D
                else:
                    r[label + "_max_rel_diff"] = float(np.max(np.abs(D - ref[n]) / np.abs(ref[n])))
        finally:
            sol.close()
    print(json.dumps(out), flush=True)
    return out


res = {"n_gpus": a.ngpu, "what": "us per host-buffer call (wall clock incl. H2D/D2H), device group of one process"}
res["c4"] = probe("C4 test_kperp at k_perp=3, k_par=1e-3 (120x240, nmax 88/29)", tables.config_kpar_fast,
                  dict(emulate_nproc=4), 3.0, 1.0e-3,
                  lambda n: 9.9e-4 * (1.0 + 0.05 * rng.uniform(-1, 1, n)) + 1j * 2e-5 * rng.uniform(-1, 1, n),
                  (1, 64, 1024, 16384))
if not a.skip_c5:
    w = bench.WORKLOADS["c5"]
    res["c5"] = probe("C5 " + w["desc"], lambda: bench.build_plasma(w), dict(nmax_force=200), w["kperp"], w["kpar"],
                      lambda n: bench.map_omegas(w, 0, 1, n), (1, 8, 296 * a.ngpu))
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
