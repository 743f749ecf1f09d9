#!/usr/bin/env python
"""bench.py -- D(omega,k) evaluations per second of the disp() hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c5|c5small]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...       (one rank per GPU)

Workload: BASELINE config 5 (C5: synthetic 3-species bi-kappa f0 on 1024x2048, nmax = 200, k = (15.5, 1e-2), omegas of
the 512x512 complex-omega map).  A "step" is one pass of the hot path over one batch of 296 omegas per GPU.

The headline keys (all direct quadrature, the formulation north_star prescribes):
  value      device-resident omegas in, D out (alps_b200_disp_batch_dev), CUDA-event timed; every rank its own batch
             (weak scaling, no data-path collective)
  e2e        the same batch through the host-buffer call alps_b200_disp_batch (H2D of the omegas and D2H of D inside the
             timed region)
  roofline   k_quad_mma (DMMA.8x8x4 on the FP64 units) against the DMMA micro-benchmark of the same job: useful flops
             12 per (|n|, iperp, ipar) / CUDA-event time of that kernel
  cpu_baseline  the CPU oracle (restated reference, `mpirun -np <cores>` emulated) on a bounded sample: strided harmonic
             subsets that together make up whole D evaluations
Reported beside them, never mixed into value / roofline:
  fast_path     the k-hoisted map mode (alps_b200_set_mode(1)): D/s incl. the per-k table build, and its own roofline
                (k_fast_tiled against the DFMA micro-benchmark, flop model stated in the line)
  strong_map    the PRODUCT's multi-GPU path: complete complex-omega maps through alps_b200_map_search, collective over
                the library-owned NCCL communicator for N > 1 (slices per rank, one ncclAllGather, sentinels, .map-less
                find_minima inside the wall time): the full 512x512 map in hoisted mode and a 256x128 sub-map (or the
                full map with --full-map) in direct mode.  Strong scaling: the same map on 1..N GPUs.
  harmonic_shard  (N > 1) C4 (tests/test_kperp.in at k_perp = 3) batches of 1 / 64 / 1024 omegas: one GPU vs the OMEGA
                partition (ncclAllGather) vs the HARMONIC partition (ncclAllReduce of the chi partials inside the call)
  extra         C1..C4 of BASELINE.json on this GPU: batched D/s (direct, hoisted) and one disp() at a time
`--impl reference` times the restated reference (oracle/) alone on the host cores: the Fortran/MPI reference cannot be
built in this image (no gfortran, no MPI).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOPS_PER_POINT_HARMONIC = 34.0     # per (signed n, iperp, ipar), +n and -n separately
# useful flops of k_quad's formulation (DESIGN.md "flop model"): per (|n|, iperp, ipar) 3 weight types x
# 2 real tables (A', C') x one FMA = 12 flops, shared by +n and -n.  Tile padding and the epilogue are
# NOT counted (ncu's executed count is ~5 % higher).
FLOPS_EXECUTED_PER_ABSN_POINT = 12.0
# k_fast_tiled (hoisted mode), per (|n|, ipar, omega): 64 FP64 instructions = 54 FMA (48 moment updates, 2 squared
# denominators, 4 Newton steps of the shared reciprocal) + 10 DADD/DMUL  ->  118 flops (alps_b200/csrc/fast_kernel.cu)
FLOPS_FAST_PER_ABSN_PAR = 118.0
FP64_NOMINAL_TFLOPS = 37.2          # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz
TRAFFIC_PROFILE = "profiles/r02_k_quad_traffic.json"   # ncu --set full capture of k_quad_mma (regular-tile launch: 94 % of the step)

WORKLOADS = {
    # name: (description, builder kwargs)
    "c5": dict(desc="C5 synthetic 3-species bi-kappa f0, 1024x2048 (p_perp,p_par) grid, nmax=200 forced, "
                    "k=(15.5,1e-2), omegas from the 512x512 map om_r in [0.05,3.05] x gamma in [-0.05,0.05]",
               nperp=1024, npar=2048, nmax_force=200, kperp=15.5, kpar=1.0e-2,
               omr=(0.05, 3.05), omi=(-0.05, 0.05), nr=512, ni=512, batch=296, sub=(256, 128)),
    "c5small": dict(desc="reduced C5 (256x512 grid, nmax=48) -- smoke runs only",
                    nperp=256, npar=512, nmax_force=48, kperp=5.0, kpar=1.0e-2,
                    omr=(0.05, 3.05), omi=(-0.05, 0.05), nr=512, ni=512, batch=296, sub=(64, 32)),
}


def map_omegas(w, rank, world, batch):
    """this rank's omegas: a strided sample of its contiguous block of map rows
    (map grid of map_search, src/ALPS_fns.f90:3684-3712, linear in both directions)"""
    nr, ni = w["nr"], w["ni"]
    wr = w["omr"][0] + (w["omr"][1] - w["omr"][0]) * np.arange(nr) / (nr - 1)
    wi = w["omi"][0] + (w["omi"][1] - w["omi"][0]) * np.arange(ni) / (ni - 1)
    rows = np.array_split(np.arange(nr), world)[rank]
    grid = (wr[rows][:, None] + 1j * wi[None, :]).ravel()
    idx = (np.arange(batch) * (grid.size // batch) + (grid.size // (2 * batch))) % grid.size
    return np.ascontiguousarray(grid[idx])


def build_plasma(w):
    from alps_b200 import tables
    return tables.config_kappa3(w["nperp"], w["npar"])


class ClockSampler:
    """nvidia-smi sampled every 50 ms during the timed region; the reported clock is the median over the samples taken
    while the GPU was busy (utilisation >= 50 %), with the minimum and the power draw beside it, so a power- or
    thermally-limited clock under the FP64 load shows up instead of hiding behind the idle boost clock."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        rows, smax, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            try:
                pw = float(f[3])
            except ValueError:
                pw = None
            try:
                ut = float(f[4])
            except ValueError:
                ut = None
            rows.append((sm, pw, ut))
            smax.append(mx)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [r for r in rows if r[2] is not None and r[2] >= 50.0] or rows
        sm_b = [r[0] for r in busy]
        pw_b = [r[1] for r in busy if r[1] is not None]
        return {"sm_mhz": float(np.median(sm_b)) if sm_b else None, "sm_mhz_min": min(sm_b) if sm_b else None,
                "sm_max_mhz": max(smax) if smax else None, "power_w": float(np.median(pw_b)) if pw_b else None,
                "reasons": sorted(reasons), "samples": len(rows), "samples_busy": len(busy) if rows else 0}


# ------------------------------------------------------------------------------------------------ CPU arm
def sample_stride(steps):
    """The timed steps of a CPU run are strided harmonic subsets (|n| % stride == offset, offset cycling over the steps):
    a strided subset holds the expensive resonant harmonics in proportion, and stride consecutive steps add up to
    exactly one full D(omega,k).  stride = the number of timed steps when that is small, else its largest divisor <= 12."""
    if steps <= 12:
        return max(1, steps)
    for s in range(12, 1, -1):
        if steps % s == 0:
            return s
    return 8


def run_reference(args, w, rank, world):
    """restated reference (CPU oracle, `mpirun -np <cores>` emulated: split_processes over cores - 1 workers, one OpenMP
    thread per core) on the host cores; rank 0 alone."""
    if rank != 0:
        return
    from oracle.oracle import Oracle
    cores = os.cpu_count() or 1
    nproc = max(4, cores - cores % 2)
    plasma = build_plasma(w)
    om = map_omegas(w, 0, 1, w["batch"])
    orc = Oracle(plasma, nproc=nproc, threads=cores, nmax_force=w["nmax_force"])
    nmax = orc.set_k(w["kperp"], w["kpar"])
    stride = sample_stride(args.steps)
    times = []
    for i in range(args.warmup + args.steps):
        timed = i >= args.warmup
        # warm-up steps are thin samples (threads, caches); timed step j evaluates the harmonics |n| % stride == j % stride
        orc.set_sample(stride if timed else 4 * stride, (i - args.warmup) % stride if timed else 0)
        t0 = time.perf_counter()
        orc.disp(complex(om[(7 * i) % om.size]))
        if timed:
            times.append(time.perf_counter() - t0)
    orc.set_sample(0, 0)
    # stride consecutive steps = one D: seconds per D = stride x mean step (exact when steps is a multiple of stride)
    sec_per_D = stride * float(np.mean(times))
    dps = 1.0 / sec_per_D
    sample = ("each timed step evaluates the harmonics |n| %% %d == step %% %d of one D (C5: nmax = %s) -- %d steps = %.2f "
              "complete D evaluations, %.1f s of CPU work; restated reference (CPU oracle, OpenMP, split_processes for "
              "mpirun -np %d emulated); the Fortran/MPI build is impossible here (no gfortran, no MPI)"
              % (stride, stride, [int(n) for n in nmax], args.steps, args.steps / float(stride), sum(times), nproc))
    line = {"impl": "reference", "metric": "D(omega,k) evals/sec", "value": dps, "unit": "D/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": w["desc"]},
            "cpu_baseline": {"value": dps, "unit": "D/s", "cores": cores, "kind": "port", "mpirun_np_emulated": nproc,
                             "seconds_per_D": sec_per_D, "sample": sample},
            "e2e": {"value": dps, "unit": "D/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_leg(args, steps=3):
    """cpu_baseline of our arm: the reference arm in a fresh interpreter (no torch / CUDA threads beside OpenMP)"""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload",
                              args.workload, "--steps", str(steps), "--warmup", "1"], capture_output=True, text=True,
                             timeout=1200, env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"))
        return json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception as e:     # the baseline is reported, never required
        return {"value": None, "unit": "D/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}


# ------------------------------------------------------------------------------------------------ GPU arm
def extra_configs():
    """C1..C4 of BASELINE.json on this GPU (milliseconds each): batched D/s through alps_b200_disp_batch (direct and
    hoisted) and microseconds per alps_b200_disp call (fresh omegas: no memo hits)."""
    from alps_b200 import tables
    from alps_b200.solver import Solver
    rng = np.random.default_rng(5)
    out = []
    cfgs = [("C1 test_kpar_fast (120x240, nmax 21/13)", tables.config_kpar_fast, dict(emulate_nproc=4), 1e-2, 1e-2, 9.98811e-3, 4096),
            ("C2 test_bimax (150x300, protons NHDS k_nhds, electrons table)", tables.config_bimax, {}, 1e-3, 0.03, 3.0e-2, 4096),
            ("C3 test_relativistic (rel grid 500x500, nmax 14/14)", lambda: tables.config_relativistic(rel_backend="device"),
             {}, 1e-3, 1e-1, 6.2713e-2, 2048),
            ("C4 test_kperp at k_perp=3 (120x240, nmax 88/29)", tables.config_kpar_fast, dict(emulate_nproc=4), 3.0, 1e-3, 9.9e-4, 2048)]
    for name, make, kw, kperp, kpar, om0, nb in cfgs:
        pl = make()
        rel = any(s.relativistic for s in pl.species)
        oms = om0 * (1.0 + 0.05 * rng.uniform(-1, 1, nb)) + 1j * abs(om0) * 0.02 * rng.uniform(-1, 1, nb)
        sol = Solver(pl, **kw)
        try:
            r = {"config": name, "nmax": [int(n) for n in sol.set_k(kperp, kpar)], "omegas_per_batch": nb}
            for mode, key in ((0, "direct"), (1, "hoisted")):
                if mode == 1 and rel:
                    continue
                sol.set_mode(mode)
                sol.set_k(kperp, kpar)
                sol.disp_batch(oms)
                t = time.perf_counter()
                for _ in range(3):
                    sol.disp_batch(oms)
                r["batched_%s_D_per_s" % key] = 3 * nb / (time.perf_counter() - t)
            sol.set_mode(0)
            sol.set_k(kperp, kpar)
            for i in range(20):
                sol.disp(complex(oms[i]))
            t = time.perf_counter()
            for i in range(200):
                sol.disp(complex(oms[20 + i]))
            r["single_disp_us"] = (time.perf_counter() - t) / 200 * 1e6
            out.append(r)
        finally:
            sol.close()
    return out


def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from alps_b200 import _lib, tables
    from alps_b200.solver import Solver

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; alps_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    plasma = build_plasma(w)
    B = args.batch or w["batch"]
    sol = Solver(plasma, device=local_rank, nmax_force=w["nmax_force"], batch_max=B)
    nmax = sol.set_k(w["kperp"], w["kpar"])
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sol.set_stream(stream.cuda_stream)
    om_h = map_omegas(w, rank, world, B)
    om_d = torch.from_numpy(om_h.view(np.float64).copy()).cuda()
    D_d = torch.zeros(2 * B, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    ph = sol.info(_lib.INFO_POINT_HARMONICS)
    flops_per_D = FLOPS_PER_POINT_HARMONIC * ph
    absn = sum(int(n) + 1 for n in nmax)
    absn_points = absn * (w["nperp"] - 1.0) * (w["npar"] - 1.0)
    flops_exec_per_D = FLOPS_EXECUTED_PER_ABSN_POINT * absn_points
    variant = int(sol.info(_lib.INFO_QUAD_VARIANT))
    dmma = variant >= 9     # quadrature on the FP64 tensor pipe (same FP64 units, DMMA.8x8x4 issue)
    peak_dfma = sol.dfma_peak() if rank == 0 else None
    peak_dmma = sol.info(_lib.INFO_DMMA_PEAK) if rank == 0 else None
    peak_meas = peak_dmma if dmma else peak_dfma

    # ---- value: device-resident omegas, CUDA events around every step, this rank's own batch (no collective)
    def step_dev():
        sol.disp_batch_dev(B, om_d.data_ptr(), D_d.data_ptr())

    for _ in range(args.warmup):
        step_dev()
    sol.sync()
    l0 = sol.info(_lib.INFO_LAUNCHES)
    clocks = ClockSampler(local_rank) if rank == 0 else None   # one nvidia-smi loop per job, on rank 0's GPU
    barrier()
    if clocks:
        clocks.start()
    tot_ms, kern_ms = 0.0, 0.0
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_dev()
        e1.record()
        e1.synchronize()
        sol.sync()
        tot_ms += e0.elapsed_time(e1)
        kern_ms += sol.info(_lib.INFO_LAST_KERNEL_MS)
    barrier()
    clk = clocks.stop() if clocks else None
    launches = int(sol.info(_lib.INFO_LAUNCHES) - l0)
    D_first = D_d.cpu().numpy().copy()

    # ---- e2e: the same batch through the host-buffer call
    D_h = np.zeros(B, dtype=np.complex128)
    for _ in range(min(args.warmup, 2)):
        D_h = sol.disp_batch(om_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        D_h = sol.disp_batch(om_h)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    assert np.array_equal(D_h.view(np.float64), D_first), "device-resident and host-buffer paths disagree"
    assert np.all(np.isfinite(D_first)), "non-finite D in the benchmark batch"
    D_chk64 = sol.disp_batch(om_h[:64])
    sol.close()

    # ---- fast_path: the k-hoisted map mode, its own solver (internal chunk = the library's choice), set_k (table
    # build) inside the timed region of every step
    fast = None
    solf = Solver(plasma, device=local_rank, nmax_force=w["nmax_force"])
    solf.set_stream(stream.cuda_stream)
    solf.set_mode(1)
    solf.set_k(w["kperp"], w["kpar"])
    BF = int(solf.info(_lib.INFO_BATCH))          # one internal chunk: the kernel time of a step is one launch
    if not args.no_fast:
        om_f = map_omegas(w, rank, world, BF)
        om_fd = torch.from_numpy(om_f.view(np.float64).copy()).cuda()
        D_fd = torch.zeros(2 * BF, dtype=torch.float64, device="cuda")
        for _ in range(max(1, min(args.warmup, 3))):
            solf.disp_batch_dev(BF, om_fd.data_ptr(), D_fd.data_ptr())
        solf.sync()
        barrier()
        fast_ms, fast_kern = 0.0, 0.0
        for _ in range(args.steps):
            flush.zero_()
            f0_, f1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0_.record()
            solf.set_k(w["kperp"], w["kpar"])          # rebuilds the k tables every step
            solf.disp_batch_dev(BF, om_fd.data_ptr(), D_fd.data_ptr())
            f1_.record()
            f1_.synchronize()
            solf.sync()
            fast_ms += f0_.elapsed_time(f1_)
            fast_kern += solf.info(_lib.INFO_LAST_KERNEL_MS)
        D_f = D_fd.cpu().numpy().view(np.complex128)
        # the first 64 omegas of the direct batch are map points too: compare the two formulations there
        solf.disp_batch_dev(64, om_d.data_ptr(), D_fd.data_ptr())
        solf.sync()
        rel = float(np.max(np.abs(D_fd.cpu().numpy().view(np.complex128)[:64] - D_chk64) / np.abs(D_chk64)))
        fast_ms, fast_kern = max_over_ranks([fast_ms, fast_kern])
        flops_fast = FLOPS_FAST_PER_ABSN_PAR * absn * (w["npar"] - 1.0)
        ach = flops_fast * BF * args.steps / (fast_kern * 1e-3) / 1e12
        fast = {"value": world * BF * args.steps / (fast_ms * 1e-3), "value_per_gpu": BF * args.steps / (fast_ms * 1e-3),
                "unit": "D/s", "omegas_per_step_per_gpu": BF, "ms_per_step": fast_ms / args.steps,
                "max_rel_diff_vs_direct": rel, "finite": bool(np.all(np.isfinite(D_f.view(np.float64)))),
                "roofline": {"bound": "fp64_fma", "kernel": "k_fast_tiled", "achieved": ach, "peak": peak_dfma,
                             "unit": "TFLOP/s", "frac": ach / peak_dfma if peak_dfma else None,
                             "frac_nominal": ach / FP64_NOMINAL_TFLOPS,
                             "peak_source": "DFMA micro-benchmark run in this job",
                             "flop_model": "118 flops per (|n|, ipar, omega): 54 DFMA + 10 DADD/DMUL of k_fast_tiled's "
                                           "non-resonant loop (64 FP64 instructions; fast_kernel.cu)",
                             "flops_per_D": flops_fast, "kernel_ms_per_step": fast_kern / args.steps,
                             "kernel_share_of_step": fast_kern / fast_ms,
                             "algorithmic_bytes_per_launch": 96.0 * absn * (w["npar"] - 1.0) + BF * 2 * absn * (32 + 96)},
                "note": "k-hoisted p_perp sums (GA, GB and their moment tables rebuilt by set_k inside the timed region), "
                        "O(nmax*npar) per omega; reported separately, never mixed into value/roofline"}

    # ---- strong_map: complete maps through the product's map_search; N > 1: collective over the library communicator
    strong = None
    if not args.no_map:
        if world > 1:
            solf.comm_init_torch()
        solf.set_map_mode(0)          # the formulation of each map below is chosen explicitly with set_mode
        margs = (w["omr"][0], w["omr"][1], w["omi"][0], w["omi"][1])
        strong = {"parallelism": ("alps_b200_map_search collective over %d ranks: OMEGA partition, one ncclAllGather of D "
                                  "inside alps_b200_disp_batch, every rank finishes the map" % world) if world > 1
                  else "alps_b200_map_search on one GPU",
                  "wall_time_includes": "omega grid, H2D / D2H, the gather, NaN / infinity sentinels, log10|D|, find_minima"}
        sub = (w["nr"], w["ni"]) if args.full_map else w["sub"]
        for mode, key, (nr, ni) in ((1, "hoisted_full_map", (w["nr"], w["ni"])), (0, "direct_map", sub)):
            solf.set_mode(mode)
            solf.set_k(w["kperp"], w["kpar"])
            solf.map_search(*margs, 16, 9 * world)           # warm-up: allocations, communicator
            barrier()
            t0 = time.perf_counter()
            if mode == 1:
                solf.set_k(w["kperp"], w["kpar"])             # the per-k table build belongs to the map
            om, val, cal, roots = solf.map_search(*margs, nr, ni, numroots=1000)
            torch.cuda.synchronize()
            dt, = max_over_ranks([time.perf_counter() - t0])
            strong[key] = {"nr": nr, "ni": ni, "seconds": dt, "D_per_s": nr * ni / dt, "minima_found": len(roots),
                           "finite": bool(np.all(np.isfinite(val))), "checksum_log10absD": float(np.sum(val))}
        strong["note"] = ("direct_map is the %dx%d map%s; the same map on 1, 2, 4, 8 GPUs = strong scaling"
                          % (sub[0], sub[1], "" if args.full_map else " (sub-grid of the 512x512 map with the same "
                             "omega range; --full-map runs all 262144 points, ~2 min on one GPU)"))
    solf.close()

    # ---- harmonic_shard (N > 1): C4 batches, one GPU vs the two partitions of the library communicator
    harm = None
    if world > 1 and not args.no_map:
        pl4 = tables.config_kpar_fast()
        s4 = Solver(pl4, device=local_rank, emulate_nproc=4)
        s4.set_stream(stream.cuda_stream)
        s4.comm_init_torch()
        rng = np.random.default_rng(5)
        harm = {"config": "C4 tests/test_kperp.in at k_perp=3, k_par=1e-3 (120x240, nmax 88/29)",
                "unit": "us per host-buffer call, max over ranks",
                "limiter": "one D is ~50 us of work for a whole GPU: the partial sums of 88+29 harmonics are latency, "
                           "the 768 B/species/omega exchange is latency; the OMEGA partition needs one collective per "
                           "call, the HARMONIC one keeps plan / resonant / assemble replicated on every rank"}

        def timed(fn, reps):
            for _ in range(3):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            e1.synchronize()
            return max_over_ranks([(time.perf_counter() - t0) / reps * 1e6, e0.elapsed_time(e1) / reps * 1e3])

        for n in (1, 64, 1024):
            om = 9.9e-4 * (1.0 + 0.05 * rng.uniform(-1, 1, n)) + 1j * 2e-5 * rng.uniform(-1, 1, n)
            omd = torch.from_numpy(om.view(np.float64).copy()).cuda()
            Dd = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
            reps = 200 if n <= 64 else 30
            s4.set_partition(_lib.PARTITION_OMEGA)
            s4.set_k(3.0, 1e-3)

            def one():
                s4.disp_batch_dev(n, omd.data_ptr(), Dd.data_ptr())
                s4.sync()
            t_one = timed(one, reps)
            Dref = Dd.cpu().numpy().view(np.complex128).copy()
            t_om = timed(lambda: s4.disp_batch(om), reps)
            s4.set_partition(_lib.PARTITION_HARMONIC)
            s4.set_k(3.0, 1e-3)
            t_h = timed(lambda: s4.disp_batch(om), reps)
            Dh = s4.disp_batch(om)
            harm["n%d" % n] = {"one_gpu_wall": t_one[0], "one_gpu_cuda_events": t_one[1],
                               "omega_partition_wall": t_om[0], "omega_partition_cuda_events": t_om[1],
                               "harmonic_partition_wall": t_h[0], "harmonic_partition_cuda_events": t_h[1],
                               "harmonic_max_rel_diff": float(np.max(np.abs(Dh - Dref) / np.abs(Dref)))}
        s4.close()

    tot_ms, e2e_ms, kern_ms = max_over_ranks([tot_ms, e2e_s * 1e3, kern_ms])
    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        extra = extra_configs()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    n_total = world * B * args.steps
    value = n_total / (tot_ms * 1e-3)
    e2e = n_total / (e2e_ms * 1e-3)
    achieved = flops_exec_per_D * B * args.steps / (kern_ms * 1e-3) / 1e12   # one GPU's kernel
    achieved34 = flops_per_D * B * args.steps / (kern_ms * 1e-3) / 1e12
    cpu = None if args.no_cpu else cpu_leg(args)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, TRAFFIC_PROFILE)
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_launch"] / tj["omegas_per_launch"] * B   # scaled to this launch size
            traffic_src = ("NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one "
                           "`ncu --set full` capture of this kernel's regular-tile launch, 36 of the 39 tiles of a step "
                           "(%s); the packed remainder launch is not in it" % TRAFFIC_PROFILE)
        except Exception:
            traffic = None
    # one read of A', C' (fragment order, padded) and W per launch + plan (32 B) and moment sums (96 B) per item
    alg_bytes = float(sum(2 * 8 * (w["nperp"] - 1) * (w["npar"] - 1) + 8 * (w["nperp"] - 1) * 3 * (int(n) + 1)
                          for n in nmax) + B * sum(2 * (int(n) + 1) for n in nmax) * (32 + 96))
    line = {"metric": "D(omega,k) evals/sec", "value": value, "unit": "D/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": w["desc"], "omegas_per_step_per_gpu": B,
                       "nmax": [int(n) for n in nmax], "point_harmonics_per_D": ph,
                       "flops_per_D_survey_model": flops_per_D, "flops_per_D_executed": flops_exec_per_D,
                       "parallelism": "omega-shard x%d (weak: every rank its own batch, no collective)" % world,
                       "l2": "flushed between steps (256 MiB write)", "mode": "direct quadrature"},
            "e2e": {"value": e2e, "unit": "D/s", "h2d_bytes_per_step": 16 * B, "d2h_bytes_per_step": 16 * B},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor" if dmma else "fp64_fma",
                         "pipe": ("FP64 tensor pipe (DMMA.8x8x4: the 64 FMA/clk/SM FP64 units)" if dmma
                                  else "FP64 FMA pipe (DFMA)"),
                         "achieved": achieved, "peak": peak_meas, "unit": "TFLOP/s",
                         "frac": achieved / peak_meas if peak_meas else None, "traffic": traffic,
                         "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes,
                         "peak_source": ("DMMA (mma.sync.m8n8k4.f64)" if dmma else "DFMA") +
                                        " micro-benchmark run in this job (MEASURED_PEAKS.json has no FP64 entry)",
                         "peak_dfma_microbench": peak_dfma, "peak_dmma_microbench": peak_dmma,
                         "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_nominal": achieved / FP64_NOMINAL_TFLOPS,
                         "flop_model": "useful FP64 flops of the quadrature kernel's formulation: 12 per (|n|, iperp, ipar); "
                                       "+n and -n share the p_perp sums; padding/epilogue not counted (DESIGN.md)",
                         "achieved_survey_34flop_model": achieved34,
                         "frac_survey_34flop_model": achieved34 / peak_meas if peak_meas else None,
                         "kernel": "k_quad_mma" if dmma else "k_quad", "quad_variant": variant,
                         "kernel_ms_per_step": kern_ms / args.steps,
                         "kernel_share_of_step": kern_ms / tot_ms},
            "cpu_baseline": cpu, "clocks": clk, "fast_path": fast, "strong_map": strong, "harmonic_shard": harm,
            "extra": extra}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-fast", action="store_true", help="skip the k-hoisted fast-path leg")
    ap.add_argument("--no-map", action="store_true", help="skip the strong_map / harmonic_shard legs")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1..C4 rows")
    ap.add_argument("--full-map", action="store_true", help="strong_map: the complete 512x512 map in direct mode too")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
