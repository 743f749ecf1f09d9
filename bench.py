#!/usr/bin/env python
"""bench.py -- D(omega,k) evaluations per second of the disp() hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c5|c1]

A "step" is one pass of the hot path over one batch of omegas taken from the complex-omega map of
the workload (per GPU; N GPUs shard the map with no communication: weak scaling).
  value      device-resident omegas in, D out (alps_b200_disp_batch_dev), CUDA-event timed
  e2e        the same batch through the host-buffer call alps_b200_disp_batch (H2D of the omegas
             and D2H of D inside the timed region)
  roofline   the quadrature kernel (k_quad_mma: DMMA.8x8x4 on the FP64 units) against the DMMA micro-benchmark of
             the same job: useful flops 12 per (|n|, iperp, ipar) (the survey's 34-flop figure is
             reported beside it) / CUDA-event time of that kernel
  cpu_baseline  the CPU oracle (restated reference, OpenMP) on a bounded sample of the same workload
  clocks     nvidia-smi every 50 ms during the timed region: median / minimum SM clock over the busy samples, power
`--impl reference` times the restated reference (oracle/) alone on the host cores: the Fortran/MPI
reference cannot be built in this image (no gfortran, no MPI).
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
FP64_NOMINAL_TFLOPS = 37.2          # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz

WORKLOADS = {
    # name: (description, builder kwargs)
    "c5": dict(desc="C5 synthetic 3-species bi-kappa f0, 1024x2048 (p_perp,p_par) grid, nmax=200 forced, "
                    "k=(15.5,1e-2), omegas from the 512x512 map om_r in [0.05,3.05] x gamma in [-0.05,0.05]",
               nperp=1024, npar=2048, nmax_force=200, kperp=15.5, kpar=1.0e-2,
               omr=(0.05, 3.05), omi=(-0.05, 0.05), nr=512, ni=512, batch=296),
    "c5small": dict(desc="reduced C5 (256x512 grid, nmax=48) -- smoke runs only",
                    nperp=256, npar=512, nmax_force=48, kperp=5.0, kpar=1.0e-2,
                    omr=(0.05, 3.05), omi=(-0.05, 0.05), nr=512, ni=512, batch=296),
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


CPU_SAMPLE_TARGET_S = 12.0     # CPU work per sample (the task asks for a bounded sample of about 10-30 s)


def cpu_sample(w, plasma, om, ncap=None, threads=0):
    """Bounded sample of the workload on the host cores with the CPU oracle: one D evaluation
    restricted to harmonics |n| <= ncap, scaled to all harmonics by the signed-harmonic count.  ncap is calibrated
    once per process on these cores (a |n| <= 8 evaluation is timed, then ncap is chosen for about
    CPU_SAMPLE_TARGET_S seconds per sample), so the run stays bounded on slow hosts too."""
    from oracle.oracle import Oracle
    cores = threads or (os.cpu_count() or 1)
    orc = Oracle(plasma, nproc=0, threads=cores, nmax_force=w["nmax_force"])
    nmax = orc.set_k(w["kperp"], w["kpar"])
    if ncap is None:
        ncap = getattr(cpu_sample, "_ncap", None)
    if ncap is None:
        c0 = int(min(8, min(nmax)))
        orc.set_ncap(c0)
        t0 = time.perf_counter()
        orc.disp(complex(om))
        t8 = max(time.perf_counter() - t0, 1e-3)
        ncap = int(((CPU_SAMPLE_TARGET_S / t8) * (2 * c0 + 1) - 1) / 2)
        ncap = cpu_sample._ncap = max(c0, min(ncap, int(min(nmax))))
    ncap = int(min(ncap, min(nmax)))
    orc.set_ncap(ncap)
    t0 = time.perf_counter()
    orc.disp(complex(om))
    dt = time.perf_counter() - t0
    frac = sum(2 * ncap + 1 for _ in nmax) / float(sum(2 * int(n) + 1 for n in nmax))
    if getattr(cpu_sample, "_ncap", None) == ncap and dt < 0.6 * CPU_SAMPLE_TARGET_S:
        # the small calibration run over-estimates the cost per harmonic (fixed work, load balance): grow the next sample
        cpu_sample._ncap = max(ncap, min(int((2 * ncap + 1) * CPU_SAMPLE_TARGET_S / dt - 1) // 2, int(min(nmax))))
    return {"seconds_sample": dt, "fraction": frac, "d_per_s": frac / dt, "cores": cores, "ncap": ncap,
            "nmax": [int(n) for n in nmax]}


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    plasma = build_plasma(w)
    om = map_omegas(w, 0, 1, w["batch"])
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_sample(w, plasma, om[(7 * i) % om.size])
        if i >= args.warmup:
            vals.append(r)
    dps = float(np.mean([v["d_per_s"] for v in vals]))
    ms = float(np.mean([v["seconds_sample"] for v in vals])) * 1e3
    sample = ("one D evaluation restricted to |n|<=%d (%.2f%% of the signed harmonics), scaled; restated "
              "reference (CPU oracle, OpenMP over harmonics); Fortran/MPI build impossible here"
              % (vals[-1]["ncap"], 100 * vals[-1]["fraction"]))
    if len({v["ncap"] for v in vals}) > 1:
        sample += "; |n| cap per timed step: %s" % [v["ncap"] for v in vals]
    line = {"impl": "reference", "metric": "D(omega,k) evals/sec", "value": dps, "unit": "D/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": w["desc"]},
            "cpu_baseline": {"value": dps, "unit": "D/s", "cores": vals[0]["cores"], "kind": "port",
                             "sample": sample},
            "e2e": {"value": dps, "unit": "D/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from alps_b200 import _lib
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
    absn_points = sum(int(n) + 1 for n in nmax) * (w["nperp"] - 1.0) * (w["npar"] - 1.0)
    flops_exec_per_D = FLOPS_EXECUTED_PER_ABSN_POINT * absn_points
    variant = int(sol.info(_lib.INFO_QUAD_VARIANT))
    dmma = variant >= 9     # quadrature on the FP64 tensor pipe (same FP64 units, DMMA.8x8x4 issue)
    peak_dfma = sol.dfma_peak() if rank == 0 else None
    peak_dmma = sol.info(_lib.INFO_DMMA_PEAK) if rank == 0 else None
    peak_meas = peak_dmma if dmma else peak_dfma

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

    # ---- end to end through the host-buffer call
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

    # ---- separately reported: the k-hoisted "map fast path" (alps_b200_set_mode(1)); its set_k
    # (table build) is inside the timed region, amortised over the omegas of the step
    fast = None
    if not args.no_fast:
        BF = 16 * B
        om_f = map_omegas(w, rank, world, BF)
        om_fd = torch.from_numpy(om_f.view(np.float64).copy()).cuda()
        D_fd = torch.zeros(2 * BF, dtype=torch.float64, device="cuda")
        sol.set_mode(1)
        sol.set_k(w["kperp"], w["kpar"])
        sol.disp_batch_dev(BF, om_fd.data_ptr(), D_fd.data_ptr())
        sol.sync()
        barrier()
        f0_, f1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0_.record()
        for _ in range(args.steps):
            sol.set_k(w["kperp"], w["kpar"])          # rebuilds the k tables every step
            sol.disp_batch_dev(BF, om_fd.data_ptr(), D_fd.data_ptr())
        f1_.record()
        f1_.synchronize()
        sol.sync()
        fast_ms = f0_.elapsed_time(f1_)
        # same omegas as the direct batch are a subset: compare the first B of a direct run
        D_f = D_fd.cpu().numpy().view(np.complex128)
        sol.set_mode(0)
        sol.set_k(w["kperp"], w["kpar"])
        D_chk = sol.disp_batch(om_f[:64])
        rel = float(np.max(np.abs(D_f[:64] - D_chk) / np.abs(D_chk)))
        fast = {"value_per_gpu": BF * args.steps / (fast_ms * 1e-3), "unit": "D/s", "omegas_per_step": BF,
                "ms_per_step": fast_ms / args.steps, "max_rel_diff_vs_direct": rel,
                "note": "k-hoisted p_perp sums (GA, GB tables rebuilt by set_k inside the timed region), "
                        "O(nmax*npar) per omega; reported separately, never mixed into value/roofline"}

    times = torch.tensor([tot_ms, e2e_s * 1e3, kern_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    tot_ms, e2e_ms, kern_ms = [float(x) for x in times.cpu()]
    if rank == 0:
        n_total = world * B * args.steps
        value = n_total / (tot_ms * 1e-3)
        e2e = n_total / (e2e_ms * 1e-3)
        achieved = flops_exec_per_D * B * args.steps / (kern_ms * 1e-3) / 1e12   # one GPU's kernel
        achieved34 = flops_per_D * B * args.steps / (kern_ms * 1e-3) / 1e12
        cpu = None
        if world == 1 and not args.no_cpu:
            # the CPU leg runs in a fresh interpreter (no torch / CUDA threads competing with OpenMP)
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload",
                                      args.workload, "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                                     timeout=900, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
                cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception as e:     # the baseline is reported, never required
                cpu = {"value": None, "unit": "D/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_k_quad_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj["dram_bytes_per_launch"] / tj["omegas_per_launch"] * B   # scaled to this launch size
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
                           "parallelism": "omega-shard x%d" % world,
                           "l2": "flushed between steps (256 MiB write)", "mode": "direct quadrature"},
                "e2e": {"value": e2e, "unit": "D/s", "h2d_bytes_per_step": 16 * B, "d2h_bytes_per_step": 16 * B},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor" if dmma else "fp64_fma",
                             "pipe": ("FP64 tensor pipe (DMMA.8x8x4: the 64 FMA/clk/SM FP64 units)" if dmma
                                      else "FP64 FMA pipe (DFMA)"),
                             "achieved": achieved, "peak": peak_meas, "unit": "TFLOP/s",
                             "frac": achieved / peak_meas if peak_meas else None, "traffic": traffic,
                             "algorithmic_bytes": alg_bytes,
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
                "cpu_baseline": cpu, "clocks": clk, "fast_path": fast}
        print(json.dumps(line), flush=True)
    sol.close()
    if world > 1:
        dist.destroy_process_group()


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
