"""Twin of the reference's main program (src/ALPS.f90:19-147) for the disp() path:

    python -m alps_b200.run path/to/<runname>.in [--dist path/to/<arrayName>_dist.in] [--out solution]

reads the `.in` namelists (omitted keys: the defaults of src/ALPS_var.f90), obtains the f0 tables (`distribution/<arrayName>.<is>.array` if present,
else regenerated from the `_dist.in` closed forms), then map_search or refine_guess and the k scans,
writing `<out>/<runname>.map / .roots / .scan_* / .eigen_* / .heat_* / .heat_mech_*` in the
reference's formats.  Every D(omega,k) comes from the GPU (libalps_b200.so).

Analytic-continuation parameters: the twin of determine_param_fit (alps_b200/fits.py: Levenberg-Marquardt rows
started from the &ffit blocks, Chebyshev series for ac_method = 2) runs like in the reference whenever the tables
come from distribution/*.array files, or with --fit; tables regenerated from a _dist.in use the generator's ideal
parameters unless --fit is given.  NHDS calc_chi for use_bM species runs on the device
(csrc/nhds_kernel.cu).

Several GPUs of one box (replaces `mpirun -np N`): `python -m torch.distributed.run --nproc-per-node N
--master-addr 127.0.0.1 -m alps_b200.run x.in --emulate-nproc 4 ...` -- one process per GPU; the ranks join the library-owned NCCL communicator
(alps_b200_comm_init) and map_search is collective (OMEGA partition: a slice of the nr x ni grid per rank, one ncclAllGather
of D inside the library); rank 0 alone writes the files and runs the sequential root refinement and k scans.  Or one
process for all GPUs: `python -m alps_b200.run x.in --ngpu N` (device group, include/alps_b200.h)."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import tables
from .namelist import read_namelists
from .solver import Solver


def plasma_from_inputs(nl, dist_nl=None, base_dir=".", fit=None, rel_backend="host"):
    """fit: run the twin of determine_param_fit (the reference always does, src/ALPS.f90:114); None = only when the
    tables are read from distribution/*.array files (regenerated tables come with their ideal fit parameters).
    Omitted &system keys take the defaults of src/ALPS_var.f90.
    rel_backend: where the spline of the relativistic regrid is evaluated (relativistic.derivative_f0_rel):
    "device" in the twin main program, "host" (numpy statement of the same loop) for set-up checks without a GPU."""
    s = nl["system"]
    nspec, nperp, npar = int(s["nspec"]), int(s["nperp"]), int(s["npar"])
    vA = float(s["va"])
    name = s.get("arrayname", "")
    species, fits_in = [], []
    for i in range(1, nspec + 1):
        sp = nl["spec_%d" % i]
        nf = int(sp.get("ff", 1))
        ft, pc, par = [], [], []
        for j in range(1, nf + 1):
            f = nl.get("ffit_%d_%d" % (i, j), {})
            ft.append(int(f.get("fit_type_in", 1)))
            pc.append(float(f.get("perpcorr", 1.0)))
            par.append([float(f.get("fit_%d" % k, 0.0)) for k in range(1, 6)])
        po = nl.get("poly_spec_%d" % i, {})
        bm = nl.get("bm_spec_%d" % i, {})
        species.append(tables.Species(bM_nmaxs=int(bm.get("bm_nmaxs", 500)),
                                      bM_Bessel_zeros=float(bm.get("bm_bessel_zeros", 1.0e-50)),
                                      bM_betas=float(bm.get("bm_betas", 1.0)), bM_alphas=float(bm.get("bm_alphas", 1.0)),
                                      bM_pdrifts=float(bm.get("bm_pdrifts", 0.0)),
                                      ns=float(sp["nn"]), qs=float(sp["qq"]), ms=float(sp["mm"]),
                                      relativistic=bool(sp.get("relat", False)), usebM=bool(sp.get("use_bm", False)),
                                      ACmethod=int(sp.get("ac_method", 1)), fit_type=ft, perp_correction=pc,
                                      logfit=bool(sp.get("log_fit", True)), poly_kind=int(po.get("kind", 1)),
                                      poly_order=int(po.get("order", 0)) if int(sp.get("ac_method", 1)) == 2 else 0,
                                      poly_log_max=float(po.get("log_max", 18.0))))
        fits_in.append(par)
    pp = np.zeros((nspec, nperp + 1, npar + 1, 2), order="F")
    f0 = np.zeros((nspec, nperp + 1, npar + 1), order="F")
    maxfits = max(len(sp.fit_type) for sp in species)
    ngamma = int(s.get("ngamma", 100))      # src/ALPS_var.f90:128, 132
    pf = np.zeros((nspec, max(nperp, ngamma) + 1, 5, maxfits), order="F")
    have_files = all(os.path.exists(os.path.join(base_dir, "distribution", "%s.%d.array" % (name, i + 1)))
                     or species[i].usebM for i in range(nspec))
    if fit is None:
        fit = have_files
    if have_files and not fit:
        raise SystemExit("distribution/%s.<is>.array tables need the fit producers (the &ffit blocks only hold start "
                         "values; the reference always runs determine_param_fit): drop --no-fit" % name)
    if have_files:
        for i in range(nspec):
            if species[i].usebM:
                continue
            a = np.loadtxt(os.path.join(base_dir, "distribution", "%s.%d.array" % (name, i + 1)))
            a = a.reshape((nperp + 1, npar + 1, 3))
            pp[i, :, :, 0], pp[i, :, :, 1], f0[i] = a[:, :, 0], a[:, :, 1], a[:, :, 2]
            for j, par in enumerate(fits_in[i]):
                for k in range(5):
                    pf[i, :, k, j] = par[k]
    else:
        if dist_nl is None:
            raise SystemExit("no distribution/%s.<is>.array files and no --dist file" % name)
        ds = dist_nl["system"]
        specs = []
        for i in range(1, nspec + 1):
            d = dist_nl["spec_%d" % i]
            specs.append(tables.DistSpec(ms=float(d["ms_read"]), tau=float(d["taus"]), alph=float(d["alphs"]),
                                         drift=float(d["ps"]), kappa=float(d["kappas"]),
                                         distribution=int(d["distributions"]), autoscale=bool(d["autoscales"]),
                                         maxPperp=float(d["maxpperps"]), maxPpar=float(d["maxppars"])))
        pp, f0, fits = tables.generate_distribution(specs, nperp, npar, beta=float(ds["beta"]), vA=float(ds["va"]),
                                                    maxP=float(ds["maxp"]))
        for i in range(nspec):
            if species[i].usebM:       # read_f0 zeroes the tables of use_bM species (src/ALPS_io.f90:672-676)
                pp[i] = 0.0
                f0[i] = 0.0
                continue
            if fit:     # the reference's flow: fit types, perpcorr and start values come from the .in file
                continue
            species[i].fit_type = [fits[i]["fit_type"]] if not species[i].relativistic else species[i].fit_type
            species[i].perp_correction = [fits[i]["perpcorr"]]
            for k in range(5):
                pf[i, :, k, 0] = fits[i]["params"][k]
    pl = tables.Plasma(nperp=nperp, npar=npar, vA=vA, species=species, pp=pp, f0=f0, param_fit=pf,
                       ngamma=ngamma, npparbar=int(s.get("npparbar", 200)),
                       Bessel_zero=float(s.get("bessel_zero", 1.0e-45)), Tlim=float(s.get("tlim", 0.01)),
                       positions_principal=int(s.get("positions_principal", 5)),
                       n_resonance_interval=int(s.get("n_resonance_interval", 100)),
                       kperp_norm=bool(s.get("kperp_norm", True)))
    if any(sp.relativistic for sp in species):
        from .relativistic import derivative_f0_rel
        rel = [i for i, sp in enumerate(species) if sp.relativistic]
        shape = (len(rel), pl.ngamma + 1, pl.npparbar + 1)
        pl.f0_rel, pl.gamma_rel, pl.pparbar_rel = (np.zeros(shape, order="F") for _ in range(3))
        pl.df0_rel = np.zeros(shape + (2,), order="F")
        done = []      # species with identical tables and mass share one regrid (pair plasmas)
        for r, i in enumerate(rel):
            hit = [c for j, c in done if species[j].ms == species[i].ms and np.array_equal(pp[j], pp[i])
                   and np.array_equal(f0[j], f0[i])]
            if not hit:
                done.append((i, derivative_f0_rel(pp[i], f0[i], species[i].ms, vA, pl.ngamma, pl.npparbar,
                                                  backend=rel_backend)))
                hit = [done[-1][1]]
            g, p, f, d, integ = hit[0]
            pl.gamma_rel[r], pl.pparbar_rel[r], pl.f0_rel[r], pl.df0_rel[r] = g, p, f, d
            if not have_files and not fit:
                pf[i, :, 0, 0] = pf[i, 0, 0, 0] / integ
    if fit:
        from .fits import FitOptions, determine_param_fit
        opt = FitOptions(maxsteps_fit=int(s.get("maxsteps_fit", 500)),
                         lambda_initial_fit=float(s.get("lambda_initial_fit", 1.0)),
                         lambdafac_fit=float(s.get("lambdafac_fit", 10.0)), epsilon_fit=float(s.get("epsilon_fit", 1.0e-8)))
        initial = np.zeros((nspec, 5, maxfits))
        for i in range(nspec):
            for j, par in enumerate(fits_in[i]):
                initial[i, :, j] = par
        rel_in = None
        if any(sp.relativistic for sp in species):
            rel_in = {i: (pl.f0_rel[r], pl.gamma_rel[r], pl.pparbar_rel[r])
                      for r, i in enumerate(k for k, sp in enumerate(species) if sp.relativistic)}
        pl.param_fit, poly, pl.fit_quality = determine_param_fit(pl, initial, opt, rel_in)
        pl.fit_ran = True
        if poly is not None:
            pl.poly_fit_coeffs = poly
    return pl


def main(argv=None):
    ap = argparse.ArgumentParser(prog="alps_b200.run")
    ap.add_argument("input")
    ap.add_argument("--dist", default=None)
    ap.add_argument("--out", default="solution")
    ap.add_argument("--emulate-nproc", "--nproc", dest="nproc", type=int, default=0,
                    help="MPI size of the reference run to emulate (under torchrun spell it --emulate-nproc: "
                         "torchrun's own parser claims --nproc as an abbreviation of --nproc-per-node)")
    ap.add_argument("--map-mode", choices=["direct", "hoisted"], default="hoisted",
                    help="map_search only: 'hoisted' (default) evaluates the map with the k-hoisted p_perp sums "
                         "(alps_b200_set_map_mode(1): O(nmax*npar) per omega instead of O(nmax*nperp*npar), same D to "
                         "rounding, DESIGN.md 4b); the root refinement that follows always uses the direct quadrature")
    ap.add_argument("--ngpu", type=int, default=1,
                    help="single process: devices the library drives itself (alps_b200_cfg.ngpu, device group); under "
                         "torchrun every rank drives one GPU and the ranks join the library's NCCL communicator instead")
    ap.add_argument("--fit", dest="fit", action="store_true", default=None,
                    help="run the twin of determine_param_fit (LM / Chebyshev fits) like the reference always does "
                         "(src/ALPS.f90:114: determine_param_fit before the first disp).  Default whenever the f0 tables "
                         "are read from distribution/*.array files")
    ap.add_argument("--no-fit", dest="fit", action="store_false",
                    help="skip the fits: only allowed when the tables are regenerated from the _dist.in closed forms, whose "
                         "ideal fit parameters are then used; refused for file-based tables (their &ffit blocks hold "
                         "start values, not fitted parameters)")
    a = ap.parse_args(argv)
    nl = read_namelists(a.input)
    runname = os.path.splitext(os.path.basename(a.input))[0]
    dist_nl = read_namelists(a.dist) if a.dist else None
    pl = plasma_from_inputs(nl, dist_nl, base_dir=os.getcwd(), fit=a.fit, rel_backend="device")
    s = nl["system"]
    if getattr(pl, "fit_ran", False):
        # what output_fit prints (src/ALPS_analyt.f90:942-943)
        q = pl.fit_quality
        print(" Sum of all least-squares: %14.4E" % q)
        print(" Standard error of the estimate: %14.4E" % np.sqrt(q / (1.0 * pl.nspec * pl.nperp * pl.npar)))
    os.makedirs(a.out, exist_ok=True)
    prefix = os.path.join(a.out, runname)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sol = Solver(pl, emulate_nproc=a.nproc, device=local_rank if world > 1 else -1, ngpu=a.ngpu if world == 1 else 1)
    if world > 1:
        # one process per GPU (the reference's ranks): library-owned NCCL communicator, map_search is collective
        sol.comm_init_torch()
    try:
        kperp, kpar = float(s["kperp"]), float(s["kpar"])
        nmax = sol.set_k(kperp, kpar)
        print("nmax:", list(map(int, nmax)))
        opts = sol.opts(numiter=int(s.get("numiter", 50)), D_threshold=float(s.get("d_threshold", 1e-5)),
                        D_prec=float(s.get("d_prec", 1e-5)), D_tol=float(s.get("d_tol", 1e-7)),
                        D_gap=float(s.get("d_gap", 1e-5)), secant_method=int(s.get("secant_method", 2)))
        nroots = int(s.get("nroots", 1))
        if bool(s.get("use_map", False)):
            m = nl["maps_1"]
            sol.set_map_mode(1 if a.map_mode == "hoisted" else 0)
            om, val, cal, roots = sol.map_search(float(m["omi"]), float(m["omf"]), float(m["gami"]), float(m["gamf"]),
                                                 int(m["nr"]), int(m["ni"]), bool(m.get("loggridw", False)),
                                                 bool(m.get("loggridg", False)),
                                                 bool(s.get("determine_minima", True)),
                                                 map_path=prefix + ".map" if rank == 0 else None)
            guesses = roots[:min(nroots, len(roots))] if bool(s.get("determine_minima", True)) else []
        else:
            guesses = [complex(float(nl["guess_%d" % i]["g_om"]), float(nl["guess_%d" % i]["g_gam"]))
                       for i in range(1, nroots + 1)]
        if world > 1:
            sol.comm_finalize()     # nothing collective from here on
        if rank != 0:      # the root refinement and the k scans are sequential: rank 0 alone
            return 0
        w, D = sol.refine_guess(guesses, opts, roots_path=prefix + ".roots") if guesses else (np.zeros(0, complex), None)
        for r, d in zip(w, D if D is not None else []):
            print("root %s  D=%s" % (r, d))
        if int(s.get("n_scan", 0)) > 0 and int(s.get("scan_option", 1)) == 1 and len(w):
            for ik in range(1, int(s["n_scan"]) + 1):
                sc = nl["scan_input_%d" % ik]
                rows, w = sol.om_scan(w, opts, int(sc["scan_type"]), float(sc["swi"]), float(sc["swf"]),
                                      bool(sc["swlog"]), int(sc["ns"]), int(sc.get("nres", 1)),
                                      bool(sc.get("eigen", False)), bool(sc.get("heating", False)), prefix, ik)
                print("scan %d done: k=(%g,%g)" % (ik, sol.kperp, sol.kpar))
        elif int(s.get("n_scan", 0)) == 2 and int(s.get("scan_option", 1)) == 2 and len(w):
            def blk(ik):
                sc = nl["scan_input_%d" % ik]
                return dict(scan_type=sc["scan_type"], swi=sc["swi"], swf=sc["swf"], swlog=sc["swlog"], ns=sc["ns"],
                            nres=sc.get("nres", 1), eigen=sc.get("eigen", False), heat=sc.get("heating", False))
            rows, w = sol.om_double_scan(w, opts, blk(1), blk(2), prefix)
            print("double scan done: %d x %d points" % rows.shape[:2])
    finally:
        from . import _lib
        try:
            stats = (int(sol.info(_lib.INFO_D_EVALS)), int(sol.info(_lib.INFO_SET_K_CALLS)),
                     int(sol.info(_lib.INFO_MEMO_HITS)), int(sol.info(_lib.INFO_PREFETCHED)))
            print("D(omega,k) evaluations: %d (%d of them batched ahead of their request), set_k calls: %d, "
                  "disp() calls answered from the memo: %d" % (stats[0], stats[3], stats[1], stats[2]))
            main.last_stats = stats
        except Exception:
            pass
        sol.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
