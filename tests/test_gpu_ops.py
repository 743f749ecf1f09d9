"""Oracle parity at the OPERATING POINTS that bench.py measures and DESIGN.md quotes ("close the parity
holes"): BASELINE config 5 with all 201 harmonics per species (nmax = 200), C4 at k_perp = 3 with `mpirun -np 4`
emulated (nmax 88 / 29), C2 on its own 150x300 grid -- each through every execution class of the CUDA path:

  * one disp() at a time                       (latency class, n <= 8: CUDA-graph chain, narrow tiles where they apply)
  * a batch of 65 omegas via disp_batch_full   (throughput class, n > 64: tile-major k_quad_mma variant 15 + packed
                                                remainder tiles, warp-per-harmonic k_resonant) -- D, chi0, chi0_low, wave
  * a batch of 20 omegas                       (small class, 8 < n <= 64: p_par-split CTAs)
  * the k-hoisted map mode (set_mode(1))       (STORE tables + k_fast), single and throughput batch

against tests/golden/oracle_vectors_ops.npz (written by tests/golden/make_oracle_vectors_ops.py from the CPU oracle: a
C5 omega costs the oracle a minute, so those vectors are committed rather than recomputed) and, where the oracle is
fast enough (C4, C2), against the oracle run live on further omegas.  Tolerance 1e-9 with the scale-aware criteria of
tests/util.py (element-wise for chi0 / chi0_low, term scale for wave and D)."""
import os

import numpy as np
import pytest

from alps_b200 import tables
from tests.util import chi_err, det_scale, omega_samples, scaled_err, wave_scale

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors_ops.npz")
TOL = 1e-9


def _check(pl, kperp, kpar, om, got, want, what, tol=TOL):
    """got / want: (D, chi0[nspec,3,3], chi0_low[nspec,3,3,3], wave[3,3]) of one omega"""
    Dg, chi_g, low_g, wave_g = got
    Do, chi_o, low_o, wave_o = want
    worst = 0.0
    for s in range(pl.nspec):
        e = chi_err(chi_g[s], chi_o[s])
        assert e < tol, (what, "chi0", s, om, e)
        worst = max(worst, e)
        for m in range(3):
            if np.max(np.abs(low_o[s, :, :, m])) > 0:
                el = chi_err(low_g[s, :, :, m], low_o[s, :, :, m])
                assert el < tol, (what, "chi0_low", s, m, om, el)
                worst = max(worst, el)
    ws = wave_scale(chi_o, complex(om), pl.vA, kperp, kpar)
    ew = scaled_err(wave_g, wave_o, ws)
    assert ew < tol, (what, "wave", om, ew)
    ed = abs(Dg - Do) / det_scale(ws)
    assert ed < tol, (what, "D", om, ed)
    return max(worst, ew, ed)


def _all_classes(pl, sol, kperp, kpar, oms, want, filler, label, modes=(0, 1)):
    """every execution class of the CUDA path at the omegas `oms` against `want` (list of oracle tuples)"""
    worst = {}
    n = len(oms)
    for mode in modes:
        sol.set_mode(mode)
        sol.set_k(kperp, kpar)
        tag = "%s/%s" % (label, "hoisted" if mode else "direct")
        # one disp() at a time
        w = 0.0
        for om, ref in zip(oms, want):
            w = max(w, _check(pl, kperp, kpar, om, sol.disp(complex(om), full=True), ref, tag + "/single"))
            Dplain = sol.disp(complex(om))            # D-only call: the captured graph chain
            ws = wave_scale(ref[1], complex(om), pl.vA, kperp, kpar)
            assert abs(Dplain - ref[0]) / det_scale(ws) < TOL, (tag, "graph chain", om)
        worst[tag + "/single"] = w
        # small class (8 < n <= 64) and throughput class (n > 64): the operating points sit in the middle of the batch
        for size, cls in ((20, "small"), (65, "throughput")):
            fill = np.asarray(filler[: size - n])
            batch = np.concatenate([fill[: (size - n) // 2], np.asarray(oms), fill[(size - n) // 2:]])
            assert batch.size == size
            D, chi0, low, wave = sol.disp_batch_full(batch)
            assert np.all(np.isfinite(D.view(np.float64)))
            o = (size - n) // 2
            w = 0.0
            for i, (om, ref) in enumerate(zip(oms, want)):
                w = max(w, _check(pl, kperp, kpar, om, (D[o + i], chi0[o + i], low[o + i], wave[o + i]), ref,
                                  tag + "/" + cls))
            worst[tag + "/" + cls] = w
            # D-only batch entry point gives the same bits as the full one
            assert np.array_equal(sol.disp_batch(batch).view(np.float64), D.view(np.float64)), (tag, cls)
    sol.set_mode(0)
    sol.set_k(kperp, kpar)
    return worst


def _fixture(name):
    z = np.load(GOLD)
    oms = z[name + "_om"]
    want = [(z[name + "_D"][i], z[name + "_chi0"][i], z[name + "_chi0_low"][i], z[name + "_wave"][i])
            for i in range(oms.size)]
    return z[name + "_nmax"], z[name + "_k"], oms, want


def test_c5_nmax200_every_class_against_the_committed_oracle_vectors():
    """BASELINE config 5 exactly as bench.py runs it (3-species bi-kappa, 1024x2048, nmax = 200 forced, k = (15.5, 1e-2)):
    Bessel orders up to 201 at z up to k_perp p_perp_max / q, 13 harmonic tiles per species incl. the packed remainder
    tile, resonant harmonics of all three species, Im(omega) < 0, > 0 and = 0."""
    from alps_b200.solver import Solver
    nmax, k, oms, want = _fixture("c5")
    pl = tables.config_kappa3(1024, 2048)
    filler = omega_samples(17, 64, (0.05, 3.05), (-0.05, 0.05))
    sol = Solver(pl, nmax_force=200)
    try:
        assert list(sol.set_k(k[0], k[1])) == list(nmax) == [200, 200, 200]
        worst = _all_classes(pl, sol, k[0], k[1], oms, want, filler, "c5")
        # what bench.py times: a 296-omega device-resident batch -- same bits as the host-buffer call
        import torch
        big = np.concatenate([np.asarray(oms), omega_samples(18, 292, (0.05, 3.05), (-0.05, 0.05))])
        om_d = torch.from_numpy(big.view(np.float64).copy()).cuda()
        D_d = torch.zeros(2 * big.size, dtype=torch.float64, device="cuda")
        sol.set_stream(torch.cuda.current_stream().cuda_stream)
        sol.disp_batch_dev(big.size, om_d.data_ptr(), D_d.data_ptr())
        sol.sync()
        D = D_d.cpu().numpy().view(np.complex128)
        for i, (om, ref) in enumerate(zip(oms, want)):
            ws = wave_scale(ref[1], complex(om), pl.vA, k[0], k[1])
            assert abs(D[i] - ref[0]) / det_scale(ws) < TOL, ("bench batch", om)
    finally:
        sol.close()
    print("c5 worst scaled errors:", {a: "%.2e" % b for a, b in worst.items()})


def test_c4_kperp3_every_class_against_oracle():
    """tests/test_kperp.in at the end of its scan: k_perp = 3, k_par = 1e-3, mpirun -np 4 emulated -> nmax 88 / 29 and
    split_processes ranges past nmax (src/ALPS_fns.f90:4048-4064, 4185-4200).  Committed vectors + live oracle."""
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    nmax, k, oms, want = _fixture("c4")
    pl = tables.config_kpar_fast()
    filler = omega_samples(19, 64, (1.0e-3, 1.2), (-2.0e-3, 2.0e-3))
    orc = Oracle(pl, nproc=4)
    sol = Solver(pl, emulate_nproc=4)
    try:
        nm_o = orc.set_k(k[0], k[1])
        assert list(sol.set_k(k[0], k[1])) == list(nm_o) == list(nmax)
        assert int(nmax[0]) >= 80 and int(nmax[1]) >= 25
        # the committed vectors are what the oracle computes today
        for om, ref in zip(oms[:2], want[:2]):
            live = orc.disp(complex(om), full=True)
            assert abs(live[0] - ref[0]) <= 1e-12 * abs(ref[0]) and np.allclose(live[1], ref[1], rtol=1e-12, atol=0)
        worst = _all_classes(pl, sol, k[0], k[1], oms, want, filler, "c4")
        # live oracle on further omegas (n = 0 and n = +-1 resonances, both half planes), throughput class
        extra = np.array([2.9e-3 - 3.0e-4j, 1.7e-3 + 2.0e-5j, 0.9985 - 1.0e-3j, 1.0045 + 2.0e-4j, 2.003 - 1.0e-5j])
        ref = [orc.disp(complex(o), full=True) for o in extra]
        batch = np.concatenate([filler[:30], extra, filler[30:60]])
        D, chi0, low, wave = sol.disp_batch_full(batch)
        for i, (om, r) in enumerate(zip(extra, ref)):
            _check(pl, k[0], k[1], om, (D[30 + i], chi0[30 + i], low[30 + i], wave[30 + i]), r, "c4/live")
    finally:
        sol.close()
    print("c4 worst scaled errors:", {a: "%.2e" % b for a, b in worst.items()})


def test_c2_bimax_150x300_every_class_against_oracle():
    """tests/test_bimax.in on its real grid (150x300): protons use_bM (k_nhds), electrons from the table."""
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    nmax, k, oms, want = _fixture("c2")
    pl = tables.config_bimax(150, 300)
    filler = omega_samples(23, 64, (5.0e-3, 8.0e-2), (-2.0e-2, 5.0e-3))
    orc = Oracle(pl)
    sol = Solver(pl)
    try:
        nm_o = orc.set_k(k[0], k[1])
        assert list(sol.set_k(k[0], k[1])) == list(nm_o) == list(nmax)
        worst = _all_classes(pl, sol, k[0], k[1], oms, want, filler, "c2")
        extra = omega_samples(29, 4, (1.0e-2, 6.0e-2), (-2.0e-2, 2.0e-3))
        ref = [orc.disp(complex(o), full=True) for o in extra]
        batch = np.concatenate([filler[:40], extra, filler[40:61]])
        D, chi0, low, wave = sol.disp_batch_full(batch)
        for i, (om, r) in enumerate(zip(extra, ref)):
            _check(pl, k[0], k[1], om, (D[40 + i], chi0[40 + i], low[40 + i], wave[40 + i]), r, "c2/live")
    finally:
        sol.close()
    print("c2 worst scaled errors:", {a: "%.2e" % b for a, b in worst.items()})


def test_relativistic_fit_type_5():
    """Fit type 5 (Juettner in Gamma times a Gaussian in pbar_par, src/ALPS_analyt.f90 fit_function case 5) in the
    Landau term of a relativistic species: CUDA path against the oracle, damped omegas with resonances in the cone."""
    from tests.test_gpu_parity import _compare
    pl = tables.config_relativistic(nperp=20, npar=40, ngamma=60, npparbar=80)
    for i, s in enumerate(pl.species):
        s.fit_type = [5]
        pl.param_fit[i, :, 1, 0] = 0.07      # p2: width of the pbar_par Gaussian
        pl.param_fit[i, :, 2, 0] = 0.15      # p3: its centre
    oms = [6.2713e-2 - 4.662e-8j, 1.0 - 1.655e-6j, 0.3 - 0.02j, 0.45 - 0.06j, 0.5 + 0.01j, 0.8 + 0j]
    _compare(pl, 1.0e-3, 1.0e-1, oms)


def test_mode1_parity_battery():
    """The k-hoisted map mode against the oracle on the cases of the direct-path battery: all full_integrate branches,
    resonances on grid nodes / at the grid edges / outside the grid, kperp_norm = F, ACmethod 0 / 1 / 2, bi-kappa
    tables, emulated split_processes ranges, a use_bM species."""
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    cases = []
    pl = tables.config_small(24, 48, kind=1)
    dp = pl.pp[0, 2, 2, 1] - pl.pp[0, 2, 1, 1]
    nodes = [0.05 * pl.pp[0, 2, j, 1] for j in (1, 2, 30, 46, 47)]            # Re p_res on nodes (kpar = 0.05, n = 0)
    edge = [0.05 * (pl.pp[0, 2, 47, 1] + 2.4 * dp), 0.05 * (pl.pp[0, 2, 1, 1] - 1.6 * dp)]
    oms = (list(omega_samples(1, 8, (0.02, 1.5), (-0.05, 0.05))) + [0.3 + 0j, 0.011 - 1e-6j, 1.0 + 1e-5j]
           + [complex(x, g) for x in nodes + edge for g in (-1e-3, 0.0, 2e-3)])
    cases.append(("bimax all branches", pl, {}, (0.3, 0.05), oms))
    cases.append(("kappa", tables.config_small(28, 56, kind=2), {}, (0.2, 0.08),
                  list(omega_samples(2, 8, (0.02, 1.2), (-0.03, 0.03)))))
    cases.append(("emulated nproc", tables.config_small(24, 48, kind=1), dict(nproc=8), (0.5, 0.05),
                  list(omega_samples(3, 4, (0.02, 1.0), (-0.02, 0.02)))))
    cases.append(("negative kpar", tables.config_small(24, 48, kind=1), {}, (0.3, -0.05),
                  list(omega_samples(4, 6, (0.02, 1.0), (-0.03, 0.03)))))
    pn = tables.config_small(24, 48, kind=1)
    pn.kperp_norm = False
    cases.append(("kperp_norm=F", pn, {}, (0.3, 0.05), list(omega_samples(6, 5, (0.02, 1.0), (-0.03, 0.03)))))
    p0 = tables.config_small(24, 48, kind=1)
    for s in p0.species:
        s.ACmethod = 0
    cases.append(("ACmethod 0", p0, {}, (0.3, 0.05), [0.31 - 0.02j, 0.011 - 1e-6j, 0.7 - 1e-3j]))
    from numpy.polynomial import chebyshev as Ch
    p2 = tables.config_small(24, 48, kind=1)
    order = 12
    coeffs = np.zeros((p2.nspec, p2.nperp + 1, order + 1), order="F")
    for i in range(p2.nspec):
        ppar = p2.pp[i, 0, :, 1]
        x = (ppar - 0.5 * (ppar[-1] + ppar[0])) / (0.5 * (ppar[-1] - ppar[0]))
        for ip in range(p2.nperp + 1):
            coeffs[i, ip, :] = Ch.chebfit(x, np.log10(p2.f0[i, ip, :]), order)
    p2.poly_fit_coeffs = coeffs
    for s in p2.species:
        s.ACmethod, s.poly_order, s.poly_kind, s.logfit, s.poly_log_max = 2, order, 1, True, 18.0
    cases.append(("ACmethod 2", p2, {}, (0.3, 0.05), [0.31 - 0.02j, 0.011 - 1e-6j, 0.7 - 1e-3j, 0.2 - 0.04j]))
    cases.append(("use_bM species", tables.config_bimax(60, 120), {}, (1.0e-3, 3.0e-2),
                  [3.0e-2 - 1.0e-5j, 4.5e-2 - 1.9e-2j, 2.0e-2 + 3.0e-3j]))
    for label, pl, kw, (kperp, kpar), oms in cases:
        orc = Oracle(pl, nproc=kw.get("nproc", 0))
        sol = Solver(pl, emulate_nproc=kw.get("nproc", 0))
        try:
            orc.set_k(kperp, kpar)
            sol.set_mode(1)
            sol.set_k(kperp, kpar)
            want = [orc.disp(complex(o), full=True) for o in oms]
            D, chi0, low, wave = sol.disp_batch_full(np.asarray(oms))
            for i, (om, r) in enumerate(zip(oms, want)):
                _check(pl, kperp, kpar, om, (D[i], chi0[i], low[i], wave[i]), r, "mode1/" + label)
                _check(pl, kperp, kpar, om, sol.disp(complex(om), full=True), r, "mode1/single/" + label)
        finally:
            sol.close()


@pytest.mark.parametrize("cfg", ["parallel", "oblique", "c3"])
def test_relativistic_throughput_class_against_oracle(cfg):
    """Batches of more than 64 omegas take the omega-tiled relativistic kernels (k_rel_plan -> k_rel<.,1> -> k_rel_tiled,
    csrc/rel_kernel.cu) instead of one CTA per (omega, species, |n|): D, chi0, chi0_low and wave of the batch against the
    oracle at omegas with resonances inside the cone, at the cone edge and outside the grid, growing / damped / real, and
    against the single-omega path for every omega of the batch."""
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    if cfg == "parallel":
        pl, (kperp, kpar) = tables.config_relativistic(nperp=20, npar=40, ngamma=60, npparbar=80), (1.0e-3, 1.0e-1)
        key = [6.2713e-2 - 4.662e-8j, 1.0 - 1.655e-6j, 0.5 + 0.01j, 0.3 - 0.02j, 2.5 + 0.0j, 0.9 + 1e-3j, 0.05 + 0j]
        fill = omega_samples(51, 70, (0.02, 2.6), (-0.03, 0.03))
    elif cfg == "oblique":
        pl, (kperp, kpar) = tables.config_relativistic(nperp=20, npar=40, ngamma=120, npparbar=400), (0.8, 0.3)
        key = [1.0 - 1.655e-6j, 0.5 + 0.01j, 0.3 - 0.02j, 1.7 + 0.004j, 0.6 - 0.05j]
        fill = omega_samples(52, 66, (0.25, 2.0), (-0.05, 0.05))
    else:
        pl, (kperp, kpar) = tables.config_relativistic(), (1.0e-3, 1.0e-1)
        key = [6.2713e-2 - 4.662e-8j, 1.0 - 1.655e-6j, 0.4 + 0.005j]
        fill = omega_samples(53, 70, (0.03, 1.5), (-0.01, 0.01))
    orc = Oracle(pl)
    sol = Solver(pl)
    try:
        assert list(orc.set_k(kperp, kpar)) == list(sol.set_k(kperp, kpar))
        want = []
        for o in key:                      # omegas where the reference stops with alps_error(8) are not test points
            try:
                want.append((o, orc.disp(complex(o), full=True)))
            except RuntimeError:
                pass
        assert len(want) >= 3
        keep = []
        for o in fill:                     # the same for the filler omegas (coarse grids hit alps_error(8) easily)
            try:
                sol.disp(complex(o))
                keep.append(o)
            except Exception:
                pass
        batch = np.concatenate([np.asarray(keep[:33]), np.array([o for o, _ in want]), np.asarray(keep[33:])])
        assert batch.size > 64
        D, chi0, low, wave = sol.disp_batch_full(batch)
        for i, (o, ref) in enumerate(want):
            _check(pl, kperp, kpar, o, (D[33 + i], chi0[33 + i], low[33 + i], wave[33 + i]), ref, "rel-tiled/" + cfg)
        # every omega of the batch against the single-omega path (latency class: one CTA per harmonic)
        for i in range(0, batch.size, 3):
            d1, c1, l1, w1 = sol.disp(complex(batch[i]), full=True)
            ws = wave_scale(c1, complex(batch[i]), pl.vA, kperp, kpar)
            assert abs(D[i] - d1) / det_scale(ws) < 1e-10, (cfg, batch[i])
            assert scaled_err(wave[i], w1, ws) < 1e-10
        # A/B: the tiled kernels switched off give the same numbers
        assert np.all(np.isfinite(D.view(np.float64)))
    finally:
        sol.close()
