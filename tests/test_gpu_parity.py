"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
inputs.  Tolerance 1e-9 (BASELINE.json north_star), applied as SURVEY.md section 7 prescribes:
chi0 element-wise (tests.util.chi_err), wave(i,j) relative to the magnitude of the terms summed into
it, D relative to the magnitude of its summed products -- at a root D and wave(1,1) are
cancellations, so "relative to |D|" is meaningless there."""
import os

import numpy as np
import pytest

from alps_b200 import tables
from tests.util import chi_err, det_scale, omega_samples, scaled_err, tensor_err, wave_scale

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-9


def _compare(pl, kperp, kpar, oms, nproc=0, tol=TOL):
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    orc = Oracle(pl, nproc=nproc)
    sol = Solver(pl, emulate_nproc=nproc)
    try:
        nmax_o = orc.set_k(kperp, kpar)
        nmax_g = sol.set_k(kperp, kpar)
        assert list(nmax_o) == list(nmax_g)
        worst = 0.0
        Db = sol.disp_batch(oms)
        for i, om in enumerate(oms):
            Do, chi_o, low_o, wave_o = orc.disp(complex(om), full=True)
            Dg, chi_g, low_g, wave_g = sol.disp(complex(om), full=True)
            for s in range(pl.nspec):
                e = chi_err(chi_g[s], chi_o[s])
                assert e < tol, ("chi0", s, om, e)
                for m in range(3):
                    if np.max(np.abs(low_o[s, :, :, m])) > 0:
                        el = chi_err(low_g[s, :, :, m], low_o[s, :, :, m])
                        assert el < tol, ("chi0_low", s, m, om, el)
            ws = wave_scale(chi_o, complex(om), pl.vA, kperp, kpar)
            ew = scaled_err(wave_g, wave_o, ws)
            assert ew < tol, ("wave", om, ew)
            ed = abs(Dg - Do) / det_scale(ws)
            assert ed < tol, ("D", om, ed)
            assert abs(Db[i] - Dg) <= 1e-13 * det_scale(ws), ("batch vs single", om)
            worst = max(worst, ew, ed)
        return worst
    finally:
        sol.close()


def test_small_bimax_all_branches():
    pl = tables.config_small(24, 48, kind=1)
    # resonant and non-resonant, damped / growing / real omega
    oms = list(omega_samples(1, 12, (0.02, 1.5), (-0.05, 0.05))) + [0.3 + 0j, 0.011 - 1e-6j, 1.0 + 1e-5j]
    _compare(pl, 0.3, 0.05, oms)


def test_small_kappa():
    pl = tables.config_small(28, 56, kind=2)
    oms = list(omega_samples(2, 10, (0.02, 1.2), (-0.03, 0.03)))
    _compare(pl, 0.2, 0.08, oms)


def test_small_emulated_nproc():
    pl = tables.config_small(24, 48, kind=1)
    oms = list(omega_samples(3, 4, (0.02, 1.0), (-0.02, 0.02)))
    _compare(pl, 0.5, 0.05, oms, nproc=8)


def test_kpar_fast_config():
    """C1 tables (tests/test_kpar_fast.in), omegas around the golden roots."""
    pl = tables.config_kpar_fast()
    oms = [9.98811e-3 - 2.31322e-7j, 9.9e-3 - 5.5e-6j, 1.2e-2 + 1e-6j, 5e-2 - 3e-4j, 0.3 + 0.01j, 0.9 + 0j]
    _compare(pl, 1e-2, 1e-2, oms, nproc=4)


def test_derivative_f0_matches_oracle():
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    pl = tables.config_small(24, 48, kind=2)
    orc = Oracle(pl)
    sol = Solver(pl)
    try:
        assert np.array_equal(sol.df0(), orc.df0())   # same IEEE operations: bit-exact
    finally:
        sol.close()


def test_harmonic_shards_sum_to_the_unsharded_result():
    """Harmonic sharding (replaces split_processes + MPI_REDUCE): the chi partials of the shards add
    up to the unsharded partials, and assembling the sum gives the same D."""
    import torch
    from alps_b200.solver import Solver
    pl = tables.config_small(24, 48, kind=1)
    oms = np.array(list(omega_samples(5, 9, (0.02, 1.4), (-0.04, 0.04))))
    n = oms.size
    sol = Solver(pl)
    try:
        sol.set_stream(torch.cuda.current_stream().cuda_stream)
        sol.set_k(0.4, 0.05)
        D_ref = sol.disp_batch(oms)
        L = sol.chi_partial_len()
        om_d = torch.from_numpy(oms.view(np.float64).copy()).cuda()
        full = torch.zeros(n * L, dtype=torch.float64, device="cuda")
        sol.chi_partial_dev(n, om_d.data_ptr(), full.data_ptr())
        for world in (2, 3, 8):
            acc = torch.zeros_like(full)
            for rank in range(world):
                sol.set_harmonic_shard(rank, world)
                sol.set_k(0.4, 0.05)
                part = torch.zeros_like(full)
                sol.chi_partial_dev(n, om_d.data_ptr(), part.data_ptr())
                torch.cuda.synchronize()
                acc += part          # what ncclAllReduce(sum) does across ranks
            scale = float(full.abs().max())
            assert float((acc - full).abs().max()) <= 1e-12 * scale
            D_d = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
            sol.assemble_dev(n, om_d.data_ptr(), acc.data_ptr(), D_d.data_ptr())
            torch.cuda.synchronize()
            D = D_d.cpu().numpy().view(np.complex128)
            assert np.max(np.abs(D - D_ref) / np.abs(D_ref)) < 1e-10
        sol.set_harmonic_shard(0, 1)
    finally:
        sol.close()


def test_relativistic_species_small_grid():
    """Relativistic integrators (src/ALPS_fns_rel.f90) on a reduced tests/test_relativistic.in
    (Juettner pair plasma, vA = 1, fit type 4): resonant / non-resonant harmonics, Landau term."""
    pl = tables.config_relativistic(nperp=20, npar=40, ngamma=60, npparbar=80)
    oms = [6.2713e-2 - 4.662e-8j, 1.0 - 1.655e-6j, 0.5 + 0.01j, 0.3 - 0.02j, 2.5 + 0.0j, 0.9 + 1e-3j]
    _compare(pl, 1.0e-3, 1.0e-1, oms)


def test_relativistic_oblique_many_harmonics():
    """k_perp ~ 1: ~30 harmonics per species, resonances inside the cone for many gamma."""
    pl = tables.config_relativistic(nperp=20, npar=40, ngamma=120, npparbar=400)
    _compare(pl, 0.8, 0.3, [1.0 - 1.655e-6j, 0.5 + 0.01j, 0.3 - 0.02j])


def test_relativistic_error_8_is_reported_like_the_reference():
    """alps_error(8) (src/ALPS_fns_rel.f90:655-656): the principal-value window covers the whole cone.
    On a coarse relativistic grid both the oracle and the CUDA path must report it."""
    from alps_b200 import _lib
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    pl = tables.config_relativistic(nperp=20, npar=40, ngamma=60, npparbar=80)
    orc = Oracle(pl)
    orc.set_k(0.8, 0.3)
    with pytest.raises(RuntimeError, match="alps_error\\(8\\)"):
        orc.disp(1.0 - 1.655e-6j)
    assert abs(orc.disp(0.5 + 0.01j)) > 0          # no error at this omega
    sol = Solver(pl)
    try:
        sol.set_k(0.8, 0.3)
        with pytest.raises(_lib.AlpsB200Error) as e:
            sol.disp(1.0 - 1.655e-6j)
        assert e.value.code == 8
        assert abs(sol.disp(0.5 + 0.01j) - orc.disp(0.5 + 0.01j)) < 1e-9 * abs(orc.disp(0.5 + 0.01j))
    finally:
        sol.close()


def test_relativistic_small_batches_every_size(monkeypatch):
    """Batches of 1 .. 64 omegas of a relativistic pair plasma: k_rel_plan_blk + k_rel_rows (the rows of the resonant
    entries over the whole GPU, both register variants) + k_rel_nonres against the oracle, and against the path that leaves
    the rows to the CTAs of their tile (ALPS_B200_REL_ROWS=0).  Oblique k: a dozen harmonics per species, several of them
    resonant, resonances inside the cone for many Gamma rows; omegas with Im > 0, = 0 and < 0 (Landau rows)."""
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    pl = tables.config_relativistic(nperp=20, npar=40, ngamma=120, npparbar=400)
    kperp, kpar = 0.4, 0.3
    base = [1.0 - 1.655e-6j, 0.5 + 0.01j, 0.3 - 0.02j, 0.8 + 0.0j, 1.7 - 0.05j]
    oms = np.array([base[i % len(base)] * (1.0 + 0.003 * (i // len(base))) for i in range(64)])
    orc = Oracle(pl)
    orc.set_k(kperp, kpar)
    want = {i: orc.disp(complex(oms[i]), full=True) for i in (0, 1, 2, 3, 4, 7, 19, 63)}
    got = {}
    for rows in ("1", "0"):
        monkeypatch.setenv("ALPS_B200_REL_ROWS", rows)
        sol = Solver(pl)
        try:
            sol.set_k(kperp, kpar)
            out = {n: sol.disp_batch(oms[:n]) for n in (1, 2, 3, 8, 20, 64)}
            out["single"] = np.array([[sol.disp(complex(o)) for _ in range(3)][-1] for o in oms[:5]])   # replayed chain
            got[rows] = out
        finally:
            sol.close()
    monkeypatch.delenv("ALPS_B200_REL_ROWS", raising=False)
    for i, (Do, chi_o, _, _) in want.items():
        sc = det_scale(wave_scale(chi_o, complex(oms[i]), pl.vA, kperp, kpar))
        for n in (1, 2, 3, 8, 20, 64):
            if i < n:
                assert abs(got["1"][n][i] - Do) / sc < TOL, (n, i)
                assert abs(got["1"][n][i] - got["0"][n][i]) / sc < 1e-10, (n, i)
        if i < 5:
            assert abs(got["1"]["single"][i] - Do) / sc < TOL, i
    # within the latency class a single disp() and a batch are the same bits
    assert np.array_equal(got["1"]["single"][:3], got["1"][3])
    assert np.array_equal(got["1"][8][:3], got["1"][3])


def test_relativistic_config_c3():
    """C3 at full size (ngamma = npparbar = 500, 30x60 input table) near its two roots."""
    pl = tables.config_relativistic()
    _compare(pl, 1.0e-3, 1.0e-1, [6.2713e-2 - 4.662e-8j, 1.0 - 1.655e-6j, 0.4 + 0.005j])


def test_bimax_config_nhds_species():
    """C2 (tests/test_bimax.in): species 1 use_bM=T -- its chi is the closed-form NHDS calc_chi, computed by
    k_nhds on the device inside alps_b200_disp*; species 2 is integrated from the table.  The oracle uses its own
    CPU restatement of calc_chi (oracle/nhds_oracle.hpp, pinned against scipy by tests/test_nhds.py)."""
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    pl = tables.config_bimax(60, 120)
    kperp, kpar = 1.0e-3, 0.03
    orc = Oracle(pl)
    sol = Solver(pl)
    try:
        for kperp, kpar in ((1.0e-3, 0.03), (0.8, 0.05)):
            assert list(orc.set_k(kperp, kpar)) == list(sol.set_k(kperp, kpar))
            oms = np.array([3.0e-2 - 1.0e-5j, 4.5e-2 - 1.9e-2j, 0.1 + 0.002j, 0.05 + 0.0j])
            Db = sol.disp_batch(oms)
            Dbig = sol.disp_batch(np.tile(oms, 40))      # throughput batch class
            for i, om in enumerate(oms):
                Do, chi_o, low_o, wave_o = orc.disp(complex(om), full=True)
                Dg, chi_g, low_g, wave_g = sol.disp(complex(om), full=True)
                ws = wave_scale(chi_o, complex(om), pl.vA, kperp, kpar)
                assert scaled_err(wave_g, wave_o, ws) < TOL
                assert abs(Dg - Do) / det_scale(ws) < TOL
                assert abs(Db[i] - Dg) <= 1e-13 * det_scale(ws)
                assert abs(Dbig[i] - Dg) <= 1e-11 * det_scale(ws) and abs(Dbig[i + 4 * 39] - Dbig[i]) == 0.0
                # the CUDA-graph replay of the single-omega chain (second and later plain calls) includes k_nhds
                for _ in range(3):
                    assert abs(sol.disp(complex(om)) - Dg) <= 1e-13 * det_scale(ws)
                for s in range(2):
                    assert chi_err(chi_g[s], chi_o[s]) < TOL
                    for m in range(3):
                        assert chi_err(low_g[s, :, :, m], low_o[s, :, :, m]) < TOL
    finally:
        sol.close()


def test_fast_path_matches_direct_path_and_oracle():
    """alps_b200_set_mode(1): the k-hoisted tables give the same chi / D as the direct quadrature
    (rounding-level differences only) and the same parity against the oracle."""
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    pl = tables.config_small(28, 56, kind=2)
    kperp, kpar = 0.35, 0.06
    oms = np.array(list(omega_samples(11, 24, (0.02, 1.5), (-0.05, 0.05))) + [0.3 + 0j, 0.011 - 1e-6j])
    sol = Solver(pl)
    try:
        sol.set_k(kperp, kpar)
        D0, chi_d = sol.disp_batch(oms, want_chi0=True)
        sol.set_mode(1)
        sol.set_k(kperp, kpar)
        D1, chi_f = sol.disp_batch(oms, want_chi0=True)
        full1 = [sol.disp(complex(om), full=True) for om in oms[:6]]
        sol.set_mode(0)
        sol.set_k(kperp, kpar)
        D2 = sol.disp_batch(oms)
    finally:
        sol.close()
    assert np.array_equal(D0, D2)                      # switching modes back is clean
    for i in range(oms.size):
        for s in range(pl.nspec):
            assert chi_err(chi_f[i, s], chi_d[i, s]) < 1e-9
    orc = Oracle(pl)
    orc.set_k(kperp, kpar)
    for om, (Dg, chi_g, low_g, wave_g) in zip(oms[:6], full1):
        Do, chi_o, low_o, wave_o = orc.disp(complex(om), full=True)
        ws = wave_scale(chi_o, complex(om), pl.vA, kperp, kpar)
        assert scaled_err(wave_g, wave_o, ws) < TOL
        assert abs(Dg - Do) / det_scale(ws) < TOL
        for s in range(pl.nspec):
            assert chi_err(chi_g[s], chi_o[s]) < TOL
            for m in range(3):
                if np.max(np.abs(low_o[s, :, :, m])) > 0:
                    assert chi_err(low_g[s, :, :, m], low_o[s, :, :, m]) < TOL


def _damped(seed, n, re_range):
    """omegas with Im <= 0 so that the Landau term (eval_fit) is exercised on resonant harmonics"""
    return list(omega_samples(seed, n, re_range, (-0.05, 0.0))) + [0.5 * (re_range[0] + re_range[1]) + 0j]


def test_kperp_norm_false():
    pl = tables.config_small(24, 48, kind=1)
    pl.kperp_norm = False
    _compare(pl, 0.3, 0.05, _damped(21, 5, (0.02, 1.2)) + [0.4 + 0.01j])


def test_analytic_continuation_method_0_hardcoded_maxwellians():
    """ACmethod = 0: distribution_analyt (distribution/distribution_analyt.f90:66-85)"""
    specs = [tables.DistSpec(ms=1.0), tables.DistSpec(ms=1.0 / 1836.0)]
    pl = tables.make_plasma(specs, ns=[1.0, 1.0], qs=[1.0, -1.0], nperp=24, npar=48, Bessel_zero=1.0e-30)
    for s in pl.species:
        s.ACmethod = 0
    _compare(pl, 0.3, 0.05, _damped(22, 5, (0.02, 1.2)))


def test_analytic_continuation_method_2_chebyshev():
    """ACmethod = 2: Chebyshev series of log10 f0 per p_perp row (fit_function_poly,
    src/ALPS_analyt.f90:262-363).  The coefficients are an input (the GLLS fit is out of scope); here they
    come from numpy's chebfit."""
    from numpy.polynomial import chebyshev as Ch
    pl = tables.config_small(24, 48, kind=1)
    order = 12
    coeffs = np.zeros((pl.nspec, pl.nperp + 1, order + 1), order="F")
    for i in range(pl.nspec):
        ppar = pl.pp[i, 0, :, 1]
        x = (ppar - 0.5 * (ppar[-1] + ppar[0])) / (0.5 * (ppar[-1] - ppar[0]))
        for ip in range(pl.nperp + 1):
            coeffs[i, ip, :] = Ch.chebfit(x, np.log10(pl.f0[i, ip, :]), order)
    pl.poly_fit_coeffs = coeffs
    for s in pl.species:
        s.ACmethod, s.poly_order, s.poly_kind, s.logfit, s.poly_log_max = 2, order, 1, True, 18.0
    _compare(pl, 0.3, 0.05, _damped(23, 5, (0.02, 1.2)))


def test_fit_type_6_bi_moyal():
    pl = tables.config_small(24, 48, kind=4)
    _compare(pl, 0.3, 0.05, _damped(24, 4, (0.02, 1.0)))


def test_fit_type_3_juettner_nonrelativistic_species():
    specs = [tables.DistSpec(ms=1.0, distribution=3), tables.DistSpec(ms=0.2, distribution=3)]
    pl = tables.make_plasma(specs, ns=[1.0, 1.0], qs=[1.0, -1.0], nperp=24, npar=48, vA=0.3, maxP=4.0,
                            Bessel_zero=1.0e-30)
    _compare(pl, 0.3, 0.2, _damped(25, 4, (0.05, 1.0)))


def test_two_fits_per_species_and_drift():
    """n_fits = 2 (core + beam Maxwellians summed in fit_function) on a drifting table"""
    specs = [tables.DistSpec(ms=1.0, drift=0.5), tables.DistSpec(ms=5.44662e-4)]
    pl = tables.make_plasma(specs, ns=[1.0, 1.0], qs=[1.0, -1.0], nperp=24, npar=48, Bessel_zero=1.0e-30)
    pf = np.zeros((2, pl.nperp + 1, 5, 2), order="F")
    pf[:, :, :, 0] = pl.param_fit[:, :, :, 0]
    pf[0, :, 0, 0] *= 0.7                       # split the proton Maxwellian into two components
    pf[0, :, :, 1] = pl.param_fit[0, :, :, 0]
    pf[0, :, 0, 1] *= 0.3
    pl.param_fit = pf
    pl.species[0].fit_type = [1, 1]
    pl.species[0].perp_correction = [pl.species[0].perp_correction[0]] * 2
    pl.species[1].fit_type = [1]
    _compare(pl, 0.3, 0.05, _damped(26, 4, (0.02, 1.2)))


def test_batch_chunking_and_api_errors():
    """alps_b200_disp_batch splits large batches into internal chunks (batch_max) -- same D; and the
    entry points report usage / grid / nmax errors instead of computing garbage."""
    from alps_b200 import _lib
    from alps_b200.solver import Solver
    pl = tables.config_small(24, 48, kind=1)
    oms = np.array(list(omega_samples(31, 23, (0.02, 1.4), (-0.04, 0.04))))
    sol = Solver(pl)
    try:
        with pytest.raises(_lib.AlpsB200Error) as e:
            sol.disp(0.3 + 0.01j)                       # set_k not called yet
        assert e.value.code == -2
        sol.set_k(0.3, 0.05)
        D_ref, chi_ref = sol.disp_batch(oms, want_chi0=True)
    finally:
        sol.close()
    sol = Solver(pl, batch_max=5)
    try:
        sol.set_k(0.3, 0.05)
        D, chi = sol.disp_batch(oms, want_chi0=True)    # 5 chunks of <= 5 omegas
        assert np.array_equal(D, D_ref) and np.array_equal(chi, chi_ref)
    finally:
        sol.close()
    sol = Solver(pl, nmax_cap=8)
    try:
        with pytest.raises(_lib.AlpsB200Error) as e:
            sol.set_k(0.3, 0.05)                        # needs nmax ~ 29 > cap
        assert e.value.code == -5
    finally:
        sol.close()
    bad = tables.config_small(24, 48, kind=1)
    bad.pp = bad.pp.copy(order="F")
    bad.pp[0, 3, 7, 1] *= 1.0000001                     # no longer a separable grid
    with pytest.raises(_lib.AlpsB200Error) as e:
        Solver(bad)
    assert e.value.code == -4
    _lib.lib().alps_b200_finalize()


def test_real_omega_exactly_on_a_resonant_node():
    """Im(om) = 0 and Re(p_res) exactly on a p_par node: the node itself sits in the excluded window
    (weight 0), its 1/den would be infinite.  Both GPU modes must stay finite and agree with the oracle."""
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    pl = tables.config_small(24, 48, kind=1)
    kperp, kpar = 0.3, 0.05
    ppar = pl.pp[0, 0, :, 1]
    oms = [complex(kpar * ppar[30] / pl.species[0].ms, 0.0), complex((kpar * ppar[20] + 1.0) / pl.species[0].ms, 0.0)]
    orc = Oracle(pl)
    orc.set_k(kperp, kpar)
    sol = Solver(pl)
    try:
        for mode in (0, 1):
            sol.set_mode(mode)
            sol.set_k(kperp, kpar)
            for om in oms:
                Do, chi_o, _, wave_o = orc.disp(om, full=True)
                Dg, chi_g, _, wave_g = sol.disp(om, full=True)
                assert np.all(np.isfinite(chi_g)) and np.isfinite(Dg.real)
                ws = wave_scale(chi_o, om, pl.vA, kperp, kpar)
                assert scaled_err(wave_g, wave_o, ws) < TOL, (mode, om)
                assert abs(Dg - Do) / det_scale(ws) < TOL, (mode, om)
    finally:
        sol.close()


def test_resonance_at_the_grid_edges_and_outside_the_grid():
    """integrate_res edge fall-backs (src/ALPS_fns.f90:971-992: the near-pole piece is dropped when the window
    touches either end of the p_par grid), resonances up to positions_principal cells outside the grid
    (determine_resonances :641-745), the dp/2 rule of upperlimit (:1002-1006) and funct_g's mid-cell ties:
    Re(p_res) of the protons' n = 0 and n = 1 harmonics swept over nodes next to both grid ends, with exact
    node / mid-cell / generic offsets, for damped, growing and real omega."""
    pl = tables.config_small(24, 48, kind=1)
    kperp, kpar = 0.3, 0.05
    ms, qs = pl.species[0].ms, pl.species[0].qs
    ppar = pl.pp[0, 0, :, 1]
    dp = ppar[1] - ppar[0]
    npar = pl.npar
    M_I = pl.positions_principal
    nodes = sorted(set([0, 1, 2, 3, M_I + 1, M_I + 2, M_I + 3, npar - M_I - 3, npar - M_I - 2, npar - 3, npar - 2,
                        npar - 1, npar]))
    oms = []
    for j in nodes:
        for off in (0.0, 0.31, 0.5, -0.5):
            pres = ppar[j] + off * dp
            for n in (0, 1):
                re = (kpar * pres + n * qs) / ms
                for im in (-3e-4, 0.0, 2e-4):
                    oms.append(complex(re, im))
    for c in (0.5, 1.0, M_I - 0.5, M_I, M_I + 0.5, M_I + 1.5):     # outside the grid, both sides
        for pres in (ppar[0] - c * dp, ppar[npar] + c * dp):
            oms.append(complex(kpar * pres / ms, -1e-4))
            oms.append(complex((kpar * pres + qs) / ms, 1e-4))
    oms = [om for om in oms if abs(om) > 1e-6]
    _compare(pl, kperp, kpar, oms)


def test_negative_kpar_and_drifting_species():
    """k_par < 0 flips the side of the Landau contour (abs(kpar) and sign(kpar) in landau_integrate /
    integrate_res, src/ALPS_fns.f90:1088-1165, 1327-1452)."""
    pl = tables.config_small(24, 48, kind=1)
    oms = list(omega_samples(7, 8, (0.02, 1.4), (-0.04, 0.04))) + [0.3 + 0j, 0.011 - 1e-6j]
    _compare(pl, 0.3, -0.05, oms)


def test_quadrature_variants_and_batch_classes_agree(monkeypatch):
    """The same D from every execution path of the non-relativistic chain: CUDA-graph replay of disp()
    (third call on), plain disp(), batches of the latency class (<= 8), the small-batch class (<= 64) and
    the throughput class (> 64, warp-per-harmonic k_resonant), the DMMA quadrature (default) and the DFMA
    quadrature (ALPS_B200_QUAD_VARIANT=8).  They differ only in summation order; bound: the 1e-9 parity tolerance on |D|."""
    from alps_b200.solver import Solver
    from alps_b200 import _lib
    pl = tables.config_kpar_fast()
    kperp, kpar = 1e-2, 2e-2
    oms = np.array(list(omega_samples(11, 150, (5e-3, 0.4), (-2e-3, 2e-3))) + [2e-2 + 0j])
    res = {}
    for variant in ("15", "8", "9"):
        monkeypatch.setenv("ALPS_B200_QUAD_VARIANT", variant)
        sol = Solver(pl, emulate_nproc=4)
        try:
            assert int(sol.info(_lib.INFO_QUAD_VARIANT)) == int(variant)
            sol.set_k(kperp, kpar)
            big = sol.disp_batch(oms)                       # > 64
            mid = sol.disp_batch(oms[:40])                  # <= 64
            tiny = sol.disp_batch(oms[:6])                  # <= 8
            single = np.array([sol.disp(complex(o)) for o in oms[:6]])   # graph replay from the 2nd call on
            again = np.array([sol.disp(complex(o)) for o in oms[:6]])
            assert np.array_equal(single, again)            # replay is deterministic
            assert np.array_equal(single, tiny)             # same batch class: bitwise
            res[variant] = big
            scale = np.abs(big[:40]) + 1e-300
            assert np.max(np.abs(mid - big[:40]) / scale) < 1e-9
            assert np.max(np.abs(tiny - big[:6]) / scale[:6]) < 1e-9
        finally:
            sol.close()
    scale = np.abs(res["15"])
    assert np.max(np.abs(res["8"] - res["15"]) / scale) < 1e-9
    assert np.max(np.abs(res["9"] - res["15"]) / scale) < 1e-9
    # set_k with another k_par between replays: the graph is re-used or rebuilt, never stale
    monkeypatch.delenv("ALPS_B200_QUAD_VARIANT")
    sol = Solver(pl, emulate_nproc=4)
    try:
        for kp in (2e-2, 3e-2, 2e-2):
            sol.set_k(kperp, kp)
            ref = sol.disp_batch(oms[:40])
            got = np.array([sol.disp(complex(o)) for o in oms[:4]])
            assert np.max(np.abs(got - ref[:4]) / np.abs(ref[:4])) < 1e-9, kp
    finally:
        sol.close()


KNOBS = ("ALPS_B200_ZC", "ALPS_B200_FUSE", "ALPS_B200_PDL", "ALPS_B200_EARLY", "ALPS_B200_SPIN", "ALPS_B200_FORK")


def test_single_omega_graph_knobs_are_bitwise_neutral(monkeypatch):
    """The latency chain of alps_b200_disp -- fused plan kernel reading omega from pinned host memory, fused
    harmonic-sum + determinant kernel writing D and the error words to pinned host memory -- against the plain
    chain (H2D / memset / D2H nodes, k_chi_partial and k_assemble as two launches): bitwise the same D, chi0,
    chi0_low and wave, for table species, a use_bM species and a relativistic pair plasma."""
    from alps_b200.solver import Solver
    cases = [(tables.config_kpar_fast(), dict(emulate_nproc=4), (1e-2, 2e-2),
              list(omega_samples(5, 12, (5e-3, 0.4), (-2e-3, 2e-3))) + [2e-2 + 0j]),
             (tables.config_bimax(60, 120), {}, (1.0e-3, 0.03),
              [3.0e-2 - 1.0e-5j, 4.5e-2 - 1.9e-2j, 0.1 + 0.002j, 0.05 + 0.0j]),
             (tables.config_relativistic(nperp=20, npar=40, ngamma=60, npparbar=80), {}, (1.0e-3, 1.0e-1),
              [6.2713e-2 - 4.662e-8j, 1.0 - 1.655e-6j, 0.5 + 0.01j, 0.3 - 0.02j, 2.5 + 0.0j, 0.9 + 1e-3j])]
    for icase, (pl, kw, k, oms) in enumerate(cases):
        out = {}
        for tag, env in (("fused", "110000"), ("plain", "000000"), ("pdl", "111000"), ("early", "111111")):
            for name, v in zip(KNOBS, env):
                monkeypatch.setenv(name, v)
            sol = Solver(pl, **kw)
            try:
                sol.set_k(*k)
                singles = np.array([sol.disp(complex(o)) for o in oms])      # graph replay from the 2nd call on
                again = np.array([sol.disp(complex(o)) for o in oms])
                assert np.array_equal(singles, again)
                full = [sol.disp(complex(o), full=True) for o in oms[:3]]
                batch = sol.disp_batch(np.array(oms[:8]))
                out[tag] = (singles, full, batch)
            finally:
                sol.close()
        assert np.array_equal(out["fused"][0], out["plain"][0])
        assert np.array_equal(out["fused"][0], out["pdl"][0])       # programmatic dependent launches in the graph
        assert np.array_equal(out["fused"][0], out["early"][0])     # ... with the Landau blocks started on k_plan's flag
        assert np.array_equal(out["fused"][2], out["early"][2])
        assert np.array_equal(out["fused"][2], out["plain"][2])
        if icase == 0:
            assert np.array_equal(out["fused"][0][:8], out["fused"][2])      # disp() and a small disp_batch(): same class
        for a, b in zip(out["fused"][1], out["plain"][1]):
            assert a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))
    for name in KNOBS:
        monkeypatch.delenv(name, raising=False)


def test_single_omega_chain_with_many_resonant_harmonics(monkeypatch):
    """k_par = 20 at k_perp = 3: dozens of harmonics are resonant at once -- more resonant (item, moment) units than the
    records of the harmonic sums hold (resonant.cu: CHI_REC), so the captured chain completes them the general way, and
    k_resonant_lat runs its wide grid.  D of the replayed chain against the oracle, and bitwise against the chain
    with the flag-driven starts switched off."""
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    pl = tables.config_kpar_fast()
    kperp, kpar = 3.0, 20.0
    oms = [1.5 - 0.05j, 30.0 + 0.01j, 12.3 - 0.4j, 55.0 + 0.0j]
    orc = Oracle(pl, nproc=4)
    orc.set_k(kperp, kpar)
    want = [orc.disp(om, full=True) for om in oms]
    got = {}
    for early in ("1", "0"):
        monkeypatch.setenv("ALPS_B200_EARLY", early)
        sol = Solver(pl, emulate_nproc=4)
        try:
            sol.set_k(kperp, kpar)
            rows = []
            for om in oms:
                d = [sol.disp(om) for _ in range(3)]      # plain launches, capture, replay
                assert d[0] == d[1] == d[2]
                rows.append(d[2])
            got[early] = rows
        finally:
            sol.close()
    monkeypatch.delenv("ALPS_B200_EARLY", raising=False)
    assert got["1"] == got["0"]
    for om, d, (Do, chi_o, _, _) in zip(oms, got["1"], want):
        ws = wave_scale(chi_o, complex(om), pl.vA, kperp, kpar)
        assert abs(d - Do) / det_scale(ws) < TOL, (om, d, Do)


def test_packed_remainder_tiles_are_bitwise_neutral(monkeypatch):
    """Throughput batches send a species' last harmonic tile through the packed instantiation of k_quad_mma when its
    upper group of 8 harmonics holds at most two harmonics of the summed range (their weight rows share one M-tile,
    an empty group issues no DMMA).  Same products in the same order: bitwise the D of the regular layout
    (ALPS_B200_NO_PACK=1), for 1, 2, 9 and 10 harmonics in the last tile."""
    import subprocess, sys, json
    code = r"""
import json, sys, numpy as np
sys.path.insert(0, %r)
from alps_b200 import tables
from alps_b200.solver import Solver
from tests.util import omega_samples
pl = tables.config_small(48, 96, kind=2)
oms = np.array(list(omega_samples(21, 96, (0.05, 2.0), (-0.03, 0.03))))
out = {}
for nmax in (16, 17, 24, 25, 31):
    sol = Solver(pl, nmax_force=nmax)
    sol.set_k(1.5, 0.05)
    out[nmax] = sol.disp_batch(oms).view(np.float64).tolist()
    sol.close()
print(json.dumps(out))
""" % ROOT
    res = {}
    for tag, env in (("packed", {}), ("regular", {"ALPS_B200_NO_PACK": "1"})):
        e = {k: v for k, v in os.environ.items() if k != "ALPS_B200_NO_PACK"}
        e.update(env)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        res[tag] = json.loads(p.stdout.strip().splitlines()[-1])
    for nmax in res["packed"]:
        a, b = np.array(res["packed"][nmax]), np.array(res["regular"][nmax])
        assert np.all(np.isfinite(a)) and np.array_equal(a, b), nmax


def test_committed_oracle_vectors():
    """The CUDA path against the committed oracle fixture tests/golden/oracle_vectors.npz (no oracle call here):
    chi0 / chi0_low element-wise, wave and D with the scale-aware 1e-9 criterion of the live parity tests."""
    from alps_b200.solver import Solver
    ref = np.load(os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz"))
    makers = {"small_bimax": (lambda: tables.config_small(24, 48, kind=1), 0),
              "small_kappa": (lambda: tables.config_small(24, 48, kind=2), 0),
              "kpar_fast": (tables.config_kpar_fast, 4)}
    for name, (make, nproc) in makers.items():
        pl = make()
        kperp, kpar = (float(x) for x in ref[name + "_k"])
        sol = Solver(pl, emulate_nproc=nproc)
        try:
            assert list(sol.set_k(kperp, kpar)) == list(ref[name + "_nmax"])
            oms = ref[name + "_om"]
            Db = sol.disp_batch(oms)
            for i, om in enumerate(oms):
                Dg, chi_g, low_g, wave_g = sol.disp(complex(om), full=True)
                chi_o, low_o, wave_o, Do = ref[name + "_chi0"][i], ref[name + "_chi0_low"][i], ref[name + "_wave"][i], ref[name + "_D"][i]
                ws = wave_scale(chi_o, complex(om), pl.vA, kperp, kpar)
                assert scaled_err(wave_g, wave_o, ws) < TOL, (name, om)
                assert abs(Dg - Do) / det_scale(ws) < TOL and abs(Db[i] - Do) / det_scale(ws) < TOL, (name, om)
                for s in range(pl.nspec):
                    assert chi_err(chi_g[s], chi_o[s]) < TOL, (name, om, s)
                    for m in range(3):
                        assert chi_err(low_g[s, :, :, m], low_o[s, :, :, m]) < TOL, (name, om, s, m)
        finally:
            sol.close()
