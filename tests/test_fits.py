"""Twin of determine_param_fit (alps_b200/fits.py) against the reference's own log and against the ideal
parameters of the table generator.  CPU only (set-up code)."""
import os

import numpy as np

from alps_b200 import fits, tables
from alps_b200.namelist import read_namelists
from alps_b200.run import plasma_from_inputs

HERE = os.path.dirname(os.path.abspath(__file__))


def test_lm_fit_reproduces_the_reference_log():
    """tests/test_kpar_fast.out:71-81 of the reference: param_fit(:,1,:,1) to its 5 printed digits,
    'Sum of all least-squares: 3.3601E-008', 'Standard error of the estimate: 7.6378E-007'."""
    nl = read_namelists(os.path.join(HERE, "inputs", "test_kpar_fast.in"))
    dl = read_namelists(os.path.join(HERE, "inputs", "test_kpar_fast_dist.in"))
    pl = plasma_from_inputs(nl, dl, base_dir=os.path.join(HERE, "inputs"), fit=True)
    assert "%.4E" % pl.fit_quality == "3.3601E-08"
    assert "%.4E" % np.sqrt(pl.fit_quality / (pl.nspec * pl.nperp * pl.npar)) == "7.6378E-07"
    assert ["%.4E" % v for v in pl.param_fit[0, 1, :2, 0]] == ["1.7959E-01", "1.0000E+00"]
    assert ["%.4E" % v for v in pl.param_fit[1, 1, :2, 0]] == ["1.4128E+04", "1.8360E+03"]
    assert abs(pl.param_fit[0, 1, 2, 0]) < 1e-15 and abs(pl.param_fit[1, 1, 2, 0]) < 1e-15   # drift: 2.1E-17 there
    # every row: amplitude follows exp(-(perpcorr_ideal - perpcorr_in) p_perp^2) of the ideal fit
    ideal = plasma_from_inputs(nl, dl, base_dir=os.path.join(HERE, "inputs"))
    for i in range(2):
        a = pl.param_fit[i, :, 0, 0] * np.exp(-pl.species[i].perp_correction[0] * pl.pp[i, :, 0, 0] ** 2)
        b = ideal.param_fit[i, :, 0, 0] * np.exp(-ideal.species[i].perp_correction[0] * ideal.pp[i, :, 0, 0] ** 2)
        assert np.max(np.abs(a / b - 1.0)) < 2e-5
        assert np.max(np.abs(pl.param_fit[i, :, 1, 0] / ideal.param_fit[i, :, 1, 0] - 1.0)) < 1e-5


def test_lm_fit_recovers_kappa_and_drifting_maxwellian_rows():
    """Start 3 % off on every parameter; the fit must come back to the generating values."""
    rng = np.random.default_rng(0)
    ppar = np.linspace(-4.0, 4.0, 161)
    opt = fits.FitOptions(epsilon_fit=1e-14)
    # kappa (fit type 2): f = A (1 + b (p-d)^2 + pc e pperp^2)^(-kappa-1)
    true = np.array([0.3, 0.12, 0.25, -9.0, 0.12])
    for first, pperp in ((True, 0.0), (False, 0.7)):
        f = fits.fit_function([2], [1.0], true, pperp, ppar, 1.0, 1e-4)
        start = true * (1.0 + 0.03 * rng.uniform(-1, 1, 5))
        if first:
            start[4] = true[4]          # frozen on the first row
        got, q = fits.LM_nonlinear_fit(np.log(f), [2], [1.0], start, pperp, ppar, 1.0, 1e-4, True, first, opt)
        assert q < 1e-10
        if first:     # off axis A, b and e are degenerate ((1 + e') factors out): only the function is determined
            assert np.max(np.abs(got / true - 1.0)) < 1e-6
        fit = fits.fit_function([2], [1.0], got, pperp, ppar, 1.0, 1e-4)
        assert np.max(np.abs(fit / f - 1.0)) < 1e-5
    # two Maxwellians, linear fit (logfit = F)
    true = np.array([1.0, 0.8, -0.5, 0.2, 2.0, 1.5])
    f = fits.fit_function([1, 1], [1.0, 1.0], true, 0.3, ppar, 1.0, 1e-4)
    got, q = fits.LM_nonlinear_fit(f, [1, 1], [1.0, 1.0], true * 1.03, 0.3, ppar, 1.0, 1e-4, False, False, opt)
    assert np.max(np.abs(got / true - 1.0)) < 1e-5


def test_jacobian_matches_finite_differences():
    ppar = np.linspace(-2.0, 2.5, 37)
    cases = {1: [0.7, 0.9, 0.2], 2: [0.3, 0.12, 0.25, -5.0, 0.2], 3: [2.0, 30.0, 0.1], 4: [1.3], 5: [1.3, 0.4, -0.2],
             6: [0.5, 0.3, 0.1, 0.2]}
    for ft, p in cases.items():
        p = np.array(p)
        JT = fits.determine_JT([ft], [1.1], p, 0.6, ppar, 1.0, 0.3, first_row=False)
        assert JT.shape == (len(p), ppar.size)
        for k in range(len(p)):
            h = 1e-6 * max(1.0, abs(p[k]))
            pp_, pm_ = p.copy(), p.copy()
            pp_[k] += h
            pm_[k] -= h
            fd = (fits.fit_function([ft], [1.1], pp_, 0.6, ppar, 1.0, 0.3) -
                  fits.fit_function([ft], [1.1], pm_, 0.6, ppar, 1.0, 0.3)) / (2 * h)
            assert np.max(np.abs(JT[k] - fd)) <= 1e-6 * np.max(np.abs(fd)) + 1e-12, (ft, k)
    # first row: the perpendicular parameter of kappa / bi-Moyal fits has no Jacobian row
    assert fits.determine_JT([2], [1.0], np.array(cases[2]), 0.0, ppar, 1.0, 0.3, True).shape[0] == 4
    assert fits.determine_JT([6], [1.0], np.array(cases[6]), 0.0, ppar, 1.0, 0.3, True).shape[0] == 3


def test_chebyshev_series_fit():
    """determine_GLLS (ACmethod = 2): log10 of a Maxwellian row is a parabola -> exact at order >= 2, and the
    zero floor of the reference (1 % of the smallest positive value)."""
    npar, order = 120, 8
    ppar = np.linspace(-3.0, 3.0, npar + 1)
    f0 = np.exp(-np.outer(np.linspace(0, 2, 5) ** 2, np.ones(npar + 1))) * np.exp(-ppar ** 2)[None, :]
    c = fits.determine_GLLS(f0, order, logfit=True)
    yy = -1.0 + np.arange(npar + 1) * (2.0 / npar)
    rec = np.polynomial.chebyshev.chebval(yy, c.T).T if False else np.array([np.polynomial.chebyshev.chebval(yy, r) for r in c])
    assert np.max(np.abs(rec - np.log10(f0))) < 1e-9
    assert np.max(np.abs(c[:, 3:])) < 1e-9
    f0z = f0.copy()
    f0z[2, :7] = 0.0
    cz = fits.determine_GLLS(f0z, order, logfit=True)
    rz = np.polynomial.chebyshev.chebval(yy[7:], cz[2])
    floor = np.log10(0.01 * f0z[2][f0z[2] > 0].min())
    assert np.isfinite(cz).all() and abs(np.polynomial.chebyshev.chebval(yy[0], cz[2]) - floor) < 1.5
    assert np.max(np.abs(rz - np.log10(f0z[2, 7:]))) < 2.0


def test_chebyshev_glls_against_an_independent_least_squares_solve():
    """determine_GLLS / least_squares_fit (src/ALPS_analyt.f90:731-848: normal equations through dgemm / dgesv) at the
    order the shipped inputs use (30, tests/test_chebyshev.in) on bi-kappa rows, whose log10 is not a polynomial: the
    coefficients must be the least-squares solution that an orthogonal-factorisation solver (numpy lstsq, SVD) finds
    on the Chebyshev-Vandermonde matrix of the same nodes, and the series must reproduce the rows."""
    from numpy.polynomial import chebyshev as Ch
    npar, order = 300, 30
    pl = tables.make_plasma([tables.DistSpec(ms=1.0, kappa=4.0, distribution=2)], ns=[1.0], qs=[1.0], nperp=12,
                            npar=npar)
    f0 = pl.f0[0]
    c = fits.determine_GLLS(f0, order, logfit=True)
    yy = -1.0 + np.arange(npar + 1) * (2.0 / npar)
    V = Ch.chebvander(yy, order)
    assert np.array_equal(V, fits.polynomial_basis(npar, order)) or np.max(np.abs(V - fits.polynomial_basis(npar, order))) < 1e-12
    for iperp in range(f0.shape[0]):
        ref = np.linalg.lstsq(V, np.log10(f0[iperp]), rcond=None)[0]
        # the normal equations square the condition number (~1e3 here): 1e-9 of the largest coefficient
        assert np.max(np.abs(c[iperp] - ref)) < 1e-9 * np.max(np.abs(ref)), iperp
        rec = Ch.chebval(yy, c[iperp])
        assert np.max(np.abs(rec - Ch.chebval(yy, ref))) < 1e-9          # the same series ...
        assert np.max(np.abs(rec - np.log10(f0[iperp]))) < 1e-3         # ... whose truncation error is ~1e-4 in log10
    # linear fit (logfit = F) goes through the same solver
    cl = fits.determine_GLLS(f0, order, logfit=False)
    refl = np.linalg.lstsq(V, f0[3], rcond=None)[0]
    assert np.max(np.abs(cl[3] - refl)) < 1e-9 * np.max(np.abs(refl))


def test_relativistic_rows_fit_type_4():
    """Juettner species on the (Gamma, pbar_par) grid (fit type 4, one amplitude per Gamma row, cone limits of
    lines 577-592): started from the .in value 0.62719 the fit lands on the amplitude that normalises f0_rel."""
    nl = read_namelists(os.path.join(HERE, "inputs", "test_relativistic_small.in"))
    dl = read_namelists(os.path.join(HERE, "inputs", "test_relativistic_dist.in"))
    pl = plasma_from_inputs(nl, dl, base_dir=os.path.join(HERE, "inputs"), fit=True)
    assert pl.fit_quality < 1e-12
    for i, sp in enumerate(pl.species):
        assert sp.fit_type == [4]
        a = pl.param_fit[i, :pl.ngamma + 1, 0, 0]
        g = pl.gamma_rel[i, :, 1]
        inside = pl.f0_rel[i] > -1.0
        rows = np.where(inside.sum(axis=1) > 3)[0]
        assert rows.size > pl.ngamma // 2
        for ig in rows[::7]:
            ip = np.where(inside[ig])[0][len(np.where(inside[ig])[0]) // 2]
            assert abs(a[ig] * np.exp(-sp.perp_correction[0] * g[ig]) / pl.f0_rel[i, ig, ip] - 1.0) < 1e-9
