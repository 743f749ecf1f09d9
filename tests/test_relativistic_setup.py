"""Host-side relativistic set-up (alps_b200/relativistic.py, twin of derivative_f0_rel,
src/ALPS_fns_rel.f90:36-426): thin-plate-spline regrid of a Juettner table, cone sentinel,
normalisation and derivatives against the analytic distribution."""
import numpy as np

from alps_b200 import tables


def test_regrid_of_a_juettner_table():
    pl = tables.config_relativistic(nperp=20, npar=40, ngamma=40, npparbar=60)
    g, p, f, d = pl.gamma_rel[0], pl.pparbar_rel[0], pl.f0_rel[0], pl.df0_rel[0]
    # uniform separable grid
    assert np.allclose(np.diff(g[:, 0]), g[1, 0] - g[0, 0]) and np.all(g == g[:, :1])
    assert np.allclose(np.diff(p[0, :]), p[0, 1] - p[0, 0]) and np.all(p == p[:1, :])
    # cone sentinel exactly where gamma^2 - 1 < pparbar^2
    outside = (g ** 2 - 1.0) < p ** 2
    assert np.all(f[outside] == -1.0) and np.all(f[~outside] > 0.0)
    # normalisation: sum gamma f 2 pi dgamma dpparbar (ms/vA)^3 = 1
    dg, dp = g[2, 2] - g[1, 2], p[2, 2] - p[2, 1]
    assert abs(np.sum(g[~outside] * f[~outside]) * 2 * np.pi * dg * dp - 1.0) < 1e-12
    # f ~ exp(-2 gamma): d ln f / d gamma = -2, no pparbar dependence (interior of the cone)
    inner = np.zeros_like(outside)
    inner[2:-2, 2:-2] = True
    inner &= ~outside
    inner[1:, :] &= ~outside[:-1, :]
    inner[:-1, :] &= ~outside[1:, :]
    inner[:, 1:] &= ~outside[:, :-1]
    inner[:, :-1] &= ~outside[:, 1:]
    ratio = d[:, :, 0][inner] / f[inner]
    assert np.max(np.abs(ratio + 2.0)) < 0.05
    assert np.max(np.abs(d[:, :, 1][inner] / f[inner])) < 0.05
    # the fit amplitude handed to eval_fit reproduces the table: f = p1 exp(-2 gamma)
    p1 = pl.param_fit[0, 0, 0, 0]
    model = p1 * np.exp(-pl.species[0].perp_correction[0] * g)
    assert np.max(np.abs(model[inner] / f[inner] - 1.0)) < 0.02


import pytest


@pytest.mark.gpu
def test_device_spline_evaluation_matches_the_host_statement():
    """alps_b200_tps_eval (k_tps_eval: the evaluation loop of polyharmonic_spline, src/ALPS_fns_rel.f90:300-331,
    407-423) against the numpy statement of the same loop: log f0_rel to 1e-11 absolute (the sums run over ~900
    nodes with weights of both signs; numpy's matmul blocks the sum, the kernel adds in node order), the cone
    sentinel identical, normalised f0_rel and its derivatives to 1e-9 relative."""
    from alps_b200.relativistic import derivative_f0_rel
    pl = tables.config_relativistic(nperp=20, npar=40, ngamma=40, npparbar=60)
    host = derivative_f0_rel(pl.pp[0], pl.f0[0], 1.0, 1.0, 40, 60, backend="host")
    dev = derivative_f0_rel(pl.pp[0], pl.f0[0], 1.0, 1.0, 40, 60, backend="device")
    assert np.array_equal(host[0], dev[0]) and np.array_equal(host[1], dev[1])
    fh, fd = host[2], dev[2]
    assert np.array_equal(fh == -1.0, fd == -1.0)
    inside = fh > 0.0
    assert np.max(np.abs(np.log(fh[inside]) - np.log(fd[inside]))) < 1e-9
    scale = np.max(np.abs(host[3]))
    assert np.max(np.abs(host[3] - dev[3])) < 1e-8 * scale
    assert abs(host[4] - dev[4]) < 1e-9 * abs(host[4])


def test_regrid_against_scipy_rbf_interpolator():
    """Independent pin of the thin-plate-spline regrid (polyharmonic_spline, src/ALPS_fns_rel.f90:300-426, restated in
    alps_b200/relativistic.py: r^2 log r kernel + linear polynomial, dense LAPACK solve): scipy's RBFInterpolator with
    kernel='thin_plate_spline', degree=1 solves the same interpolation problem with its own assembly, scaling and
    solver.  Both interpolate log f0 of the same Juettner table onto the same (Gamma, pbar_par) grid; inside the cone the
    two splines agree to ~1e-8 (they are THE interpolant: the problem is unisolvent), and both reproduce the analytic
    log f0 = const - a Gamma to the spline's own accuracy."""
    from scipy.interpolate import RBFInterpolator
    from alps_b200.relativistic import derivative_f0_rel
    pl = tables.config_relativistic(nperp=20, npar=40, ngamma=40, npparbar=60)
    pp, f0, ms, vA = pl.pp[0], pl.f0[0], 1.0, 1.0
    g, p, f, d, integ = derivative_f0_rel(pp, f0, ms, vA, 40, 60, backend="host")
    gc = np.sqrt(1.0 + (pp[:, :, 0] ** 2 + pp[:, :, 1] ** 2) * vA * vA / (ms * ms)).ravel()
    pc = (pp[:, :, 1] * vA / ms).ravel()
    rbf = RBFInterpolator(np.column_stack([gc, pc]), np.log(f0).ravel(), kernel="thin_plate_spline", degree=1)
    inside = f > 0.0
    ref = rbf(np.column_stack([g[inside], p[inside]]))
    ours = np.log(f[inside] * integ)            # undo the renormalisation of derivative_f0_rel
    assert np.max(np.abs(ours - ref)) < 1e-7, np.max(np.abs(ours - ref))
    # and against the analytic table: log f0 = log C - a Gamma with a = perp_correction
    a = pl.species[0].perp_correction[0]
    lin = ours + a * g[inside]
    assert np.max(np.abs(lin - np.median(lin))) < 2e-2
