"""Host-side set-up of the relativistic (Gamma, pbar_par) tables -- the input producer of the
relativistic integrators (one-off rank-0 work in the reference).  The dense 1894^2
solve is LAPACK dgesv on the host like in the reference; the evaluation of the spline on the (Gamma, pbar_par)
grid -- 4.7e8 kernel evaluations at C3 -- runs on the device with backend="device" (alps_b200_tps_eval,
csrc/setup_kernels.cu: k_tps_eval), which is what the twin main program uses; backend="host" is the numpy
statement of the same loop (set-up checks without a GPU).

Follows derivative_f0_rel and polyharmonic_spline, src/ALPS_fns_rel.f90:36-426: thin-plate-spline
regrid of log f0 from the (p_perp,p_par) table onto a uniform (Gamma, pbar_par) grid, cone sentinel
f0_rel = -1 outside the sub-luminal cone, renormalisation, centred / one-sided differences.
"""
from __future__ import annotations

import numpy as np


def _tps_kernel(r):
    """r^2 log r for r >= 1, r*log(r**r) for 0 < r < 1, 0 at r = 0 (lines 385-391)"""
    out = np.zeros_like(r)
    big = r >= 1.0
    out[big] = r[big] * r[big] * np.log(r[big])
    small = (r > 0.0) & ~big
    out[small] = r[small] * np.log(r[small] ** r[small])
    return out


def _tps_eval_device(gc, pc, w, gx, px):
    import ctypes as C

    from . import _lib
    gc, pc, w, gx, px = (np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (gc, pc, w, gx, px))
    out = np.zeros(gx.size)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.lib().alps_b200_tps_eval(gc.size, p(gc), p(pc), p(w), gx.size, p(gx), p(px), p(out)))
    return out


def derivative_f0_rel(pp_s, f0_s, ms, vA, ngamma, npparbar, smoothing=0.0, backend="host"):
    """One species: pp_s (nperp+1, npar+1, 2), f0_s (nperp+1, npar+1) ->
    gamma_rel, pparbar_rel, f0_rel (ngamma+1, npparbar+1), df0_rel (ngamma+1, npparbar+1, 2)."""
    from scipy.linalg import solve
    pperp, ppar = pp_s[:, :, 0], pp_s[:, :, 1]
    gamma = np.sqrt((pperp ** 2 + ppar ** 2) * vA ** 2 / ms ** 2 + 1.0)
    gamma_min = float(gamma.min())
    pparbar_min = float(ppar.min()) * vA / ms
    pparbar_max = float(ppar.max()) * vA / ms
    gamma_max_use = float(np.sqrt(1.0 + pperp[-1, 1] ** 2 * vA ** 2 / ms ** 2))
    gc = np.sqrt(1.0 + (pperp ** 2 + ppar ** 2) * vA * vA / (ms * ms)).ravel()     # iperp outer, ipar inner
    pc = (ppar * vA / ms).ravel()
    grid = np.log(f0_s).ravel()
    n = gc.size
    g1 = gamma_min + ((gamma_max_use - gamma_min) * np.arange(ngamma + 1)) / (1.0 * ngamma)
    p1 = pparbar_min + ((pparbar_max - pparbar_min) * np.arange(npparbar + 1)) / (1.0 * npparbar)
    gamma_rel, pparbar_rel = np.meshgrid(g1, p1, indexing="ij")
    # polyharmonic_spline: [[K + smoothing I, P], [P^T, 0]] w = [grid, 0]
    M = np.zeros((n + 3, n + 3))
    r = np.sqrt((gc[:, None] - gc[None, :]) ** 2 + (pc[:, None] - pc[None, :]) ** 2)
    M[:n, :n] = _tps_kernel(r)
    M[np.arange(n), np.arange(n)] += smoothing
    M[:n, n], M[:n, n + 1], M[:n, n + 2] = 1.0, gc, pc
    M[n, :n], M[n + 1, :n], M[n + 2, :n] = 1.0, gc, pc
    rhs = np.zeros(n + 3)
    rhs[:n] = grid
    w = solve(M, rhs)          # LAPACK dgesv, like the reference (line 402)
    f0_rel = np.zeros_like(gamma_rel)
    if backend == "device":
        f0_rel = _tps_eval_device(gc, pc, w, gamma_rel, pparbar_rel).reshape(gamma_rel.shape)
    elif backend == "host":
        for i in range(ngamma + 1):
            rr = np.sqrt((gamma_rel[i, :, None] - gc[None, :]) ** 2 + (pparbar_rel[i, :, None] - pc[None, :]) ** 2)
            f0_rel[i] = _tps_kernel(rr) @ w[:n] + w[n] + w[n + 1] * gamma_rel[i] + w[n + 2] * pparbar_rel[i]
    else:
        raise ValueError("backend must be 'host' or 'device'")
    f0_rel = np.exp(f0_rel)
    f0_rel[(gamma_rel ** 2 - 1.0) < pparbar_rel ** 2] = -1.0           # outside the cone
    dgamma = gamma_rel[2, 2] - gamma_rel[1, 2]
    dpparbar = pparbar_rel[2, 2] - pparbar_rel[2, 1]
    inside = f0_rel > -1.0
    integrate = float(np.sum(gamma_rel[inside] * f0_rel[inside]) * 2.0 * np.pi * dgamma * dpparbar * (ms / vA) ** 3)
    f0_rel[f0_rel != -1.0] /= integrate
    return gamma_rel, pparbar_rel, f0_rel, rel_derivatives(f0_rel, gamma_rel, pparbar_rel), integrate


def rel_derivatives(f0_rel, gamma_rel, pparbar_rel):
    """df0_rel(:, :, 1:2) = d f0_rel / d Gamma, d f0_rel / d pbar_par by centred differences, one-sided next to the cone
    (f0_rel = -1 outside it): the last block of derivative_f0_rel, src/ALPS_fns_rel.f90:230-270."""
    df = np.zeros(f0_rel.shape + (2,))
    F = f0_rel
    c = (slice(1, -1), slice(1, -1))
    up, dn = F[2:, 1:-1], F[:-2, 1:-1]
    ok = (dn > 0.0) & (up > 0.0)
    d1 = np.zeros_like(up)
    d1[ok] = ((up - dn) / (gamma_rel[2:, 1:-1] - gamma_rel[:-2, 1:-1]))[ok]
    rt, lf, mid = F[1:-1, 2:], F[1:-1, :-2], F[1:-1, 1:-1]
    pr, pl_, pm = pparbar_rel[1:-1, 2:], pparbar_rel[1:-1, :-2], pparbar_rel[1:-1, 1:-1]
    d2 = np.zeros_like(mid)
    ok = (rt > 0.0) & (lf > 0.0)
    d2[ok] = ((rt - lf) / (pr - pl_))[ok]
    e1 = (mid >= 0.0) & (rt <= 0.0) & (lf > 0.0)          # right neighbour outside the cone
    d2[e1] = ((mid - lf) / (pm - pl_))[e1]
    e2 = (mid >= 0.0) & (rt > 0.0) & (lf <= 0.0)          # left neighbour outside the cone
    d2[e2] = ((rt - mid) / (pr - pm))[e2]
    df[c + (0,)] = d1
    df[c + (1,)] = d2
    return df
