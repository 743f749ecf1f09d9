"""Host twin of the reference's analytic-continuation fit producers:

    determine_param_fit   src/ALPS_analyt.f90:365-728    Levenberg-Marquardt fit of every p_perp (Gamma) row
    determine_JT          src/ALPS_analyt.f90:1183-1389  transposed Jacobian of the fit function
    LM_nonlinear_fit      src/ALPS_analyt.f90:1395-1636
    fit_function          src/ALPS_analyt.f90:185-266    (real p_par: what the LM loop evaluates)
    set_polynomial_basis / determine_GLLS / least_squares_fit   src/ALPS_analyt.f90:731-889 (ACmethod = 2)

It is set-up code, run once per f0 table on the host like the reference runs it on rank 0; the products
(`param_fit`, `poly_fit_coeffs`) are inputs of the GPU path (`alps_b200_upload`).  The small dense solves go
through the same LAPACK routines the reference calls (dgetrf/dgetri, dgesv) via scipy.  Rows are sequential
by construction: row iperp starts from the converged parameters of row iperp-1 (line 489).

Known answers in the reference (tests/test_kpar_fast.out:71-81): param_fit(:,1,:,1) and
"Sum of all least-squares: 3.3601E-008" -- reproduced by tests/test_fits.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.linalg import lapack

N_PARAMS = {1: 3, 2: 5, 3: 3, 4: 1, 5: 3, 6: 4}   # src/ALPS_analyt.f90:466-472


@dataclass
class FitOptions:
    """&system fit controls (src/ALPS_io.f90:81-85; defaults of tests/test_kpar_fast.in)."""
    maxsteps_fit: int = 500
    lambda_initial_fit: float = 1.0
    lambdafac_fit: float = 10.0
    epsilon_fit: float = 1.0e-8


class FitError(RuntimeError):
    pass


def fit_function(fit_type, perp_correction, params, pperp, ppar, ms, vA):
    """fit_function for real p_par (vectorised over ppar), src/ALPS_analyt.f90:185-266."""
    f = np.zeros_like(np.asarray(ppar, dtype=np.float64))
    k = 0
    for ft, pc in zip(fit_type, perp_correction):
        p = params[k:k + N_PARAMS[ft]]
        if ft == 1:
            f = f + p[0] * np.exp(-pc * pperp ** 2) * np.exp(-p[1] * (ppar - p[2]) ** 2)
        elif ft == 2:
            kp = 1.0 + p[1] * (ppar - p[2]) ** 2 + pc * p[4] * pperp ** 2
            f = f + p[0] * kp ** p[3]
        elif ft == 3:
            sq = np.sqrt(1.0 + (pperp ** 2 + (ppar - p[2]) ** 2) * vA * vA / (ms * ms))
            f = f + p[0] * np.exp(-p[1] * sq)
        elif ft == 4:
            f = f + p[0] * np.exp(-pc * pperp)
        elif ft == 5:
            f = f + p[0] * np.exp(-pc * pperp) * np.exp(-p[1] * (ppar - p[2]) ** 2)
        elif ft == 6:
            a = p[3] * pc * pperp ** 2 + p[1] * (ppar - p[2]) ** 2
            f = f + p[0] * np.exp(0.5 * (a - np.exp(a)))
        else:
            raise FitError("unknown fit type %d" % ft)
        k += N_PARAMS[ft]
    return f


def determine_JT(fit_type, perp_correction, params, pperp, ppar, ms, vA, first_row):
    """Rows of the transposed Jacobian for the *fitted* parameters only: on the first row (iperp = 0) the
    perpendicular parameter of kappa (5th) and bi-Moyal (4th) fits is frozen and has no row
    (src/ALPS_analyt.f90:1183-1389 with the JT_ind bookkeeping of lines 1289-1296, 1375-1383)."""
    rows = []
    k = 0
    for ft, pc in zip(fit_type, perp_correction):
        p = params[k:k + N_PARAMS[ft]]
        d = ppar - p[2] if ft in (1, 2, 3, 5, 6) else None
        if ft == 1:
            e = np.exp(-p[1] * d ** 2 - pc * pperp ** 2)
            rows += [e, -(d ** 2) * p[0] * e, 2.0 * p[1] * d * p[0] * e]
        elif ft == 2:
            kp = 1.0 + p[1] * d ** 2 + pc * p[4] * pperp ** 2
            rows += [kp ** p[3],
                     p[0] * p[3] * kp ** (p[3] - 1.0) * d ** 2,
                     p[0] * p[3] * kp ** (p[3] - 1.0) * 2.0 * p[1] * (p[2] - ppar),
                     np.log(kp) * p[0] * kp ** p[3]]
            if not first_row:
                rows.append(p[0] * p[3] * kp ** (p[3] - 1.0) * pc * pperp ** 2)
        elif ft == 3:
            sq = np.sqrt(1.0 + (pperp ** 2 + d ** 2) * vA * vA / (ms * ms))
            e = np.exp(-p[1] * sq)
            rows += [e, -p[0] * e * sq, p[0] * e * (p[1] / sq) * d * vA * vA / (ms * ms)]
        elif ft == 4:
            rows += [np.exp(-pc * pperp) * np.ones_like(ppar)]
        elif ft == 5:
            e = np.exp(-p[1] * d ** 2) * np.exp(-pc * pperp)
            rows += [e, -p[0] * e * d ** 2, 2.0 * p[0] * d * p[1] * e]
        elif ft == 6:
            a = p[3] * pc * pperp ** 2 + p[1] * d ** 2
            e = np.exp(0.5 * (a - np.exp(a)))
            rows += [e, p[0] * e * 0.5 * d ** 2 * (1.0 - np.exp(a)), p[0] * e * p[1] * (p[2] - ppar) * (1.0 - np.exp(a))]
            if not first_row:
                rows.append(p[0] * e * 0.5 * pc * pperp ** 2 * (1.0 - np.exp(a)))
        k += N_PARAMS[ft]
    return np.array(rows, dtype=np.float64)


def param_mask(fit_type, first_row):
    """param_mask of determine_param_fit (lines 493-548)."""
    mask = []
    for ft in fit_type:
        m = [True] * N_PARAMS[ft]
        if first_row and ft == 2:
            m[4] = False
        if first_row and ft == 6:
            m[3] = False
        mask += m
    return np.array(mask, dtype=bool)


def LM_nonlinear_fit(g, fit_type, perp_correction, params, pperp, ppar, ms, vA, logfit, first_row, opt: FitOptions):
    """One row.  Returns (params, quality) with quality = the sum of squared residuals *before* the last
    parameter update, exactly what the reference accumulates (src/ALPS_analyt.f90:1395-1636)."""
    params = np.array(params, dtype=np.float64)
    mask = param_mask(fit_type, first_row)
    if g.size < params.size:
        raise FitError("alps_error(2): fewer points than fit parameters")
    lam = opt.lambda_initial_fit
    counter = 0
    with np.errstate(all="ignore"):
        while True:
            counter += 1
            JT = determine_JT(fit_type, perp_correction, params, pperp, ppar, ms, vA, first_row)
            f = fit_function(fit_type, perp_correction, params, pperp, ppar, ms, vA)
            if logfit:
                JT = JT / f
                res = g - np.log(f)
            else:
                res = g - f
            LSQ = float(np.sum(res * res))
            JTJ = JT @ JT.T
            A = JTJ + lam * np.diag(np.diag(JTJ))
            lu, piv, info = lapack.dgetrf(A)
            if info != 0:
                raise FitError("Fit matrix is numerically singular.")
            Ainv, info = lapack.dgetri(lu, piv)
            if info != 0:
                raise FitError("Fit matrix inversion failed.")
            delta = Ainv @ (JT @ res)
            params[mask] += delta
            f = fit_function(fit_type, perp_correction, params, pperp, ppar, ms, vA)
            res = g - (np.log(f) if logfit else f)
            LSQnew = float(np.sum(res * res))
            if LSQnew > LSQ:
                params[mask] -= delta
                lam *= opt.lambdafac_fit
            else:
                lam /= opt.lambdafac_fit
            # NaN compares false like in Fortran: the loop then runs to maxsteps_fit
            if abs(LSQnew - LSQ) < opt.epsilon_fit or counter == opt.maxsteps_fit:
                return params, LSQ


def _pack(row, fit_type):
    """param_fit(is,iperp,1:5,ifit) -> flat params (lines 495-548)."""
    out = []
    for j, ft in enumerate(fit_type):
        out += [row[q, j] for q in range(N_PARAMS[ft])]
    return np.array(out, dtype=np.float64)


def _unpack(params, fit_type, row):
    k = 0
    for j, ft in enumerate(fit_type):
        for q in range(N_PARAMS[ft]):
            row[q, j] = params[k + q]
        k += N_PARAMS[ft]


def determine_param_fit(plasma, initial, opt: FitOptions | None = None, rel=None):
    """Twin of determine_param_fit for all species of `plasma` (alps_b200.tables.Plasma).

    initial: (nspec, 5, maxfits) start values of row 0 (the &ffit_is_ifit namelists).
    rel: for relativistic species, dict is -> (f0_rel, gamma_rel, pparbar_rel) on the (Gamma, pbar_par) grid
         as produced by alps_b200.relativistic (cone sentinel f0_rel = -1).
    Returns (param_fit (nspec, max(nperp,ngamma)+1, 5, maxfits) Fortran order, poly_fit_coeffs or None,
    qualitytotal).  Species with use_bM, ACmethod 0 or 2 get zero parameters like the reference (lines 447-461)."""
    opt = opt or FitOptions()
    nspec, nperp, npar = plasma.nspec, plasma.nperp, plasma.npar
    maxfits = max(len(s.fit_type) for s in plasma.species)
    nrow = nperp
    if rel:
        nrow = max([nperp] + [v[0].shape[0] - 1 for v in rel.values()])
    pf = np.zeros((nspec, nrow + 1, 5, maxfits), order="F")
    maxorder = max(s.poly_order for s in plasma.species)
    poly = np.zeros((nspec, nperp + 1, maxorder + 1), order="F") if maxorder > 0 else None
    total = 0.0
    for i, sp in enumerate(plasma.species):
        if sp.usebM or sp.ACmethod == 0:
            continue
        if sp.ACmethod == 2:
            poly[i, :, :sp.poly_order + 1] = determine_GLLS(plasma.f0[i], sp.poly_order, sp.logfit, sp.poly_kind)
            continue
        ft, pc = list(sp.fit_type), list(sp.perp_correction)
        upper = nperp
        if sp.relativistic:
            f0r, gr, pbr = rel[i]
            upper = f0r.shape[0] - 1
        row = np.array(initial[i], dtype=np.float64)          # (5, maxfits)
        for iperp in range(upper + 1):
            first = iperp == 0
            params = _pack(row, ft)
            if sp.relativistic:
                npb = f0r.shape[1] - 1
                lo, up = 1, npb - 1
                found_lo = found_up = False
                for ip in range(1, npb):
                    if not found_lo and f0r[iperp, ip - 1] <= -1.0 and f0r[iperp, ip] > -1.0:
                        lo, found_lo = ip, True
                    if not found_up and f0r[iperp, ip] > -1.0 and f0r[iperp, ip + 1] <= -1.0:
                        up, found_up = ip, True
                if up - lo > 2:
                    vals = f0r[iperp, lo:up + 1]
                    g = np.log(vals) if sp.logfit else vals.copy()
                    params, q = LM_nonlinear_fit(g, ft, pc, params, gr[iperp, lo:up + 1], pbr[iperp, lo:up + 1], sp.ms,
                                                 plasma.vA, sp.logfit, first, opt)
                else:
                    # too few points inside the cone: amplitude from the mid point (lines 613-624)
                    q = 0.0 if first else q   # the reference re-adds the previous row's quality here (line 654)
                    k = 0
                    for j, t in enumerate(ft):
                        params[k] = f0r[iperp, (up + lo) // 2] / np.exp(-pc[j] * gr[iperp, 1])
                        if t == 5:
                            params[k + 1] = 1.0e-12
                            params[k + 2] = 0.0
                        k += N_PARAMS[t]
            else:
                vals = plasma.f0[i, iperp, :]
                with np.errstate(all="ignore"):
                    g = np.log(vals) if sp.logfit else np.array(vals)
                params, q = LM_nonlinear_fit(g, ft, pc, params, plasma.pp[i, iperp, :, 0], plasma.pp[i, iperp, :, 1],
                                             sp.ms, plasma.vA, sp.logfit, first, opt)
            total += q
            _unpack(params, ft, row)
            pf[i, iperp, :, :len(ft)] = row[:, :len(ft)]
        # rows above this species' grid keep the last fitted row (param_fit(is,iperp,:,:) = row iperp-1 is only
        # applied inside the loop; the reference leaves them at their initial value 0)
    return pf, poly, total


# ---------------------------------------------------------------------------------- ACmethod = 2
def polynomial_basis(npar, order, kind=1):
    """set_polynomial_basis, src/ALPS_analyt.f90:850-889: Chebyshev T_n on yy = -1 + ipar * 2/npar."""
    if kind != 1:
        raise FitError("alps_error(10): unknown polynomial kind")
    yy = -1.0 + np.arange(npar + 1) * (2.0 / npar)
    P = np.zeros((npar + 1, order + 1))
    P[:, 0] = 1.0
    if order >= 1:
        P[:, 1] = yy
    for n in range(2, order + 1):
        P[:, n] = 2.0 * yy * P[:, n - 1] - P[:, n - 2]
    return P


def least_squares_fit(AA, BB):
    """Normal equations through dgemm / dgemv / dgesv, src/ALPS_analyt.f90:795-848."""
    alpha = AA.T @ AA
    beta = AA.T @ BB
    _, _, x, info = lapack.dgesv(alpha, beta)
    if info != 0:
        raise FitError("Error in dgesv: %d" % info)
    return x


def determine_GLLS(f0_is, order, logfit, kind=1):
    """determine_GLLS, src/ALPS_analyt.f90:731-793: one series per p_perp row; log10 of the row with the
    reference's floor for zeros (1 % of the smallest positive value, or 1e-40)."""
    nperp, npar = f0_is.shape[0] - 1, f0_is.shape[1] - 1
    P = polynomial_basis(npar, order, kind)
    out = np.zeros((nperp + 1, order + 1))
    for iperp in range(nperp + 1):
        row = np.array(f0_is[iperp, :], dtype=np.float64)
        if logfit:
            if row.min() > 0.0:
                fit = np.log10(row)
            else:
                min_val = row[row > 0.0].min() if np.any(row > 0.0) else 1.0e-40
                row[row == 0.0] = 0.01 * min_val
                with np.errstate(all="ignore"):
                    fit = np.log10(row)
        else:
            fit = row
        out[iperp] = least_squares_fit(P, fit)
    return out


def write_fit_parameters(path, params_rows):
    """distribution/<runname>.fit_parameters.<is>.out: `write (unit,*) iperp, params` per row (line 656)."""
    with open(path, "w") as fh:
        for iperp, p in enumerate(params_rows):
            fh.write(" %11d " % iperp + " ".join("%24.16E" % v for v in p) + "\n")
