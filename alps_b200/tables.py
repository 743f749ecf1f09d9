"""Synthetic f0(p_perp, p_par) tables and the plasma description handed to the solver.

The closed forms follow the reference's table generator
(distribution/generate_distribution.f90:160-340): bi-Maxwellian (1), bi-kappa (2),
Juettner (3) and bi-Moyal (4) on the uniform grid p_perp[i] = i*dp_perp,
p_par[j] = j*dp_par - p_par_max + drift, together with the "ideal" analytic-continuation
fit parameters the generator prints (lines 195-263).  Arrays are kept in the reference's
Fortran layout (column-major, species index fastest; src/ALPS_var.f90:174-246) because that
is what the disp() boundary hands over (include/alps_b200.h).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


@dataclass
class DistSpec:
    """One &spec_j block of a *_dist.in file (distribution/generate_distribution.f90:424-470)."""
    ms: float = 1.0
    tau: float = 1.0
    alph: float = 1.0
    drift: float = 0.0
    kappa: float = 8.0
    distribution: int = 1
    autoscale: bool = True
    maxPperp: float = 1.0
    maxPpar: float = 1.0


@dataclass
class Species:
    """One &spec_j block of an ALPS .in file (src/ALPS_io.f90:281-340) plus its fits."""
    ns: float = 1.0
    qs: float = 1.0
    ms: float = 1.0
    relativistic: bool = False
    usebM: bool = False
    ACmethod: int = 1
    fit_type: List[int] = field(default_factory=lambda: [1])
    perp_correction: List[float] = field(default_factory=lambda: [1.0])
    logfit: bool = True
    poly_kind: int = 1
    poly_order: int = 0
    poly_log_max: float = 18.0
    # &bM_spec_j (src/ALPS_io.f90:342-372), used when usebM
    bM_nmaxs: int = 500
    bM_Bessel_zeros: float = 1.0e-50
    bM_betas: float = 1.0
    bM_alphas: float = 1.0
    bM_pdrifts: float = 0.0


@dataclass
class Plasma:
    """Everything disp() reads from module alps_var (src/ALPS_var.f90)."""
    nperp: int
    npar: int
    vA: float
    species: List[Species]
    pp: np.ndarray            # (nspec, nperp+1, npar+1, 2)  Fortran order
    f0: np.ndarray            # (nspec, nperp+1, npar+1)     Fortran order
    param_fit: np.ndarray     # (nspec, max(nperp,ngamma)+1, 5, maxfits) Fortran order
    df0: Optional[np.ndarray] = None  # (nspec, nperp-1, npar-1, 2) Fortran order
    poly_fit_coeffs: Optional[np.ndarray] = None
    # relativistic species only (src/ALPS_fns.f90:230-233): (nspec_rel, ngamma+1, npparbar+1[, 2])
    f0_rel: Optional[np.ndarray] = None
    df0_rel: Optional[np.ndarray] = None
    gamma_rel: Optional[np.ndarray] = None
    pparbar_rel: Optional[np.ndarray] = None
    ngamma: int = 0
    npparbar: int = 0
    Bessel_zero: float = 1.0e-50
    Tlim: float = 0.01
    positions_principal: int = 3
    n_resonance_interval: int = 100
    kperp_norm: bool = True
    fit_quality: float = 0.0   # qualitytotal of determine_param_fit when the fits were run (alps_b200.fits)

    @property
    def nspec(self) -> int:
        return len(self.species)

    @property
    def maxfits(self) -> int:
        return max(1, max(len(s.fit_type) for s in self.species))


def _bessk2(x: float) -> float:
    from scipy.special import kn
    return float(kn(2, x))


def generate_distribution(specs: Sequence[DistSpec], nperp: int, npar: int, beta: float = 1.0,
                          vA: float = 1.0e-4, maxP: float = 6.0):
    """f0 tables + ideal fit parameters, distribution/generate_distribution.f90:160-340.

    Returns (pp, f0, fits) with pp (nspec,nperp+1,npar+1,2) and f0 (nspec,nperp+1,npar+1) in
    Fortran order and fits[is] = dict(fit_type, params[5], perpcorr).
    """
    nspec = len(specs)
    pp = np.zeros((nspec, nperp + 1, npar + 1, 2), order="F")
    f0 = np.zeros((nspec, nperp + 1, npar + 1), order="F")
    fits = []
    pi = math.atan(1.0) * 4.0
    kappa1 = specs[0].kappa
    for i, sp in enumerate(specs):
        ms, tau, alph, drift, kap = sp.ms, sp.tau, sp.alph, sp.drift, sp.kappa
        ifit = [0.0] * 5
        iperpcorr = 1.0
        if sp.distribution == 0:
            # table from distribution/distribution_analyt.f90:66-85 (hard-coded beta = 1 Maxwellians of
            # species 1 and 2), no normalisation, limits always maxPperpS / maxPparS (lines 177-181)
            if i > 1:
                raise ValueError("distribution_analyt defines species 1 and 2 only")
            norm = 1.0
            pperp_max, ppar_max = sp.maxPperp, sp.maxPpar
            ftype = 1
        elif sp.distribution == 1:
            norm = pi ** (-1.5) / ((ms * beta * tau) ** 1.5 * alph)
            pperp_max = maxP * math.sqrt(ms * tau * alph)
            ppar_max = maxP * math.sqrt(ms * tau)
            ifit = [norm, 1.0 / (beta * ms * tau), drift, 0.0, 0.0]
            iperpcorr = 1.0 / (tau * beta * ms * alph)
            ftype = 1
        elif sp.distribution == 2:
            a = math.sqrt((2.0 * kap - 3.0) / (2.0 * kap))
            norm = 1.0 / ((ms * beta * tau * pi * kap) ** 1.5 * alph)
            norm = norm * math.gamma(kap + 1.0) / (math.gamma(kap - 0.5) * a ** 3)
            scale = math.sqrt((kap - 1.5) / (kappa1 - 1.5))
            pperp_max = maxP * math.sqrt(ms * tau * alph) * scale
            ppar_max = maxP * math.sqrt(ms * tau) * scale
            ifit = [norm, 1.0 / (beta * ms * kap * a * a * tau), drift, -1.0 - kap, 1.0]
            iperpcorr = 1.0 / (tau * beta * ms * kap * a * a * alph)
            ftype = 2
        elif sp.distribution == 3:
            norm = vA / (2.0 * pi * math.sqrt(alph) * ms ** 2 * beta * tau)
            norm = norm / _bessk2(2.0 * ms / (vA * vA * alph * beta * tau))
            pperp_max = math.sqrt(maxP * maxP * tau + (tau - ms * ms) / (vA * vA)) * math.sqrt(alph)
            ppar_max = math.sqrt(maxP * maxP * tau + (tau - ms * ms) / (vA * vA))
            ifit = [norm, vA * vA * alph / (ms * ms), drift, 0.0, 0.0]
            iperpcorr = 2.0 * ms / (vA * vA * beta * tau * alph)
            ftype = 3
        elif sp.distribution == 4:
            norm = 1.0
            pperp_max = maxP * math.sqrt(ms * tau * alph)
            ppar_max = maxP * math.sqrt(ms * tau)
            ifit = [1.0, 1.0 / (beta * ms * tau), drift, 1.0, 0.0]
            iperpcorr = 1.0 / (tau * beta * ms * alph)
            ftype = 6
        else:
            raise ValueError("distribution type %d not supported" % sp.distribution)
        if not sp.autoscale and sp.distribution != 0:
            pperp_max, ppar_max = sp.maxPperp, sp.maxPpar
        dpperp = pperp_max / float(nperp)
        dppar = 2.0 * ppar_max / float(npar)
        pperp = np.arange(nperp + 1, dtype=np.float64) * dpperp
        ppar = np.arange(npar + 1, dtype=np.float64) * dppar - ppar_max + drift
        P, Q = np.meshgrid(pperp, ppar, indexing="ij")
        dq = Q - drift
        if sp.distribution == 0:
            ms_a = 1.0 if i == 0 else 1.0 / 1836.0
            f = (pi ** (-1.5) / ((ms_a * 1.0) ** 1.5)) * np.exp(-(Q * Q / (1.0 * ms_a) + (P * P) / (1.0 * ms_a)))
        elif sp.distribution == 1:
            f = np.exp(-((dq * dq) / (beta * ms * tau) + (P * P) / (tau * beta * ms * alph)))
        elif sp.distribution == 2:
            f = (1.0 + (dq * dq) / (beta * ms * kap * a * a * tau)
                 + (P * P) / (tau * beta * ms * kap * a * a * alph)) ** (-1.0 - kap)
        elif sp.distribution == 3:
            f = np.exp(-(2.0 * ms / (vA * vA * beta * tau * alph)) *
                       np.sqrt(1.0 + P * P * vA * vA / (ms * ms) + dq * dq * vA * vA * alph / (ms * ms)))
        else:
            e = (dq * dq) / (beta * ms * tau) + (P * P) / (tau * beta * ms * alph)
            f = np.exp(0.5 * (e - np.exp(e)))
        f = f * norm
        if sp.distribution == 4:
            integ = float(np.sum(dpperp * dppar * 2.0 * pi * P * f))
            norm = 1.0 / integ
            ifit[0] = norm
            f = f * norm
        pp[i, :, :, 0] = P
        pp[i, :, :, 1] = Q
        f0[i] = f
        fits.append(dict(fit_type=ftype, params=ifit, perpcorr=iperpcorr))
    return pp, f0, fits


def ideal_param_fit(fits, nperp: int, ngamma: int = 0) -> np.ndarray:
    """param_fit(nspec,0:max(nperp,ngamma),5,maxfits) filled with the ideal parameters for every
    iperp (what determine_param_fit converges to, tests/test_kpar_fast.out:71-77)."""
    nspec = len(fits)
    n = max(nperp, ngamma) + 1
    pf = np.zeros((nspec, n, 5, 1), order="F")
    for i, ft in enumerate(fits):
        for k in range(5):
            pf[i, :, k, 0] = ft["params"][k]
    return pf


def make_plasma(specs: Sequence[DistSpec], ns: Sequence[float], qs: Sequence[float], nperp: int,
                npar: int, beta: float = 1.0, vA: float = 1.0e-4, maxP: float = 6.0, **kw) -> Plasma:
    pp, f0, fits = generate_distribution(specs, nperp, npar, beta, vA, maxP)
    species = [Species(ns=ns[i], qs=qs[i], ms=specs[i].ms, ACmethod=1,
                       fit_type=[fits[i]["fit_type"]], perp_correction=[fits[i]["perpcorr"]])
               for i in range(len(specs))]
    return Plasma(nperp=nperp, npar=npar, vA=vA, species=species, pp=pp, f0=f0,
                  param_fit=ideal_param_fit(fits, nperp), **kw)


# ---------------------------------------------------------------- named configurations
def config_kpar_fast() -> Plasma:
    """C1: tests/test_kpar_fast.in + distribution/test_kpar_fast_dist.in (bi-Maxwellian p+e)."""
    specs = [DistSpec(ms=1.0), DistSpec(ms=5.44662e-4)]
    return make_plasma(specs, ns=[1.0, 1.0], qs=[1.0, -1.0], nperp=120, npar=240,
                       Bessel_zero=1.0e-50)


def config_kappa3(nperp: int = 1024, npar: int = 2048, kappa: float = 8.0) -> Plasma:
    """C5: synthetic 3-species bi-kappa plasma."""
    specs = [DistSpec(ms=1.0, kappa=kappa, distribution=2),
             DistSpec(ms=5.44662e-4, kappa=kappa, distribution=2),
             DistSpec(ms=4.0, kappa=kappa, distribution=2)]
    return make_plasma(specs, ns=[1.0, 1.04, 0.02], qs=[1.0, -1.0, 2.0], nperp=nperp, npar=npar,
                       Bessel_zero=1.0e-45)


def config_bimax(nperp: int = 150, npar: int = 300) -> Plasma:
    """C2: tests/test_bimax.in -- protons use_bM=T (closed-form NHDS chi), electrons from the f0 table."""
    specs = [DistSpec(ms=1.0), DistSpec(ms=5.44662e-4)]
    pl = make_plasma(specs, ns=[1.0, 1.0], qs=[1.0, -1.0], nperp=nperp, npar=npar, Bessel_zero=1.0e-45)
    pl.species[0].usebM = True
    pl.pp[0] = 0.0      # read_f0 zeroes the tables of use_bM species (src/ALPS_io.f90:672-676)
    pl.f0[0] = 0.0
    return pl


def config_small(nperp: int = 24, npar: int = 48, kind: int = 1) -> Plasma:
    """Small two-species case for fast parity tests."""
    specs = [DistSpec(ms=1.0, distribution=kind, alph=1.3), DistSpec(ms=5.44662e-4, distribution=kind)]
    return make_plasma(specs, ns=[1.0, 1.0], qs=[1.0, -1.0], nperp=nperp, npar=npar,
                       Bessel_zero=1.0e-30)


def config_relativistic(nperp: int = 30, npar: int = 60, ngamma: int = 500, npparbar: int = 500,
                        rel_backend: str = "host") -> Plasma:
    """C3: tests/test_relativistic.in + distribution/test_relativistic_dist.in (Juettner pair plasma,
    vA = 1, both species relativistic, fit type 4).  The (Gamma, pbar_par) tables come from the host-side
    restatement of derivative_f0_rel (alps_b200/relativistic.py)."""
    from .relativistic import derivative_f0_rel
    vA = 1.0
    specs = [DistSpec(ms=1.0, distribution=3), DistSpec(ms=1.0, distribution=3)]
    pp, f0, fits = generate_distribution(specs, nperp, npar, beta=1.0, vA=vA, maxP=5.0)
    nrel = 2
    shape = (nrel, ngamma + 1, npparbar + 1)
    f0_rel, gam, pb = (np.zeros(shape, order="F") for _ in range(3))
    df0_rel = np.zeros(shape + (2,), order="F")
    pf = np.zeros((2, max(nperp, ngamma) + 1, 5, 1), order="F")
    species = []
    cache = None
    for i in range(2):
        if cache is None:     # both species have the same f0 table: one thin-plate-spline solve
            cache = derivative_f0_rel(pp[i], f0[i], specs[i].ms, vA, ngamma, npparbar, backend=rel_backend)
        g, p, f, d, integ = cache
        gam[i], pb[i], f0_rel[i], df0_rel[i] = g, p, f, d
        pf[i, :, 0, 0] = fits[i]["params"][0] / integ          # amplitude of the renormalised table
        species.append(Species(ns=1.0, qs=1.0 if i == 0 else -1.0, ms=specs[i].ms, relativistic=True,
                               ACmethod=1, fit_type=[4], perp_correction=[fits[i]["perpcorr"]]))
    return Plasma(nperp=nperp, npar=npar, vA=vA, species=species, pp=pp, f0=f0, param_fit=pf,
                  f0_rel=f0_rel, df0_rel=df0_rel, gamma_rel=gam, pparbar_rel=pb, ngamma=ngamma,
                  npparbar=npparbar, Bessel_zero=1.0e-45, positions_principal=5)
