"""NHDS (calc_chi for use_bM species).  CPU: the oracle restatement (oracle/nhds_oracle.hpp) against an
independent evaluation of the same hot bi-Maxwellian susceptibility with scipy's Faddeeva and Bessel
functions, and against the table quadrature of the oracle for the same bi-Maxwellian.  GPU: the product's
device kernels (alps_b200/csrc/nhds_kernel.cu, through the C ABI) against the oracle restatement."""
import numpy as np
import pytest
from scipy.special import ive, wofz

from alps_b200 import tables
from oracle.oracle import nhds_calc_chi


def _stix_chi(sp, om, kperp, kz):
    """Stix (10-57) with exp(-z)-scaled I_n, NHDS normalisation (kperp_norm), n in [-60, 60]."""
    Omega, ell2 = sp.qs / sp.ms, sp.ms / (sp.ns * sp.qs * sp.qs)
    vth, vd, al = np.sqrt(sp.bM_betas / (sp.ns * sp.ms)), sp.bM_pdrifts / sp.ms, sp.bM_alphas
    z = 0.5 * (kperp * vth / Omega) ** 2 * al
    Y = np.zeros((3, 3), dtype=complex)
    for n in range(-60, 61):
        zeta = (om - kz * vd - n * Omega) / (kz * vth)
        Z = 1j * np.sqrt(np.pi) * wofz(zeta)
        res = om - kz * vd - n * Omega
        An = (al - 1.0) + (al * res + n * Omega) * Z / (kz * vth)
        Bn = (al * (om - n * Omega) - (kz * vd - n * Omega)) / kz + (om - n * Omega) * (al * res + n * Omega) * Z / (kz * kz * vth)
        I = ive(abs(n), z)
        dI = 0.5 * (ive(abs(n + 1), z) + ive(abs(n - 1), z))
        Y[0, 0] += n * n * I * An / z
        Y[0, 1] += -1j * n * (I - dI) * An
        Y[0, 2] += kperp * n * I * Bn / (Omega * z)
        Y[1, 1] += (n * n * I / z + 2 * z * I - 2 * z * dI) * An
        Y[1, 2] += 1j * kperp * (I - dI) * Bn / Omega
        Y[2, 2] += 2 * (om - n * Omega) * I * Bn / (kz * vth * vth * al)
    chi = Y / ell2
    chi[2, 2] += 2 * om * vd / (ell2 * kz * vth * vth * al)
    return chi


@pytest.mark.parametrize("qs,ms,alpha,drift", [(1.0, 1.0, 1.0, 0.0), (-1.0, 5.44662e-4, 1.0, 0.0), (1.0, 1.0, 2.0, 0.3)])
def test_calc_chi_against_scipy(qs, ms, alpha, drift):
    sp = tables.Species(ns=1.0, qs=qs, ms=ms, usebM=True, bM_betas=1.0, bM_alphas=alpha, bM_pdrifts=drift)
    for om in (0.3 + 0.01j, 0.9 - 0.05j, 1.7 + 0.0j):
        for kperp, kz in ((0.1, 0.2), (1.0, 0.5)):
            chi, low = nhds_calc_chi(sp, om, kperp, kz)
            ref = _stix_chi(sp, om, kperp, kz)
            for (i, j) in ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)):
                # BESSI is the ~1e-7 Numerical-Recipes polynomial, WOFZ is 13-14 digits
                assert abs(chi[i, j] - ref[i, j]) <= 2e-6 * np.max(np.abs(ref)), (om, kperp, kz, i, j)
            assert np.allclose(chi[1, 0], -chi[0, 1]) and np.allclose(chi[2, 0], chi[0, 2])
            # n = 0, +-1 pieces add up to less than the whole and are finite
            assert np.all(np.isfinite(low))


def test_cold_plasma_limit():
    sp = tables.Species(ns=1.0, qs=1.0, ms=1.0, usebM=True, bM_betas=0.0)
    chi, low = nhds_calc_chi(sp, 0.5 + 0.0j, 0.1, 0.2)
    R, L = -0.5 / 1.5, -0.5 / (0.5 - 1.0)
    assert np.isclose(chi[0, 0], 0.5 * (R + L)) and np.isclose(chi[2, 2], -1.0)
    assert np.all(low == 0)


def test_table_quadrature_agrees_with_the_closed_form():
    """Physics cross-check of the whole chain: the oracle's table quadrature of a bi-Maxwellian f0 and
    the NHDS closed form describe the same plasma.  Real omega: the Landau term carries the whole
    resonant part there (for Im(om) != 0 the reference's linearly interpolated funct_g underestimates
    the zz pole term of the electrons, whose resonance sits within half a cell of p_par = 0 -- a
    property of the reference's algorithm that parity reproduces)."""
    from oracle.oracle import Oracle
    pl = tables.config_kpar_fast()
    orc = Oracle(pl)
    kperp, kpar, om = 1.0e-2, 1.0e-2, 9.98811e-3 + 0.0j
    orc.set_k(kperp, kpar)
    _, chi0, _, _ = orc.disp(om, full=True)
    for i, s in enumerate(pl.species):
        sp = tables.Species(ns=s.ns, qs=s.qs, ms=s.ms, usebM=True, bM_betas=1.0, bM_alphas=1.0)
        chi, _ = nhds_calc_chi(sp, om, kperp, kpar)
        chi_tab = chi0[i] * (om * om * pl.vA * pl.vA)
        assert abs(chi_tab[2, 2] - chi[2, 2]) < 1e-2 * abs(chi[2, 2])
        assert abs(chi_tab[2, 2].imag - chi[2, 2].imag) < 1e-2 * abs(chi[2, 2].imag)
        assert abs(chi_tab[0, 1] - chi[0, 1]) < 1e-2 * abs(chi[0, 1])


@pytest.mark.gpu
def test_device_calc_chi_matches_the_oracle():
    """k_nhds_bessel + k_nhds (alps_b200_nhds_calc_chi) against the CPU restatement: all nine entries of chi and
    the n = 0, +-1 pieces, hot / drifting / anisotropic / cold species, both signs of k_par, kperp_norm on and off,
    upper and lower half plane and real omega, harmonic cut-offs from 2 to the bM_nmaxs cap."""
    from alps_b200.solver import nhds_calc_chi as dev_chi
    cases = [dict(qs=1.0, ms=1.0), dict(qs=-1.0, ms=5.44662e-4), dict(qs=1.0, ms=1.0, bM_alphas=2.0, bM_pdrifts=0.3),
             dict(qs=2.0, ms=4.0, ns=0.05, bM_betas=0.4, bM_alphas=0.5, bM_pdrifts=-0.2),
             dict(qs=1.0, ms=1.0, bM_nmaxs=7, bM_Bessel_zeros=1e-300), dict(qs=1.0, ms=1.0, bM_betas=0.0, bM_pdrifts=0.1)]
    worst = 0.0
    for kw in cases:
        sp = tables.Species(usebM=True, **{"ns": 1.0, "bM_betas": 1.0, **kw})
        for om in (0.3 + 0.01j, 0.9 - 0.05j, 1.7 + 0.0j, 3.0e-2 - 1.0e-5j, 12.0 + 3.0j, 0.2 - 0.6j):
            for kperp, kz in ((0.1, 0.2), (1.0, 0.5), (1e-3, -0.03), (6.0, 0.05)):
                for norm in (True, False):
                    if sp.bM_betas == 0.0 and not norm:
                        continue
                    chi_o, low_o = nhds_calc_chi(sp, om, kperp, kz, norm)
                    chi_d, low_d = dev_chi(sp, om, kperp, kz, norm)
                    scale = np.max(np.abs(chi_o))
                    assert np.all(np.isfinite(chi_o))
                    err = max(np.max(np.abs(chi_d - chi_o)), np.max(np.abs(low_d - low_o))) / scale
                    worst = max(worst, err)
                    assert err < 1e-12, (kw, om, kperp, kz, norm, err)
    print("worst relative difference device vs oracle: %.2e" % worst)
