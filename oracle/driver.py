"""TEST INFRASTRUCTURE ONLY: Python restatement of the reference's root drivers, used to pin
the oracle's disp() against the reference's golden scan files.

secant_osc  -> src/ALPS_fns.f90:1919-2101
secant      -> src/ALPS_fns.f90:1815-1917
om_scan     -> src/ALPS_fns.f90:2198-2600 (wave-vector stepping only; scan types 3 and 4)
"""
from __future__ import annotations

import math

import numpy as np

_F01 = float(np.float32(0.1))      # REAL*4 literals promoted to double in the reference
_F1EM3 = float(np.float32(1.0e-3))


def secant_osc(disp, om: complex, numiter: int, D_threshold: float, D_prec: float) -> complex:
    delta = complex(1.0e-6, 1.0e-8)
    lam = _F01
    osc_threshold = _F1EM3
    D = disp(om)
    minom, minD = om, D
    prevom = om * (1.0 - D_prec)
    prev2om = prev3om = prev4om = om
    prevD = disp(prevom)
    if abs(prevD) < abs(minD):
        minom, minD = prevom, prevD
    it = 0
    go = True
    damping = 1.0
    osc = 0
    while it <= numiter - 1 and go:
        it += 1
        D = disp(om)
        if abs(D - prevD) < 1.0e-80:
            prevom = prevom + 1.0e-8
            prevD = disp(prevom)
        if abs(D) < D_threshold:
            go = False
        else:
            if it > 4:
                def close(a):
                    return (abs(om.real - a.real) < abs(om.real) * osc_threshold and
                            abs(om.imag - a.imag) < abs(om.imag) * osc_threshold)
                if close(prevom) or close(prev2om) or close(prev3om) or close(prev4om):
                    osc += 1
                    damping = min(0.5, damping * 0.75)
            if osc > 1:
                Dprime = (disp(om * (1.0 + delta)) - disp(om * (1.0 - delta))) / (2.0 * om * delta)
                jump = D / (Dprime + lam * D)
            else:
                jump = damping * D * (om - prevom) / (D - prevD)
            if abs(jump) > _F01 * abs(om):
                jump = (_F01 * abs(om)) * (jump / abs(jump))
            if abs(D) > abs(prevD):
                jump = 0.5 * jump
            prev4om, prev3om, prev2om, prevom = prev3om, prev2om, prevom, om
            prevD = D
            if abs(D) < abs(minD):
                minom, minD = om, D
            om = om - jump
    if it >= numiter:
        om = minom
    return om


def secant(disp, om: complex, numiter: int, D_threshold: float, D_prec: float) -> complex:
    prevom = om * (1.0 - D_prec)
    Dprev = disp(prevom)
    minD = disp(om)
    minom = om
    if abs(Dprev) < abs(minD):
        minom, minD = prevom, Dprev
    it = 0
    go = True
    while it <= numiter - 1 and go:
        it += 1
        D = disp(om)
        if abs(D - Dprev) < 1.0e-80:
            prevom = prevom + 1.0e-8
            Dprev = disp(prevom)
        if abs(D) < D_threshold:
            jump = 0.0
            go = False
        else:
            jump = D * (om - prevom) / (D - Dprev)
        prevom = om
        om = om - jump
        Dprev = D
        if abs(D) < abs(minD):
            minom, minD = om, D
    if it >= numiter:
        om = minom
    return om


def scan_k(oracle, kperp0: float, kpar0: float, scan_type: int, swf: float, nsteps: int, log: bool,
           guess: complex, numiter: int, D_threshold: float, D_prec: float, method=secant_osc):
    """refine_guess at (kperp0,kpar0) then om_scan of type 3 (kperp) / 4 (kpar), one root.
    Returns rows (kperp, kpar, omega) like the .scan_* file."""
    oracle.set_k(kperp0, kpar0)
    om = method(oracle.disp, guess, numiter, D_threshold, D_prec)
    rows = [(kperp0, kpar0, om)]
    last = kperp0 if scan_type == 3 else kpar0
    diff = (math.log10(swf) - math.log10(last)) / nsteps if log else (swf - last) / nsteps
    for it in range(1, nsteps + 1):
        k = 10.0 ** (math.log10(last) + diff * it) if log else last + diff * it
        kperp, kpar = (k, kpar0) if scan_type == 3 else (kperp0, k)
        oracle.set_k(kperp, kpar)
        om = method(oracle.disp, om, numiter, D_threshold, D_prec)
        rows.append((kperp, kpar, om))
    return rows


def calc_eigen(oracle, omega: complex, kperp: float, kpar: float, vA: float, ns, qs, current_int=None, split=False):
    """calc_eigen, src/ALPS_fns.f90:2605-2899, on top of the oracle's disp(full=True): electric and magnetic
    eigenfunctions (E_x = 1), species velocity and density fluctuations, heating rates P_s and the wave energy W_EM
    (chi at real omega and at 1.000001 omega for d(chi_h)/d(omega)).  Returns (ef[3], bf[3], Us[nspec,3], ds[nspec],
    Ps[nspec], W_EM) -- the columns of the .eigen_* and .heat_* files (:2334-2340)."""
    nspec = len(ns)
    _, chi0, _, w = oracle.disp(omega, full=True)
    e = np.zeros(3, dtype=complex)
    e[0] = 1.0
    e[2] = -e[0] * (w[1, 0] * w[2, 1] - w[2, 0] * w[1, 1]) / (w[1, 2] * w[2, 1] - w[2, 2] * w[1, 1])
    if abs(w[2, 1]) != 0.0:
        e[1] = (-e[2] * w[2, 2] - e[0] * w[2, 0]) / w[2, 1]
    else:
        e[1] = (w[1, 0] * w[0, 2] - w[0, 0] * w[1, 2]) / (w[1, 2] * w[0, 1] - w[1, 1] * w[0, 2])
    b = np.array([-kpar * e[1], -(kperp * e[2] - kpar * e[0]), kperp * e[1]]) / (omega * vA)
    pflow = np.zeros(nspec) if current_int is None else np.asarray(current_int) / (np.asarray(ns) * np.asarray(qs))
    Us = np.zeros((nspec, 3), dtype=complex)
    ds = np.zeros(nspec, dtype=complex)
    for s in range(nspec):
        base = [-(vA * vA / (qs[s] * ns[s])) * 1j * omega * sum(e[q] * chi0[s, j, q] for q in range(3)) for j in range(3)]
        Us[s] = base
        if pflow[s] != 0.0:
            Us[s, 2] = (base[2] - pflow[s] * kperp * Us[s, 0] / (omega - kpar * pflow[s])) / \
                       (1.0 + (kpar * pflow[s]) / (omega - kpar * pflow[s]))
        ds[s] = (1.0 / vA) * (Us[s, 0] * kperp + Us[s, 2] * kpar) / (omega - kpar * pflow[s])
    # heating: anti-Hermitian part of chi at real omega, Hermitian part at omega and 1.000001 omega
    _, c_r, _, _ = oracle.disp(complex(omega.real, 0.0), full=True)
    chia = np.array([-0.5j * (c_r[s] - c_r[s].conj().T) for s in range(nspec)])
    chih_old = 0.5 * sum(c_r[s] + c_r[s].conj().T for s in range(nspec))
    Psc = np.array([np.conj(e) @ chia[s] @ e for s in range(nspec)])
    _, c_p, low, _ = oracle.disp(complex((omega * 1.000001).real, 0.0), full=True)
    chih = 0.5 * sum(c_p[s] + c_p[s].conj().T for s in range(nspec))
    dchih = (1.000001 * chih - chih_old) / 0.000001
    W_EM = float((np.conj(e) @ dchih @ e + np.sum(b * np.conj(b))).real)
    if not split:
        return e, b, Us, ds, Psc.real / W_EM, W_EM
    # heating split by mechanism (:2800-2890) from chi0_low of the LAST disp call (the one at 1.000001 Re omega):
    # n = 0 yy and yz parts (transit-time damping), n = 0 zy and zz parts (Landau damping), n = +1 and n = -1
    # (cyclotron) with the perpendicular field only; columns of the .heat_mech_* file, four per species
    Ps_split = np.zeros((nspec, 4))
    exy = np.array([e[0], e[1], 0.0])
    for s in range(nspec):
        l0 = low[s, :, :, 1]
        Ps_split[s, 0] = ((-0.5j * np.conj(e[1]) * e[1] * (l0[1, 1] - np.conj(l0[1, 1]))).real +
                          (-0.5j * (e[2] * np.conj(e[1]) * l0[1, 2] - np.conj(e[2]) * e[1] * np.conj(l0[1, 2]))).real)
        Ps_split[s, 1] = ((-0.5j * (e[1] * np.conj(e[2]) * l0[2, 1] - np.conj(e[1]) * e[2] * np.conj(l0[2, 1]))).real +
                          (-0.5j * np.conj(e[2]) * e[2] * (l0[2, 2] - np.conj(l0[2, 2]))).real)
        for col, m in ((2, 2), (3, 0)):          # n = +1, n = -1
            lm = low[s, :, :, m]
            Ps_split[s, col] = (np.conj(exy) @ (-0.5j * (lm - lm.conj().T)) @ exy).real
    return e, b, Us, ds, Psc.real / W_EM, W_EM, Ps_split / W_EM
