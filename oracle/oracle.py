"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of the CPU oracle (oracle/alps_oracle.c, and
oracle/nhds_oracle.cpp for use_bM species).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by the product package alps_b200/.
PARITY UNPINNED at 1e-9 (see alps_oracle.h): pinned to the reference by its 5-digit goldens only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_NHDS = None


class OracleCfg(C.Structure):
    _fields_ = [("nspec", C.c_int), ("nperp", C.c_int), ("npar", C.c_int), ("ngamma", C.c_int),
                ("npparbar", C.c_int), ("vA", C.c_double), ("Bessel_zero", C.c_double),
                ("Tlim", C.c_double), ("positions_principal", C.c_int),
                ("n_resonance_interval", C.c_int), ("kperp_norm", C.c_int), ("nproc", C.c_int),
                ("maxfits", C.c_int), ("maxorder", C.c_int), ("nmax_force", C.c_int)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libalps_oracle.so")
    src = os.path.join(_HERE, "alps_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_bessj.restype = C.c_double
        _LIB.oracle_bessj.argtypes = [C.c_int, C.c_double]
        _LIB.oracle_int_ee.restype = C.c_double
        _LIB.oracle_int_ee_rel.restype = C.c_double
        _LIB.oracle_upload_rel.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.oracle_set_k.argtypes = [C.c_double, C.c_double, C.c_void_p]
        _LIB.oracle_set_species.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.c_int, C.c_double]
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    """flat float64 copy in Fortran element order"""
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


class Oracle:
    """disp() of the reference restated on the CPU.  The C library mirrors module alps_var in ONE static state: the
    newest Oracle owns it, and using an older instance afterwards raises instead of silently answering for the
    wrong plasma."""
    _owner = None

    def _check_owner(self):
        if Oracle._owner is not self:
            raise RuntimeError("this Oracle was superseded by a newer one (the C oracle has one global state); "
                               "create it again")

    def __init__(self, plasma, nproc: int = 0, threads: int = 0, nmax_force: int = 0):
        L = lib()
        Oracle._owner = self
        self.L = L
        self.pl = plasma
        maxorder = max(s.poly_order for s in plasma.species)
        cfg = OracleCfg(plasma.nspec, plasma.nperp, plasma.npar, plasma.ngamma, plasma.npparbar,
                        plasma.vA, plasma.Bessel_zero, plasma.Tlim, plasma.positions_principal,
                        plasma.n_resonance_interval, int(plasma.kperp_norm), nproc, plasma.maxfits,
                        maxorder, nmax_force)
        L.oracle_init(C.byref(cfg))
        for i, s in enumerate(plasma.species):
            ft = np.asarray(s.fit_type, dtype=np.int32)
            pc = np.asarray(s.perp_correction, dtype=np.float64)
            rc = L.oracle_set_species(i + 1, s.ns, s.qs, s.ms, int(s.relativistic), int(s.usebM),
                                      s.ACmethod, len(s.fit_type), _p(ft), _p(pc), int(s.logfit),
                                      s.poly_kind, s.poly_order, s.poly_log_max)
            if rc:
                raise RuntimeError("oracle_set_species failed: %d" % rc)
        pp = _f(plasma.pp)
        if plasma.df0 is None:
            df0 = np.zeros(plasma.nspec * (plasma.nperp - 1) * (plasma.npar - 1) * 2)
            L.oracle_derivative_f0(_p(_f(plasma.f0)), _p(pp), _p(df0), plasma.nspec, plasma.nperp,
                                   plasma.npar)
        else:
            df0 = _f(plasma.df0)
        self.df0_flat = df0
        rc = L.oracle_upload(_p(pp), _p(df0), _p(_f(plasma.param_fit)), _p(_f(plasma.poly_fit_coeffs)))
        if rc:
            raise RuntimeError("oracle_upload failed")
        if plasma.f0_rel is not None:
            L.oracle_upload_rel(plasma.f0_rel.shape[0], _p(_f(plasma.f0_rel)), _p(_f(plasma.df0_rel)),
                                _p(_f(plasma.gamma_rel)), _p(_f(plasma.pparbar_rel)))
        if threads:
            L.oracle_set_threads(threads)
        self.nmax = None

    def df0(self):
        pl = self.pl
        return self.df0_flat.reshape((pl.nspec, pl.nperp - 1, pl.npar - 1, 2), order="F")

    def set_k(self, kperp: float, kpar: float):
        self._check_owner()
        nmax = np.zeros(self.pl.nspec, dtype=np.int32)
        rc = self.L.oracle_set_k(kperp, kpar, _p(nmax))
        if rc:
            raise RuntimeError("oracle_set_k failed")
        self.nmax = nmax
        self.kperp, self.kpar = kperp, kpar
        return nmax

    def set_external_chi(self, is_: int, chi, chi_low):
        c = np.ascontiguousarray(np.asarray(chi, dtype=np.complex128).ravel(order="F"))
        cl = np.ascontiguousarray(np.asarray(chi_low, dtype=np.complex128).ravel(order="F"))
        self.L.oracle_set_external_chi(is_, _p(c.view(np.float64)), _p(cl.view(np.float64)))

    def set_ncap(self, ncap: int):
        self.L.oracle_set_ncap(ncap)

    def set_sample(self, stride: int, offset: int = 0):
        """bench sampling only: evaluate the harmonics |n| % stride == offset (stride <= 1: all)"""
        self.L.oracle_set_sample(int(stride), int(offset))

    def disp(self, om: complex, full: bool = False, nhds: bool = True):
        """disp(om) of the reference; use_bM species get their chi from the NHDS restatement
        (src/ALPS_fns.f90:344-362) unless nhds=False (the caller then feeds set_external_chi itself)."""
        self._check_owner()
        n = self.pl.nspec
        if nhds:
            for i, s in enumerate(self.pl.species):
                if s.usebM:
                    chi, low = nhds_calc_chi(s, om, self.kperp, self.kpar, bool(self.pl.kperp_norm))
                    self.set_external_chi(i + 1, chi, low)
        omv = np.array([om.real, om.imag])
        D = np.zeros(2)
        if not full:
            rc = self.L.oracle_disp(_p(omv), _p(D), None, None, None)
            if rc:
                raise RuntimeError("oracle_disp: alps_error(%d)" % rc)
            return complex(D[0], D[1])
        chi0 = np.zeros(n * 9 * 2)
        low = np.zeros(n * 27 * 2)
        wave = np.zeros(18)
        rc = self.L.oracle_disp(_p(omv), _p(D), _p(chi0), _p(low), _p(wave))
        if rc:
            raise RuntimeError("oracle_disp: alps_error(%d)" % rc)
        c = lambda a, shape: (a[0::2] + 1j * a[1::2]).reshape(shape, order="F")
        return (complex(D[0], D[1]), c(chi0, (n, 3, 3)), c(low, (n, 3, 3, 3)), c(wave, (3, 3)))

    def full_integrate(self, is_: int, nn: int, mode: int, om: complex):
        omv = np.array([om.real, om.imag])
        out = np.zeros(2)
        fr = C.c_int(0)
        self.L.oracle_full_integrate(is_, nn, mode, _p(omv), _p(out), C.byref(fr))
        return complex(out[0], out[1]), bool(fr.value)

    def eval_fit(self, is_: int, iperp: int, p: complex):
        pv = np.array([p.real, p.imag])
        out = np.zeros(2)
        self.L.oracle_eval_fit(is_, iperp, _p(pv), _p(out))
        return complex(out[0], out[1])

    def int_ee(self, is_: int) -> float:
        return self.L.oracle_int_ee(is_)

    def int_ee_rel(self, is_: int) -> float:
        return self.L.oracle_int_ee_rel(is_)

    def nlim(self):
        cap = 4096
        n = C.c_int(0)
        a = np.zeros(cap, dtype=np.int32)
        b = np.zeros(cap, dtype=np.int32)
        c = np.zeros(cap, dtype=np.int32)
        self.L.oracle_get_nlim(C.byref(n), _p(a), _p(b), _p(c), cap)
        return [(int(a[i]), int(b[i]), int(c[i])) for i in range(n.value)]


def bessj(n: int, x: float) -> float:
    return lib().oracle_bessj(n, x)


def nhds_calc_chi(species, om: complex, kperp: float, kpar: float, kperp_norm: bool = True):
    """calc_chi of src/ALPS_NHDS.f90:59-242 for a use_bM species, CPU restatement (oracle/nhds_oracle.hpp):
    returns chi(3,3), chi_low(3,3,-1:1)."""
    global _NHDS
    if _NHDS is None:
        build()
        so = os.path.join(_HERE, "libnhds_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _NHDS = C.CDLL(so)
        _NHDS.oracle_nhds_calc_chi.argtypes = ([C.c_double] * 3 + [C.c_int] + [C.c_double] * 6
                                               + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p])
    om = complex(om)
    x = np.array([om.real, om.imag])
    chi = np.zeros(9, dtype=np.complex128)
    low = np.zeros(27, dtype=np.complex128)
    s = species
    rc = _NHDS.oracle_nhds_calc_chi(s.ns, s.qs, s.ms, s.bM_nmaxs, s.bM_Bessel_zeros, s.bM_betas, s.bM_alphas,
                                    s.bM_pdrifts, kpar, kperp, _p(x), int(kperp_norm),
                                    _p(chi.view(np.float64)), _p(low.view(np.float64)))
    if rc:
        raise RuntimeError("oracle_nhds_calc_chi: cold-plasma species need kperp_norm=.true.")
    return chi.reshape((3, 3), order="F"), low.reshape((3, 3, 3), order="F")
