"""Host-side mirror of the reference's disp() interface over the C ABI.

Names follow the reference (src/ALPS_fns.f90): `disp`, `derivative_f0`, `determine_nmax`
(inside `set_k`), `map_search`, `refine_guess`, `secant_osc`, `om_scan`.  All numerical work
happens in libalps_b200.so on the GPU; this module only marshals arrays.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .tables import Plasma


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    """flat float64 copy in Fortran element order (what the Fortran side would hand over)"""
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


class Solver:
    """One plasma (set of f0 tables) resident on one GPU.  Not re-entrant: the C ABI holds one
    global instance, like the reference's module state."""

    def __init__(self, plasma: Plasma, emulate_nproc: int = 0, device: int = -1, batch_max: int = 0,
                 nmax_cap: int = 0, nmax_force: int = 0, ngpu: int = 1):
        """ngpu > 1: device group -- this process drives `ngpu` devices starting at `device`; the library keeps the
        tables on every one and partitions disp_batch / map_search itself (set_partition)."""
        self.L = _lib.lib()
        self.pl = plasma
        maxorder = max(s.poly_order for s in plasma.species)
        cfg = _lib.Cfg(plasma.nspec, plasma.nperp, plasma.npar, plasma.ngamma, plasma.npparbar,
                       plasma.vA, plasma.Bessel_zero, plasma.Tlim, plasma.positions_principal,
                       plasma.n_resonance_interval, int(plasma.kperp_norm), emulate_nproc,
                       plasma.maxfits, maxorder, device, nmax_cap, batch_max, nmax_force, ngpu)
        _lib.check(self.L.alps_b200_init(C.byref(cfg)))
        for i, s in enumerate(plasma.species):
            ft = np.asarray(s.fit_type, dtype=np.int32)
            pc = np.asarray(s.perp_correction, dtype=np.float64)
            _lib.check(self.L.alps_b200_set_species(
                i + 1, s.ns, s.qs, s.ms, int(s.relativistic), int(s.usebM), s.ACmethod,
                len(s.fit_type), _p(ft), _p(pc), int(s.logfit), s.poly_kind, s.poly_order,
                s.poly_log_max))
            if s.usebM:
                _lib.check(self.L.alps_b200_set_bm_species(i + 1, s.bM_nmaxs, s.bM_Bessel_zeros, s.bM_betas,
                                                           s.bM_alphas, s.bM_pdrifts))
        pp = _f(plasma.pp)
        df0 = _f(plasma.df0)
        _lib.check(self.L.alps_b200_upload(_p(pp), _p(df0), _p(_f(plasma.param_fit)),
                                           _p(_f(plasma.poly_fit_coeffs))))
        if plasma.f0_rel is not None:
            _lib.check(self.L.alps_b200_upload_rel(plasma.f0_rel.shape[0], _p(_f(plasma.f0_rel)),
                                                   _p(_f(plasma.df0_rel)), _p(_f(plasma.gamma_rel)),
                                                   _p(_f(plasma.pparbar_rel))))
        self._df0 = None
        if df0 is None:
            self._df0 = self.derivative_f0(plasma.f0)
        self.nmax = None
        self.kperp = self.kpar = None

    # ---- derivative_f0, src/ALPS_fns.f90:34-248
    def derivative_f0(self, f0) -> np.ndarray:
        pl = self.pl
        out = np.zeros(pl.nspec * (pl.nperp - 1) * (pl.npar - 1) * 2)
        _lib.check(self.L.alps_b200_derivative_f0(_p(_f(f0)), _p(out)))
        return out.reshape((pl.nspec, pl.nperp - 1, pl.npar - 1, 2), order="F")

    def df0(self):
        return self._df0

    # ---- determine_nmax / split_processes / determine_bessel_array, src/ALPS_fns.f90:3971-4255
    def set_k(self, kperp: float, kpar: float) -> np.ndarray:
        nmax = np.zeros(self.pl.nspec, dtype=np.int32)
        _lib.check(self.L.alps_b200_set_k(float(kperp), float(kpar), _p(nmax)))
        self.nmax = nmax
        self.kperp, self.kpar = float(kperp), float(kpar)
        return nmax

    def set_mode(self, mode: int):
        """0 = direct quadrature per omega (default); 1 = k-hoisted tables ("map fast path").
        Call set_k afterwards."""
        _lib.check(self.L.alps_b200_set_mode(mode))

    def set_partition(self, kind: int):
        """_lib.PARTITION_OMEGA (default: batches are cut into one slice per GPU) or _lib.PARTITION_HARMONIC (every GPU
        sums a block of harmonics, the chi partials are summed over NVLink).  Call set_k afterwards."""
        _lib.check(self.L.alps_b200_set_partition(kind))

    def comm_init(self, rank: int, world: int, broadcast):
        """One process per GPU: join the library-owned NCCL communicator.  `broadcast(buf)` must overwrite the 128-byte
        numpy uint8 array `buf` on every rank with rank 0's content (torch.distributed.broadcast, MPI_Bcast)."""
        idb = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.check(self.L.alps_b200_comm_unique_id(_p(idb)))
        broadcast(idb)
        _lib.check(self.L.alps_b200_comm_init(rank, world, _p(idb)))

    def comm_finalize(self):
        _lib.check(self.L.alps_b200_comm_finalize())

    def comm_init_torch(self, group=None):
        """comm_init over an initialised torch.distributed process group (any backend)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)

        def bcast(buf):
            t = torch.from_numpy(buf)
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=0, group=group)
            buf[:] = t.cpu().numpy()
        self.comm_init(rank, world, bcast)

    def set_map_mode(self, mode: int):
        """formulation of map_search's batch: 1 = k-hoisted tables (default), 0 = the current mode (direct)"""
        _lib.check(self.L.alps_b200_set_map_mode(mode))

    def set_harmonic_shard(self, rank: int, nranks: int):
        _lib.check(self.L.alps_b200_set_harmonic_shard(rank, nranks))

    # ---- disp(om), src/ALPS_fns.f90:252-636
    def disp(self, om: complex, full: bool = False):
        n = self.pl.nspec
        omv = np.array([om.real, om.imag])
        D = np.zeros(2)
        if not full:
            _lib.check(self.L.alps_b200_disp(_p(omv), _p(D), None, None, None))
            return complex(D[0], D[1])
        chi0 = np.zeros(n * 18)
        low = np.zeros(n * 54)
        wave = np.zeros(18)
        _lib.check(self.L.alps_b200_disp(_p(omv), _p(D), _p(chi0), _p(low), _p(wave)))
        c = lambda a, shape: (a[0::2] + 1j * a[1::2]).reshape(shape, order="F")
        return (complex(D[0], D[1]), c(chi0, (n, 3, 3)), c(low, (n, 3, 3, 3)), c(wave, (3, 3)))

    def disp_batch(self, om: Sequence[complex], want_chi0: bool = False):
        """n independent omegas (host buffers in, host buffers out)."""
        om = np.ascontiguousarray(np.asarray(om, dtype=np.complex128).ravel())
        n = om.size
        D = np.zeros(n, dtype=np.complex128)
        chi0 = np.zeros(n * self.pl.nspec * 9, dtype=np.complex128) if want_chi0 else None
        _lib.check(self.L.alps_b200_disp_batch(n, _p(om.view(np.float64)), _p(D.view(np.float64)),
                                               _p(chi0.view(np.float64)) if want_chi0 else None))
        if want_chi0:
            return D, chi0.reshape((n, 3, 3, self.pl.nspec)).transpose(0, 3, 1, 2).swapaxes(2, 3)
        return D

    def disp_batch_full(self, om: Sequence[complex]):
        """n independent omegas with every side output of disp(): (D[n], chi0[n,nspec,3,3], chi0_low[n,nspec,3,3,3],
        wave[n,3,3]) -- element [i] is what Solver.disp(om[i], full=True) returns."""
        om = np.ascontiguousarray(np.asarray(om, dtype=np.complex128).ravel())
        n, ns = om.size, self.pl.nspec
        D = np.zeros(n, dtype=np.complex128)
        chi0 = np.zeros(n * ns * 9, dtype=np.complex128)
        low = np.zeros(n * ns * 27, dtype=np.complex128)
        wave = np.zeros(n * 9, dtype=np.complex128)
        f = lambda a: _p(a.view(np.float64))
        _lib.check(self.L.alps_b200_disp_batch_full(n, f(om), f(D), f(chi0), f(low), f(wave)))
        return (D, chi0.reshape((n, 3, 3, ns)).transpose(0, 3, 2, 1), low.reshape((n, 3, 3, 3, ns)).transpose(0, 4, 3, 2, 1),
                wave.reshape((n, 3, 3)).transpose(0, 2, 1))

    def disp_batch_dev(self, n: int, d_om_ptr: int, d_D_ptr: int):
        """Device-resident omegas / D (raw device pointers, e.g. torch .data_ptr())."""
        _lib.check(self.L.alps_b200_disp_batch_dev(n, C.c_void_p(d_om_ptr), C.c_void_p(d_D_ptr)))

    def chi_partial_len(self) -> int:
        return self.L.alps_b200_chi_partial_len()

    def chi_partial_dev(self, n: int, d_om_ptr: int, d_partial_ptr: int):
        _lib.check(self.L.alps_b200_chi_partial_dev(n, C.c_void_p(d_om_ptr), C.c_void_p(d_partial_ptr)))

    def assemble_dev(self, n: int, d_om_ptr: int, d_partial_ptr: int, d_D_ptr: int):
        _lib.check(self.L.alps_b200_assemble_dev(n, C.c_void_p(d_om_ptr), C.c_void_p(d_partial_ptr),
                                                 C.c_void_p(d_D_ptr)))

    def add_external_chi(self, is_: int, chi, chi_low=None):
        c = np.ascontiguousarray(np.asarray(chi, dtype=np.complex128).ravel(order="F"))
        cl = None if chi_low is None else np.ascontiguousarray(
            np.asarray(chi_low, dtype=np.complex128).ravel(order="F"))
        _lib.check(self.L.alps_b200_add_external_chi(is_, _p(c.view(np.float64)),
                                                     None if cl is None else _p(cl.view(np.float64))))

    # ---- omega-point generators (C++ twins in csrc/drivers.cpp; D always comes from the GPU)
    @staticmethod
    def opts(numiter=50, D_threshold=1.0e-5, D_prec=1.0e-5, D_tol=1.0e-7, D_gap=1.0e-5, secant_method=2):
        return _lib.SolverOpts(numiter, D_threshold, D_prec, D_tol, D_gap, secant_method)

    def _root(self, fn, om, opts):
        v = np.array([om.real, om.imag])
        it = C.c_int(0)
        _lib.check(fn(_p(v), C.byref(opts), C.byref(it)))
        return complex(v[0], v[1]), it.value

    def secant(self, om, opts):          # src/ALPS_fns.f90:1815-1917
        return self._root(self.L.alps_b200_secant, om, opts)

    def secant_osc(self, om, opts):      # src/ALPS_fns.f90:1919-2101
        return self._root(self.L.alps_b200_secant_osc, om, opts)

    def rtsec(self, om, opts):           # src/ALPS_fns.f90:2105-2195
        return self._root(self.L.alps_b200_rtsec, om, opts)

    def refine_guess(self, wroots, opts, roots_path=None):   # src/ALPS_fns.f90:3793-3856
        w = np.ascontiguousarray(np.asarray(wroots, dtype=np.complex128).ravel())
        D = np.zeros(w.size, dtype=np.complex128)
        _lib.check(self.L.alps_b200_refine_guess(w.size, _p(w.view(np.float64)), C.byref(opts),
                                                 roots_path.encode() if roots_path else None,
                                                 _p(D.view(np.float64))))
        return w, D

    def map_search(self, omi, omf, gami, gamf, nr, ni, loggridw=False, loggridg=False,
                   determine_minima=True, numroots=100, map_path=None, shard=None):   # src/ALPS_fns.f90:3595-3788
        """shard = (rank, world, all_gather): omega sharding over one process per GPU (sharding.py) -- this
        rank evaluates its contiguous slice of the nr x ni grid, `all_gather(padded_slice)` returns every rank's
        slice (torch.distributed.all_gather semantics), every rank finishes the map; only rank 0 writes
        map_path.  The gathered map is bitwise the single-GPU map (batch classes)."""
        if shard is not None:
            from . import sharding
            rank, world, all_gather = shard
            return sharding.map_search_sharded(self.disp_batch, rank, world, all_gather, omi, omf, gami, gamf, nr, ni,
                                               loggridw, loggridg, determine_minima, numroots, map_path)
        m = _lib.MapCfg(omi, omf, gami, gamf, nr, ni, int(loggridw), int(loggridg), int(determine_minima))
        n = nr * ni
        om = np.zeros(n, dtype=np.complex128)
        cal = np.zeros(n, dtype=np.complex128)
        val = np.zeros(n)
        iroots = np.zeros(2 * numroots, dtype=np.int32)
        nfound = C.c_int(0)
        _lib.check(self.L.alps_b200_map_search(C.byref(m), map_path.encode() if map_path else None,
                                               _p(om.view(np.float64)), _p(val), _p(cal.view(np.float64)),
                                               numroots, _p(iroots), C.byref(nfound)))
        shape = (nr, ni)
        om, cal, val = (a.reshape(shape, order="F") for a in (om, cal, val))
        k = min(nfound.value, numroots)
        ir = iroots[0:2 * k:2] - 1
        ii = iroots[1:2 * k:2] - 1
        return om, val, cal, [complex(om[a, b]) for a, b in zip(ir, ii)]

    def current_int(self):
        """n_s q_s <p_par>/m_s of derivative_f0 (src/ALPS_fns.f90:161-189), host bookkeeping"""
        pl = self.pl
        out = np.zeros(pl.nspec)
        for i, s in enumerate(pl.species):
            if s.usebM:     # src/ALPS_fns.f90:161-166: the drift of the closed-form bi-Maxwellian
                out[i] = s.ns * s.qs * s.bM_pdrifts / s.ms
                continue
            dpperp = pl.pp[i, 2, 2, 0] - pl.pp[i, 1, 2, 0]
            dppar = abs(pl.pp[i, 2, 2, 1] - pl.pp[i, 2, 1, 1])
            out[i] = np.sum((s.ns * s.qs / s.ms) * pl.pp[i, :, :, 0] * pl.pp[i, :, :, 1] * pl.f0[i]
                            * 2.0 * np.pi * dpperp * dppar)
        return out

    def calc_eigen(self, om, eigen=True, heat=True, current_int=None):   # src/ALPS_fns.f90:2605-2899
        pl = self.pl
        n = pl.nspec
        ns = np.array([s.ns for s in pl.species])
        qs = np.array([s.qs for s in pl.species])
        ci = self.current_int() if current_int is None else np.asarray(current_int, dtype=np.float64)
        omv = np.array([om.real, om.imag])
        ef, bf = np.zeros(3, np.complex128), np.zeros(3, np.complex128)
        Us, ds = np.zeros(3 * n, np.complex128), np.zeros(n, np.complex128)
        Ps, Pss, W = np.zeros(n), np.zeros(4 * n), C.c_double(0.0)
        f = lambda a: _p(a.view(np.float64))
        _lib.check(self.L.alps_b200_calc_eigen(_p(omv), n, _p(ns), _p(qs), _p(ci), self.kperp, self.kpar,
                                               pl.vA, int(eigen), int(heat), f(ef), f(bf), f(Us), f(ds),
                                               _p(Ps), _p(Pss), C.byref(W)))
        return dict(ef=ef, bf=bf, Us=Us.reshape((3, n), order="F"), ds=ds, Ps=Ps,
                    Ps_split=Pss.reshape((4, n), order="F"), W_EM=W.value)

    def om_scan(self, wroots, opts, scan_type, swi, swf, swlog, ns_steps, nres=1, eigen=False, heat=False,
                prefix=None, ik=1):   # src/ALPS_fns.f90:2198-2600 (+ scan_read, src/ALPS_io.f90:455-549)
        pl = self.pl
        n = pl.nspec
        kpl, kql = C.c_double(self.kperp), C.c_double(self.kpar)
        sc = _lib.ScanCfg()
        _lib.check(self.L.alps_b200_scan_setup(scan_type, swi, swf, int(swlog), ns_steps, nres, int(eigen),
                                               int(heat), C.byref(kpl), C.byref(kql), C.byref(sc)))
        w = np.ascontiguousarray(np.asarray(wroots, dtype=np.complex128).ravel())
        ns = np.array([s.ns for s in pl.species])
        qs = np.array([s.qs for s in pl.species])
        ci = self.current_int()
        kperp, kpar = C.c_double(self.kperp), C.c_double(self.kpar)
        rows = np.zeros((ns_steps + 1, w.size, 4))
        _lib.check(self.L.alps_b200_om_scan(C.byref(sc), w.size, _p(w.view(np.float64)), C.byref(opts), n,
                                            _p(ns), _p(qs), _p(ci), pl.vA, C.byref(kperp), C.byref(kpar),
                                            prefix.encode() if prefix else None, ik, _p(rows)))
        self.kperp, self.kpar = kperp.value, kpar.value
        return rows, w

    def set_root_batching(self, on: bool):
        """advance all roots of a k step concurrently (one disp_batch per iteration); same results"""
        _lib.check(self.L.alps_b200_set_root_batching(int(on)))

    def om_double_scan(self, wroots, opts, scan1, scan2, prefix=None):
        """scan_option=2 (src/ALPS_fns.f90:2904-3591).  scan1/scan2: dicts with scan_type, swi, swf, swlog,
        ns, nres, eigen, heat (the &scan_input_1/2 blocks, read in that order like scan_read does)."""
        pl = self.pl
        kpl, kql = C.c_double(self.kperp), C.c_double(self.kpar)
        scs = []
        for sc in (scan1, scan2):
            c = _lib.ScanCfg()
            _lib.check(self.L.alps_b200_scan_setup(int(sc["scan_type"]), float(sc["swi"]), float(sc["swf"]),
                                                   int(sc["swlog"]), int(sc["ns"]), int(sc.get("nres", 1)),
                                                   int(sc.get("eigen", False)), int(sc.get("heat", False)),
                                                   C.byref(kpl), C.byref(kql), C.byref(c)))
            scs.append(c)
        w = np.ascontiguousarray(np.asarray(wroots, dtype=np.complex128).ravel())
        ns = np.array([s.ns for s in pl.species])
        qs = np.array([s.qs for s in pl.species])
        ci = self.current_int()
        kperp, kpar = C.c_double(self.kperp), C.c_double(self.kpar)
        rows = np.zeros((int(scan1["ns"]) + 1, int(scan2["ns"]) + 1, w.size, 4))
        _lib.check(self.L.alps_b200_om_double_scan(C.byref(scs[0]), C.byref(scs[1]), w.size, _p(w.view(np.float64)),
                                                   C.byref(opts), pl.nspec, _p(ns), _p(qs), _p(ci), pl.vA,
                                                   C.byref(kperp), C.byref(kpar),
                                                   prefix.encode() if prefix else None, _p(rows)))
        self.kperp, self.kpar = kperp.value, kpar.value
        return rows, w

    # ---- plumbing
    def set_stream(self, stream_ptr: Optional[int]):
        _lib.check(self.L.alps_b200_set_stream(C.c_void_p(stream_ptr or 0)))

    def sync(self):
        _lib.check(self.L.alps_b200_sync())

    def info(self, what: int) -> float:
        out = C.c_double(0.0)
        _lib.check(self.L.alps_b200_get_info(what, C.byref(out)))
        return out.value

    def dfma_peak(self) -> float:
        out = C.c_double(0.0)
        _lib.check(self.L.alps_b200_dfma_peak(C.byref(out)))
        return out.value

    def close(self):
        self.L.alps_b200_finalize()


def nhds_calc_chi(species, om: complex, kperp: float, kpar: float, kperp_norm: bool = True):
    """calc_chi of ALPS_NHDS.f90 for a use_bM species and one omega, evaluated by the device kernels the hot
    path uses (k_nhds_bessel + k_nhds; stateless, needs a GPU): returns chi(3,3), chi_low(3,3,-1:1)."""
    L = _lib.lib()
    x = np.array([om.real, om.imag])
    chi = np.zeros(9, dtype=np.complex128)
    low = np.zeros(27, dtype=np.complex128)
    s = species
    _lib.check(L.alps_b200_nhds_calc_chi(s.ns, s.qs, s.ms, s.bM_nmaxs, s.bM_Bessel_zeros, s.bM_betas, s.bM_alphas,
                                         s.bM_pdrifts, kpar, kperp, _p(x), int(kperp_norm),
                                         _p(chi.view(np.float64)), _p(low.view(np.float64))))
    return chi.reshape((3, 3), order="F"), low.reshape((3, 3, 3), order="F")
