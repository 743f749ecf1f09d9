"""Host-only pieces of the C++ twin drivers (alps_b200/csrc/drivers.cpp) that need no GPU, against independent
Python restatements of the reference: scan_read's step sizes (src/ALPS_io.f90:481-549)."""
import ctypes as C
import math

import numpy as np

from alps_b200 import _lib


def _scan_read(scan_type, swi, swf, swlog, ns, nres, kperp_last, kpar_last):
    """src/ALPS_io.f90:481-549, literally (including case 2 leaving k_last at k_0, lines 527-528)"""
    den = 1.0 * ns * nres
    pi = 4.0 * math.atan(1.0)
    diff = diff2 = 0.0
    if scan_type == 0:
        if swlog:
            diff = (math.log10(swi) - math.log10(kperp_last)) / den
            diff2 = (math.log10(swf) - math.log10(kpar_last)) / den
        else:
            diff = (swi - kperp_last) / den
            diff2 = (swf - kpar_last) / den
        kperp_last, kpar_last = swi, swf
    elif scan_type == 1:
        theta_0 = math.atan(kperp_last / kpar_last)
        k_0 = math.sqrt(kperp_last ** 2 + kpar_last ** 2)
        if swlog:
            diff = (math.log10(swf * pi / 180.0) - math.log10(theta_0)) / den
        else:
            diff = ((swf * pi / 180.0) - theta_0) / den
        kpar_last = k_0 * math.cos(swf * pi / 180.0)
        kperp_last = k_0 * math.sin(swf * pi / 180.0)
    elif scan_type == 2:
        theta_0 = math.atan(kperp_last / kpar_last)
        k_0 = math.sqrt(kperp_last ** 2 + kpar_last ** 2)
        diff = (math.log10(swf) - math.log10(k_0)) / den if swlog else (swf - k_0) / den
        kpar_last = k_0 * math.cos(theta_0)
        kperp_last = k_0 * math.sin(theta_0)
    elif scan_type == 3:
        diff = (math.log10(swf) - math.log10(kperp_last)) / den if swlog else (swf - kperp_last) / den
        kperp_last = swf
    else:
        diff = (math.log10(swf) - math.log10(kpar_last)) / den if swlog else (swf - kpar_last) / den
        kpar_last = swf
    return diff, diff2, kperp_last, kpar_last


def test_scan_setup_matches_scan_read(built_lib):
    L = _lib.lib()
    rng = np.random.default_rng(3)
    for scan_type in range(5):
        for swlog in (0, 1):
            for _ in range(5):
                kperp0, kpar0 = (float(x) for x in rng.uniform(1e-3, 2.0, 2))
                swi, swf = (float(x) for x in rng.uniform(1e-3, 60.0 if scan_type == 1 else 3.0, 2))
                ns, nres = int(rng.integers(1, 300)), int(rng.integers(1, 5))
                kp, kq = C.c_double(kperp0), C.c_double(kpar0)
                sc = _lib.ScanCfg()
                rc = L.alps_b200_scan_setup(scan_type, swi, swf, swlog, ns, nres, 1, 0, C.byref(kp), C.byref(kq),
                                            C.byref(sc))
                assert rc == 0
                want = _scan_read(scan_type, swi, swf, bool(swlog), ns, nres, kperp0, kpar0)
                got = (sc.diff, sc.diff2, kp.value, kq.value)
                for g, w in zip(got, want):
                    assert g == w or abs(g - w) <= 4e-16 * abs(w), (scan_type, swlog, got, want)   # libm last bit
                assert (sc.type, sc.n_out, sc.n_res, sc.log_scan, sc.eigen, sc.heat) == (scan_type, ns, nres, swlog, 1, 0)
    sc = _lib.ScanCfg()
    assert L.alps_b200_scan_setup(7, 1.0, 1.0, 0, 1, 1, 0, 0, C.byref(C.c_double(1.0)), C.byref(C.c_double(1.0)),
                                  C.byref(sc)) != 0


def test_omega_slices_partition_the_batch(built_lib):
    """alps_b200_omega_slice (host only): the contiguous slices of the OMEGA partition tile [0, n) exactly, differ in
    size by at most one, and agree with alps_b200/sharding.py's omega_shard (the gloo tests' twin)."""
    from alps_b200 import sharding
    L = _lib.lib()
    for n in (0, 1, 7, 8, 9, 64, 65, 1000, 262144):
        for parts in (1, 2, 3, 8):
            prev, sizes = 0, []
            for r in range(parts):
                lo, hi = C.c_int(-1), C.c_int(-1)
                assert L.alps_b200_omega_slice(n, r, parts, C.byref(lo), C.byref(hi)) == 0
                assert lo.value == prev and hi.value >= lo.value
                assert (lo.value, hi.value) == sharding.omega_shard(n, r, parts)
                sizes.append(hi.value - lo.value)
                prev = hi.value
            assert prev == n and max(sizes) - min(sizes) <= 1
    lo, hi = C.c_int(0), C.c_int(0)
    assert L.alps_b200_omega_slice(10, 3, 3, C.byref(lo), C.byref(hi)) != 0       # rank out of range


def test_partition_and_communicator_calls_need_an_initialised_library(built_lib):
    """no GPU here: the multi-GPU entry points refuse to act before alps_b200_init instead of touching a device"""
    L = _lib.lib()
    assert L.alps_b200_set_partition(_lib.PARTITION_HARMONIC) == -2
    buf = (C.c_char * 128)()
    assert L.alps_b200_comm_init(0, 2, buf) == -2
    assert L.alps_b200_comm_finalize() == 0
    assert b"alps_b200_init" in L.alps_b200_last_error()
