"""ctypes binding of the C ABI in include/alps_b200.h (libalps_b200.so, built in-tree by
alps_b200/csrc/Makefile).  There is no fallback: a missing library or a failing call raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ALPS_B200_LIB") or os.path.join(_HERE, "libalps_b200.so")   # override: A/B builds only
_LIB = None

SYMBOLS = [
    "alps_b200_init", "alps_b200_finalize", "alps_b200_last_error", "alps_b200_set_species",
    "alps_b200_upload", "alps_b200_upload_rel", "alps_b200_derivative_f0", "alps_b200_set_k", "alps_b200_disp",
    "alps_b200_disp_batch", "alps_b200_disp_batch_full", "alps_b200_disp_batch_dev", "alps_b200_disp_prefetch", "alps_b200_add_external_chi",
    "alps_b200_set_bm_species", "alps_b200_nhds_calc_chi",
    "alps_b200_set_harmonic_shard", "alps_b200_chi_partial_len", "alps_b200_chi_partial_dev",
    "alps_b200_assemble_dev", "alps_b200_set_mode", "alps_b200_set_stream", "alps_b200_sync",
    "alps_b200_get_info", "alps_b200_dfma_peak", "alps_b200_emulate_split",
    "alps_b200_secant", "alps_b200_secant_osc", "alps_b200_rtsec", "alps_b200_refine_guess",
    "alps_b200_map_search", "alps_b200_map_grid", "alps_b200_map_finish", "alps_b200_calc_eigen", "alps_b200_scan_setup", "alps_b200_om_scan",
    "alps_b200_om_double_scan", "alps_b200_set_root_batching", "alps_b200_tps_eval",
    "alps_b200_set_partition", "alps_b200_comm_unique_id", "alps_b200_comm_init", "alps_b200_comm_finalize",
    "alps_b200_omega_slice", "alps_b200_set_map_mode", "alps_b200_map_eval",
]

INFO_POINT_HARMONICS, INFO_LAUNCHES, INFO_SM_COUNT, INFO_LAST_KERNEL_MS, INFO_BATCH, INFO_DFMA_NOREUSE, \
    INFO_DMMA_PEAK, INFO_QUAD_VARIANT, INFO_D_EVALS, INFO_SET_K_CALLS, INFO_MEMO_HITS, INFO_PREFETCHED, INFO_NGPU = range(13)
PARTITION_OMEGA, PARTITION_HARMONIC = 0, 1


class Cfg(C.Structure):
    _fields_ = [("nspec", C.c_int), ("nperp", C.c_int), ("npar", C.c_int), ("ngamma", C.c_int),
                ("npparbar", C.c_int), ("vA", C.c_double), ("Bessel_zero", C.c_double),
                ("Tlim", C.c_double), ("positions_principal", C.c_int),
                ("n_resonance_interval", C.c_int), ("kperp_norm", C.c_int),
                ("emulate_nproc", C.c_int), ("maxfits", C.c_int), ("maxorder", C.c_int),
                ("device", C.c_int), ("nmax_cap", C.c_int), ("batch_max", C.c_int),
                ("nmax_force", C.c_int), ("ngpu", C.c_int)]


class SolverOpts(C.Structure):
    _fields_ = [("numiter", C.c_int), ("D_threshold", C.c_double), ("D_prec", C.c_double),
                ("D_tol", C.c_double), ("D_gap", C.c_double), ("secant_method", C.c_int)]


class MapCfg(C.Structure):
    _fields_ = [("omi", C.c_double), ("omf", C.c_double), ("gami", C.c_double), ("gamf", C.c_double),
                ("nr", C.c_int), ("ni", C.c_int), ("loggridw", C.c_int), ("loggridg", C.c_int),
                ("determine_minima", C.c_int)]


class ScanCfg(C.Structure):
    _fields_ = [("type", C.c_int), ("n_out", C.c_int), ("n_res", C.c_int), ("log_scan", C.c_int),
                ("eigen", C.c_int), ("heat", C.c_int), ("diff", C.c_double), ("diff2", C.c_double)]


class AlpsB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("alps_b200 error %d: %s" % (code, msg))
        self.code = code


def build(force: bool = False) -> str:
    """Compile the CUDA extension for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "-s", "clean"])
    subprocess.check_call(["make", "-C", csrc, "-s", "-j", str(min(8, os.cpu_count() or 1))])
    return SO_PATH


def _sig(L, name, argtypes):
    """argument types of one entry point; an A/B build given with ALPS_B200_LIB may predate the newest entry points"""
    if hasattr(L, name):
        getattr(L, name).argtypes = argtypes
    elif not os.environ.get("ALPS_B200_LIB"):
        raise AlpsB200Error(-1, "%s does not export %s" % (SO_PATH, name))


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise AlpsB200Error(-1, "%s is missing: run `make -C alps_b200/csrc` (or "
                                "__graft_entry__.build()); there is no CPU fallback" % SO_PATH)
        L = C.CDLL(SO_PATH)
        L.alps_b200_last_error.restype = C.c_char_p
        _sig(L, "alps_b200_set_species", [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double])
        _sig(L, "alps_b200_upload", [C.c_void_p] * 4)
        _sig(L, "alps_b200_derivative_f0", [C.c_void_p] * 2)
        _sig(L, "alps_b200_upload_rel", [C.c_int] + [C.c_void_p] * 4)
        _sig(L, "alps_b200_set_k", [C.c_double, C.c_double, C.c_void_p])
        _sig(L, "alps_b200_disp", [C.c_void_p] * 5)
        _sig(L, "alps_b200_disp_batch", [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p])
        _sig(L, "alps_b200_disp_batch_full", [C.c_int] + [C.c_void_p] * 5)
        _sig(L, "alps_b200_disp_batch_dev", [C.c_int, C.c_void_p, C.c_void_p])
        _sig(L, "alps_b200_disp_prefetch", [C.c_int, C.c_void_p])
        _sig(L, "alps_b200_add_external_chi", [C.c_int, C.c_void_p, C.c_void_p])
        _sig(L, "alps_b200_set_harmonic_shard", [C.c_int, C.c_int])
        _sig(L, "alps_b200_set_bm_species", [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double])
        _sig(L, "alps_b200_nhds_calc_chi", [C.c_double] * 3 + [C.c_int] + [C.c_double] * 6 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p])
        _sig(L, "alps_b200_chi_partial_dev", [C.c_int, C.c_void_p, C.c_void_p])
        _sig(L, "alps_b200_assemble_dev", [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p])
        _sig(L, "alps_b200_set_mode", [C.c_int])
        _sig(L, "alps_b200_set_stream", [C.c_void_p])
        _sig(L, "alps_b200_get_info", [C.c_int, C.c_void_p])
        _sig(L, "alps_b200_dfma_peak", [C.c_void_p])
        _sig(L, "alps_b200_emulate_split", [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p])
        V = C.c_void_p
        _sig(L, "alps_b200_secant", [V, V, V])
        _sig(L, "alps_b200_secant_osc", [V, V, V])
        _sig(L, "alps_b200_rtsec", [V, V, V])
        _sig(L, "alps_b200_refine_guess", [C.c_int, V, V, C.c_char_p, V])
        _sig(L, "alps_b200_map_search", [V, C.c_char_p, V, V, V, C.c_int, V, V])
        _sig(L, "alps_b200_map_grid", [V, V])
        _sig(L, "alps_b200_map_finish", [V, V, C.c_char_p, V, C.c_int, V, V])
        _sig(L, "alps_b200_calc_eigen", [V, C.c_int, V, V, V, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, V, V, V, V, V, V, V])
        _sig(L, "alps_b200_scan_setup", [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, V, V, V])
        _sig(L, "alps_b200_om_scan", [V, C.c_int, V, V, C.c_int, V, V, V, C.c_double, V, V, C.c_char_p, C.c_int, V])
        _sig(L, "alps_b200_om_double_scan", [V, V, C.c_int, V, V, C.c_int, V, V, V, C.c_double, V, V, C.c_char_p, V])
        _sig(L, "alps_b200_set_root_batching", [C.c_int])
        _sig(L, "alps_b200_tps_eval", [C.c_int, V, V, V, C.c_int, V, V, V])
        _sig(L, "alps_b200_set_partition", [C.c_int])
        _sig(L, "alps_b200_comm_unique_id", [V])
        _sig(L, "alps_b200_comm_init", [C.c_int, C.c_int, V])
        _sig(L, "alps_b200_omega_slice", [C.c_int, C.c_int, C.c_int, V, V])
        _sig(L, "alps_b200_set_map_mode", [C.c_int])
        _sig(L, "alps_b200_map_eval", [C.c_int, V, V])
        _LIB = L
    return _LIB


def check(rc: int):
    if rc != 0:
        raise AlpsB200Error(rc, lib().alps_b200_last_error().decode())
