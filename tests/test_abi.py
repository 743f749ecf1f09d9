"""The C-ABI library loads and exports every symbol include/alps_b200.h declares; without a GPU the
entry points fail loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "alps_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(alps_b200_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported(built_lib):
    from alps_b200 import _lib
    L = C.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert set(_lib.SYMBOLS) == set(names)


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from alps_b200 import _lib, tables
    from alps_b200.solver import Solver
    with pytest.raises(_lib.AlpsB200Error) as e:
        Solver(tables.config_small(16, 32))
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    L = _lib.lib()
    D = np.zeros(2)
    om = np.array([0.1, 0.0])
    rc = L.alps_b200_disp(om.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p), None, None, None)
    assert rc != 0


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "alps_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), os.path.join(dirpath, f)


def test_emulate_split_matches_oracle_split_processes(built_lib):
    """host logic of split_processes: C++ twin (alps_b200_emulate_split) vs the oracle's restatement"""
    from alps_b200 import _lib, tables
    from oracle.oracle import Oracle
    L = _lib.lib()
    pl = tables.config_small(24, 48)
    for nproc, k in [(4, 0.3), (6, 0.3), (8, 0.5), (12, 0.1), (40, 0.05), (64, 0.3)]:
        orc = Oracle(pl, nproc=0)
        base = orc.set_k(k, 0.05).copy()
        orc = Oracle(pl, nproc=nproc)
        adj = orc.set_k(k, 0.05)
        want_hi = [max(n2 for s, n1, n2 in orc.nlim() if s == i + 1) for i in range(pl.nspec)]
        nmax = base.astype(np.int32).copy()
        nhi = np.zeros(pl.nspec, dtype=np.int32)
        usebm = np.zeros(pl.nspec, dtype=np.int32)
        rc = L.alps_b200_emulate_split(nproc, pl.nspec, usebm.ctypes.data_as(C.c_void_p),
                                       nmax.ctypes.data_as(C.c_void_p), nhi.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert list(nmax) == list(adj), (nproc, list(nmax), list(adj))
        assert list(nhi) == want_hi, (nproc, list(nhi), want_hi)


def test_fortran_shim_binds_only_declared_entry_points():
    """integration/alps_b200_shim.f90 (source only -- no Fortran compiler here): every bind(c) interface it declares
    is an entry point of include/alps_b200.h, with the same number of arguments."""
    shim = open(os.path.join(ROOT, "integration", "alps_b200_shim.f90")).read()
    shim = re.sub(r"&\s*\n\s*", " ", shim)                      # join continuation lines
    hdr = open(os.path.join(ROOT, "include", "alps_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    bound = re.findall(r"(?:function|subroutine)\s+(alps_b200_\w+)\s*\(([^)]*)\)\s*bind\(c\)", shim, flags=re.I)
    assert len(bound) >= 10
    for name, args in bound:
        m = re.search(r"\b%s\s*\(([^)]*)\)" % name, hdr)
        assert m, name
        n_f = len([a for a in args.split(",") if a.strip()])
        c_args = m.group(1).strip()
        n_c = 0 if c_args in ("", "void") else len(c_args.split(","))
        assert n_f == n_c, (name, n_f, n_c)


def test_cfg_struct_layout_is_the_same_in_header_ctypes_and_shim():
    """alps_b200_cfg: field order and C types in include/alps_b200.h, alps_b200/_lib.py (ctypes) and the Fortran
    bind(c) type of the shim."""
    from alps_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "alps_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} alps_b200_cfg;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        typ, names = decl.split(None, 1)
        fields += [(n.strip(), typ) for n in names.split(",")]
    ctypes_fields = [(n, "int" if t is C.c_int else "double") for n, t in _lib.Cfg._fields_]
    assert fields == ctypes_fields
    shim = open(os.path.join(ROOT, "integration", "alps_b200_shim.f90")).read()
    tbody = re.search(r"type, bind\(c\) :: alps_b200_cfg(.*?)end type", shim, flags=re.S | re.I).group(1)
    shim_fields = []
    for line in tbody.splitlines():
        line = line.split("!")[0].strip()
        if "::" not in line:
            continue
        typ, names = line.split("::")
        kind = "int" if "c_int" in typ else "double"
        shim_fields += [(n.strip(), kind) for n in names.split(",")]
    assert shim_fields == fields


def test_every_header_struct_matches_its_ctypes_mirror():
    from alps_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "alps_b200.h")).read()
    mirrors = {"alps_b200_cfg": _lib.Cfg, "alps_b200_solver_opts": _lib.SolverOpts, "alps_b200_map": _lib.MapCfg,
               "alps_b200_scan": _lib.ScanCfg}
    found = re.findall(r"typedef struct \{(.*?)\} (\w+);", hdr, flags=re.S)
    assert sorted(n for _, n in found) == sorted(mirrors)
    for body, name in found:
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                typ, names = decl.split(None, 1)
                fields += [(n.strip(), typ) for n in names.split(",")]
        got = [(n, "int" if t is C.c_int else "double") for n, t in mirrors[name]._fields_]
        assert fields == got, name


def test_ctypes_argument_counts_match_the_header(built_lib):
    from alps_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "alps_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    checked = 0
    for name, args in re.findall(r"\b(alps_b200_\w+)\s*\(([^)]*)\)\s*;", hdr):
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue
        args = args.strip()
        n_c = 0 if args in ("", "void") else len(args.split(","))
        assert len(fn.argtypes) == n_c, (name, len(fn.argtypes), n_c)
        checked += 1
    assert checked >= 25
