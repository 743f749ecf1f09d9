"""Pins the CPU oracle (oracle/) against the reference's own golden vectors for the path:
tests/test_kpar_fast.scan_kpara_1.root_1 (33 roots of the k_par scan, 5 significant digits) and the
known answers in tests/test_kpar_fast.out (nmax 21/13, density integral 9.9958E-001)."""
import os

import numpy as np

from alps_b200 import tables
from oracle import driver
from oracle.oracle import Oracle, bessj

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fmt5(x):
    """the reference prints es14.4e3: 5 significant digits"""
    return float("%.4E" % x)


def test_bessj_is_the_numerical_recipes_approximation():
    from scipy.special import jv
    # x >= 0 only: for x < -n the reference's BESSJ (no |x| / parity fix-up, Miller start index tuned
    # for |x| <= n, src/ALPS_fns_rel.f90:1601-1614) is simply wrong, and parity means reproducing that.
    xs = np.linspace(0.0, 30.0, 121)
    for n in (0, 1, 2, 5, 13, 21):
        ours = np.array([bessj(n, float(x)) for x in xs])
        exact = jv(n, xs)
        # ~1e-8 accurate, not better: parity needs the reference's approximation, not the exact J_n
        assert np.max(np.abs(ours - exact)) < 5e-7
    assert abs(bessj(1, 20.0) - float(jv(1, 20.0))) > 1e-12


def test_nmax_and_density_of_test_kpar_fast():
    pl = tables.config_kpar_fast()
    orc = Oracle(pl, nproc=4)
    nmax = orc.set_k(1.0e-2, 1.0e-2)
    assert list(nmax) == [21, 13]                       # tests/test_kpar_fast.out:99-100
    assert orc.nlim() == [(1, 0, 10), (1, 11, 21), (2, 0, 13)]
    for i in range(2):
        dpperp = pl.pp[i, 1, 0, 0] - pl.pp[i, 0, 0, 0]
        dppar = pl.pp[i, 0, 1, 1] - pl.pp[i, 0, 0, 1]
        dens = float(np.sum(pl.pp[i, :, :, 0] * pl.f0[i]) * 2 * np.pi * dpperp * dppar)
        assert "%.4E" % dens == "9.9958E-01"            # tests/test_kpar_fast.out:53,58


def test_golden_kpar_scan_33_rows():
    """refine_guess + om_scan (scan_type 4, log, 32 steps, secant_osc) with the oracle's disp()."""
    gold = np.loadtxt(os.path.join(GOLD, "test_kpar_fast.scan_kpara_1.root_1"))
    pl = tables.config_kpar_fast()
    orc = Oracle(pl, nproc=4)
    rows = driver.scan_k(orc, 1.0e-2, 1.0e-2, 4, 1.0e-1, 32, True, complex(9.9e-3, -5.5e-6),
                         numiter=50, D_threshold=1.0e-15, D_prec=1.0e-5)
    assert len(rows) == gold.shape[0] == 33
    for (kperp, kpar, om), g in zip(rows, gold):
        assert _fmt5(kperp) == g[0] and _fmt5(kpar) == g[1]
        assert _fmt5(om.real) == g[2], (kpar, om, g)
        assert _fmt5(om.imag) == g[3], (kpar, om, g)
    # the .eigen_* and .heat_* goldens of the same run pin the side outputs of disp() -- wave (E, B), chi0 per species
    # (velocity / density fluctuations, heating rates at real omega) and d(chi_h)/d(omega) (W_EM) -- through the
    # restatement of calc_eigen (oracle/driver.py): every column of every row to the 5 printed digits (last digit free)
    ge = np.loadtxt(os.path.join(GOLD, "test_kpar_fast.eigen_kpara_1.root_1"))
    gh = np.loadtxt(os.path.join(GOLD, "test_kpar_fast.heat_kpara_1.root_1"))
    gm = np.loadtxt(os.path.join(GOLD, "test_kpar_fast.heat_mech_kpara_1.root_1"))
    ns = [sp.ns for sp in pl.species]
    qs = [sp.qs for sp in pl.species]
    for row, (kperp, kpar, om) in enumerate(rows):
        if row % 4 and row != 32:
            continue                                  # every fourth row and the last: keeps the CPU suite short
        orc.set_k(kperp, kpar)
        e, b, Us, ds, Ps, W, Psplit = driver.calc_eigen(orc, om, kperp, kpar, pl.vA, ns, qs, split=True)
        ri = lambda z: np.array([np.ravel(z).real, np.ravel(z).imag]).T.ravel()
        mine = np.concatenate([[kperp, kpar, om.real, om.imag], ri(e), ri(b), ri(Us), ri(ds)])
        assert mine.shape == ge[row].shape
        tol = 2e-4 * np.maximum(np.abs(ge[row]), 1e-3 * np.max(np.abs(ge[row])))
        assert np.all(np.abs(mine - ge[row]) <= tol), (row, mine, ge[row])
        heat = np.concatenate([[kperp, kpar, om.real, om.imag], Ps, [W]])
        assert np.all(np.abs(heat[4:] - gh[row][4:]) <= 2e-4 * np.abs(gh[row][4:])), (row, heat, gh[row])
        # heating by mechanism (chi0_low, the n = 0, +-1 parts): the shipped .heat_mech_* file is from another build
        # than the other goldens (its own gamma column differs by 0.36 %), hence only the 1e-2 level of the
        # reference's own pytest, with margin
        gm_row = gm[row][4:]
        assert np.all(np.abs(Psplit.ravel() - gm_row) <= 3e-2 * np.maximum(np.abs(gm_row), 1e-3 * np.max(np.abs(gm_row))))


def test_chi_known_answers_of_the_survey_probe():
    """un-normalised chi diagonals at the last golden root (k = (1e-2, 1e-1)), computed there
    independently of this repository with scipy's Bessel functions (trusted to ~7 digits): pins chi itself, not only
    the roots.  Only n = 0 is resonant for both species at this omega."""
    pl = tables.config_kpar_fast()
    orc = Oracle(pl, nproc=4)
    assert list(orc.set_k(1.0e-2, 1.0e-1)) == [21, 13]
    om = complex(9.2594146e-2, -2.8659349e-4)
    D, chi0, low, wave = orc.disp(om, full=True)
    un = chi0 * (om ** 2 * pl.vA ** 2)
    want = [[8.79481e-3 - 5.4906e-5j, 8.6933e-3 + 1.4282e-5j, 1.58714e-3 + 1.1828035j],
            [8.83923e-6 - 2.8932e-8j, 8.34002e-6 + 3.78767e-6j, 1.7251811 + 1.2189672e-1j]]
    for s in range(2):
        for i in range(3):
            assert abs(un[s, i, i] - want[s][i]) <= 5e-6 * abs(want[s][i]), (s, i, un[s, i, i], want[s][i])
    # and omega is the root of the golden scan's last row to its printed digits: |D| is ~1e-13 of its term scale
    assert abs(D) < 1e-9 * float(np.max(np.abs(wave))) ** 3


def test_oracle_reproduces_its_committed_vectors():
    """tests/golden/oracle_vectors.npz (written by tests/golden/make_oracle_vectors.py): the oracle of today gives the
    D, chi0, chi0_low and wave it gave when the fixture was committed -- the CUDA path is compared with the same file
    on the GPU (tests/test_gpu_parity.py::test_committed_oracle_vectors)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_oracle_vectors", os.path.join(GOLD, "make_oracle_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.compute()
    ref = np.load(os.path.join(GOLD, "oracle_vectors.npz"))
    assert sorted(now) == sorted(ref.files)
    for k in ref.files:
        a, b = np.asarray(now[k]), ref[k]
        assert a.shape == b.shape, k
        scale = np.max(np.abs(b)) if b.size else 1.0
        assert np.max(np.abs(a - b)) <= 1e-12 * scale, k        # OpenMP reduction order only


def test_oracle_reproduces_its_operating_point_vectors():
    """tests/golden/oracle_vectors_ops.npz (tests/golden/make_oracle_vectors_ops.py): the C4 (k_perp = 3, nmax 88 / 29,
    mpirun -np 4 emulated) and C2 (150x300, use_bM protons) vectors are recomputed here; the C5 nmax = 200 vectors cost
    the oracle minutes per omega and are only checked for shape and finiteness on the CPU (the CUDA path is compared
    with all of them on the GPU, tests/test_gpu_ops.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_ops", os.path.join(GOLD, "make_oracle_vectors_ops.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = np.load(os.path.join(GOLD, "oracle_vectors_ops.npz"))
    now = mod.compute(["c4", "c2"])
    for k, a in now.items():
        b = ref[k]
        assert np.asarray(a).shape == b.shape, k
        scale = np.max(np.abs(b)) if b.size else 1.0
        assert np.max(np.abs(np.asarray(a) - b)) <= 1e-12 * scale, k
    assert list(ref["c4_nmax"]) == [88, 29]
    assert list(ref["c5_nmax"]) == [200, 200, 200] and ref["c5_om"].size >= 3
    for k in ("c5_D", "c5_chi0", "c5_chi0_low", "c5_wave"):
        assert ref[k].shape[0] == ref["c5_om"].size and np.all(np.isfinite(ref[k].view(np.float64))), k
    im = ref["c5_om"].imag
    assert np.any(im < 0) and np.any(im > 0) and np.any(im == 0)
