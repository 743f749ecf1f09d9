"""BASELINE.json's full sizes (C5: 3-species bi-kappa, 1024x2048 grid): the oracle cannot finish a D
with nmax = 200 in reasonable time, so full-size coverage is (i) oracle parity on the full grid with
the harmonics capped at |n| <= 1, and (ii) size-independent properties with nmax = 200: batch == single,
direct quadrature == k-hoisted fast path, sum of harmonic shards == unsharded."""
import numpy as np
import pytest

from alps_b200 import tables
from tests.util import chi_err, det_scale, scaled_err, wave_scale

pytestmark = pytest.mark.gpu
KPERP, KPAR = 15.5, 1.0e-2


@pytest.fixture(scope="module")
def c5():
    return tables.config_kappa3(1024, 2048)


def test_full_grid_parity_with_capped_harmonics(c5):
    from alps_b200.solver import Solver
    from oracle.oracle import Oracle
    orc = Oracle(c5, nmax_force=1)
    sol = Solver(c5, nmax_force=1)
    try:
        orc.set_k(KPERP, KPAR)
        sol.set_k(KPERP, KPAR)
        for om in (0.31 - 0.02j, 1.04 + 0.01j):
            Do, chi_o, low_o, wave_o = orc.disp(om, full=True)
            Dg, chi_g, low_g, wave_g = sol.disp(om, full=True)
            ws = wave_scale(chi_o, om, c5.vA, KPERP, KPAR)
            assert scaled_err(wave_g, wave_o, ws) < 1e-9
            assert abs(Dg - Do) / det_scale(ws) < 1e-9
            for s in range(3):
                assert chi_err(chi_g[s], chi_o[s]) < 1e-9
    finally:
        sol.close()


def test_full_size_properties_nmax_200(c5):
    import torch
    from alps_b200.solver import Solver
    rng = np.random.default_rng(3)
    oms = rng.uniform(0.05, 3.05, 12) + 1j * rng.uniform(-0.05, 0.05, 12)
    sol = Solver(c5, nmax_force=200)
    try:
        sol.set_stream(torch.cuda.current_stream().cuda_stream)
        assert list(sol.set_k(KPERP, KPAR)) == [200, 200, 200]
        D = sol.disp_batch(oms)
        assert np.all(np.isfinite(D.view(np.float64)))
        # batch == single, bitwise
        for i in (0, 5, 11):
            assert sol.disp(complex(oms[i])) == D[i]
        # harmonic shards add up (what the NCCL all-reduce does for C4-style runs)
        n, L = oms.size, sol.chi_partial_len()
        om_d = torch.from_numpy(oms.view(np.float64).copy()).cuda()
        full = torch.zeros(n * L, dtype=torch.float64, device="cuda")
        sol.chi_partial_dev(n, om_d.data_ptr(), full.data_ptr())
        acc = torch.zeros_like(full)
        for rank in range(4):
            sol.set_harmonic_shard(rank, 4)
            sol.set_k(KPERP, KPAR)
            part = torch.zeros_like(full)
            sol.chi_partial_dev(n, om_d.data_ptr(), part.data_ptr())
            torch.cuda.synchronize()
            acc += part
        assert float((acc - full).abs().max()) <= 1e-12 * float(full.abs().max())
        sol.set_harmonic_shard(0, 1)
        # k-hoisted fast path == direct quadrature up to rounding
        sol.set_mode(1)
        sol.set_k(KPERP, KPAR)
        Df = sol.disp_batch(oms)
        assert np.max(np.abs(Df - D) / np.abs(D)) < 1e-10
    finally:
        sol.close()


def test_harmonic_shard_after_an_unsharded_small_batch(c5):
    """the p_par split of small batches depends on the tile list, which a harmonic shard shrinks without
    changing NI -- the partial-sum rows (Sbulk) allocated for the unsharded split were too few for the sharded one
    (64 omegas x 16 splits against 64 x 7: an out-of-bounds device write).  Sequence of the report, at full size: set_k,
    a 64-omega batch, set_harmonic_shard(r, 4), set_k, chi partials of 64 omegas; the shards must add up to the
    unsharded partials (and compute-sanitizer stays silent, scripts/gpu_sanitizer.sh)."""
    import torch
    from alps_b200.solver import Solver
    rng = np.random.default_rng(9)
    oms = rng.uniform(0.05, 3.05, 64) + 1j * rng.uniform(-0.05, 0.05, 64)
    sol = Solver(c5, nmax_force=200)
    try:
        sol.set_stream(torch.cuda.current_stream().cuda_stream)
        sol.set_k(KPERP, KPAR)
        D = sol.disp_batch(oms)                         # allocates the batch buffers for the unsharded split
        n, L = oms.size, sol.chi_partial_len()
        om_d = torch.from_numpy(oms.view(np.float64).copy()).cuda()
        full = torch.zeros(n * L, dtype=torch.float64, device="cuda")
        sol.chi_partial_dev(n, om_d.data_ptr(), full.data_ptr())
        acc = torch.zeros_like(full)
        for rank in range(4):
            sol.set_harmonic_shard(rank, 4)
            sol.set_k(KPERP, KPAR)
            part = torch.zeros_like(full)
            sol.chi_partial_dev(n, om_d.data_ptr(), part.data_ptr())
            torch.cuda.synchronize()
            acc += part
        assert float((acc - full).abs().max()) <= 1e-12 * float(full.abs().max())
        sol.set_harmonic_shard(0, 1)
        sol.set_k(KPERP, KPAR)
        assert np.array_equal(sol.disp_batch(oms).view(np.float64), D.view(np.float64))
    finally:
        sol.close()
