"""Several GPUs behind the C ABI (VERDICT r01 "put the multi-GPU split inside the library"): the device group
(alps_b200_cfg.ngpu, one process) and the library-owned NCCL communicator (one process per GPU), each under the OMEGA and
the HARMONIC partition -- replaces split_processes and the MPI_REDUCE pair of disp() (src/ALPS_fns.f90:4079-4207,
519-523).  Skipped on a box with one GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from alps_b200 import tables
from tests.util import omega_samples

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


needs2 = pytest.mark.skipif("_ngpu() < 2", reason="needs two GPUs")


@needs2
def test_device_group_omega_partition_is_bitwise_the_single_gpu_result(tmp_path):
    from alps_b200 import _lib
    from alps_b200.solver import Solver
    pl = tables.config_small(48, 96, kind=2)
    kperp, kpar = 1.5, 0.05
    sizes = (1, 5, 9, 37, 64, 65, 203)
    oms = {n: omega_samples(40 + n, n, (0.05, 2.0), (-0.03, 0.03)) for n in sizes}
    margs = (0.05, 2.0, -0.03, 0.03, 40, 32)
    ref = {}
    sol = Solver(pl)
    try:
        sol.set_k(kperp, kpar)
        for n in sizes:
            ref[n] = sol.disp_batch_full(oms[n])
        ref["map"] = sol.map_search(*margs, map_path=str(tmp_path / "one.map"))
        ref["single"] = sol.disp(complex(oms[9][3]), full=True)
    finally:
        sol.close()
    for ngpu in sorted({2, min(_ngpu(), 8)}):
        sol = Solver(pl, ngpu=ngpu)
        try:
            assert int(sol.info(_lib.INFO_NGPU)) == ngpu
            sol.set_k(kperp, kpar)
            for n in sizes:
                got = sol.disp_batch_full(oms[n])
                for a, b, what in zip(got, ref[n], ("D", "chi0", "chi0_low", "wave")):
                    assert np.array_equal(np.ascontiguousarray(a).view(np.float64),
                                          np.ascontiguousarray(b).view(np.float64)), (ngpu, n, what)
                assert np.array_equal(sol.disp_batch(oms[n]).view(np.float64), ref[n][0].view(np.float64))
            m = sol.map_search(*margs, map_path=str(tmp_path / ("g%d.map" % ngpu)))
            assert np.array_equal(m[2], ref["map"][2]) and np.array_equal(m[1], ref["map"][1]) and m[3] == ref["map"][3]
            assert open(tmp_path / ("g%d.map" % ngpu)).read() == open(tmp_path / "one.map").read()
            s = sol.disp(complex(oms[9][3]), full=True)
            assert s[0] == ref["single"][0] and np.array_equal(s[1], ref["single"][1])
            # the hoisted map mode through the group
            sol.set_mode(1)
            sol.set_k(kperp, kpar)
            Dh = sol.disp_batch(oms[203])
            assert np.max(np.abs(Dh - ref[203][0]) / np.abs(ref[203][0])) < 1e-10
        finally:
            sol.close()


@needs2
@pytest.mark.parametrize("reduce", ["p2p", "nccl"])
def test_device_group_harmonic_partition(reduce, monkeypatch):
    """every device sums a block of harmonics, device 0 adds the partial rows of its peers (peer memory over NVLink, or
    ncclAllReduce) and assembles: the single-GPU D up to the order of the harmonic sum"""
    from alps_b200 import _lib
    from alps_b200.solver import Solver
    monkeypatch.setenv("ALPS_B200_REDUCE", reduce)
    pl = tables.config_kpar_fast()
    kperp, kpar = 3.0, 1.0e-3                         # C4: nmax 88 / 29
    oms = np.concatenate([[2.5e-3 - 1.0e-4j, 1.002 - 1.0e-4j, 2.5e-3 + 0j],
                          omega_samples(3, 97, (1.0e-3, 1.2), (-2.0e-3, 2.0e-3))])
    sol = Solver(pl, emulate_nproc=4)
    try:
        sol.set_k(kperp, kpar)
        ref = sol.disp_batch_full(oms)
        ref1 = [sol.disp(complex(o), full=True) for o in oms[:3]]
    finally:
        sol.close()
    for ngpu in sorted({2, min(_ngpu(), 8)}):
        sol = Solver(pl, emulate_nproc=4, ngpu=ngpu)
        try:
            sol.set_partition(_lib.PARTITION_HARMONIC)
            assert list(sol.set_k(kperp, kpar)) == [88, 29]
            got = sol.disp_batch_full(oms)
            assert np.max(np.abs(got[0] - ref[0]) / np.abs(ref[0])) < 1e-11
            for a, b in zip(got[1:], ref[1:]):
                assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b))
            for o, r in zip(oms[:3], ref1):
                # a few omegas of a small configuration are evaluated unsharded on device 0: bitwise the one-GPU value
                g = sol.disp(complex(o), full=True)
                assert g[0] == r[0] and np.array_equal(g[2], r[2])
                assert sol.disp(complex(o)) == r[0]
            d8 = sol.disp_batch(oms[:8])
            assert np.max(np.abs(d8 - ref[0][:8]) / np.abs(ref[0][:8])) < 1e-11
            # back to the OMEGA partition: bitwise the single-GPU result again
            sol.set_partition(_lib.PARTITION_OMEGA)
            sol.set_k(kperp, kpar)
            assert np.array_equal(sol.disp_batch(oms).view(np.float64), ref[0].view(np.float64))
        finally:
            sol.close()


@needs2
def test_one_process_per_gpu_library_communicator(tmp_path):
    """torchrun, 2 ranks: alps_b200_comm_init + both partitions (scripts/multi_gpu_check.py asserts bitwise / 1e-10)"""
    out = tmp_path / "check.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "scripts", "multi_gpu_check.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OMEGA partition" in r.stdout and "HARMONIC partition" in r.stdout
    assert out.exists()
