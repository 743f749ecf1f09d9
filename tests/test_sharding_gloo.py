"""N>1 host logic on CPU: world_size-2 gloo run of the omega-shard gather and the harmonic-shard
all-reduce that replace the reference's MPI split (src/ALPS_fns.f90:4079-4207, 519-523)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alps_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_disp(om):
    return (om * om - 0.25) * np.exp(1j * om.real)


def _fake_map_disp(om):
    """two zeros inside the test map: strict minima of log10|D| next to them"""
    om = np.asarray(om)
    return (om - (0.31 - 0.11j)) * (om - (0.74 + 0.06j))


def _fake_harmonic(n, om):
    return (1.0 / (n + 1.0)) * (om + n) ** 2


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- omega sharding: no data-path collective, one all_gather at the end
        n = 37
        om = np.linspace(0.1, 2.0, n) + 0.01j
        lo, hi = sharding.omega_shard(n, rank, world)
        local = _fake_disp(om[lo:hi])

        def all_gather(pad):
            t = torch.from_numpy(pad.view(np.float64).copy())
            outs = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            return [o.numpy().view(np.complex128) for o in outs]

        full = sharding.gather_omega_shards(local, n, rank, world, all_gather)
        ok1 = np.array_equal(full, _fake_disp(om))
        # --- harmonic sharding: partial sums + all_reduce
        nhi = 21
        nlo, nhr = sharding.harmonic_shard(nhi, rank, world)
        part = np.zeros(n, dtype=np.complex128)
        for h in range(nlo, nhr + 1):
            part += _fake_harmonic(h, om)
        t = torch.from_numpy(part.view(np.float64).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot = t.numpy().view(np.complex128)
        ref = sum(_fake_harmonic(h, om) for h in range(nhi + 1))
        ok2 = np.allclose(tot, ref, rtol=1e-14, atol=0)
        # --- map_search sharded over the ranks (host halves of the C ABI + gather) against one process
        args = (0.05, 1.0, -0.3, 0.2, 23, 17)
        om_s, val_s, cal_s, roots_s = sharding.map_search_sharded(_fake_map_disp, rank, world,
                                                                  sharding.torch_all_gather(), *args)
        om_1, val_1, cal_1, roots_1 = sharding.map_search_sharded(_fake_map_disp, 0, 1, lambda pad: [pad], *args)
        ok3 = (np.array_equal(om_s, om_1) and np.array_equal(val_s, val_1) and np.array_equal(cal_s, cal_1)
               and roots_s == roots_1 and len(roots_1) >= 2)
        q.put((rank, bool(ok1), bool(ok2), bool(ok3)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_shard_gather_and_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res


def test_shards_partition_exactly():
    for n in (1, 7, 148, 262144):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = sharding.omega_shard(n, r, world)
                cover += list(range(lo, hi)) if n < 1000 else [lo, hi]
            if n < 1000:
                assert cover == list(range(n))
    for nhi in (0, 1, 13, 21, 200):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                a, b = sharding.harmonic_shard(nhi, r, world)
                got += list(range(a, b + 1))
            assert got == list(range(nhi + 1))


def _find_minima_py(val):
    """src/ALPS_fns.f90:3860-3966: strict minima over the existing 4-neighbours, scanned ii = ni..1, ir = 1..nr"""
    nr, ni = val.shape
    out = []
    for ii in range(ni - 1, -1, -1):
        for ir in range(nr):
            nb = [val[a, b] for a, b in ((ir - 1, ii), (ir + 1, ii), (ir, ii - 1), (ir, ii + 1))
                  if 0 <= a < nr and 0 <= b < ni]
            if all(val[ir, ii] < v for v in nb):
                out.append((ir, ii))
    return out


def test_map_grid_and_finish_are_host_only_and_follow_map_search(tmp_path):
    """alps_b200_map_grid / alps_b200_map_finish (the halves of map_search around the disp loop) need no GPU:
    grid formulas (src/ALPS_fns.f90:3684-3712), sentinels (:3726-3742), .map rows (5es16.6e3, blank line per ir),
    find_minima order."""
    path = str(tmp_path / "t.map")

    def disp(om):
        D = _fake_map_disp(om)
        D[5] = complex(np.nan, 0.0)
        D[6] = complex(1.0e101, 0.0)
        D[7] = complex(np.nan, 1.0)
        return D

    for logw, logg, gam in ((False, False, (-0.3, 0.2)), (True, True, (1e-3, 0.2))):
        nr, ni = 23, 17
        om, val, cal, roots = sharding.map_search_sharded(disp, 0, 1, lambda pad: [pad], 0.05, 1.0, gam[0], gam[1],
                                                          nr, ni, loggridw=logw, loggridg=logg, map_path=path)
        ir = np.arange(nr)
        ii = np.arange(ni)
        wr = 0.05 * (1.0 / 0.05) ** (ir / (nr - 1.0)) if logw else 0.05 + (1.0 - 0.05) / (nr - 1.0) * ir
        wi = gam[0] * (gam[1] / gam[0]) ** (ii / (ni - 1.0)) if logg else gam[0] + (gam[1] - gam[0]) / (ni - 1.0) * ii
        assert np.allclose(om.real, wr[:, None] * np.ones(ni), rtol=1e-15, atol=0)
        assert np.allclose(om.imag, np.ones(nr)[:, None] * wi, rtol=1e-15, atol=0)
        flat_val = val.ravel(order="F")
        flat_cal = cal.ravel(order="F")
        assert flat_val[5] == 999999.0 and flat_cal[5] == 999999.0      # NaN sentinel
        assert flat_val[6] == 899999.0 and flat_cal[6] == 899999.0      # infinity sentinel
        # the reference's NaN test only fires for Im(D) == 0 exactly: with Im(D) /= 0, `tmp .ne. cal` compares
        # (Re D, 0) with D and is always true (:3728-3731), so a complex NaN stays in the map
        assert np.isnan(flat_val[7]) and np.isnan(flat_cal[7].real)
        keep = np.ones(nr * ni, dtype=bool)
        keep[[5, 6, 7]] = False
        D = _fake_map_disp(om.ravel(order="F"))
        assert np.allclose(flat_val[keep], np.log10(np.abs(D))[keep], rtol=1e-14, atol=0)   # libm vs numpy: last bit
        want = _find_minima_py(val)
        assert [(int(np.argmin(np.abs(wr - r.real))), int(np.argmin(np.abs(wi - r.imag)))) for r in roots] == want
        lines = open(path).read().split("\n")
        assert len(lines[0]) == 5 * 16 and lines[ni] == ""
        assert len([l for l in lines if l.strip()]) == nr * ni
        def es16(v):      # Fortran es16.6e3
            m, ex = ("%.6E" % v).split("E")
            return ("%sE%s%03d" % (m, "-" if int(ex) < 0 else "+", abs(int(ex)))).rjust(16)
        assert lines[0] == "".join(es16(v) for v in (om[0, 0].real, om[0, 0].imag, val[0, 0], cal[0, 0].real, cal[0, 0].imag))
        first = [float(x) for x in lines[0].split()]
        assert abs(first[0] - wr[0]) <= 1e-6 * wr[0] and abs(first[2] - val[0, 0]) <= 1e-6 * abs(val[0, 0])
