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
        q.put((rank, bool(ok1), bool(ok2)))
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
    assert all(r[1] and r[2] for r in res), res


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
