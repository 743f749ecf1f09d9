"""Partitioning of the path over the GPUs of one box (one process per GPU).

The reference's only parallelism is the MPI harmonic split of split_processes
(src/ALPS_fns.f90:4079-4207) with two MPI_REDUCEs per D (:519-523).  Here:
  * omega sharding  -- map_search grids / batches of roots: contiguous slices of the omega list per
    rank, no communication during compute, one all_gather of D at the end;
  * harmonic sharding -- few omegas, large nmax: each rank sums a contiguous block of |n| per species
    (alps_b200_set_harmonic_shard), the un-normalised chi partials are all-reduced (NCCL on GPUs,
    gloo in the CPU tests), then every rank assembles D.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np


def omega_shard(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the omega list owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def harmonic_shard(nhi: int, rank: int, world: int) -> Tuple[int, int]:
    """[nlo, nhi_rank] (inclusive; empty if nlo > nhi_rank) of the harmonics |n| in [0, nhi] owned by
    `rank` -- the same blocks alps_b200_set_harmonic_shard uses (api.cu, alps_b200_set_k)."""
    tot = nhi + 1
    per = ((tot + world - 1) // world + 1) & ~1   # even blocks: TMA needs 16-byte aligned W columns
    nlo = min(rank * per, tot)
    return nlo, min(nlo + per, tot) - 1


def gather_omega_shards(local: np.ndarray, n: int, rank: int, world: int, all_gather: Callable) -> np.ndarray:
    """Reassemble the full D array from per-rank slices.  `all_gather(padded_local)` returns the list of
    every rank's padded slice (torch.distributed.all_gather semantics)."""
    size = max(omega_shard(n, r, world)[1] - omega_shard(n, r, world)[0] for r in range(world))
    pad = np.zeros(size, dtype=np.complex128)
    pad[: local.size] = local
    parts: List[np.ndarray] = all_gather(pad)
    out = np.zeros(n, dtype=np.complex128)
    for r in range(world):
        lo, hi = omega_shard(n, r, world)
        out[lo:hi] = parts[r][: hi - lo]
    return out


def map_search_sharded(disp_batch: Callable, rank: int, world: int, all_gather: Callable, omi, omf, gami, gamf,
                       nr: int, ni: int, loggridw=False, loggridg=False, determine_minima=True, numroots=100,
                       map_path=None):
    """map_search (src/ALPS_fns.f90:3595-3788) with its nr x ni loop of disp calls sharded over `world`
    processes (one per GPU): every rank builds the grid (alps_b200_map_grid), evaluates its contiguous slice
    with `disp_batch` (Solver.disp_batch), the slices are gathered with `all_gather`, and every rank runs the
    rest of map_search (sentinels, find_minima; the .map file on rank 0) on the full D array
    (alps_b200_map_finish).  Returns (om, val, cal, roots) like Solver.map_search."""
    import ctypes as C

    from . import _lib
    L = _lib.lib()
    m = _lib.MapCfg(omi, omf, gami, gamf, nr, ni, int(loggridw), int(loggridg), int(determine_minima))
    n = nr * ni
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    om = np.zeros(n, dtype=np.complex128)
    val = np.zeros(n)
    iroots = np.zeros(2 * numroots, dtype=np.int32)
    nfound = C.c_int(0)
    _lib.check(L.alps_b200_map_grid(C.byref(m), p(om.view(np.float64))))
    lo, hi = omega_shard(n, rank, world)
    local = np.asarray(disp_batch(om[lo:hi]), dtype=np.complex128) if hi > lo else np.zeros(0, dtype=np.complex128)
    cal = np.ascontiguousarray(gather_omega_shards(local, n, rank, world, all_gather))
    path = map_path.encode() if (map_path and rank == 0) else None
    _lib.check(L.alps_b200_map_finish(C.byref(m), p(cal.view(np.float64)), path, p(val), numroots, p(iroots),
                                      C.byref(nfound)))
    om, cal, val = (a.reshape((nr, ni), order="F") for a in (om, cal, val))
    k = min(nfound.value, numroots)
    ir = iroots[0:2 * k:2] - 1
    ii = iroots[1:2 * k:2] - 1
    return om, val, cal, [complex(om[a, b]) for a, b in zip(ir, ii)]


def torch_all_gather(group=None) -> Callable:
    """`all_gather` callable for gather_omega_shards / map_search_sharded over torch.distributed (NCCL on GPUs:
    the padded slice goes through the current CUDA device; gloo: host tensors)."""
    import torch
    import torch.distributed as dist

    def all_gather(pad: np.ndarray):
        t = torch.from_numpy(pad.view(np.float64).copy())
        nccl = dist.get_backend(group) == "nccl"
        if nccl:
            t = t.cuda()
        outs = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
        dist.all_gather(outs, t, group=group)
        return [o.cpu().numpy().view(np.complex128) for o in outs]

    return all_gather
