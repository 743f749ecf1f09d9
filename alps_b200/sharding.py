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
