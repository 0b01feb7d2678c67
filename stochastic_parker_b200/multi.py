"""One process per GPU: rank identity, particle sharding and the diagnostics reduction.

The reference's multi-rank mode for this path is `size_mpi_sub = 1`: every rank holds the
whole field and its own particles (mhd_data_parallel.f90:246-267), ranks meet only in the
MPI_REDUCE calls of the diagnostics (diagnostics.f90:881-905, 143-151, 1707-1708).  Here a
rank is a GPU.  The library reduces with NCCL when it has a communicator
(`bootstrap_comm`); `reduce_diagnostics` is the same reduction on a host communicator
(torch.distributed, any backend) for drivers that keep their own.
"""
from __future__ import annotations

import os


def rank_info() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_count(total: int, world: int, rank: int) -> int:
    """Particles of `rank` when `total` are split evenly by index (low ranks take the rest)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the communicator")
    return total // world + (1 if rank < total % world else 0)


def bootstrap_comm(sim, dist) -> None:
    """Create the library's NCCL communicator: rank 0 makes the 128-byte id, the host
    communicator broadcasts it (the Fortran driver would MPI_BCAST it)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    obj = [sim.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    sim.comm_init(obj[0], world, rank)


def reduce_diagnostics(d: dict, dist) -> dict:
    """All-reduce one diagnostics record over the host communicator, like the reference's
    MPI_REDUCE calls: SUM for fglobal / flocalK / var_local(1:6) (diagnostics.f90:882-903,
    144-145), MIN for pdt_min, MAX for pdt_max and pmax (diagnostics.f90:146-151, 1707).
    Weights are dyadic, so the FP64 sums do not depend on the reduction order."""
    import numpy as np
    import torch

    out = dict(d)

    def red(a, op):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        dist.all_reduce(t, op=op)
        return t.numpy()

    out["fglobal"] = red(d["fglobal"], dist.ReduceOp.SUM)
    out["flocal"] = [None if a is None else red(a, dist.ReduceOp.SUM) for a in d["flocal"]]
    q = np.array(d["quick"], dtype=np.float64)
    q[:6] = red(q[:6], dist.ReduceOp.SUM)
    q[6:7] = red(q[6:7], dist.ReduceOp.MIN)
    q[7:8] = red(q[7:8], dist.ReduceOp.MAX)
    out["quick"] = q
    out["pmax"] = float(red(np.array([d["pmax"]]), dist.ReduceOp.MAX)[0])
    if "fescaped" in d:
        out["fescaped"] = red(d["fescaped"], dist.ReduceOp.SUM)
    if "fescaped_local" in d:  # the per-face arrays, diagnostics.f90:1174-1230
        out["fescaped_local"] = [None if s is None else
                                 {f: (None if a is None else red(a, dist.ReduceOp.SUM)) for f, a in s.items()}
                                 for s in d["fescaped_local"]]
    return out
