"""Worker of tests/test_cpu_multirank.py: one rank of a world_size-N gloo job.

Each rank is what one GPU is in production: the full field, its own shard of the particles
(origin = rank), its own Philox streams.  The oracle stands in for the GPU library (same
method names); the reduction is stochastic_parker_b200.multi.reduce_diagnostics over gloo.
"""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import make_case  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from stochastic_parker_b200 import rank_info, reduce_diagnostics, run_intervals, shard_count  # noqa: E402


def main():
    outdir, total = sys.argv[1], int(sys.argv[2])
    rank, world, _ = rank_info()
    dist.init_process_group("gloo")
    w, P, frames, ts = make_case("c1", grid=32, nptl=total)
    P.mpi_rank = rank
    n = shard_count(total, world, rank)
    o = Oracle(P, 4 * total)
    rec, steps = run_intervals(o, frames, ts, nptl=n, particle_v0=w.particle_v0, pmin_split=1.05, split_ratio=1.05)
    red = reduce_diagnostics(rec[-1], dist)
    np.save(os.path.join(outdir, f"ptl_{rank}.npy"), o.download_particles())
    if rank == 0:
        np.savez(os.path.join(outdir, "reduced.npz"), fglobal=red["fglobal"], quick=red["quick"], pmax=red["pmax"],
                 **{f"flocal{k}": a for k, a in enumerate(red["flocal"]) if a is not None})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
