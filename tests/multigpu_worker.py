"""Worker of test_gpu_parity.py::test_two_gpus_nccl_allreduce: one rank = one GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import make_case  # noqa: E402
from stochastic_parker_b200 import GpatSim, bootstrap_comm, rank_info, run_intervals, shard_count  # noqa: E402


def main():
    outdir, total = sys.argv[1], int(sys.argv[2])
    rank, world, local = rank_info()
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # rendezvous only; the data path is the library's own NCCL communicator
    w, P, frames, ts = make_case("c1", grid=64, nptl=total)
    P.mpi_rank = rank
    n = shard_count(total, world, rank)
    g = GpatSim(P, 4 * total, device=local)
    bootstrap_comm(g, dist)
    rec, steps = run_intervals(g, frames, ts, nptl=n, particle_v0=w.particle_v0, pmin_split=1.05, split_ratio=1.05)
    red = rec[-1]
    np.save(os.path.join(outdir, f"ptl_{rank}.npy"), g.download_particles())
    np.savez(os.path.join(outdir, f"reduced_{rank}.npz"), fglobal=red["fglobal"], quick=red["quick"], pmax=red["pmax"],
             **{f"flocal{k}": a for k, a in enumerate(red["flocal"]) if a is not None})
    g.comm_destroy()
    g.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
