"""N > 1 on CPU (gloo, world_size 2): the particle path shards with no data-path collective.

Ranks hold disjoint particle shards (`origin` = rank enters the Philox key), run the same
intervals independently, and meet only in the diagnostics reduction.  The reduced histograms
must equal -- bit for bit, weights being dyadic -- the histograms of the union of the shards
computed by one process.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from helpers import make_case
from oracle.oracle import Oracle
from stochastic_parker_b200 import shard_count

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_count():
    assert [shard_count(10, 4, r) for r in range(4)] == [3, 3, 2, 2]
    assert sum(shard_count(1_000_003, 8, r) for r in range(8)) == 1_000_003
    assert shard_count(5, 1, 0) == 5
    with pytest.raises(ValueError):
        shard_count(5, 2, 2)


@pytest.mark.timeout(300)
def test_two_ranks_reduce_to_the_union(tmp_path):
    total = 601  # odd: the shards differ in size
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multirank_worker.py"), str(tmp_path), str(total)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    shards = [np.load(tmp_path / f"ptl_{k}.npy") for k in range(2)]
    assert [int(np.unique(s["origin"])[0]) for s in shards] == [0, 1]
    # injected tags overlap (each rank counts from 0) but the streams do not
    assert not np.any(np.isin(shards[0]["x"], shards[1]["x"]))
    red = np.load(tmp_path / "reduced.npz")
    w, P, frames, ts = make_case("c1", grid=32, nptl=total)
    o = Oracle(P, 8 * total)
    o.upload_particles(np.concatenate(shards))
    d = o.diagnostics(True)
    assert np.array_equal(red["fglobal"], d["fglobal"])
    for k in range(3):
        assert np.array_equal(red[f"flocal{k}"], d["flocal"][k])
    assert red["quick"][0] == len(shards[0]) + len(shards[1])
    assert red["quick"][2] == d["quick"][2]                       # sum of weights
    assert red["quick"][6] == d["quick"][6] and red["quick"][7] == d["quick"][7]  # min / max dt
    assert float(red["pmax"]) == d["pmax"]


def test_reduce_diagnostics_covers_the_escaped_face_arrays():
    """reduce_diagnostics walks every array the MPI_REDUCEs of diagnostics.f90:881-905 and 1174-1230 touch;
    a stand-in communicator of two identical ranks (SUM doubles, MIN / MAX keep) makes that visible."""
    from stochastic_parker_b200 import reduce_diagnostics

    class TwoEqualRanks:
        class ReduceOp:
            SUM, MIN, MAX = "sum", "min", "max"

        @staticmethod
        def all_reduce(t, op):
            if op == "sum":
                t.mul_(2)

    d = dict(fglobal=np.ones((4, 1)), flocal=[np.ones((1, 2, 2, 3, 1)), None, None, None],
             quick=np.arange(1.0, 9.0), pmax=3.0, fescaped=np.ones((4, 4, 1)),
             fescaped_local=[dict(x=np.ones((2, 1, 2, 3, 1)), y=np.full((2, 1, 2, 3, 1), 0.5), z=None), None, None, None])
    r = reduce_diagnostics(d, TwoEqualRanks)
    assert r["fglobal"].sum() == 8 and r["flocal"][0].sum() == 24 and r["flocal"][1] is None
    assert list(r["quick"]) == [2, 4, 6, 8, 10, 12, 7, 8] and r["pmax"] == 3.0
    assert r["fescaped"].sum() == 32
    assert r["fescaped_local"][0]["x"].sum() == 24 and r["fescaped_local"][0]["y"].sum() == 12
    assert r["fescaped_local"][0]["z"] is None and r["fescaped_local"][1] is None
    assert d["fescaped_local"][0]["x"].sum() == 12          # the input record is left alone
