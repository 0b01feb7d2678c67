"""Host side of particle tracking: the tag table of the reference's two-run workflow.

`select_tags` is the post-processing step docs/source/development/particle_module.rst:51-99
describes (written from that description): pick particles from a dump of the first run, trace
every one back through its splits to the injected particle, and sort the table by origin,
tag_injected and the tag_splitted chain -- the order `is_particle_selected`
(particle_module.f90:5920-5959) relies on when it searches the table with findloc.
"""
from __future__ import annotations

import numpy as np


def select_tags(ptl: np.ndarray, index: np.ndarray) -> np.ndarray:
    """(nselected, nsplit_max + 2) int32: origin, tag_injected, tag_splitted after split 1, 2, ...

    `ptl` is a particle dump (abi.PARTICLE_DTYPE), `index` the rows to track."""
    origin = ptl["origin"][index].astype(np.int64)
    tag_injected = np.abs(ptl["tag_injected"][index].astype(np.int64))
    tag_splitted = np.abs(ptl["tag_splitted"][index].astype(np.int64))
    split_times = ptl["split_times"][index].astype(np.int64)
    n = len(index)
    nsplit_max = int(split_times.max()) if n else 0
    tags = np.zeros((n, nsplit_max + 2), dtype=np.int32)
    tags[:, 0] = origin
    tags[:, 1] = tag_injected
    for i in range(n):
        ns = int(split_times[i])
        if ns > 0:
            tags[i, ns + 1] = tag_splitted[i]
            for k in range(ns, 1, -1):  # the parent either kept its tag or the child added 2**(k-1)
                cur = int(tags[i, k + 1])
                tags[i, k] = cur - 2 ** (k - 1) if cur > 2 ** (k - 1) else cur
    order = np.lexsort(tags[:, ::-1].T)
    return np.ascontiguousarray(tags[order])
