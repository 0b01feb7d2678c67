import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
from helpers import make_case, box_of
from oracle.oracle import Oracle
from stochastic_parker_b200.abi import rng_steps
for key, grid in (("c1", 1024), ("c3", 0)):
    kw = dict(key=key, nptl=20000, nframes=4)
    if grid: kw["grid"] = grid
    else: kw["grid"] = 2048
    w, P, frames, ts = make_case(**kw)
    o = Oracle(P, w.nptl_max)
    o.upload_fields(0, frames[0]); o.upload_fields(1, frames[1])
    o.inject_uniform(20000, 0.0, 1, w.particle_v0, 0.0, 0.0, box_of(P), w.power_index)
    prev = None
    for i in (1, 2, 3):
        before = rng_steps(o.download_particles()).astype(np.int64)
        keys0 = o.download_particles()["tag_injected"].copy()
        t = time.time()
        o.particle_mover((i - 1) * w.dt_out, w.dt_out, 100, 1, 0)
        p = o.download_particles()
        # particles may be reordered by removal; match by tag
        order0 = np.argsort(keys0); order1 = np.argsort(p["tag_injected"])
        if len(p) == len(keys0):
            st = np.empty(len(p), dtype=np.int64)
            st[order1] = (rng_steps(p).astype(np.int64)[order1] - before[order0])
            st_by_tag = (rng_steps(p).astype(np.int64)[order1] - before[order0])
            print(key, 'interval', i, 'steps/particle mean %.0f min %d p10 %.0f p50 %.0f p90 %.0f p99 %.0f max %d' % (st_by_tag.mean(), st_by_tag.min(), *np.quantile(st_by_tag, [0.1, 0.5, 0.9, 0.99]), st_by_tag.max()), '%.1fs' % (time.time() - t))
            if prev is not None and len(prev) == len(st_by_tag):
                print('   correlation with the previous interval: %.3f' % np.corrcoef(prev, st_by_tag)[0, 1])
            prev = st_by_tag
        else:
            print(key, 'interval', i, 'population changed', len(keys0), len(p)); prev = None
        o.swap_fields(); o.upload_fields(1, frames[min(i + 1, 3)])
