import sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
from helpers import make_case, CASES, sort_by_key, rel_err
from oracle.oracle import Oracle
from stochastic_parker_b200 import GpatSim
from stochastic_parker_b200.driver import run_intervals
for name, nptl in [("c1_2d_focused_transport", 200), ("c5_3d_ft", 300), ("c4_2d_focused_transport_dpp", 200)]:
    w, P, frames, ts = make_case(**CASES[name], nptl=nptl)
    Pg = P.copy(); Pg.strict_math = 0
    g, o = GpatSim(Pg, w.nptl_max), Oracle(P, w.nptl_max)
    kw = dict(nptl=nptl, dist_flag=1, particle_v0=w.particle_v0, split_flag=1, num_fine_steps=2)
    rg, sg = run_intervals(g, frames, ts, **kw); ro, so = run_intervals(o, frames, ts, **kw)
    a, b = sort_by_key(g.download_particles()), sort_by_key(o.download_particles())
    print(name, 'steps', sg, so, 'n', len(a), len(b))
    if len(a) == len(b):
        for f in ("x", "y", "z", "p", "mu", "v", "t"):
            e = rel_err(a[f], b[f]) if f not in "xyz" else np.abs(a[f] - b[f]) / max(1.0, np.abs(b[f]).max())
            print('  ', f, 'max %.2e' % e.max(), 'p90 %.2e' % np.quantile(e, 0.9), 'p50 %.2e' % np.quantile(e, 0.5), 'n>1e-9', int((e > 1e-9).sum()))
    print('   fglobal diff', [float(np.abs(x["fglobal"] - y["fglobal"]).sum()) for x, y in zip(rg, ro)])
