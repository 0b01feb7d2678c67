"""C5 (3-D flux rope) A/B probe: the frames are generated ONCE, then every variant (environment switches read by the
library at launch time) pushes the same population through two MHD intervals.
usage: python scripts/r02/c5_probe.py <grid> <nptl> 'name1:ENV=VAL;ENV=VAL' name2: ...      (run on the GPU box)"""
import os
import sys
import time

sys.path.insert(0, ".")
from stochastic_parker_b200 import GpatSim, config, mhd  # noqa: E402

grid, nptl = int(sys.argv[1]), int(sys.argv[2])
variants = []
for a in sys.argv[3:]:
    name, _, envs = a.partition(":")
    variants.append((name, dict(e.split("=", 1) for e in envs.split(";") if e)))
w = config.WORKLOADS["c5"].scaled(grid=grid, nptl=nptl)
cfg = mhd.mhd_config(w.nx, w.ny, w.nz, w.lx, w.ly, w.lz, w.dt_out, w.ndim)
P = config.build_params(w.conf_text(), cfg, w.ndim, nframes=200, cli=w.cli)
t = time.time()
frames = [mhd.make_frame(w.kind, w.nx, w.ny, w.nz, f, w.dt_out) for f in range(3)]
print(f"c5 grid {grid} nptl {nptl}: frames in {time.time() - t:.1f} s", flush=True)
box = [P.xmin, P.ymin, P.zmin, P.xmax, P.ymax, P.zmax]
for name, env in variants:
    for k in ("GPAT_SORT_TILE", "GPAT_NO_SPEC3D", "GPAT_NO_L3D", "GPAT_PUSH_MAXCTAS", "GPAT_PUSH_SORT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = GpatSim(P, w.nptl_max)
    g.upload_fields(0, frames[0])
    for i in (1, 2):
        g.upload_fields(1, frames[i])
        if i == 1:
            g.inject_uniform(nptl, 0.0, 1, w.particle_v0, 0.0, w.dt_out, box, w.power_index)
        steps = g.particle_mover((i - 1) * w.dt_out, w.dt_out, 100, 1, 0)
        tm = g.timings()
        if i == 2:
            print(f"{name:28s} {env}: {steps} steps, push {tm.push_ms:.1f} ms, mover {tm.mover_ms:.1f} ms = "
                  f"{steps / (tm.mover_ms * 1e-3):.3e} steps/s", flush=True)
        g.swap_fields()
    g.close()
