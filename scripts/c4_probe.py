"""C4 (D_pp layout L2E) throughput probe: same physics on two grid sizes and, optionally, C2-like base
physics on the big grid -- separates 'the D_pp kernel is slower' from 'the 4096^2 store misses the L2'.
usage: python scripts/c4_probe.py <workload> <grid> <nptl> [intervals]      (run on the GPU box)"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from stochastic_parker_b200 import GpatSim, config, mhd  # noqa: E402

key, grid, nptl = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
nint = int(sys.argv[4]) if len(sys.argv) > 4 else 2
w = config.WORKLOADS[key].scaled(grid=grid, nptl=nptl)
cfg = mhd.mhd_config(w.nx, w.ny, w.nz, w.lx, w.ly, w.lz, w.dt_out, w.ndim)
P = config.build_params(w.conf_text(), cfg, w.ndim, nframes=200, cli=w.cli)
t = time.time()
frames = [mhd.make_frame(w.kind, w.nx, w.ny, w.nz, f, w.dt_out) for f in range(nint + 1)]
print(f"{key} grid {grid} nptl {nptl}: frames in {time.time() - t:.1f} s", flush=True)
g = GpatSim(P, w.nptl_max)
g.upload_fields(0, frames[0])
box = [P.xmin, P.ymin, P.zmin, P.xmax, P.ymax, P.zmax]
for i in range(1, nint + 1):
    g.upload_fields(1, frames[i])
    if i == 1:  # one population, all at the start of the interval: every interval pushes the same work
        g.inject_uniform(nptl, 0.0, 1, w.particle_v0, 0.0, 0.0, box, w.power_index)
    t = time.time()
    steps = g.particle_mover((i - 1) * w.dt_out, w.dt_out, 100, 1, 0)
    dt = time.time() - t
    n = len(g.download_particles())
    print(f"interval {i}: {steps} steps ({steps / max(n, 1):.0f} per particle, {n} particles) in {dt * 1e3:.1f} ms"
          f" = {steps / dt:.3e} steps/s", flush=True)
    g.swap_fields()
g.close()
