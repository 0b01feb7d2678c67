"""Throughput of the paths outside the five named configs -- 1-D, focused transport, turbulence maps -- through
gpat_debug_push_n (a fixed number of push_particle_* calls per particle: identical work for every routing).
usage: python scripts/r02/alt_probe.py [nsteps]      (run on the GPU box)
GPAT_ALT_STRICT=1 routes these runs to the reference-order kernels (the state before the kSpecAlt instantiations)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from helpers import box_of, make_case  # noqa: E402
from stochastic_parker_b200 import GpatSim, mhd  # noqa: E402

NSTEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 200
CASES = [
    ("ft_2d", dict(key="c1", grid=1024, conf=dict(dt_min_rel=1e-4), cli=dict(focused_transport=1, duu_init=5.0)), 1_000_000, False),
    ("ft_2d_dpp_nlgc", dict(key="c4", grid=1024, conf=dict(dt_min_rel=1e-3),
                            cli=dict(focused_transport=1, duu_init=5.0, nlgc=1, kperp_kpara=0.05)), 1_000_000, False),
    ("ft_2d_3rd", dict(key="c1", grid=1024, conf=dict(dt_min_rel=1e-3),
                       cli=dict(focused_transport=1, duu_init=5.0, include_3rd_dim=1)), 1_000_000, False),
    ("ft_3d", dict(key="c5", grid=128, conf=dict(dt_min_rel=1e-3), cli=dict(focused_transport=1, duu_init=5.0)), 1_000_000, False),
    ("shock_1d", dict(key="s1", grid=4096), 1_000_000, False),
    ("maps_2d", dict(key="c1", grid=1024, conf=dict(dt_min_rel=1e-4)), 1_000_000, True),
    ("maps_3d", dict(key="c5", grid=128, conf=dict(dt_min_rel=1e-4)), 1_000_000, True),
    ("plain_2d (C1 kernel, for scale)", dict(key="c1", grid=1024), 1_000_000, False),
]
only = os.environ.get("ALT_PROBE_ONLY")
for name, kw, nptl, maps in CASES:
    if only and only not in name:
        continue
    w, P, frames, _ = make_case(**kw, nptl=nptl, nframes=2)
    if maps:
        P.deltab_flag = 1
        P.correlation_flag = 1
    line = f"{name:34s}"
    for route in os.environ.get("ALT_PROBE_ROUTES", "1,0").split(","):
        os.environ["GPAT_ALT_STRICT"] = route
        g = GpatSim(P, w.nptl_max)
        g.upload_fields(0, frames[0])
        g.upload_fields(1, frames[1])
        if maps:
            for slot in (0, 1):
                m = mhd.make_turbulence_maps(P.nx, P.ny, P.nz, slot, ndim=P.ndim)
                g.upload_turbulence(0, slot, m[0], m[1])
                g.upload_turbulence(1, slot, m[2], m[3])
        best = 0.0
        if os.environ.get("ALT_PROBE_MODE") == "mover":
            # whole MHD intervals through gpat_particle_mover (cell sort, roll-back, compaction): what a run does.
            # One population injected at the start of the first interval; the second interval is the one timed.
            n = int(os.environ.get("ALT_PROBE_MOVER_NPTL", "200000"))
            g.inject_uniform(n, 0.0, 1, w.particle_v0, 0.0, 0.0, box_of(P), w.power_index)
            for i in (1, 2):
                steps = g.particle_mover((i - 1) * w.dt_out, w.dt_out, 100, 1, 0)
                tm = g.timings()
                best = steps / (tm.mover_ms * 1e-3)
                g.swap_fields()
                g.upload_fields(1, frames[(i + 1) % 2])   # any valid frame: throughput only
                if maps:
                    m = mhd.make_turbulence_maps(P.nx, P.ny, P.nz, (i + 1) % 2, ndim=P.ndim)
                    g.upload_turbulence(0, 1, m[0], m[1])
                    g.upload_turbulence(1, 1, m[2], m[3])
            line += f"  [{steps / max(1, n):.0f} steps/particle]"
        else:
            g.inject_uniform(nptl, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
            for rep in range(3):
                t = time.time()
                steps = g.debug_push_n(0.0, w.dt_out, NSTEPS)
                dt = time.time() - t
                best = max(best, steps / dt)
        line += f"  {'reference-order' if route == '1' else 'production'} {best:.3e}"
        g.close()
    print(line, flush=True)
