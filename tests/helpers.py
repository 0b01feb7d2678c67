"""Shared builders for the parity tests: small versions of BASELINE.json's configs."""
from __future__ import annotations

import numpy as np

from stochastic_parker_b200 import config, mhd
from stochastic_parker_b200.abi import PARTICLE_DTYPE, rng_steps

KEY_FIELDS = ("origin", "tag_injected", "tag_splitted")


def make_case(key="c1", grid=64, nptl=256, nframes=3, conf=None, cli=None, nz=None, **wl):
    """(workload, Params, frames, tstamps) for a scaled-down named config."""
    w = config.WORKLOADS[key].scaled(grid=grid, nptl=nptl)
    if nz is not None:
        w.nz = nz
    for k, v in wl.items():
        setattr(w, k, v)
    if conf:
        w.conf = dict(w.conf, **conf)
    if cli:
        w.cli = dict(w.cli, **cli)
    cfg = mhd.mhd_config(w.nx, w.ny, w.nz, w.lx, w.ly, w.lz, w.dt_out, w.ndim)
    P = config.build_params(w.conf_text(), cfg, w.ndim, nframes=200, cli=w.cli)
    frames = [mhd.make_frame(w.kind, w.nx, w.ny, w.nz, f, w.dt_out) for f in range(nframes)]
    tstamps = [f * w.dt_out for f in range(nframes)]
    return w, P, frames, tstamps


def box_of(P):
    return [P.xmin, P.ymin, P.zmin, P.xmax, P.ymax, P.zmax]


def sort_by_key(ptl: np.ndarray) -> np.ndarray:
    order = np.lexsort((ptl["tag_splitted"], ptl["tag_injected"], ptl["origin"]))
    return ptl[order]


def rel_err(a, b, floor=1e-300):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)


def assert_particles_close(got, ref, rtol, what="", int_exact=True, frac_outliers=0.0, atol_pos=0.0):
    """Compare two particle sets in the SAME order.  Positions are compared with a tolerance
    relative to the box scale (they can be close to 0), p/t/dt relative to their value."""
    assert len(got) == len(ref), f"{what}: {len(got)} vs {len(ref)} particles"
    if len(got) == 0:
        return
    if int_exact:
        for f in ("split_times", "count_flag", "origin", "nsteps_pushed", "tag_injected", "tag_splitted"):
            assert np.array_equal(got[f], ref[f]), f"{what}: integer field {f} differs"
        assert np.array_equal(rng_steps(got), rng_steps(ref)), f"{what}: RNG step counters differ"
    bad = np.zeros(len(got), dtype=bool)
    worst = {}
    for f in ("x", "y", "z"):
        scale = max(1.0, float(np.max(np.abs(ref[f]))))
        e = np.abs(got[f] - ref[f]) / scale
        worst[f] = float(e.max())
        bad |= e > max(rtol, atol_pos)
    for f in ("p", "t", "dt", "weight", "mu", "v"):
        e = rel_err(got[f], ref[f])
        worst[f] = float(e.max())
        bad |= e > rtol
    nbad = int(bad.sum())
    assert nbad <= frac_outliers * len(got), f"{what}: {nbad}/{len(got)} particles beyond {rtol:g}; worst {worst}"
    return worst


def assert_particles_identical(got, ref, what=""):
    """Field-by-field bit equality (the 2 alignment bytes of the record are not data)."""
    assert len(got) == len(ref), f"{what}: {len(got)} vs {len(ref)} particles"
    for f in PARTICLE_DTYPE.names:
        a, b = np.ascontiguousarray(got[f]), np.ascontiguousarray(ref[f])
        if a.dtype.kind == "f":
            a, b = a.view(np.uint64), b.view(np.uint64)
        bad = np.nonzero(a != b)[0]
        assert len(bad) == 0, f"{what}: field {f} differs at {len(bad)} particles, first {bad[:5]}: {got[f][bad[:3]]} vs {ref[f][bad[:3]]}"
