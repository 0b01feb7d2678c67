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


# Scaled-down variants of BASELINE.json's configs, one per model switch of the path (shared by the GPU
# parity tests and by the golden vectors computed from the reference's own Fortran).
CASES = {
    "c1_2d": dict(key="c1", grid=64),
    "c1_2d_no_time_interp": dict(key="c1", grid=64, cli=dict(time_interp=0)),
    "c1_2d_check_drift": dict(key="c1", grid=64, cli=dict(check_drift_2d=1)),
    "c1_2d_nlgc": dict(key="c1", grid=64, conf=dict(dt_min_rel=1e-3), cli=dict(nlgc=1, kperp_kpara=0.05)),
    "c1_2d_acc_region": dict(key="c1", grid=64, conf=dict(acc_region_flag=1)),
    "c1_2d_include_3rd": dict(key="c1", grid=64, cli=dict(include_3rd_dim=1)),
    "c2_flare_open": dict(key="c2", grid=64),
    "c3_shock_open": dict(key="c3", grid=64),
    "c4_dpp_wave_shear": dict(key="c4", grid=64),
    "c4_dpp_strong_kret0": dict(key="c4", grid=64, conf=dict(kret=0.0), cli=dict(weak_scattering=0)),
    "c1_2d_focused_transport": dict(key="c1", grid=64, conf=dict(dt_min_rel=1e-4),
                                    cli=dict(focused_transport=1, duu_init=5.0)),
    "c4_2d_focused_transport_dpp": dict(key="c4", grid=64, conf=dict(dt_min_rel=1e-3),
                                        cli=dict(focused_transport=1, duu_init=5.0, nlgc=1, kperp_kpara=0.05)),
    "c1_2d_ft_include_3rd": dict(key="c1", grid=64, conf=dict(dt_min_rel=1e-3),
                                 cli=dict(focused_transport=1, duu_init=5.0, include_3rd_dim=1)),
    "c5_3d_ft": dict(key="c5", grid=32, conf=dict(dt_min_rel=1e-3), cli=dict(focused_transport=1, duu_init=5.0)),
    "s1_shock_1d": dict(key="s1", grid=256),
    "s1_shock_1d_dpp_nlgc": dict(key="s1", grid=256, conf=dict(dt_min_rel=1e-3),
                                 cli=dict(dpp_wave=1, dpp_shear=1, nlgc=1, kperp_kpara=0.05)),
    "c5_3d": dict(key="c5", grid=32),
    "c5_3d_acc_surfaces_union": dict(key="c5", grid=32, conf=dict(acc_region_flag=1),
                                     cli=dict(acc_by_surface=1, surface_norm1="+z", surface2_existed=1,
                                              surface_norm2="-y")),
    "c5_3d_acc_surface_no_time_interp": dict(key="c5", grid=32, conf=dict(acc_region_flag=1),
                                             cli=dict(acc_by_surface=1, surface_norm1="-x", time_interp=0)),
    "c5_3d_ft_acc_surfaces_intersection": dict(key="c5", grid=32, conf=dict(acc_region_flag=1, dt_min_rel=1e-3),
                                               cli=dict(focused_transport=1, duu_init=5.0, acc_by_surface=1,
                                                        surface_norm1="+y", surface2_existed=1,
                                                        surface_norm2="-z", is_intersection=1)),
    "c5_3d_dpp_nlgc": dict(key="c5", grid=32, conf=dict(kpara0=0.02, dt_min_rel=1e-3),
                           cli=dict(dpp_wave=1, dpp_shear=1, nlgc=1, kperp_kpara=0.05)),
}


# ---- golden vectors of the reference (tests/golden/ref_f90_*.npz) ---------------------------------
REF_GOLDEN_CASES = [n for n in sorted(CASES) if "acc_surface" not in n]
# whole-interval runs of the focused-transport pushers take ~1e5 steps per particle at the cases' dt_min_rel
_GOLDEN_NPTL = {"c1_2d_focused_transport": 2, "c4_2d_focused_transport_dpp": 6, "c1_2d_ft_include_3rd": 6, "c5_3d_ft": 6}


def _pack_sparse(out, key, a):
    a = np.asarray(a)
    idx = np.flatnonzero(a)
    out[key + ".shape"] = np.array(a.shape, dtype=np.int64)
    out[key + ".idx"] = idx.astype(np.int64)
    out[key + ".val"] = a.reshape(-1)[idx].copy()


def unpack_sparse(z, key):
    a = np.zeros(int(np.prod(z[key + ".shape"])))
    a[z[key + ".idx"]] = z[key + ".val"]
    return a.reshape(tuple(z[key + ".shape"]))


def golden_collect(make_sim, name):
    """Everything the golden file of one case holds, computed by `make_sim(P, nptl_max)` -- RefSim (the
    reference's Fortran, when the file is generated), the C oracle (CPU test) or the GPU library (GPU test).
    Three runs per case: injection with the three momentum distributions; 1 + 40 single pushes; two full MHD
    intervals of solve_transport_equation with fine steps, splitting, escapes and every histogram."""
    import hashlib
    from stochastic_parker_b200.driver import run_intervals
    out = {}
    w, P, frames, ts = make_case(**CASES[name], nptl=64)
    P.strict_math = 1
    # (1) gradients, interpolation, injection
    s = make_sim(P, w.nptl_max)
    s.upload_fields(0, frames[0])
    if P.time_interp:
        s.upload_fields(1, frames[1])
    out["grad_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(s.get_fields(0)).tobytes()).digest(), dtype=np.uint8)
    rng = np.random.default_rng(11)
    n = 48
    x = rng.uniform(P.xmin, P.xmax, n)
    y = rng.uniform(P.ymin, P.ymax, n) if P.ndim > 1 else np.full(n, P.ymin)
    z = rng.uniform(P.zmin, P.zmax, n) if P.ndim > 2 else np.full(n, P.zmin)
    rt = rng.uniform(0, 1, n)
    out["interp_x"], out["interp_y"], out["interp_z"], out["interp_rt"] = x, y, z, rt
    out["interp"] = s.interp(x, y, z, rt)
    for d in (1, 0, 2):
        s.inject_uniform(16, 0.0, d, w.particle_v0, ts[0], ts[1] - ts[0], box_of(P), w.power_index)
    out["inject"] = s.download_particles()
    # (2) single pushes from the injected population (all three momentum distributions present)
    so = s.debug_push_n(ts[0], ts[1] - ts[0], 1)
    out["steps1"], out["steps1_count"] = s.download_particles(), np.int64(so)
    so = s.debug_push_n(ts[0], ts[1] - ts[0], 40)
    out["steps41"], out["steps41_count"] = s.download_particles(), np.int64(so)
    s.close()
    # (3) two MHD intervals; open boundaries: inject next to an outflow face so that particles escape
    nptl = _GOLDEN_NPTL.get(name, 32)
    w, P, frames, ts = make_case(**CASES[name], nptl=nptl)
    P.strict_math = 1
    s = make_sim(P, w.nptl_max)
    box = box_of(P)
    if name.startswith("c2"):      # flare sheet: the outflow leaves through the high-y face around x = lx/2
        box = [P.xmin + 0.45 * P.lx, P.ymax - 1.5 * P.dy, P.zmin, P.xmin + 0.55 * P.lx, P.ymax, P.zmax]
    elif P.pbc[0] == 1:            # shocks: the flow leaves through the high-x face
        box[0] = P.xmax - 1.5 * P.dx
    kw = dict(nptl=nptl, dist_flag=1, particle_v0=w.particle_v0, inject_new_ptl=True, split_flag=1, part_box=box,
              pmin_split=1.05, split_ratio=1.05, num_fine_steps=2, dump_escaped_dist=True, dump_escaped=True)
    recs, steps = run_intervals(s, frames, ts, **kw)
    out["run_steps"] = np.int64(steps)
    out["run_particles"] = s.download_particles()
    c = s.counters()
    out["run_counters_int"] = np.array([c.nptl_current, c.nptl_split, c.nptl_escaped, c.tag_max], dtype=np.int64)
    out["run_counters_f"] = np.array([c.leak, c.leak_negp])
    for r in recs:
        f = f"run_f{r['frame']}_"
        out[f + "fglobal"] = r["fglobal"]
        q = np.array(r["quick"], dtype=np.float64)
        if not getattr(s, "quick_is_average", False) and q[0] > 0:
            q[5] = q[5] / q[0]   # the ABI returns the sum of dt, the reference writes the average (DG:153-157)
        out[f + "quick"], out[f + "pmax"] = q, np.float64(r["pmax"])
        for k, a in enumerate(r["flocal"]):
            if a is not None:
                _pack_sparse(out, f + f"flocal{k + 1}", a)
        if "fescaped" in r:
            out[f + "fescaped"] = r["fescaped"]
            out[f + "escaped_particles"] = r["escaped_particles"]
            for k, d in enumerate(r["fescaped_local"]):
                if d is not None:
                    for ax in "xyz":
                        if d[ax] is not None:
                            _pack_sparse(out, f + f"fescaped{k + 1}_{ax}", d[ax])
    s.close()
    return out


def tracked_split_population(P):
    """Six hand-made particles + a tag table for split_particle's tracking branch (PM:5452-5473).  Every tracked
    chain is distinct, as in a real run -- two particles that map to the SAME particles_tracked slot would make the
    result depend on the serial loop order of the reference, which a parallel split does not reproduce."""
    tags = np.array([[0, 5, 1, 3], [0, 5, 2, 2], [0, 9, 1, 1], [1, 5, 1, 3]], dtype=np.int32)
    ptl = np.zeros(6, dtype=PARTICLE_DTYPE)
    ptl["origin"] = [0, 0, 0, 1, 1, 0]
    ptl["tag_injected"] = [-5, -9, 7, -5, 6, -9]
    ptl["tag_splitted"] = [-1, -1, 1, -1, 1, -1]
    ptl["split_times"] = [0, 0, 0, 1, 0, 2]
    ptl["p"] = P.p0 * np.array([3.0, 3.0, 3.0, 5.0, 3.0, 1.0])     # the last one is below its threshold
    ptl["weight"] = 0.5 ** ptl["split_times"].astype(float)
    ptl["count_flag"] = 1
    ptl["nsteps_tracked"] = [3, 3, 0, 3, 0, 3]
    ptl["x"] = np.arange(6) * 0.1
    return tags, ptl
