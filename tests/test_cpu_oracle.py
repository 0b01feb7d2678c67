"""CPU tests of the oracle (oracle/gpat_oracle.c): the checker itself must be trustworthy.

The reference ships NO golden vectors for the particle path, so the
oracle is pinned here by what can be pinned without the Fortran binary:
  * the published Philox4x32-10 known-answer vectors (Random123 kat_vectors),
  * a second, independent restatement of the 2-D step in numpy (oracle/np_step.py),
  * closed-form properties of the algorithm (exact interpolation/gradients of linear fields,
    variance 2*kappa*t of the stochastic step, dyadic split weights, histogram identities),
  * the reference's own Python-side files (tests/golden/, generated from /root/reference by
    tests/golden/make_golden.py).
"""
import os

import numpy as np
import pytest

from helpers import assert_particles_close, assert_particles_identical, box_of, make_case, rel_err, sort_by_key
from oracle import np_step
from oracle.oracle import Oracle, philox4x32_10
from stochastic_parker_b200 import run_intervals
from stochastic_parker_b200.abi import PARTICLE_DTYPE, RNG_TABLE, rng_steps


# ---- RNG -----------------------------------------------------------------------------------
# Random123 kat_vectors, "philox4x32 10" lines (counter, key -> output)
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_known_answers(ctr, key, want):
    assert tuple(philox4x32_10(ctr, key)) == want


def test_injection_uses_documented_philox_stream():
    """inject_one_particle draws x, y, z, mu, t in that order (particle_module.f90:401-436)
    from counter (0xFFFFFFFF.., tag) blocks; here: positions are uniform in the box, mu in
    [-mu_max, mu_max] with mu_max = dble(0.99f), t in the frame, and two ranks never share a
    stream (the key carries `origin`)."""
    w, P, frames, _ = make_case("c1", grid=32, nptl=4000)
    o = Oracle(P, 8000)
    o.inject_uniform(4000, 1e-5, 1, w.particle_v0, 0.3, 0.1, box_of(P), 6.2)
    a = o.download_particles()
    assert len(a) == 4000
    assert a["x"].min() >= P.xmin and a["x"].max() <= P.xmax
    assert a["y"].min() >= P.ymin and a["y"].max() <= P.ymax
    assert np.all(np.abs(a["mu"]) <= float(np.float32(0.99)))
    assert np.all((a["t"] >= 0.3) & (a["t"] <= 0.4))
    assert np.all(a["p"] == P.p0) and np.all(a["weight"] == 1.0) and np.all(a["dt"] == 1e-5)
    assert np.array_equal(a["tag_injected"], np.arange(4000)) and np.all(a["tag_splitted"] == 1)
    assert abs(a["x"].mean() - 0.5 * (P.xmin + P.xmax)) < 5 * P.lx / np.sqrt(12 * 4000)
    P2 = P.copy()
    P2.mpi_rank = 1
    o2 = Oracle(P2, 8000)
    o2.inject_uniform(4000, 1e-5, 1, w.particle_v0, 0.3, 0.1, box_of(P), 6.2)
    b = o2.download_particles()
    assert np.all(b["origin"] == 1) and not np.any(a["x"] == b["x"])


# ---- gradients and interpolation ---------------------------------------------------------------
def _linear_frame(nx, ny, coef):
    """8-variable frame whose variable v is a0 + ax*i + ay*j on storage indices (exact in FP32)."""
    j, i = np.meshgrid(np.arange(ny + 4), np.arange(nx + 4), indexing="ij")
    f = np.zeros((ny + 4, nx + 4, 8), dtype=np.float32)
    for v, (a0, ax, ay) in enumerate(coef):
        f[..., v] = a0 + ax * i + ay * j
    return f


def test_gradients_of_linear_fields_are_exact():
    """calc_fields_gradients (mhd_data_parallel.f90:533-566): centred and one-sided 3-point
    differences are exact for linear data, including the ghost edges."""
    w, P, _, _ = make_case("c1", grid=16, nptl=8)
    coef = [(1.0, 0.5, -0.25), (0.0, 2.0, 1.0), (3.0, 0.0, 0.0), (1.0, 0.125, 0.0),
            (-2.0, 1.0, 1.0), (0.5, -1.0, 2.0), (0.25, 0.0, -0.5), (4.0, 0.75, 0.25)]
    f = _linear_frame(P.nx, P.ny, coef)
    o = Oracle(P, 16)
    o.upload_fields(0, f)
    g = o.get_fields(0)[0]  # (ny+4, nx+4, 32)
    for v, (_, ax, ay) in enumerate(coef):
        assert np.all(g[..., 8 + 3 * v + 0] == np.float32(ax * 0.5 / P.dx * 2.0))
        assert np.all(g[..., 8 + 3 * v + 1] == np.float32(ay * 0.5 / P.dy * 2.0))
        assert np.all(g[..., 8 + 3 * v + 2] == 0.0)  # unresolved z in 2-D


def test_gradients_match_numpy_restatement():
    w, P, frames, _ = make_case("c1", grid=24, nptl=8)
    o = Oracle(P, 16)
    o.upload_fields(0, frames[0])
    got = o.get_fields(0)[0]
    ref = np_step.gradients32(frames[0], P.dx, P.dy)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_interp_is_exact_for_bilinear_data_and_blends_in_time():
    """interp_fields (mhd_data_parallel.f90:1751-1793): bilinear weights reproduce a linear
    field, and the time blend is fields*(1-rt) + fields2*rt."""
    w, P, _, _ = make_case("c1", grid=16, nptl=8)
    coef = [(1.0, 0.5, -0.25)] * 8
    f0 = _linear_frame(P.nx, P.ny, coef)
    f1 = (3.0 * f0).astype(np.float32)
    o = Oracle(P, 16)
    o.upload_fields(0, f0)
    o.upload_fields(1, f1)
    rng = np.random.default_rng(0)
    n = 500
    x = rng.uniform(P.xmin, P.xmax, n)
    y = rng.uniform(P.ymin, P.ymax, n)
    rt = rng.uniform(0, 1, n)
    F = o.interp(x, y, np.zeros(n), rt)
    # storage index of a position: i = (x - xmin)/dx + 2 (Fortran index floor(px)+1, lower bound -1)
    exact = 1.0 + 0.5 * ((x - P.xmin) / P.dx + 2) - 0.25 * ((y - P.ymin) / P.dy + 2)
    assert np.max(np.abs(F[:, 0] - exact * (1 + 2 * rt))) < 1e-12
    # against the independent numpy restatement, all 32 slots
    fa1, fa2 = np_step.gradients32(f0, P.dx, P.dy), np_step.gradients32(f1, P.dx, P.dy)
    ref = np_step.interp32(fa1, fa2, P, x, y, rt)
    assert np.array_equal(F, ref)


# ---- one step: C oracle vs the independent numpy restatement ---------------------------------
@pytest.mark.parametrize("conf", [dict(), dict(mag_dependency=0), dict(momentum_dependency=0), dict(kret=0.0)])
def test_step_matches_numpy_restatement(conf):
    w, P, frames, _ = make_case("c1", grid=48, nptl=400, conf=conf)
    P.rng_mode = RNG_TABLE
    o = Oracle(P, w.nptl_max)
    u = np.random.default_rng(5).uniform(0, 1, (400, 2, 4))
    o.set_rng_table(u)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_uniform(400, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    assert o.debug_push_n(0.0, w.dt_out, 1) == 400
    after = o.download_particles()
    fa1 = np_step.gradients32(frames[0], P.dx, P.dy)
    fa2 = np_step.gradients32(frames[1], P.dx, P.dy)
    F = np_step.interp32(fa1, fa2, P, before["x"], before["y"], (before["t"] - 0.0) / w.dt_out)
    qdrift = float(np.float32(1.0) / np.float32(3 * P.pcharge))
    x, y, p, t, dt = np_step.push_2d(P, F, before["p"], P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out,
                                    u[before["tag_injected"], 0], before["x"], before["y"], before["t"], qdrift)
    for name, ref in (("x", x), ("y", y), ("p", p), ("t", t), ("dt", dt)):
        scale = max(1.0, np.abs(ref).max()) if name in "xy" else np.abs(ref)
        err = np.abs(after[name] - ref) / scale
        assert err.max() < 2e-15, (name, err.max())  # identical operations; pow() may differ by an ulp
    assert np.array_equal(rng_steps(after), rng_steps(before) + np.uint64(1))


# ---- statistics of the stochastic step -----------------------------------------------------------
def test_pure_diffusion_variance():
    """Uniform B along x, no flow, constant kappa: after time T the displacement variance is
    2*kpara*T along B and 2*kperp*T across it (docs/source/theory/parker_1d2d.rst; the noise
    is uniform on [-sqrt3, sqrt3], particle_module.f90:3547-3550).  With dx/dt = dp/dt = 0 the
    reference falls back to dt_min (particle_module.f90:3532-3534), so dt_min = dt_max here."""
    w, P, _, _ = make_case("c1", grid=32, nptl=20000, conf=dict(momentum_dependency=0, mag_dependency=0,
                                                               kpara0=0.002, kret=0.25, dt_min_rel=5e-3, dt_max_rel=5e-3))
    f = np.zeros((P.ny + 4, P.nx + 4, 8), dtype=np.float32)
    f[..., 3] = 1.0
    f[..., 4] = 1.0
    f[..., 7] = 1.0
    o = Oracle(P, 40000)
    o.upload_fields(0, f)
    o.upload_fields(1, f)
    n = 20000
    ptl = np.zeros(n, dtype=PARTICLE_DTYPE)
    ptl["x"] = 0.5 * (P.xmin + P.xmax)
    ptl["y"] = 0.5 * (P.ymin + P.ymax)
    ptl["p"] = P.p0
    ptl["weight"] = 1.0
    ptl["count_flag"] = 1
    ptl["tag_injected"] = np.arange(n)
    ptl["tag_splitted"] = 1
    ptl["dt"] = 1e-4
    o.upload_particles(ptl)
    T = 0.1
    o.particle_mover(0.0, T, 100, 1, 0)
    a = o.download_particles()
    assert len(a) == n and np.all(a["t"] == T)
    vx, vy = np.var(a["x"] - ptl["x"][0]), np.var(a["y"] - ptl["y"][0])
    # relative standard error of a variance estimate ~ sqrt(2/n) = 1 %; allow 5 sigma
    assert abs(vx / (2 * 0.002 * T) - 1) < 0.05
    assert abs(vy / (2 * 0.002 * 0.25 * T) - 1) < 0.05
    assert np.all(a["p"] == P.p0)  # div v = 0, no D_pp: momentum untouched


def test_adiabatic_compression_energises():
    """dp/dt = -p div(v)/3 (particle_module.f90:3477): in a uniformly converging flow every
    particle gains momentum by exp(-divv T/3)."""
    w, P, _, _ = make_case("c1", grid=32, nptl=64, conf=dict(momentum_dependency=0, mag_dependency=0,
                                                            kpara0=1e-6, kret=1.0, dt_min_rel=1e-3, dt_max_rel=1e-3))
    j, i = np.meshgrid(np.arange(P.ny + 4), np.arange(P.nx + 4), indexing="ij")
    f = np.zeros((P.ny + 4, P.nx + 4, 8), dtype=np.float32)
    c = 0.5
    f[..., 0] = -c * ((i - 2) * P.dx - 1.0)  # vx = -c (x - 1): div v = -c
    f[..., 3] = 1.0
    f[..., 6] = 1.0  # B along z: no in-plane field-aligned motion
    f[..., 7] = 1.0
    o = Oracle(P, 256)
    o.upload_fields(0, f)
    o.upload_fields(1, f)
    o.inject_uniform(64, 1e-5, 1, 1.0, 0.0, 1e-9, [0.9, 0.9, 0, 1.1, 1.1, 1], 6.2)
    o.particle_mover(0.0, 0.1, 100, 1, 0)
    a = o.download_particles()
    assert len(a) == 64
    assert np.max(np.abs(a["p"] / P.p0 / np.exp(c * 0.1 / 3.0) - 1)) < 2e-3  # first-order Euler in t


# ---- split / remove / histograms ------------------------------------------------------------------
def test_split_semantics():
    """split_particle (particle_module.f90:5430-5480): threshold pmin_split*p0*ratio**split_times,
    both copies get weight 0.5**(1+split_times), child tag = parent + 2**(split_times-1) (after the
    increment), append at the tail, silent stop at nptl_max."""
    w, P, _, _ = make_case("c1", grid=16, nptl=8)
    o = Oracle(P, 10)
    ptl = np.zeros(6, dtype=PARTICLE_DTYPE)
    ptl["p"] = P.p0 * np.array([1.0, 2.5, 4.5, 3.9, 9.0, 120.0])
    ptl["split_times"] = [0, 0, 1, 1, 2, 0]
    ptl["weight"] = 0.5 ** ptl["split_times"].astype(float)
    ptl["count_flag"] = 1
    ptl["tag_injected"] = np.arange(6)
    ptl["tag_splitted"] = 1
    o.upload_particles(ptl)
    o.split(2.0, 2.0)
    a = o.download_particles()
    # 1: below threshold; 2.5 > 2: split; 4.5 > 4: split; 3.9 < 4: no; 9 > 8: split; 120*p0 = 12 > pmax = 10: no
    assert len(a) == 9 and o.counters().nptl_split == 3
    assert list(a["split_times"][:6]) == [0, 1, 2, 1, 3, 0]
    assert list(a["weight"][:6]) == [1.0, 0.5, 0.25, 0.5, 0.125, 1.0]
    assert list(a["tag_injected"][6:]) == [1, 2, 4] and list(a["split_times"][6:]) == [1, 2, 3]
    assert list(a["weight"][6:]) == [0.5, 0.25, 0.125]
    assert list(a["tag_splitted"][6:]) == [1 + 2 ** 0, 1 + 2 ** 1, 1 + 2 ** 2]
    assert a["weight"].sum() == ptl["weight"].sum()  # splitting conserves the weight
    o.split(1.0001, 0.01)  # everything (but the particle beyond pmax) qualifies; only one slot is left
    assert o.counters().nptl_current == 10


def test_histograms_match_numpy_binning():
    """calc_particle_distributions (diagnostics.f90:757-879): global spectrum p in (pmin, pmax],
    local sets drop their top momentum bin (ip < npbins, diagnostics.f90:799)."""
    w, P, frames, ts = make_case("c1", grid=32, nptl=3000)
    o = Oracle(P, w.nptl_max)
    run_intervals(o, frames, ts, nptl=3000, particle_v0=w.particle_v0, dist_flag=2, pmin_split=1.2,
                  split_ratio=1.2)
    a = o.download_particles()
    d = o.diagnostics(True)
    pmin_log = np.log10(P.pmin)
    dp_log = (np.log10(P.pmax) - pmin_log) / P.npp_global
    sel = (a["p"] > P.pmin) & (a["p"] <= P.pmax)
    ip = np.floor((np.log10(a["p"][sel]) - pmin_log) / dp_log).astype(int)
    ref = np.bincount(np.minimum(ip, P.npp_global - 1), weights=a["weight"][sel], minlength=P.npp_global)
    assert np.array_equal(d["fglobal"][:, 0], ref)
    assert d["fglobal"].sum() == a["weight"][sel].sum()
    s = P.local[0]
    nrx, nry = P.nx // s.rx, P.ny // s.ry
    dpl = (np.log10(s.pmax) - np.log10(s.pmin)) / s.npbins
    ix = np.floor((a["x"] - P.xmin) / (P.lx / nrx)).astype(int)
    iy = np.floor((a["y"] - P.ymin) / (P.ly / nry)).astype(int)
    ipl = np.floor((np.log10(a["p"]) - np.log10(s.pmin)) / dpl).astype(int)  # 0-based: Fortran ip - 1
    ok = (ix >= 0) & (ix < nrx) & (iy >= 0) & (iy < nry) & (ipl >= 0) & (ipl < s.npbins - 1)
    ref = np.zeros((nry, nrx, s.npbins))
    np.add.at(ref, (iy[ok], ix[ok], ipl[ok]), a["weight"][ok])
    assert np.array_equal(d["flocal"][0][0, :, :, :, 0], ref)
    assert d["flocal"][0][..., s.npbins - 1, :].sum() == 0.0
    q = d["quick"]
    assert q[0] == len(a) and q[2] == a["weight"].sum() and d["pmax"] == a["p"].max()
    assert q[6] == a["dt"].min() and q[7] == a["dt"].max()
    pe, me = o.hist_edges(0)
    assert len(pe) == P.npp_global + 1 and rel_err(pe[0], P.pmin) < 1e-15 and rel_err(pe[-1], P.pmax) < 1e-14
    assert list(me) == [-1.0, 1.0]


def test_mover_ends_every_particle_on_the_frame_time_and_is_deterministic():
    """particle_mover: the roll-back + fixed-dt re-push lands every surviving particle exactly
    on t0 + dtf (particle_module.f90:1707-1826); the Philox stream makes reruns bit-identical
    and independent of the OpenMP schedule."""
    w, P, frames, ts = make_case("c1", grid=32, nptl=500)
    outs = []
    for _ in range(2):
        o = Oracle(P, w.nptl_max)
        rec, steps = run_intervals(o, frames, ts, nptl=500, particle_v0=w.particle_v0, num_fine_steps=2)
        outs.append((sort_by_key(o.download_particles()), steps, rec))
    a, b = outs[0][0], outs[1][0]
    assert outs[0][1] == outs[1][1] > 500 * 50
    assert_particles_identical(a, b, 'rerun')
    assert np.max(np.abs(a["t"] - ts[-1])) < 1e-12
    assert np.all(a["count_flag"] == 1)
    assert np.all((a["x"] >= P.xmin) & (a["x"] <= P.xmax))  # final BC pass with the un-extended box


def test_open_boundaries_leak_weight():
    """particle_boundary_condition (particle_module.f90:1984-2129) with pbc = 1: escapes are
    flagged -1..-4, removed, and their weight goes to `leak`."""
    w, P, frames, ts = make_case("c2", grid=32, nptl=600)
    o = Oracle(P, w.nptl_max)
    run_intervals(o, frames, ts, nptl=600, particle_v0=w.particle_v0, split_flag=0, inject_new_ptl=False,
                  dump_escaped_dist=False)
    c = o.counters()
    a = o.download_particles()
    assert c.leak > 0 and c.leak + c.leak_negp + a["weight"].sum() == 600.0
    assert np.all(a["count_flag"] == 1)


def test_empty_and_capacity_edges():
    w, P, frames, ts = make_case("c1", grid=16, nptl=8)
    o = Oracle(P, 12)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    assert o.particle_mover(0.0, 0.1) == 0
    d = o.diagnostics(True)
    assert d["fglobal"].sum() == 0 and d["quick"][0] == 0
    o.inject_uniform(0, 0.0, 1, 1.0, 0.0, 0.1, box_of(P), 6.2)
    assert len(o.download_particles()) == 0
    o.inject_uniform(20, 0.0, 1, 1.0, 0.0, 0.1, box_of(P), 6.2)  # overflow: slot nptl_max is overwritten
    c = o.counters()
    assert c.nptl_current == 12 and c.tag_max == 20
    assert o.download_particles()["tag_injected"][-1] == 19


# ---- 1-D (push_particle_1d, particle_module.f90:2993-3111) ---------------------------------------
def _np_step_1d(P, frames, before, u, t0, dtf):
    """Independent numpy restatement of one 1-D step written from the Fortran: 2-corner
    interpolation (particle_module.f90:649-652, mhd_data_parallel.f90:1776-1792), the 1-D kappa
    branch with mag_dependency = 0 (:2255-2292) and push_particle_1d."""
    def grads(f):  # d/dx of the 8 primaries, FP32 difference x FP64 0.5/dx -> FP32
        g = np.zeros_like(f)
        g[1:-1] = f[2:] - f[:-2]
        g[0] = (np.float32(-3) * f[0] + np.float32(4) * f[1]) - f[2]
        g[-1] = (np.float32(3) * f[-1] - np.float32(4) * f[-2]) + f[-3]
        return (g.astype(np.float64) * (0.5 / P.dx)).astype(np.float32)
    x, p, t = before["x"], before["p"], before["t"]
    px = (x - P.xmin) / P.dx
    ix = np.floor(px).astype(np.int64) + 1
    rx = px - ix + 1
    rt = (t - t0) / dtf
    F = []
    for f in frames:
        g = grads(f)
        c = ix + 1  # Fortran index -> storage index
        val = lambda a, v: a[c, v].astype(np.float64) * (1.0 - rx) + a[c + 1, v].astype(np.float64) * rx
        F.append(dict(vx=val(f, 0), rho=val(f, 3), bx=val(f, 4), by=val(f, 5), bz=val(f, 6), dvx=val(g, 0)))
    Fi = {k: F[0][k] * (1.0 - rt) + F[1][k] * rt for k in F[0]}
    knorm = (p / P.p0) ** P.pindex if P.momentum_dependency == 1 else np.ones_like(p)
    kpara = P.kpara0 * knorm
    kperp = kpara * P.kret
    if P.nlgc:  # particle_module.f90:2497-2509 with mag_dependency = 0 and no maps
        nperp = (p / P.p0) ** ((5.0 - P.gamma_turb) / 3.0) if P.momentum_dependency == 1 else np.ones_like(p)
        kperp = P.kpara0 * P.kperp_kpara * nperp * before["mu"] ** 2
    skpara, skperp = np.sqrt(2.0 * kpara), np.sqrt(2.0 * kperp)
    dx_dt = Fi["vx"] + kpara * 0.0
    dp_dt = -p * Fi["dvx"] / 3.0
    dpp = np.zeros_like(p)
    b = np.sqrt(Fi["bx"] ** 2 + Fi["by"] ** 2 + Fi["bz"] ** 2)
    if P.dpp_wave:   # calc_dpp_wave_scattering
        va = b / np.sqrt(Fi["rho"])
        dp_dt = dp_dt + ((8 * p / (27 * kpara)) if P.momentum_dependency == 1 else (4 * p / (9 * kpara))) * va ** 2
        dpp = dpp + (p * va) ** 2 / (9 * kpara)
    if P.dpp_shear:  # push_particle_1d's shear tensor (:3048-3053) + calc_dpp_flow_shear
        divv = Fi["dvx"]
        sxx, syy, szz = Fi["dvx"] - divv / 3, -divv / 3, -divv / 3
        if P.weak_scattering:
            bbs = (sxx * Fi["bx"] ** 2 + syy * Fi["by"] ** 2 + szz * Fi["bz"] ** 2) * (1.0 / b) * (1.0 / b)
            gsh = bbs ** 2 / 5
        else:
            gsh = 2 * (sxx ** 2 + syy ** 2 + szz ** 2) / 15
        on = gsh > 0
        dp_dt = np.where(on, dp_dt + (2 + P.pindex) * gsh * P.tau0 * 1.0 * p ** (P.pindex - 1) * P.p0 ** (2.0 - P.pindex), dp_dt)
        dpp = np.where(on, dpp + gsh * P.tau0 * 1.0 * p ** P.pindex * P.p0 ** (2.0 - P.pindex), dpp)
    s = np.where(skperp > 0, skperp, skpara)
    dt = np.minimum(np.minimum((0.5 * P.dx / skpara) ** 2, (s / dx_dt) ** 2),
                    float(np.float32(0.1)) * p / np.abs(dp_dt))
    dt = np.where((dx_dt != 0) & (dp_dt != 0), dt, P.dt_min_rel * dtf)
    dt = np.clip(dt, P.dt_min_rel * dtf, P.dt_max_rel * dtf)
    sdt = np.sqrt(dt)
    ran1 = (2.0 * u[:, 0] - 1.0) * np.sqrt(3.0)
    xn = x + (dx_dt * dt + ran1 * skpara * sdt)
    pn = p + (dp_dt * dt + (2.0 * u[:, 1] - 1.0) * np.sqrt(3.0) * np.sqrt(2 * dpp) * sdt)
    pn = np.maximum(pn, 0.25 * P.p0)  # particle_module.f90:3105-3109
    return xn, pn, t + dt, dt


@pytest.mark.parametrize("cli", [None, dict(dpp_wave=1, dpp_shear=1), dict(dpp_wave=1, dpp_shear=1, weak_scattering=0),
                                 dict(nlgc=1, kperp_kpara=0.05), dict(nlgc=1, kperp_kpara=0.05, dpp_wave=1, dpp_shear=1)])
def test_1d_step_matches_numpy_restatement(cli):
    w, P, frames, _ = make_case("s1", grid=256, nptl=300, cli=cli)
    assert P.ndim == 1 and frames[0].shape == (260, 8)
    P.rng_mode = RNG_TABLE
    o = Oracle(P, w.nptl_max)
    u = np.random.default_rng(11).uniform(0, 1, (300, 2, 4))
    o.set_rng_table(u)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_uniform(300, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    assert o.debug_push_n(0.0, w.dt_out, 1) == 300
    after = o.download_particles()
    x, p, t, dt = _np_step_1d(P, frames[:2], before, u[before["tag_injected"], 0], 0.0, w.dt_out)
    for name, ref in (("x", x), ("p", p), ("t", t), ("dt", dt)):
        err = np.abs(after[name] - ref) / np.maximum(np.abs(ref), 1.0 if name == "x" else 1e-300)
        assert err.max() < 4e-15, (name, err.max())
    assert np.array_equal(after["y"], before["y"]) and np.array_equal(after["z"], before["z"])


def test_1d_intervals_end_on_frame_time_and_leak_through_open_x():
    w, P, frames, ts = make_case("s1", grid=128, nptl=400, nframes=3)
    o = Oracle(P, w.nptl_max)
    res, steps = run_intervals(o, frames, ts, nptl=400, dist_flag=1, particle_v0=w.particle_v0)
    ptl = o.download_particles()
    c = o.counters()
    assert steps > 400 and len(ptl) == c.nptl_current
    assert np.all(ptl["t"] == ts[-1])
    # weight is conserved between the box and the open-x leak; y never moves in 1-D so no y escapes
    assert abs(ptl["weight"].sum() + c.leak + c.leak_negp - 800.0) < 1e-9  # 400 injected per interval
    assert res[-1]["quick"][0] == c.nptl_current


# ---- targeted injection (particle_module.f90:785-1468, mhd_data_parallel.f90:2211-2498) ----------
@pytest.mark.parametrize("mode", [1, 2, 4, 5])
def test_targeted_injection_counts_and_accepts_like_the_reference(mode):
    w, P, frames, _ = make_case("c3", grid=48, nptl=8)
    o = Oracle(P, 4000)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    fa = o.get_fields(0).reshape(P.ny + 4, P.nx + 4, 32)
    jz = np.abs(fa[..., 8 + 15] - fa[..., 8 + 13])                      # FP32, as get_ncells_large_jz
    absj = np.sqrt((fa[..., 8 + 17] - fa[..., 8 + 19]) ** 2 + (fa[..., 8 + 18] - fa[..., 8 + 14]) ** 2
                   + (fa[..., 8 + 13] - fa[..., 8 + 15]) ** 2)
    ndivv = -(fa[..., 8].astype(np.float64) + fa[..., 8 + 4].astype(np.float64))
    rho = fa[..., 3]
    box = box_of(P)
    box[0] += 0.2 * P.lx
    box[3] -= 0.1 * P.lx
    xs = P.xmin + P.dx * np.arange(P.nx)
    ys = P.ymin + P.dy * np.arange(P.ny)
    inbox = ((ys > box[1]) & (ys < box[4]))[:, None] & ((xs > box[0]) & (xs < box[3]))[None, :]
    phys = {1: jz[2:-2, 2:-2], 2: absj[2:-2, 2:-2], 5: rho[2:-2, 2:-2],
            4: ndivv[0:P.ny, 0:P.nx]}[mode]   # the divv counter reads two cells to the lower-left
    vmin = float(np.quantile(phys[inbox], 0.6))
    want_cells = int(np.count_nonzero(inbox & (phys > vmin)))
    ninj, ncells = o.inject_targeted(mode, 900, 0.0, 1, w.particle_v0, 0.0, 0.1, box, 6.2, False, vmin, 2 * want_cells)
    assert ncells == want_cells and ninj == 450           # int(900 * ncells / (2 ncells))
    ptl = o.download_particles()
    assert len(ptl) == 450 and np.array_equal(ptl["tag_injected"], np.arange(450))
    assert np.all((ptl["x"] >= box[0]) & (ptl["x"] <= box[3]))
    F = o.interp(ptl["x"], ptl["y"], ptl["z"], np.zeros(len(ptl)))
    got = {1: np.abs(F[:, 8 + 15] - F[:, 8 + 13]),
           2: np.sqrt((F[:, 8 + 17] - F[:, 8 + 19]) ** 2 + (F[:, 8 + 18] - F[:, 8 + 14]) ** 2
                      + (F[:, 8 + 13] - F[:, 8 + 15]) ** 2),
           4: -(F[:, 8] + F[:, 8 + 4]), 5: F[:, 3]}[mode]
    assert np.all(got >= vmin)                            # the loop runs while crit < vmin
    # the accepted positions favour the cells above the threshold: not uniform in the box
    assert (ptl["t"] >= 0.0).all() and (ptl["t"] <= 0.1).all() and np.all(ptl["weight"] == 1.0)


# ---- particle tracking (particle_module.f90:5825-5990, hooks at :434-440, 1697-1724, 5452-5473) ---
def _tracking_pair(sim_cls, nptl=600, nsel=25, strict=None):
    """First run -> tag table of the highest-energy particles -> tracking run (same seed)."""
    from stochastic_parker_b200.tracking import select_tags
    # dt_min_rel sizes particles_tracked: nsteps_tracking_max = ceiling(1/dt_min_rel/nsteps_interval) + 1
    w, P, frames, ts = make_case("c1", grid=48, nptl=nptl, nframes=4, conf=dict(dt_min_rel=1e-4))
    if strict is not None:
        P.strict_math = strict
    kw = dict(nptl=nptl, dist_flag=2, particle_v0=w.particle_v0, inject_new_ptl=False, split_flag=1,
              pmin_split=1.02, split_ratio=1.02, nsteps_interval=50)
    a = sim_cls(P, 8 * nptl)
    run_intervals(a, frames, ts, **kw)
    first = a.download_particles()
    assert first["split_times"].max() >= 2
    sel = np.argsort(first["p"])[-nsel:]
    tags = select_tags(first, sel)
    frames_rec = []
    b = sim_cls(P, 8 * nptl)
    run_intervals(b, frames, ts, track_tags=tags, on_tracked=lambda tf, rec: frames_rec.append(rec.copy()), **kw)
    return first, sel, tags, b.download_particles(), frames_rec


def test_tracking_replays_and_records_the_selected_particles():
    first, sel, tags, second, recs = _tracking_pair(Oracle)
    assert tags.shape[0] == 25 and tags.shape[1] == first["split_times"][sel].max() + 2
    # the tracking run moves with num_fine_steps = 1 like the first one and keys Philox by |tag|:
    # every particle of the first run exists again, same trajectory, with negated tags if tracked
    key = lambda q: (q["origin"], np.abs(q["tag_injected"]), np.abs(q["tag_splitted"]))
    oa, ob = np.lexsort(key(first)[::-1]), np.lexsort(key(second)[::-1])
    fa, fb = first[oa], second[ob]
    assert len(fa) == len(fb)
    for f in ("x", "y", "p", "t", "weight"):
        assert np.array_equal(fa[f], fb[f]), f
    tracked_now = fb["tag_splitted"] < 0
    want = np.zeros(len(fa), dtype=bool)
    want[np.isin(np.arange(len(first)), sel)[oa]] = True
    assert np.array_equal(tracked_now, want)          # exactly the selected leaves are still tagged
    # records: one array per MHD interval, (nptl_tracking, nsteps_tracking_max)
    assert len(recs) == 3 and recs[0].shape[0] == 25
    for rec in recs:
        for row in rec:
            used = row["tag_splitted"] < 0
            t = row["t"][used]
            assert np.all(np.diff(t) > 0)             # a trajectory: time increases along the row
            assert np.all(row["nsteps_pushed"][used] == 0)   # sampled every nsteps_interval pushes
    # the last sample of every tracked row belongs to the selected particle's ancestry
    last = recs[-1]
    assert np.count_nonzero((last["tag_splitted"] < 0).any(axis=1)) == 25
    assert np.all(np.abs(last["tag_injected"][last["tag_splitted"] < 0]) ==
                  np.repeat(tags[:, 1], (last["tag_splitted"] < 0).sum(axis=1)))


# ---- focused transport, 2-D (calc_duu + push_particle_2d_ft, particle_module.f90:3116-3155, 3626-3977)
def _np_step_2d_ft(P, F, ptl, u, dt_min, dt_max, dt_fixed=None, aux=None):
    """Independent numpy restatement of one 2-D focused-transport step written from the Fortran."""
    f = lambda k: F[:, k - 1]
    g = lambda k: F[:, 8 + k - 1]
    p, v, mu = ptl["p"], ptl["v"], ptl["mu"]
    vx, vy, vz, rho, bx, by, bz = f(1), f(2), f(3), f(4), f(5), f(6), f(7)
    b = np.sqrt(bx**2 + by**2 + bz**2)
    ib = 1.0 / b
    ib2, ib3 = ib * ib, ib * ib * ib
    dbx_dx, dbx_dy, dby_dx, dby_dy = g(13), g(14), g(16), g(17)
    dbz_dx, dbz_dy, db_dx, db_dy = g(19), g(20), g(22), g(23)
    # kappa with kpp = -kperp (particle_module.f90:2372-2376)
    knp = b ** (P.gamma_turb - 2.0) if P.mag_dependency == 1 else np.ones_like(b)
    if P.deltab_flag:
        knp = knp / aux[:, 0]
    if P.correlation_flag:
        knp = knp * aux[:, 8] ** (P.gamma_turb - 1.0)
    knorm = knp * (p / P.p0) ** P.pindex if P.momentum_dependency == 1 else knp
    kpara = P.kpara0 * knorm
    kperp = kpara * P.kret
    skpara, skperp = np.sqrt(2 * kpara), np.sqrt(2 * kperp)
    dkdx = db_dx * ib * (P.gamma_turb - 2.0) if P.mag_dependency == 1 else 0.0 * b
    dkdy = db_dy * ib * (P.gamma_turb - 2.0) if P.mag_dependency == 1 else 0.0 * b
    if P.deltab_flag:
        dkdx, dkdy = dkdx - aux[:, 1] / aux[:, 0], dkdy - aux[:, 2] / aux[:, 0]
    if P.correlation_flag:
        dkdx = dkdx + (P.gamma_turb - 1.0) * aux[:, 9] / aux[:, 8]
        dkdy = dkdy + (P.gamma_turb - 1.0) * aux[:, 10] / aux[:, 8]
    kpp = -kperp
    dkxx_dx = kperp * dkdx + kpp * dkdx * bx**2 * ib2 + 2.0 * kpp * bx * (dbx_dx * b - bx * db_dx) * ib3
    dkyy_dy = kperp * dkdy + kpp * dkdy * by**2 * ib2 + 2.0 * kpp * by * (dby_dy * b - by * db_dy) * ib3
    dkxy_dx = kpp * dkdx * bx * by * ib2 + kpp * ((dbx_dx * by + bx * dby_dx) * ib2 - 2.0 * bx * by * db_dx * ib3)
    dkxy_dy = kpp * dkdy * bx * by * ib2 + kpp * ((dbx_dy * by + bx * dby_dy) * ib2 - 2.0 * bx * by * db_dy * ib3)
    if P.nlgc:   # calc_spatial_diffusion_coefficients_nlgc with focused_transport = .true.
        kt = np_step.kappa_tensor(P, F, p, mu, "2d", aux, focused=True)
        kpara, kperp, skpara, skperp = kt["kpara"], kt["kperp"], kt["skpara"], kt["skperp"]
        dkxx_dx, dkyy_dy, dkxy_dx, dkxy_dy = kt["dkxx_dx"], kt["dkyy_dy"], kt["dkxy_dx"], kt["dkxy_dy"]
    vdp = float(np.float32(1.0) / np.float32(P.pcharge)) / np.sqrt((P.drift1 * P.p0 / p) ** 2 + (P.drift2 * P.p0**2 / p**2) ** 2)
    mu2 = mu**2
    muf1, muf2 = 0.5 * (1.0 - mu2), 0.5 * (3.0 * mu2 - 1.0)
    kx, ky, kz = bx * dbx_dx + by * dbx_dy, bx * dby_dx + by * dby_dy, bx * dbz_dx + by * dbz_dy
    bdc = bx * dbz_dy - by * dbz_dx + bz * (dby_dx - dbx_dy)
    vdx = vdp * (muf1 * (-bz * db_dy) * ib2 + mu2 * (by * kz - bz * ky) * ib3 + muf1 * bx * bdc * ib3)
    vdy = vdp * (muf1 * (bz * db_dx) * ib2 + mu2 * (bz * kx - bx * kz) * ib3 + muf1 * by * bdc * ib3)
    vb = v * mu * ib
    dvx_dx, dvx_dy, dvy_dx, dvy_dy, dvz_dx, dvz_dy = g(1), g(2), g(4), g(5), g(7), g(8)
    dx_dt = vx + vdx + vb * bx + dkxx_dx + dkxy_dy
    dy_dt = vy + vdy + vb * by + dkxy_dx + dkyy_dy
    divv = dvx_dx + dvy_dy
    bbg = (bx * (bx * dvx_dx + by * dvx_dy) + by * (bx * dvy_dx + by * dvy_dy) + bz * (bx * dvz_dx + by * dvz_dy)) * ib2
    bvg = (bx * (vx * dvx_dx + vy * dvx_dy) + by * (vx * dvy_dx + vy * dvy_dy) + bz * (vx * dvz_dx + vy * dvz_dy)) * ib
    dp_dt = p * -(muf1 * divv + muf2 * bbg + mu * bvg / v)
    div_bn = -(bx * db_dx + by * db_dy) * ib2
    dmu_dt = (v * div_bn + mu * divv - 3 * mu * bbg - 2 * bvg / v) * (1 - mu2) * 0.5
    dtmp = np.abs(mu) ** (P.gamma_turb - 1) + float(np.float32(0.2))
    norm = np.ones_like(b)
    if P.mag_dependency == 1:
        norm = norm * b ** (2.0 - P.gamma_turb)
    if P.deltab_flag:            # calc_duu, particle_module.f90:3143-3148
        norm = norm * aux[:, 0]
    if P.correlation_flag:
        norm = norm * aux[:, 8] ** (1.0 - P.gamma_turb)
    if P.momentum_dependency == 1:
        norm = norm * (p / P.p0) ** (P.gamma_turb - 1)
    duu = P.duu0 * (1 - mu2) * dtmp * norm
    duu_du = P.duu0 * (-2 * mu * dtmp + np.sign(mu) * (1 - mu2) * np.abs(mu) ** (P.gamma_turb - 2)) * norm
    dmu_dt = dmu_dt + duu_du
    s = np.where(skperp > 0, skperp, skpara)
    cands = [(0.5 * P.dx / s) ** 2, (0.5 * P.dy / s) ** 2, (s / dx_dt) ** 2, (s / dy_dt) ** 2,
             float(np.float32(0.1)) * p / np.abs(dp_dt), float(np.float32(0.1)) / np.abs(dmu_dt), 2.0 * duu / dmu_dt**2]
    dt = np.clip(np.minimum.reduce(cands), dt_min, dt_max)
    if dt_fixed is not None:
        dt = dt_fixed
    sdt, s3 = np.sqrt(dt), np.sqrt(3.0)
    r1, r2, rp, rm = [(2.0 * u[:, k] - 1.0) * s3 for k in range(4)]
    bxn, byn, bzn = bx * ib, by * ib, bz * ib
    ibxyn = 1.0 / np.sqrt(bxn**2 + byn**2)
    x = ptl["x"] + dx_dt * dt + skperp * ibxyn * sdt * (-bxn * bzn * r1 - by * r2)
    y = ptl["y"] + dy_dt * dt + skperp * ibxyn * sdt * (-byn * bzn * r1 + bx * r2)
    dp = dp_dt * dt
    mun = np.clip(mu + dmu_dt * dt + rm * np.sqrt(2 * duu) * sdt, -float(np.float32(0.99)), float(np.float32(0.99)))
    pn = p + dp
    vn = v + v * dp / p
    low = pn < 0.25 * P.p0
    vn = np.where(low, v * 0.25 * P.p0 / pn, vn)  # particle_module.f90:3969 divides by the NEW (too small) p
    pn = np.where(low, 0.25 * P.p0, pn)
    return x, y, pn, vn, mun, ptl["t"] + dt, dt


def test_focused_transport_2d_step_matches_numpy_restatement():
    w, P, frames, _ = make_case("c1", grid=48, nptl=400, cli=dict(focused_transport=1, duu_init=5.0))
    assert P.focused_transport == 1 and P.duu0 == 5.0 and P.nmu_global > 1
    P.rng_mode = RNG_TABLE
    o = Oracle(P, w.nptl_max)
    u = np.random.default_rng(9).uniform(0, 1, (400, 2, 4))
    o.set_rng_table(u)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_uniform(400, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    assert o.debug_push_n(0.0, w.dt_out, 1) == 400
    after = o.download_particles()
    F = o.interp(before["x"], before["y"], before["z"], (before["t"] - 0.0) / w.dt_out)
    x, y, p, v, mu, t, dt = _np_step_2d_ft(P, F, before, u[before["tag_injected"], 0],
                                           P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out)
    for name, ref in (("x", x), ("y", y), ("p", p), ("v", v), ("mu", mu), ("t", t), ("dt", dt)):
        scale = np.maximum(np.abs(ref), 1.0 if name in "xy" else 1e-300)
        err = np.abs(after[name] - ref) / scale
        assert err.max() < 1e-13, (name, err.max())
    assert np.any(after["mu"] != before["mu"]) and np.any(after["v"] != before["v"])


def test_focused_transport_2d_intervals_fill_the_pitch_angle_bins():
    w, P, frames, ts = make_case("c1", grid=48, nptl=1500, nframes=3, cli=dict(focused_transport=1, duu_init=20.0),
                                 conf=dict(dt_min_rel=1e-4))
    o = Oracle(P, w.nptl_max)
    res, steps = run_intervals(o, frames, ts, nptl=1500, dist_flag=1, particle_v0=w.particle_v0, inject_new_ptl=False)
    ptl = o.download_particles()
    assert steps > 1500 and np.all(np.abs(ptl["mu"]) <= np.float64(np.float32(0.99)))
    assert np.all(ptl["t"] == ts[-1])
    fg = res[-1]["fglobal"]                       # (npp, nmu)
    assert fg.shape[1] == P.nmu_global and np.count_nonzero(fg.sum(axis=0)) > P.nmu_global // 2
    assert abs(fg.sum() - ptl["weight"][(ptl["p"] > P.pmin) & (ptl["p"] <= P.pmax)].sum()) < 1e-9


# ---- shock injection (mhd_data_parallel.f90:1988-2045, particle_module.f90:542-633) --------------
def test_shock_injection_lands_on_the_later_frames_compression_peak():
    w, P, frames, ts = make_case("c3", grid=64, nptl=8)
    o = Oracle(P, 4000)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_at_shock(1500, 1e-5, 0, w.particle_v0, ts[0], 6.2)
    ptl = o.download_particles()
    assert len(ptl) == 1500 and np.all(ptl["t"] == ts[0]) and np.all(ptl["dt"] == 1e-5)
    # locate_shock_xpos: 1-based index of max |dvx/dx| along the ghosted x extent, per row of farray2
    fa2 = o.get_fields(1).reshape(P.ny + 4, P.nx + 4, 32)
    sx = np.argmax(np.abs(fa2[..., 8]), axis=1) + 1
    iy = np.floor(ptl["y"] / P.dy).astype(int)
    ry = ptl["y"] / P.dy - iy
    want = ((sx[iy] * (1 - ry) * (1 - ry) + sx[iy + 1] * ry * (1 - ry)) + 2) * (P.xmax - P.xmin) / (P.nx + 4)
    assert np.max(np.abs(ptl["x"] - want)) < 1e-12 * P.xmax     # rz = ry: the weights are (1-ry)^2, ry(1-ry)
    # Maxwellian envelope of the shock injector: f ~ p^2 exp(-p^2/2) in units of p0 peaks at sqrt(2) p0
    h, edges = np.histogram(ptl["p"] / P.p0, bins=20, range=(0.1, 5.0))
    assert 1.0 < 0.5 * (edges[np.argmax(h)] + edges[np.argmax(h) + 1]) < 2.0
    assert np.all(np.abs(ptl["mu"]) <= np.float64(np.float32(0.99)))


# ---- MT19937 stand-in stream (oracle only) ---------------------------------------------------------
def test_mt19937_known_answers():
    """mt19937ar.c's published test output for init_by_array({0x123, 0x234, 0x345, 0x456}) -- the very
    seed array random_number_generator.f90:16 hands to mt_stream -- starts 1067595299 955945823
    477289528 4107218783 4228976476."""
    import ctypes as C
    w, P, _, _ = make_case("c1", grid=16, nptl=8)
    o = Oracle(P, 16)
    o.lib.orc_mt19937_next.restype = C.c_uint32
    got = [o.lib.orc_mt19937_next(o.h) for _ in range(5)]
    assert got == [1067595299, 955945823, 477289528, 4107218783, 4228976476]


# ---- turbulence maps: deltab_flag / correlation_flag (particle_module.f90:2246-2254, 2364-2371) ----
def test_deltab_and_correlation_step_matches_numpy_restatement():
    from stochastic_parker_b200 import mhd
    w, P, frames, _ = make_case("c1", grid=48, nptl=400)
    P.deltab_flag = 1
    P.correlation_flag = 1
    P.rng_mode = RNG_TABLE
    o = Oracle(P, w.nptl_max)
    u = np.random.default_rng(21).uniform(0, 1, (400, 2, 4))
    o.set_rng_table(u)
    maps = [mhd.make_turbulence_maps(P.nx, P.ny, 1, f) for f in (0, 1)]
    for slot in (0, 1):
        o.upload_fields(slot, frames[slot])
        o.upload_turbulence(0, slot, maps[slot][0], maps[slot][1])
        o.upload_turbulence(1, slot, maps[slot][2], maps[slot][3])
    o.inject_uniform(400, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    assert o.debug_push_n(0.0, w.dt_out, 1) == 400
    after = o.download_particles()
    rt = (before["t"] - 0.0) / w.dt_out
    fa1 = np_step.gradients32(frames[0], P.dx, P.dy)
    fa2 = np_step.gradients32(frames[1], P.dx, P.dy)
    F = np_step.interp32(fa1, fa2, P, before["x"], before["y"], rt)
    g = [[np_step.turbulence_grad(m, P.dx, P.dy) for m in maps[s]] for s in (0, 1)]
    aux = np_step.interp_aux(g[0], g[1], P, before["x"], before["y"], rt)
    qdrift = float(np.float32(1.0) / np.float32(3 * P.pcharge))
    x, y, p, t, dt = np_step.push_2d(P, F, before["p"], P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out,
                                    u[before["tag_injected"], 0], before["x"], before["y"], before["t"], qdrift, aux=aux)
    for name, ref in (("x", x), ("y", y), ("p", p), ("t", t), ("dt", dt)):
        scale = max(1.0, np.abs(ref).max()) if name in "xy" else np.abs(ref)
        assert (np.abs(after[name] - ref) / scale).max() < 5e-15, name
    # the maps matter: the same step without them lands elsewhere
    x0 = np_step.push_2d(P.copy(), F, before["p"], P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out,
                         u[before["tag_injected"], 0], before["x"], before["y"], before["t"], qdrift,
                         aux=np.ones_like(aux) * np.array([1, 0, 0, 0] * 4))[0]
    assert np.max(np.abs(x0 - x)) > 1e-6


# ---- focused transport, 2-D + 3rd dimension and 3-D (PM:4267-4623, 4930-5320) ----------------------
def _np_step_ft_3d_like(P, F, ptl, u, u5, dt_min, dt_max):
    """Independent numpy restatement of one push_particle_2d_include_3rd_ft / push_particle_3d_ft
    step (Cartesian), with the kappa branch it is called with (PM:2294-2351 / 2392-2450, kpp = -kperp)."""
    full3d = P.ndim == 3
    f = lambda k: F[:, k - 1]
    g = lambda k: F[:, 8 + k - 1]
    Z = np.zeros(len(F))
    p, v, mu = ptl["p"], ptl["v"], ptl["mu"]
    vx, vy, vz, rho, bx, by, bz = f(1), f(2), f(3), f(4), f(5), f(6), f(7)
    b = np.sqrt(bx**2 + by**2 + bz**2)
    ib = 1.0 / b
    ib2, ib3 = ib * ib, ib * ib * ib
    dbx_dx, dbx_dy, dby_dx, dby_dy, dbz_dx, dbz_dy, db_dx, db_dy = g(13), g(14), g(16), g(17), g(19), g(20), g(22), g(23)
    dbx_dz, dby_dz, dbz_dz, db_dz = (g(15), g(18), g(21), g(24)) if full3d else (Z, Z, Z, Z)
    knp = b ** (P.gamma_turb - 2.0) if P.mag_dependency == 1 else np.ones_like(b)
    knorm = knp * (p / P.p0) ** P.pindex if P.momentum_dependency == 1 else knp
    kpara = P.kpara0 * knorm
    kperp = kpara * P.kret
    skpara, skperp = np.sqrt(2 * kpara), np.sqrt(2 * kperp)
    gm2 = P.gamma_turb - 2.0
    if P.mag_dependency == 1:
        # 3-D: no 1/B (PM:2405-2409); 2-D + 3rd: with 1/B and no z term (PM:2310-2313)
        dkdx, dkdy, dkdz = (db_dx * gm2, db_dy * gm2, db_dz * gm2) if full3d else (db_dx * ib * gm2, db_dy * ib * gm2, Z)
    else:
        dkdx = dkdy = dkdz = Z
    kpp = -kperp
    dk = lambda lead, d, bi, bj, dbi, dbj, dbm: lead + kpp * d * bi * bj * ib2 + kpp * ((dbi * bj + bi * dbj) * ib2 - 2.0 * bi * bj * dbm * ib3)
    dkxx_dx = kperp * dkdx + kpp * dkdx * bx**2 * ib2 + 2.0 * kpp * bx * (dbx_dx * b - bx * db_dx) * ib3
    dkyy_dy = kperp * dkdy + kpp * dkdy * by**2 * ib2 + 2.0 * kpp * by * (dby_dy * b - by * db_dy) * ib3
    dkzz_dz = kperp * dkdz + kpp * dkdz * bz**2 * ib2 + 2.0 * kpp * bz * (dbz_dz * b - bz * db_dz) * ib3
    dkxy_dx = dk(0.0, dkdx, bx, by, dbx_dx, dby_dx, db_dx)
    dkxy_dy = dk(0.0, dkdy, bx, by, dbx_dy, dby_dy, db_dy)
    dkxz_dx = dk(0.0, dkdx, bx, bz, dbx_dx, dbz_dx, db_dx)
    dkxz_dz = dk(0.0, dkdz, bx, bz, dbx_dz, dbz_dz, db_dz)
    dkyz_dy = dk(0.0, dkdy, by, bz, dby_dy, dbz_dy, db_dy)
    dkyz_dz = dk(0.0, dkdz, by, bz, dby_dz, dbz_dz, db_dz)
    if P.nlgc:   # calc_spatial_diffusion_coefficients_nlgc with focused_transport = .true.
        kt = np_step.kappa_tensor(P, F, p, mu, "3d" if full3d else "2d3", None, focused=True)
        kpara, kperp, skpara, skperp = kt["kpara"], kt["kperp"], kt["skpara"], kt["skperp"]
        dkxx_dx, dkyy_dy, dkzz_dz = kt["dkxx_dx"], kt["dkyy_dy"], kt["dkzz_dz"]
        dkxy_dx, dkxy_dy, dkxz_dx, dkxz_dz = kt["dkxy_dx"], kt["dkxy_dy"], kt["dkxz_dx"], kt["dkxz_dz"]
        dkyz_dy, dkyz_dz = kt["dkyz_dy"], kt["dkyz_dz"]
    vdp = float(np.float32(1.0) / np.float32(P.pcharge)) / np.sqrt((P.drift1 * P.p0 / p) ** 2 + (P.drift2 * P.p0**2 / p**2) ** 2)
    mu2 = mu**2
    muf1, muf2 = 0.5 * (1.0 - mu2), 0.5 * (3.0 * mu2 - 1.0)
    kx = bx * dbx_dx + by * dbx_dy + bz * dbx_dz
    ky = bx * dby_dx + by * dby_dy + bz * dby_dz
    kz = bx * dbz_dx + by * dbz_dy + bz * dbz_dz
    bdc = bx * (dbz_dy - dby_dz) + by * (dbx_dz - dbz_dx) + bz * (dby_dx - dbx_dy)
    vdx = vdp * (muf1 * (by * db_dz - bz * db_dy) * ib2 + mu2 * (by * kz - bz * ky) * ib3 + muf1 * bx * bdc * ib3)
    vdy = vdp * (muf1 * (bz * db_dx - bx * db_dz) * ib2 + mu2 * (bz * kx - bx * kz) * ib3 + muf1 * by * bdc * ib3)
    vdz = vdp * (muf1 * (bx * db_dy - by * db_dx) * ib2 + mu2 * (bx * ky - by * kx) * ib3 + muf1 * bz * bdc * ib3)
    vb = v * mu * ib
    dvx_dx, dvx_dy, dvy_dx, dvy_dy, dvz_dx, dvz_dy = g(1), g(2), g(4), g(5), g(7), g(8)
    dvx_dz, dvy_dz, dvz_dz = (g(3), g(6), g(9)) if full3d else (Z, Z, Z)
    dx_dt = vx + vb * bx + vdx + dkxx_dx + dkxy_dy + dkxz_dz
    dy_dt = vy + vb * by + vdy + dkxy_dx + dkyy_dy + dkyz_dz
    dz_dt = vz + vb * bz + vdz + dkxz_dx + dkyz_dy + dkzz_dz
    divv = dvx_dx + dvy_dy + dvz_dz
    bbg = (bx * (bx * dvx_dx + by * dvx_dy + bz * dvx_dz) + by * (bx * dvy_dx + by * dvy_dy + bz * dvy_dz)
           + bz * (bx * dvz_dx + by * dvz_dy + bz * dvz_dz)) * ib2
    bvg = (bx * (vx * dvx_dx + vy * dvx_dy + vz * dvx_dz) + by * (vx * dvy_dx + vy * dvy_dy + vz * dvy_dz)
           + bz * (vx * dvz_dx + vy * dvz_dy + vz * dvz_dz)) * ib
    dp_dt = p * -(muf1 * divv + muf2 * bbg + mu * bvg / v)
    div_bn = -(bx * db_dx + by * db_dy + bz * db_dz) * ib2
    dmu_dt = (v * div_bn + mu * divv - 3 * mu * bbg - 2 * bvg / v) * (1 - mu2) * 0.5
    dtmp = np.abs(mu) ** (P.gamma_turb - 1) + float(np.float32(0.2))
    norm = np.ones_like(b)
    if P.mag_dependency == 1:
        norm = norm * b ** (2.0 - P.gamma_turb)
    if P.momentum_dependency == 1:
        norm = norm * (p / P.p0) ** (P.gamma_turb - 1)
    duu = P.duu0 * (1 - mu2) * dtmp * norm
    dmu_dt = dmu_dt + P.duu0 * (-2 * mu * dtmp + np.sign(mu) * (1 - mu2) * np.abs(mu) ** (P.gamma_turb - 2)) * norm
    s = np.where(skperp > 0, skperp, skpara)
    cands = [(0.5 * P.dx / s) ** 2, (0.5 * P.dy / s) ** 2, (s / dx_dt) ** 2, (s / dy_dt) ** 2,
             float(np.float32(0.1)) * p / np.abs(dp_dt), float(np.float32(0.1)) / np.abs(dmu_dt), 2.0 * duu / dmu_dt**2]
    if full3d:
        cands += [(0.5 * P.dz / s) ** 2, (s / dz_dt) ** 2]
    dt = np.clip(np.minimum.reduce(cands), dt_min, dt_max)
    sdt, s3 = np.sqrt(dt), np.sqrt(3.0)
    r1, r2, rp = [(2.0 * u[:, k] - 1.0) * s3 for k in (0, 1, 3)]
    rm = (2.0 * u5 - 1.0) * s3
    bxn, byn, bzn = bx * ib, by * ib, bz * ib
    bxyn = np.sqrt(bxn**2 + byn**2)
    x = ptl["x"] + dx_dt * dt + (-bxn * bzn * skperp / bxyn * r1 - byn * skperp / bxyn * r2) * sdt
    y = ptl["y"] + dy_dt * dt + (-byn * bzn * skperp / bxyn * r1 + bxn * skperp / bxyn * r2) * sdt
    z = ptl["z"] + dz_dt * dt + bxyn * skperp * r1 * sdt
    mun = np.clip(mu + dmu_dt * dt + rm * np.sqrt(2 * duu) * sdt, -float(np.float32(0.99)), float(np.float32(0.99)))
    dp = dp_dt * dt
    pn = p + dp
    vn = v + v * dp / p
    low = pn < 0.25 * P.p0
    vn = np.where(low, v * 0.25 * P.p0 / pn, vn)
    pn = np.where(low, 0.25 * P.p0, pn)
    return x, y, z, pn, vn, mun, ptl["t"] + dt, dt


@pytest.mark.parametrize("key,grid,cli", [("c1", 48, dict(include_3rd_dim=1)), ("c5", 24, {}),
                                          ("c1", 48, dict(include_3rd_dim=1, nlgc=1, kperp_kpara=0.05)),
                                          ("c5", 24, dict(nlgc=1, kperp_kpara=0.05))])
def test_focused_transport_3d_like_step_matches_numpy_restatement(key, grid, cli):
    w, P, frames, _ = make_case(key, grid=grid, nptl=300, cli=dict(cli, focused_transport=1, duu_init=5.0),
                                conf=dict(r1=4, r2=8, r3=12) if key == "c5" else None)
    o = Oracle(P, w.nptl_max)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_uniform(300, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    assert o.debug_push_n(0.0, w.dt_out, 1) == 300
    after = o.download_particles()
    F = o.interp(before["x"], before["y"], before["z"], (before["t"] - 0.0) / w.dt_out)
    # the uniforms of step 0: block A = Philox((0, 0, tag, 1)), fifth = word 0 of Philox((0, 0x80000000, tag, 1))
    key0, key1 = P.seed & 0xFFFFFFFF, (P.seed >> 32) & 0xFFFFFFFF
    uA = np.array([[w_ / 4294967295.0 for w_ in philox4x32_10((0, 0, int(t), 1), (key0, key1))] for t in before["tag_injected"]])
    u5 = np.array([philox4x32_10((0, 0x80000000, int(t), 1), (key0, key1))[0] / 4294967295.0 for t in before["tag_injected"]])
    x, y, z, p, v, mu, t, dt = _np_step_ft_3d_like(P, F, before, uA, u5, P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out)
    for name, ref in (("x", x), ("y", y), ("z", z), ("p", p), ("v", v), ("mu", mu), ("t", t), ("dt", dt)):
        scale = np.maximum(np.abs(ref), 1.0 if name in "xyz" else 1e-300)
        err = np.abs(after[name] - ref) / scale
        assert err.max() < 1e-12, (name, err.max())
    assert np.any(after["z"] != before["z"])


# ---- acceleration surfaces (acc_region_surface.f90) -------------------------------------------
def _np_surface_height(P, sf0, sf1, norm, x, y, z, rt):
    """Bilinear interpolation of a surface in its own plane, written from the definition (not from the
    eight collapsed trilinear weights the oracle uses), then the linear blend in time."""
    axis = abs(norm) - 1
    c = [(x - P.xmin) / P.dx, (y - P.ymin) / P.dy, (z - P.zmin) / P.dz]
    pu, pv = [c[k] for k in range(3) if k != axis]
    iu, iv = np.floor(pu).astype(int), np.floor(pv).astype(int)
    ru, rv = pu - iu, pv - iv
    a, b = iu + 2, iv + 2  # Fortran index floor + 1, array lower bound -1

    def at(sf):
        return (sf[b, a] * (1 - ru) * (1 - rv) + sf[b, a + 1] * ru * (1 - rv)
                + sf[b + 1, a] * (1 - ru) * rv + sf[b + 1, a + 1] * ru * rv)
    return at(sf0) * (1 - rt) + at(sf1) * rt


@pytest.mark.parametrize("norm1,norm2,two,inter", [("+z", "-y", 0, 0), ("-y", "+y", 1, 1), ("+x", "-z", 1, 0),
                                                   ("-z", "+z", 1, 1)])
@pytest.mark.parametrize("ft", [0, 1])
def test_acceleration_surfaces_gate_the_momentum_update(norm1, norm2, two, inter, ft):
    """push_particle_3d / _3d_ft (particle_module.f90:4887-4892, 5297-5303): a particle gains or loses
    momentum only if its NEW position is on the right side of the surface heights interpolated at its OLD
    position.  Checked against an ungated run of the same particles (same uniforms, same displacement)."""
    from stochastic_parker_b200 import mhd
    cli = dict(acc_by_surface=1, surface_norm1=norm1, surface_norm2=norm2, surface2_existed=two,
               is_intersection=inter)
    if ft:
        cli.update(focused_transport=1, duu_init=5.0)
    conf = dict(acc_region_flag=1, r1=4, r2=8, r3=12)
    w, P, frames, _ = make_case("c5", grid=24, nptl=600, cli=cli, conf=conf)
    _, P0, _, _ = make_case("c5", grid=24, nptl=600, cli=dict(cli, acc_by_surface=0), conf=conf)
    gated, free = Oracle(P, w.nptl_max), Oracle(P0, w.nptl_max)
    surf = [[mhd.make_acc_surface(P, k, f) for f in (0, 1)] for k in range(2 if two else 1)]
    for o in (gated, free):
        o.upload_fields(0, frames[0])
        o.upload_fields(1, frames[1])
        o.inject_uniform(600, 0.0, 0, w.particle_v0, 0.03, w.dt_out, box_of(P), w.power_index)
    for k, s in enumerate(surf):
        gated.upload_acc_surface(k, 0, s[0])
        gated.upload_acc_surface(k, 1, s[1])
    before = gated.download_particles()
    assert gated.debug_push_n(0.0, w.dt_out, 1) == 600 and free.debug_push_n(0.0, w.dt_out, 1) == 600
    a, b = gated.download_particles(), free.download_particles()
    for f in ("x", "y", "z", "t", "dt", "mu"):
        assert np.array_equal(a[f], b[f]), f
    rt = (before["t"] - 0.0) / w.dt_out
    new = dict(x=a["x"], y=a["y"], z=a["z"])
    inside, margin = None, np.full(len(a), np.inf)
    for k, s in enumerate(surf):
        norm = P.surface_norm2 if k else P.surface_norm1
        h = _np_surface_height(P, s[0], s[1], norm, before["x"], before["y"], before["z"], rt)
        ph = new["xyz"[abs(norm) - 1]]
        side = (ph > h) if norm > 0 else (ph < h)
        margin = np.minimum(margin, np.abs(ph - h))
        inside = side if inside is None else ((inside & side) if inter else (inside | side))
    sure = margin > 1e-9
    assert 50 < inside[sure].sum() < sure.sum() - 50, "the surfaces must split the population"
    want = np.maximum(np.where(inside, b["p"], before["p"]), 0.25 * P.p0)  # + the momentum floor, :3601-3605
    assert np.array_equal(a["p"][sure], want[sure]), np.flatnonzero(a["p"] != want)
    if ft:
        off_floor = sure & (want > 0.25 * P.p0)  # the floor rescales v too, :5312-5320
        assert np.array_equal(a["v"][off_floor], np.where(inside, b["v"], before["v"])[off_floor])
    assert np.any(b["p"] != before["p"])


def test_acceleration_surfaces_follow_the_frames():
    """copy_acc_surface (acc_region_surface.f90:390-396) + the per-interval read: a run_intervals run with
    surfaces differs from one without, and the surface used in interval 2 is frame 1's blended to frame 2's."""
    from stochastic_parker_b200 import mhd
    cli = dict(acc_by_surface=1, surface_norm1="+z")
    conf = dict(acc_region_flag=1, r1=4, r2=8, r3=12)
    w, P, frames, ts = make_case("c5", grid=24, nptl=300, cli=cli, conf=conf, nframes=3)
    runs = []
    for shift in (0, 1):
        o = Oracle(P, w.nptl_max)
        run_intervals(o, frames, ts, nptl=300, dist_flag=0, split_flag=0, local_dist=False,
                      surfaces=lambda k, f: mhd.make_acc_surface(P, k, f + shift * (f == 2)))
        runs.append(sort_by_key(o.download_particles()))
    # the two runs share frames 0 and 1, so they agree in interval 1 and can only differ through frame 2's surface
    assert len(runs[0]) == len(runs[1]) == 600
    assert np.any(runs[0]["p"] != runs[1]["p"])
    with pytest.raises(ValueError):
        run_intervals(Oracle(P, w.nptl_max), frames, ts, nptl=10)


# ---- the other Parker pushers against numpy restatements written from the Fortran ---------------
def _table_step(key, grid, conf=None, cli=None, n=400, tweak=None, seed=11):
    """One push of n particles with tabulated uniforms; returns (P, w, frames, before, after, u)."""
    w, P, frames, _ = make_case(key, grid=grid, nptl=n, conf=conf, cli=cli)
    if tweak:
        tweak(P)
    P.rng_mode = RNG_TABLE
    o = Oracle(P, w.nptl_max)
    u = np.random.default_rng(seed).uniform(0, 1, (n, 2, 4))
    o.set_rng_table(u)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_uniform(n, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    assert o.debug_push_n(0.0, w.dt_out, 1) == n
    return P, w, frames, before, o.download_particles(), u[before["tag_injected"], 0], o


def _check(after, want, names, tol=4e-15):
    for name, ref in zip(names, want):
        scale = np.maximum(np.abs(ref), 1.0) if name in "xyz" else np.abs(ref)
        err = np.abs(after[name] - ref) / scale
        assert err.max() < tol, (name, err.max(), int(err.argmax()))


def test_3d_gradients_and_gather_match_numpy_restatement():
    w, P, frames, _ = make_case("c5", grid=24, nptl=300, conf=dict(r1=4, r2=8, r3=12))
    o = Oracle(P, w.nptl_max)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    fa1 = np_step.gradients32_3d(frames[0], P.dx, P.dy, P.dz)
    fa2 = np_step.gradients32_3d(frames[1], P.dx, P.dy, P.dz)
    assert np.array_equal(o.get_fields(0), fa1) and np.array_equal(o.get_fields(1), fa2)
    rng = np.random.default_rng(2)
    x, y, z = (rng.uniform(lo, hi, 500) for lo, hi in ((P.xmin, P.xmax), (P.ymin, P.ymax), (P.zmin, P.zmax)))
    rt = rng.uniform(0, 1, 500)
    assert np.array_equal(o.interp(x, y, z, rt), np_step.interp32_3d(fa1, fa2, P, x, y, z, rt))


def _narrow_region(P):
    for i, v in enumerate((0.2, 0.8, 0.1, 0.7, 0.3, 0.9)):
        P.acc_region[i] = v


@pytest.mark.parametrize("conf,cli,tweak", [
    (dict(), dict(), None),                                                     # C5 as configured
    (dict(mag_dependency=1, momentum_dependency=1), dict(), None),              # quirk 8: no 1/B in dk
    (dict(mag_dependency=1, momentum_dependency=1, kret=0.0), dict(dpp_wave=1, dpp_shear=1), None),
    (dict(momentum_dependency=0), dict(dpp_wave=1, dpp_shear=1, weak_scattering=0), None),
    (dict(mag_dependency=1, momentum_dependency=1), dict(nlgc=1, kperp_kpara=0.05), None),
    (dict(mag_dependency=1), dict(nlgc=1, kperp_kpara=0.05, dpp_wave=1, dpp_shear=1), None),
    (dict(acc_region_flag=1), dict(), _narrow_region),
])
def test_3d_step_matches_numpy_restatement(conf, cli, tweak):
    """push_particle_3d + both kappa routines + D_pp: the path of BASELINE's C5."""
    conf = dict(conf, r1=4, r2=8, r3=12)
    P, w, frames, before, after, u, o = _table_step("c5", 24, conf, cli, tweak=tweak)
    fa1 = np_step.gradients32_3d(frames[0], P.dx, P.dy, P.dz)
    fa2 = np_step.gradients32_3d(frames[1], P.dx, P.dy, P.dz)
    F = np_step.interp32_3d(fa1, fa2, P, before["x"], before["y"], before["z"], (before["t"] - 0.0) / w.dt_out)
    qdrift = float(np.float32(1.0) / np.float32(3 * P.pcharge))
    want = np_step.push_3d_like(P, F, before["p"], before["mu"], P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out,
                                u, before["x"], before["y"], before["z"], before["t"], qdrift, True)
    _check(after, want, ("x", "y", "z", "p", "t", "dt"))
    assert np.any(after["p"] != before["p"])


@pytest.mark.parametrize("conf,cli", [
    (dict(), dict()),
    (dict(), dict(dpp_wave=1, dpp_shear=1)),
    (dict(kret=0.0), dict(nlgc=1, kperp_kpara=0.05)),
])
def test_2d_include_3rd_step_matches_numpy_restatement(conf, cli):
    """push_particle_2d_include_3rd: 2-D fields, 3-D motion."""
    P, w, frames, before, after, u, o = _table_step("c1", 48, conf, dict(cli, include_3rd_dim=1))
    fa1 = np_step.gradients32(frames[0], P.dx, P.dy)
    fa2 = np_step.gradients32(frames[1], P.dx, P.dy)
    F = np_step.interp32(fa1, fa2, P, before["x"], before["y"], (before["t"] - 0.0) / w.dt_out)
    qdrift = float(np.float32(1.0) / np.float32(3 * P.pcharge))
    want = np_step.push_3d_like(P, F, before["p"], before["mu"], P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out,
                                u, before["x"], before["y"], before["z"], before["t"], qdrift, False)
    _check(after, want, ("x", "y", "z", "p", "t", "dt"))
    assert np.any(after["z"] != before["z"])


@pytest.mark.parametrize("key,conf,cli,tweak", [
    ("c4", dict(), dict(), None),                                               # C4 as configured: wave + shear
    ("c4", dict(kret=0.0), dict(weak_scattering=0), None),
    ("c4", dict(momentum_dependency=0, mag_dependency=0), dict(), None),
    ("c1", dict(), dict(nlgc=1, kperp_kpara=0.05), None),
    ("c4", dict(), dict(nlgc=1, kperp_kpara=0.05), None),
    ("c1", dict(acc_region_flag=1), dict(), _narrow_region),
])
def test_2d_step_with_dpp_and_nlgc_matches_numpy_restatement(key, conf, cli, tweak):
    """push_particle_2d with D_pp (BASELINE's C4), NLGC kappa and the acceleration region."""
    P, w, frames, before, after, u, o = _table_step(key, 48, conf, cli, tweak=tweak)
    fa1 = np_step.gradients32(frames[0], P.dx, P.dy)
    fa2 = np_step.gradients32(frames[1], P.dx, P.dy)
    F = np_step.interp32(fa1, fa2, P, before["x"], before["y"], (before["t"] - 0.0) / w.dt_out)
    qdrift = float(np.float32(1.0) / np.float32(3 * P.pcharge))
    want = np_step.push_2d_general(P, F, before["p"], before["mu"], P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out,
                                   u, before["x"], before["y"], before["t"], qdrift)
    _check(after, want, ("x", "y", "p", "t", "dt"))
    if key == "c4":
        assert P.dpp_wave == 1 and P.dpp_shear == 1


# ---- the mover's control flow against a plain-Python restatement ---------------------------------
def _python_interval(P, w, frames, ptls, t0, dtf, nsteps_interval, num_fine_steps, keep_cycle_flags=False):
    """One MHD interval of every particle with np_step.mover_one_particle (2-D Parker), Philox uniforms
    keyed like the library's: counter (step_lo, step_hi, tag_injected, tag_splitted), key (seed, origin)."""
    full3d = P.ndim == 3
    if full3d:
        fa1 = np_step.gradients32_3d(frames[0], P.dx, P.dy, P.dz)
        fa2 = np_step.gradients32_3d(frames[1], P.dx, P.dy, P.dz)
    else:
        fa1 = np_step.gradients32(frames[0], P.dx, P.dy)
        fa2 = np_step.gradients32(frames[1], P.dx, P.dy)
    qdrift = float(np.float32(1.0) / np.float32(3 * P.pcharge))
    dt_min, dt_max = P.dt_min_rel * dtf, P.dt_max_rel * dtf
    tally = dict(leak=0.0, leak_negp=0.0, steps=0)
    one = lambda v: np.array([v], dtype=np.float64)
    out = []
    for rec, rng0 in zip(ptls, rng_steps(ptls)):
        s = {k: float(rec[k]) for k in ("x", "y", "z", "p", "mu", "t", "dt", "weight")}
        s.update(count_flag=int(rec["count_flag"]), nsteps_pushed=int(rec["nsteps_pushed"]), rng=int(rng0))
        key = (P.seed & 0xFFFFFFFF, ((P.seed >> 32) + int(rec["origin"])) & 0xFFFFFFFF)
        tags = (int(rec["tag_injected"]), int(rec["tag_splitted"]))

        def push(s, fixed):
            blk = philox4x32_10((s["rng"] & 0xFFFFFFFF, s["rng"] >> 32) + tags, key)
            u = np.array([[b / 4294967295.0 for b in blk]])
            d = {}
            if full3d or P.include_3rd_dim:   # push_particle_3d / push_particle_2d_include_3rd
                rt = one((s["t"] - t0) / dtf)
                F = (np_step.interp32_3d(fa1, fa2, P, one(s["x"]), one(s["y"]), one(s["z"]), rt) if full3d
                     else np_step.interp32(fa1, fa2, P, one(s["x"]), one(s["y"]), rt))
                x, y, z, p, t, dt = np_step.push_3d_like(P, F, one(s["p"]), one(s["mu"]), dt_min, dt_max, u, one(s["x"]),
                                                         one(s["y"]), one(s["z"]), one(s["t"]), qdrift, full3d,
                                                         dt_fixed=one(s["dt"]) if fixed else None, deltas=d)
                s.update(x=float(x[0]), y=float(y[0]), z=float(z[0]), p=float(p[0]), t=float(t[0]), dt=float(dt[0]),
                         rng=s["rng"] + 1)
                return float(d["x"][0]), float(d["y"][0]), float(d["z"][0]), float(d["p"][0])
            F = np_step.interp32(fa1, fa2, P, one(s["x"]), one(s["y"]), one((s["t"] - t0) / dtf))
            x, y, p, t, dt = np_step.push_2d_general(P, F, one(s["p"]), one(s["mu"]), dt_min, dt_max, u, one(s["x"]),
                                                     one(s["y"]), one(s["t"]), qdrift,
                                                     dt_fixed=one(s["dt"]) if fixed else None, deltas=d)
            s.update(x=float(x[0]), y=float(y[0]), p=float(p[0]), t=float(t[0]), dt=float(dt[0]), rng=s["rng"] + 1)
            s["z"] = s["z"] + float(d["z_in_pusher"][0])   # moved inside push_particle_2d; the mover's deltaz stays 0
            return float(d["x"][0]), float(d["y"][0]), 0.0, float(d["p"][0])

        np_step.mover_one_particle(P, s, push, t0, dtf, nsteps_interval, num_fine_steps, tally)
        if keep_cycle_flags:
            s["flag_after_cycle"] = s["count_flag"]
        if s["count_flag"] == np_step.INBOX:
            np_step.final_boundary_pass(P, s, tally)
        out.append(s)
    return out, tally


@pytest.mark.parametrize("key,conf,nfine,cli", [("c1", dict(dt_min_rel=2e-3), 1, None), ("c1", dict(dt_min_rel=2e-3), 3, None),
                                                ("c3", dict(dt_min_rel=2e-3), 2, None),
                                                ("c4", dict(dt_min_rel=4e-3), 2, None),
                                                ("c1", dict(dt_min_rel=2e-3), 2, dict(check_drift_2d=1)),
                                                ("c1", dict(dt_min_rel=2e-3), 2, dict(include_3rd_dim=1)),
                                                ("c5", dict(dt_min_rel=4e-3, r1=4, r2=8, r3=12), 2, None),
                                                ("c5", dict(dt_min_rel=4e-3, r1=4, r2=8, r3=12, pbcx=1, pbcy=1, pbcz=1), 1,
                                                 None)])
def test_mover_interval_matches_python_restatement(key, conf, nfine, cli):
    """particle_mover_one_cycle + particle_mover: target times, roll-back and fixed-dt re-push, the BC test
    at the top of every step with the extended bounds, the final pass with the true ones, remove_particles."""
    n = 24
    w, P, frames, _ = make_case(key, grid=24 if key == "c5" else 48, nptl=n, conf=conf, cli=cli)
    o = Oracle(P, w.nptl_max)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_uniform(n, 0.0, 1, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    steps = o.particle_mover(0.0, w.dt_out, 100, nfine, 1)
    after = sort_by_key(o.download_particles())
    esc = sort_by_key(o.download_escaped())
    ref, tally = _python_interval(P, w, frames, before, 0.0, w.dt_out, 100, nfine)
    assert tally["steps"] == steps, (tally["steps"], steps)
    order = np.lexsort((before["tag_splitted"], before["tag_injected"], before["origin"]))
    ref = [ref[i] for i in order]
    tags = before["tag_injected"][order]
    inbox = [r for r in ref if r["count_flag"] == np_step.INBOX]
    gone = [(t, r) for t, r in zip(tags, ref) if r["count_flag"] < 0]
    assert len(inbox) == len(after) and len(gone) == len(esc)
    for name in ("x", "y", "z", "p", "t", "dt"):
        got, want = after[name], np.array([r[name] for r in inbox])
        assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), name
    if cli and cli.get("check_drift_2d"):
        assert np.any(after["z"] != 0.0)            # the out-of-plane drift accumulated in z
    assert np.array_equal(after["nsteps_pushed"], np.array([r["nsteps_pushed"] for r in inbox]))
    assert np.array_equal(rng_steps(after), np.array([r["rng"] for r in inbox], dtype=np.uint64))
    assert np.all(after["t"] == w.dt_out)                       # every survivor ends ON the frame time
    assert np.array_equal(esc["count_flag"], np.array([r["count_flag"] for _, r in gone], dtype=esc["count_flag"].dtype))
    c = o.counters()
    assert c.leak == tally["leak"] and c.leak_negp == tally["leak_negp"]
    if key == "c3":
        assert len(gone) > 0, "the open-x case must lose particles"


@pytest.mark.parametrize("key,extra", [("c1", {}), ("c2", {}), ("c5", {}), ("c5", dict(pbcx=1, pbcy=1, pbcz=1))])
def test_boundary_quirks_match_python_restatement(key, extra):
    """Particles parked around the box edges: inside the half-cell margin they are left alone by the step
    loop (extended bounds) and wrapped by L / removed by the final pass; beyond it the loop wraps them by
    L + dx or lets them escape (c2: open boundaries)."""
    n = 16
    w, P, frames, _ = make_case(key, grid=24 if key == "c5" else 48, nptl=n,
                                conf=dict(extra, dt_min_rel=5e-3, **(dict(r1=4, r2=8, r3=12) if key == "c5" else {})))
    open_box = key == "c2" or bool(extra)
    o = Oracle(P, w.nptl_max)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_uniform(n, 0.0, 1, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    ptl = o.download_particles()
    offs = np.array([0.3, 0.6, -0.3, -0.6])
    ptl["x"][0:4] = np.where(offs > 0, P.xmax + offs * P.dx, P.xmin + offs * P.dx)
    ptl["y"][4:8] = np.where(offs > 0, P.ymax + offs * P.dy, P.ymin + offs * P.dy)
    ptl["x"][8:10] = [P.xmax + 0.7 * P.dx, P.xmin - 0.2 * P.dx]      # both axes at once
    ptl["y"][8:10] = [P.ymin - 0.7 * P.dy, P.ymax + 0.2 * P.dy]
    if P.ndim == 3:
        ptl["z"][10:14] = np.where(offs > 0, P.zmax + offs * P.dz, P.zmin + offs * P.dz)
    ptl["t"][:14] = 0.02                                             # a few steps each
    o.upload_particles(ptl)
    steps = o.particle_mover(0.0, w.dt_out, 100, 1, 1)
    raw, raw_esc = o.download_particles(), o.download_escaped()
    after, esc = sort_by_key(raw), sort_by_key(raw_esc)
    ref, tally = _python_interval(P, w, frames, ptl, 0.0, w.dt_out, 100, 1, keep_cycle_flags=True)
    # remove_particles after the cycle, then again after the final pass: the ORDER of survivors and escapees
    keep1, esc1 = np_step.remove_particles([r["flag_after_cycle"] for r in ref])
    keep2, esc2 = np_step.remove_particles([ref[i]["count_flag"] for i in keep1])
    assert np.array_equal(raw["tag_injected"], ptl["tag_injected"][[keep1[i] for i in keep2]])
    assert np.array_equal(raw_esc["tag_injected"], ptl["tag_injected"][esc1 + [keep1[i] for i in esc2]])
    order = np.lexsort((ptl["tag_splitted"], ptl["tag_injected"], ptl["origin"]))
    ref = [ref[i] for i in order]
    inbox = [r for r in ref if r["count_flag"] == np_step.INBOX]
    gone = [r for r in ref if r["count_flag"] < 0]
    assert tally["steps"] == steps and len(inbox) == len(after) and len(gone) == len(esc)
    for name in ("x", "y", "z", "p", "t"):
        want = np.array([r[name] for r in inbox])
        assert np.abs(after[name] - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), name
    assert np.array_equal(esc["count_flag"], np.array([r["count_flag"] for r in gone], dtype=esc["count_flag"].dtype))
    assert o.counters().leak == tally["leak"]
    if open_box:
        assert len(gone) >= 6          # everything parked outside the true box leaves through an open boundary
        if P.ndim == 3:
            assert {-5, -6} <= set(int(f) for f in esc["count_flag"])
    else:
        assert len(gone) == 0 and np.all((after["x"] >= P.xmin) & (after["x"] <= P.xmax))


# ---- injection, value for value ------------------------------------------------------------------
def _py_inject_uniform(P, n, dt, dist_flag, particle_v0, t_frame, dt_mhd, box, power_index, tag0=0):
    """inject_particles_spatial_uniform (whole-domain branch, particle_module.f90:487-498) +
    inject_one_particle (:385-441) in plain Python.  Uniform k of particle `tag` is word k % 4 of
    Philox((k // 4, 0, tag, 0), (seed_lo, seed_hi + origin)) / 4294967295."""
    import math
    key = (P.seed & 0xFFFFFFFF, ((P.seed >> 32) + P.mpi_rank) & 0xFFFFFFFF)
    mu_max = float(np.float32(0.99))
    out = []
    for i in range(n):
        tag = tag0 + i
        state = dict(k=0, buf=None)

        def u():
            if state["k"] % 4 == 0:
                state["buf"] = philox4x32_10((state["k"] // 4, 0, tag, 0), key)
            v = state["buf"][state["k"] % 4] / 4294967295.0
            state["k"] += 1
            return v
        x = u() * (box[3] - box[0]) + box[0]
        y = u() * (box[4] - box[1]) + box[1]
        z = u() * (box[5] - box[2]) + box[2]
        mu = mu_max * (2.0 * u() - 1.0)
        if dist_flag == 0:
            ftest, fxp = 1.0, 0.5
            while ftest > fxp:
                ptmp = (u() * (P.pmax - P.pmin) + P.pmin) / P.p0
                fxp = ptmp ** 2 * math.exp(-ptmp ** 2)
                ftest = u() * float(np.float32(0.37))
            p = ptmp * P.p0
        elif dist_flag == 1:
            p = P.p0
        else:
            r01 = u()
            if int(power_index) == 1:
                p = (P.pmax / P.p0) ** r01 * P.p0
            else:
                norm = P.pmax ** (-power_index + 1) - P.p0 ** (-power_index + 1)
                p = (r01 * norm + P.p0 ** (-power_index + 1)) ** (1.0 / (-power_index + 1))
        out.append(dict(x=x, y=y, z=z, mu=mu, p=p, v=particle_v0 * p / P.p0, t=t_frame + u() * dt_mhd,
                        dt=dt, weight=1.0, tag_injected=tag, tag_splitted=1, origin=P.mpi_rank))
    return out


@pytest.mark.parametrize("dist_flag,power_index", [(0, 6.2), (1, 6.2), (2, 6.2), (2, 1.0)])
def test_injection_matches_python_restatement(dist_flag, power_index):
    w, P, _, _ = make_case("c1", grid=32, nptl=500)
    P.mpi_rank = 3
    o = Oracle(P, 2000)
    box = [0.25, 0.5, 0.0, 1.5, 1.75, 0.0]                       # -ip box: a sub-volume of the domain
    o.inject_uniform(300, 2e-6, dist_flag, 50.0, 0.4, 0.1, box, power_index)
    o.inject_uniform(200, 2e-6, dist_flag, 50.0, 0.5, 0.1, box, power_index)   # tags continue at 300
    a = o.download_particles()
    ref = (_py_inject_uniform(P, 300, 2e-6, dist_flag, 50.0, 0.4, 0.1, box, power_index)
           + _py_inject_uniform(P, 200, 2e-6, dist_flag, 50.0, 0.5, 0.1, box, power_index, tag0=300))
    assert len(a) == 500
    for f in ("x", "y", "z", "mu", "t", "dt", "weight"):
        assert np.array_equal(a[f], np.array([r[f] for r in ref])), f
    for f in ("p", "v"):                                          # exp()/pow() may differ in the last bit
        want = np.array([r[f] for r in ref])
        assert np.abs(a[f] - want).max() <= 4e-16 * np.abs(want).max(), f
    for f in ("tag_injected", "tag_splitted", "origin"):
        assert np.array_equal(a[f], np.array([r[f] for r in ref])), f
    assert np.all(a["count_flag"] == 1) and np.all(a["split_times"] == 0) and np.all(rng_steps(a) == 0)
    if dist_flag == 0:                                            # the rejection loop really ran
        assert len(np.unique(a["p"])) == 500 and a["p"].min() >= P.pmin and a["p"].max() <= P.pmax


def test_escaped_distributions_match_numpy_binning():
    """calc_escaped_distributions, global part (diagnostics.f90:934-955): fescaped(imu, ip, |count_flag|) +=
    weight for p in (pmin, pmax]; C2 has open boundaries, so several faces fill."""
    w, P, frames, ts = make_case("c2", grid=32, nptl=4000, conf=dict(dt_min_rel=1e-3))
    o = Oracle(P, w.nptl_max)
    got = []
    run_intervals(o, frames, ts, nptl=4000, particle_v0=w.particle_v0, dist_flag=2, dump_escaped_dist=True,
                  on_interval=lambda tf, d: got.append((d["fescaped"].copy(), None)))
    # the escaped list is reset after every dump: bin the last interval's escapees by hand
    o2 = Oracle(P, w.nptl_max)
    o2.upload_fields(0, frames[0])
    o2.upload_fields(1, frames[1])
    o2.inject_uniform(4000, 0.0, 2, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    o2.particle_mover(0.0, w.dt_out, 100, 1, 1)
    esc = o2.download_escaped()
    assert len(esc) > 20 and set(np.unique(esc["count_flag"])) <= {-1, -2, -3, -4}
    assert len(set(np.unique(esc["count_flag"]))) >= 2
    pmin_log = np.log10(P.pmin)
    dp_log = (np.log10(P.pmax) - pmin_log) / P.npp_global
    ref = np.zeros((2 * P.ndim, P.npp_global, P.nmu_global))
    for r in esc:
        if r["p"] > P.pmin and r["p"] <= P.pmax and -1.0 <= r["mu"] <= 1.0:
            ip = int(np.floor((np.log10(r["p"]) - pmin_log) / dp_log)) + 1
            imu = int(np.floor((r["mu"] + 1.0) / float(np.float32(2.0) / np.float32(P.nmu_global)))) + 1
            ref[abs(int(r["count_flag"])) - 1, min(ip, P.npp_global) - 1, imu - 1] += r["weight"]
    assert np.array_equal(o2.escaped_diagnostics(), ref)
    assert np.array_equal(got[0][0], ref)          # the same interval through run_intervals
    assert ref.sum() == esc["weight"].sum() == o2.counters().leak


def test_host_restart_files_continue_the_run_bit_identically(tmp_path):
    """dump_restart / read_restart + run_intervals(tmin=): a run stopped after frame 2 (quota), dumped and
    resumed in a fresh simulation gives the particles, counters and spectra of the uninterrupted run."""
    from stochastic_parker_b200 import dump_restart, read_restart
    w, P, frames, ts = make_case("c3", grid=48, nptl=600, nframes=5)
    kw = dict(nptl=600, dist_flag=2, particle_v0=w.particle_v0, pmin_split=1.05, split_ratio=1.05, num_fine_steps=2)
    a = Oracle(P, 20000)
    full, _ = run_intervals(a, frames, ts, **kw)
    b = Oracle(P, 20000)
    part1, _ = run_intervals(b, frames[:3], ts[:3], **kw)            # -te 2
    assert part1[-1]["frame"] == 2
    dump_restart(b, str(tmp_path) + "/", 2, part1[-1]["frame"])
    c = Oracle(P, 20000)
    tmin = read_restart(c, str(tmp_path) + "/")
    assert tmin == 2
    part2, _ = run_intervals(c, frames, ts, tmin=tmin, **kw)         # -rf .true. -te 4
    assert [d["frame"] for d in part1 + part2] == [d["frame"] for d in full]      # no second frame-0 record
    for x, y in zip(part1 + part2, full):
        assert np.array_equal(x["fglobal"], y["fglobal"]) and np.array_equal(x["quick"], y["quick"])
    assert_particles_identical(c.download_particles(), a.download_particles(), "restart")
    ca, cc = a.counters(), c.counters()
    assert (ca.nptl_current, ca.nptl_split, ca.tag_max, ca.leak) == (cc.nptl_current, cc.nptl_split, cc.tag_max, cc.leak)
    # quota: the loop stops after the first interval that ends beyond it
    d = Oracle(P, 20000)
    rec, _ = run_intervals(d, frames, ts, quota_seconds=0.0, **kw)
    assert [r["frame"] for r in rec] == [0, 1]


@pytest.mark.parametrize("key,grid,conf", [("c2", 32, dict(dt_min_rel=1e-3)),
                                           ("c5", 16, dict(pbcx=1, pbcy=1, pbcz=1, dt_min_rel=2e-3, r1=2, r2=4, r3=8))])
def test_local_escaped_distributions_match_numpy_binning(key, grid, conf):
    """calc_escaped_distributions, local part (diagnostics.f90:956-1170): every escapee lands in the array of
    the face it left through, binned in that face's two coordinates with the set's own p and mu bins (top
    momentum bin never filled)."""
    w, P, frames, ts = make_case(key, grid=grid, nptl=3000, conf=conf)
    o = Oracle(P, w.nptl_max)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_uniform(3000, 0.0, 2, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    o.particle_mover(0.0, w.dt_out, 100, 1, 1)
    esc = o.download_escaped()
    faces = set(int(f) for f in np.unique(esc["count_flag"]))
    assert len(esc) > 30 and len(faces) >= (4 if P.ndim == 3 else 2)
    got = o.escaped_local_diagnostics()
    total = 0.0
    for k in range(4):
        s = P.local[k]
        if not s.enabled:
            assert got[k] is None
            continue
        nrz, nry, nrx, npb, nmu = o.local_shape(k)
        ref = {"x": np.zeros((2, nrz, nry, npb, nmu)), "y": np.zeros((2, nrz, nrx, npb, nmu)),
               "z": np.zeros((2, nry, nrx, npb, nmu))}
        dpl = (np.log10(s.pmax) - np.log10(s.pmin)) / s.npbins
        dmu = float(np.float32(2.0) / np.float32(s.nmu))
        for r in esc:
            ix = int(np.floor((r["x"] - P.xmin) / (P.lx / nrx)))
            iy = int(np.floor((r["y"] - P.ymin) / (P.ly / nry)))
            iz = int(np.floor((r["z"] - P.zmin) / (P.lz / nrz)))
            ip = int(np.floor((np.log10(r["p"]) - np.log10(s.pmin)) / dpl))        # 0-based
            imu = int(np.floor((r["mu"] + 1.0) / dmu))
            if not (0 <= ip < npb - 1 and 0 <= imu < nmu):
                continue
            cx, cy, cz = 0 <= ix < nrx, 0 <= iy < nry, 0 <= iz < nrz
            face = -int(r["count_flag"])
            side = (face - 1) % 2
            if face <= 2 and cy and cz:
                ref["x"][side, iz, iy, ip, imu] += r["weight"]
            elif face in (3, 4) and cx and cz:
                ref["y"][side, iz, ix, ip, imu] += r["weight"]
            elif face >= 5 and cx and cy:
                ref["z"][side, iy, ix, ip, imu] += r["weight"]
        assert np.array_equal(got[k]["x"], ref["x"]), k
        assert np.array_equal(got[k]["y"], ref["y"]), k
        if P.ndim == 3:
            assert np.array_equal(got[k]["z"], ref["z"]), k
            total += got[k]["z"].sum()
        else:
            assert got[k]["z"] is None
        total += got[k]["x"].sum() + got[k]["y"].sum()
    assert total > 0.0


def test_focused_transport_mover_rolls_back_speed_and_pitch_angle():
    """The mover with the 2-D focused-transport pusher: the roll-back before the fixed-dt re-push also
    restores v and mu (particle_module.f90:1712-1713).  Plain-Python mover + the numpy FT step against the
    oracle's particle_mover over a whole interval with two fine steps."""
    n = 20
    w, P, frames, _ = make_case("c1", grid=48, nptl=n, conf=dict(dt_min_rel=2e-3),
                                cli=dict(focused_transport=1, duu_init=5.0))
    o = Oracle(P, w.nptl_max)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_uniform(n, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    steps = o.particle_mover(0.0, w.dt_out, 100, 2, 0)
    after = sort_by_key(o.download_particles())
    fa1 = np_step.gradients32(frames[0], P.dx, P.dy)
    fa2 = np_step.gradients32(frames[1], P.dx, P.dy)
    dt_min, dt_max = P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out
    tally = dict(leak=0.0, leak_negp=0.0, steps=0)
    one = lambda v: np.array([v], dtype=np.float64)
    out = []
    for rec, rng0 in zip(before, rng_steps(before)):
        s = {k: float(rec[k]) for k in ("x", "y", "z", "p", "v", "mu", "t", "dt", "weight")}
        s.update(count_flag=int(rec["count_flag"]), nsteps_pushed=int(rec["nsteps_pushed"]), rng=int(rng0))
        key = (P.seed & 0xFFFFFFFF, ((P.seed >> 32) + int(rec["origin"])) & 0xFFFFFFFF)
        tags = (int(rec["tag_injected"]), int(rec["tag_splitted"]))

        def push(s, fixed):
            F = np_step.interp32(fa1, fa2, P, one(s["x"]), one(s["y"]), one((s["t"] - 0.0) / w.dt_out))
            blk = philox4x32_10((s["rng"] & 0xFFFFFFFF, s["rng"] >> 32) + tags, key)
            u = np.array([[b / 4294967295.0 for b in blk]])
            st = {k: one(s[k]) for k in ("x", "y", "p", "v", "mu", "t")}
            x, y, p, v, mu, t, dt = _np_step_2d_ft(P, F, st, u, dt_min, dt_max, dt_fixed=one(s["dt"]) if fixed else None)
            d = (float(x[0]) - s["x"], float(y[0]) - s["y"], 0.0, float(p[0]) - s["p"], float(v[0]) - s["v"],
                 float(mu[0]) - s["mu"])
            s.update(x=float(x[0]), y=float(y[0]), p=float(p[0]), v=float(v[0]), mu=float(mu[0]), t=float(t[0]),
                     dt=float(dt[0]), rng=s["rng"] + 1)
            return d

        np_step.mover_one_particle(P, s, push, 0.0, w.dt_out, 100, 2, tally)
        if s["count_flag"] == np_step.INBOX:
            np_step.final_boundary_pass(P, s, tally)
        out.append(s)
    assert tally["steps"] == steps, (tally["steps"], steps)
    order = np.lexsort((before["tag_splitted"], before["tag_injected"], before["origin"]))
    ref = [out[i] for i in order if out[i]["count_flag"] == np_step.INBOX]
    assert len(ref) == len(after)
    for name in ("x", "y", "p", "v", "mu", "t", "dt"):
        want = np.array([r[name] for r in ref])
        assert np.abs(after[name] - want).max() <= 1e-10 * max(1.0, np.abs(want).max()), name
    assert np.all(after["t"] == w.dt_out) and np.any(after["mu"] != sort_by_key(before)["mu"])


@pytest.mark.parametrize("key,grid,dist_flag", [("c3", 48, 1), ("c3", 48, 0), ("c5", 16, 2)])
def test_shock_injection_matches_python_restatement(key, grid, dist_flag):
    """locate_shock_xpos + interp_shock_location + inject_particles_at_shock (mhd_data_parallel.f90:1988-2045,
    particle_module.f90:542-633) value for value, quirks included: only the LATER frame's shock positions
    survive rt = 0 (swapped time weights), rz is computed from dpy, the weights of the z pair therefore do
    not sum to one in 3-D, x = (sx + 2) * lx / nxg, the Maxwellian envelope is 0.75 with exp(-p^2/2), t is
    the frame time itself."""
    import math
    conf = dict(r1=2, r2=4, r3=8) if key == "c5" else None
    w, P, frames, _ = make_case(key, grid=grid, nptl=300, conf=conf)
    o = Oracle(P, w.nptl_max)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.inject_at_shock(300, 1e-6, dist_flag, w.particle_v0, 0.3, 6.2)
    a = o.download_particles()
    assert len(a) == 300
    if P.ndim == 2:
        fa2 = np_step.gradients32(frames[1], P.dx, P.dy)[None]          # (1, ny+4, nx+4, 32)
    else:
        fa2 = np_step.gradients32_3d(frames[1], P.dx, P.dy, P.dz)
    sx2 = np.argmax(np.abs(fa2[..., 8]), axis=-1) + 1                     # maxloc(abs(dvx_dx), dim=1), 1-based
    key_ = (P.seed & 0xFFFFFFFF, ((P.seed >> 32) + P.mpi_rank) & 0xFFFFFFFF)
    mu_max = float(np.float32(0.99))
    nxg = P.nx + 4
    for tag in range(300):
        st = dict(k=0, buf=None)

        def u():
            if st["k"] % 4 == 0:
                st["buf"] = philox4x32_10((st["k"] // 4, 0, tag, 0), key_)
            v = st["buf"][st["k"] % 4] / 4294967295.0
            st["k"] += 1
            return v
        y = u() * (P.ymax - P.ymin) + P.ymin
        dpy = y / P.dy
        iy = math.floor(dpy)
        z = u() * (P.zmax - P.zmin) + P.zmin
        iz = math.floor(z / P.dz)
        ry = dpy - iy
        rz = dpy - iy
        wts = [(1 - ry) * (1 - rz), ry * (1 - rz), (1 - ry) * rz, ry * rz]
        sx = 0.0
        if P.ndim == 2:
            for j in (0, 1):
                sx = sx + sx2[0, iy + j] * wts[j]                     # Fortran index iy+j-1, lower bound -1
        else:
            for k in (0, 1):
                for j in (0, 1):
                    sx = sx + sx2[iz + k, iy + j] * wts[k * 2 + j]
        x = ((sx * (1.0 - 0.0) + 0.0 * 0.0) + 2) * (P.xmax - P.xmin) / nxg
        if dist_flag == 0:
            ftest, fxp = 1.0, 0.5
            while ftest > fxp:
                ptmp = (u() * (P.pmax - P.pmin) + P.pmin) / P.p0
                fxp = ptmp ** 2 * math.exp(-0.5 * ptmp ** 2)
                ftest = u() * float(np.float32(0.75))
            p = ptmp * P.p0
        elif dist_flag == 1:
            p = P.p0
        else:
            r01 = u()
            norm = P.pmax ** (-6.2 + 1) - P.p0 ** (-6.2 + 1)
            p = (r01 * norm + P.p0 ** (-6.2 + 1)) ** (1.0 / (-6.2 + 1))
        mu = mu_max * (2.0 * u() - 1.0)
        r = a[tag]
        assert (r["x"], r["y"], r["z"], r["mu"], r["t"], r["dt"]) == (x, y, z, mu, 0.3, 1e-6), tag
        assert abs(r["p"] - p) <= 4e-16 * p and r["tag_injected"] == tag and r["weight"] == 1.0
    assert len(np.unique(a["x"])) > 3 or P.ndim == 2


def test_random_switch_combinations_match_numpy_restatement():
    """24 seeded random combinations of the Parker-transport switches (geometry, both dependencies, kret,
    NLGC, D_pp wave / shear, weak or strong scattering, acceleration region, time interpolation) -- one step of
    the C oracle against the numpy restatement each, so that switch INTERACTIONS are pinned too."""
    rng = np.random.default_rng(2024)
    seen = set()
    for trial in range(24):
        geom = ["2d", "2d3", "3d"][trial % 3]
        conf = dict(mag_dependency=int(rng.integers(0, 2)), momentum_dependency=int(rng.integers(0, 2)),
                    kret=float(rng.choice([0.0, 0.01, 0.3])), acc_region_flag=int(rng.integers(0, 2)))
        cli = dict(nlgc=int(rng.integers(0, 2)), kperp_kpara=0.05, dpp_wave=int(rng.integers(0, 2)),
                   dpp_shear=int(rng.integers(0, 2)), weak_scattering=int(rng.integers(0, 2)),
                   time_interp=int(rng.integers(0, 2)))
        if geom == "2d3":
            cli["include_3rd_dim"] = 1
        if geom == "3d":
            conf.update(r1=4, r2=8, r3=12)
        seen.add((geom, conf["mag_dependency"], cli["nlgc"], cli["dpp_wave"], cli["dpp_shear"]))
        key, grid = ("c5", 24) if geom == "3d" else ("c1", 48)
        tweak = _narrow_region if conf["acc_region_flag"] else None
        w, P, frames, _ = make_case(key, grid=grid, nptl=120, conf=conf, cli=cli)
        if tweak:
            tweak(P)
        P.rng_mode = RNG_TABLE
        o = Oracle(P, w.nptl_max)
        u = np.random.default_rng(100 + trial).uniform(0, 1, (120, 2, 4))
        o.set_rng_table(u)
        o.upload_fields(0, frames[0])
        if P.time_interp:
            o.upload_fields(1, frames[1])
        o.inject_uniform(120, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
        before = o.download_particles()
        assert o.debug_push_n(0.0, w.dt_out, 1) == 120
        after = o.download_particles()
        uu = u[before["tag_injected"], 0]
        rt = (before["t"] - 0.0) / w.dt_out
        if geom == "3d":
            fa1 = np_step.gradients32_3d(frames[0], P.dx, P.dy, P.dz)
            fa2 = np_step.gradients32_3d(frames[1], P.dx, P.dy, P.dz) if P.time_interp else None
            F = np_step.interp32_3d(fa1, fa2, P, before["x"], before["y"], before["z"], rt)
        else:
            fa1 = np_step.gradients32(frames[0], P.dx, P.dy)
            fa2 = np_step.gradients32(frames[1], P.dx, P.dy) if P.time_interp else None
            F = np_step.interp32(fa1, fa2, P, before["x"], before["y"], rt)
        qdrift = float(np.float32(1.0) / np.float32(3 * P.pcharge))
        dt_min, dt_max = P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out
        if geom == "2d":
            want = np_step.push_2d_general(P, F, before["p"], before["mu"], dt_min, dt_max, uu, before["x"], before["y"],
                                           before["t"], qdrift)
            names = ("x", "y", "p", "t", "dt")
        else:
            want = np_step.push_3d_like(P, F, before["p"], before["mu"], dt_min, dt_max, uu, before["x"], before["y"],
                                        before["z"], before["t"], qdrift, geom == "3d")
            names = ("x", "y", "z", "p", "t", "dt")
        for name, ref in zip(names, want):
            scale = np.maximum(np.abs(ref), 1.0) if name in "xyz" else np.abs(ref)
            err = np.abs(after[name] - ref) / scale
            assert err.max() < 4e-15, (trial, geom, conf, cli, name, err.max())
    assert len(seen) >= 15


@pytest.mark.parametrize("nlgc,third,flags", [(1, 0, (1, 1)), (1, 0, (1, 0)), (1, 0, (0, 1)), (1, 1, (1, 1)), (0, 1, (1, 1))])
def test_turbulence_maps_with_nlgc_and_third_dimension_match_numpy_restatement(nlgc, third, flags):
    """The map terms of calc_spatial_diffusion_coefficients_nlgc (particle_module.f90:2486-2500, 2588-2604:
    db2_2d and lc_2d enter k_perp) and of the include_3rd branches, which the plain 2-D test above does not reach."""
    from stochastic_parker_b200 import mhd
    cli = dict(nlgc=nlgc, kperp_kpara=0.05)
    if third:
        cli["include_3rd_dim"] = 1
    w, P, frames, _ = make_case("c1", grid=48, nptl=300, cli=cli)
    P.deltab_flag, P.correlation_flag = flags
    P.rng_mode = RNG_TABLE
    o = Oracle(P, w.nptl_max)
    u = np.random.default_rng(31).uniform(0, 1, (300, 2, 4))
    o.set_rng_table(u)
    maps = [mhd.make_turbulence_maps(P.nx, P.ny, 1, f) for f in (0, 1)]
    for slot in (0, 1):
        o.upload_fields(slot, frames[slot])
        if flags[0]:
            o.upload_turbulence(0, slot, maps[slot][0], maps[slot][1])
        if flags[1]:
            o.upload_turbulence(1, slot, maps[slot][2], maps[slot][3])
    o.inject_uniform(300, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    assert o.debug_push_n(0.0, w.dt_out, 1) == 300
    after = o.download_particles()
    rt = (before["t"] - 0.0) / w.dt_out
    fa1 = np_step.gradients32(frames[0], P.dx, P.dy)
    fa2 = np_step.gradients32(frames[1], P.dx, P.dy)
    F = np_step.interp32(fa1, fa2, P, before["x"], before["y"], rt)
    g = [[np_step.turbulence_grad(m, P.dx, P.dy) for m in maps[s]] for s in (0, 1)]
    aux = np_step.interp_aux(g[0], g[1], P, before["x"], before["y"], rt)
    if not flags[0]:
        aux[:, 0:8] = 1.0          # interp_* is not called: the locals keep their initial 1.0 (particle_module.f90:1546)
    if not flags[1]:
        aux[:, 8:16] = 1.0
    qdrift = float(np.float32(1.0) / np.float32(3 * P.pcharge))
    uu = u[before["tag_injected"], 0]
    dt_min, dt_max = P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out
    if third:
        want = np_step.push_3d_like(P, F, before["p"], before["mu"], dt_min, dt_max, uu, before["x"], before["y"],
                                    before["z"], before["t"], qdrift, False, aux=aux)
        names = ("x", "y", "z", "p", "t", "dt")
    else:
        want = np_step.push_2d_general(P, F, before["p"], before["mu"], dt_min, dt_max, uu, before["x"], before["y"],
                                       before["t"], qdrift, aux=aux)
        names = ("x", "y", "p", "t", "dt")
    _check(after, want, names, tol=1e-14)


@pytest.mark.parametrize("nlgc", [0, 1])
def test_3d_turbulence_maps_match_numpy_restatement(nlgc):
    """3-D maps: the d/dz gradients, the eight-corner gather and the dk/dz map terms of both kappa routines."""
    from stochastic_parker_b200 import mhd
    w, P, frames, _ = make_case("c5", grid=24, nptl=300, conf=dict(r1=4, r2=8, r3=12, mag_dependency=1),
                                cli=dict(nlgc=nlgc, kperp_kpara=0.05))
    P.deltab_flag, P.correlation_flag = 1, 1
    P.rng_mode = RNG_TABLE
    o = Oracle(P, w.nptl_max)
    u = np.random.default_rng(41).uniform(0, 1, (300, 2, 4))
    o.set_rng_table(u)
    maps = [mhd.make_turbulence_maps(P.nx, P.ny, P.nz, f, 3) for f in (0, 1)]
    for slot in (0, 1):
        o.upload_fields(slot, frames[slot])
        o.upload_turbulence(0, slot, maps[slot][0], maps[slot][1])
        o.upload_turbulence(1, slot, maps[slot][2], maps[slot][3])
    o.inject_uniform(300, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    assert o.debug_push_n(0.0, w.dt_out, 1) == 300
    after = o.download_particles()
    rt = (before["t"] - 0.0) / w.dt_out
    fa1 = np_step.gradients32_3d(frames[0], P.dx, P.dy, P.dz)
    fa2 = np_step.gradients32_3d(frames[1], P.dx, P.dy, P.dz)
    F = np_step.interp32_3d(fa1, fa2, P, before["x"], before["y"], before["z"], rt)
    g = [[np_step.turbulence_grad_3d(m, P.dx, P.dy, P.dz) for m in maps[s]] for s in (0, 1)]
    assert np.any(g[0][0][..., 3] != 0)                       # the maps do vary along z
    aux = np_step.interp_aux_3d(g[0], g[1], P, before["x"], before["y"], before["z"], rt)
    qdrift = float(np.float32(1.0) / np.float32(3 * P.pcharge))
    want = np_step.push_3d_like(P, F, before["p"], before["mu"], P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out,
                                u[before["tag_injected"], 0], before["x"], before["y"], before["z"], before["t"], qdrift,
                                True, aux=aux)
    _check(after, want, ("x", "y", "z", "p", "t", "dt"), tol=1e-14)
    flat = np_step.push_3d_like(P, F, before["p"], before["mu"], P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out,
                                u[before["tag_injected"], 0], before["x"], before["y"], before["z"], before["t"], qdrift,
                                True, aux=np.ones_like(aux) * np.array([1, 0, 0, 0] * 4))
    assert np.max(np.abs(flat[2] - want[2])) > 1e-7           # the maps matter for the z motion


def test_focused_transport_with_turbulence_maps_matches_numpy_restatement():
    """calc_duu's map factors (particle_module.f90:3143-3148: D_mumu ~ dB^2 * lc^(1 - gamma)) and the map terms
    of kappa under focused transport."""
    from stochastic_parker_b200 import mhd
    w, P, frames, _ = make_case("c1", grid=48, nptl=300, cli=dict(focused_transport=1, duu_init=5.0))
    P.deltab_flag, P.correlation_flag = 1, 1
    P.rng_mode = RNG_TABLE
    o = Oracle(P, w.nptl_max)
    u = np.random.default_rng(51).uniform(0, 1, (300, 2, 4))
    o.set_rng_table(u)
    maps = [mhd.make_turbulence_maps(P.nx, P.ny, 1, f) for f in (0, 1)]
    for slot in (0, 1):
        o.upload_fields(slot, frames[slot])
        o.upload_turbulence(0, slot, maps[slot][0], maps[slot][1])
        o.upload_turbulence(1, slot, maps[slot][2], maps[slot][3])
    o.inject_uniform(300, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    assert o.debug_push_n(0.0, w.dt_out, 1) == 300
    after = o.download_particles()
    rt = (before["t"] - 0.0) / w.dt_out
    F = o.interp(before["x"], before["y"], before["z"], rt)
    g = [[np_step.turbulence_grad(m, P.dx, P.dy) for m in maps[s]] for s in (0, 1)]
    aux = np_step.interp_aux(g[0], g[1], P, before["x"], before["y"], rt)
    args = (P, F, before, u[before["tag_injected"], 0], P.dt_min_rel * w.dt_out, P.dt_max_rel * w.dt_out)
    x, y, p, v, mu, t, dt = _np_step_2d_ft(*args, aux=aux)
    for name, ref in (("x", x), ("y", y), ("p", p), ("v", v), ("mu", mu), ("t", t), ("dt", dt)):
        scale = np.maximum(np.abs(ref), 1.0 if name in "xy" else 1e-300)
        assert (np.abs(after[name] - ref) / scale).max() < 1e-13, name
    P0 = P.copy()
    P0.deltab_flag, P0.correlation_flag = 0, 0
    mu0 = _np_step_2d_ft(P0, *args[1:])[4]
    assert np.max(np.abs(mu0 - mu)) > 1e-6          # the maps change the pitch-angle scattering


@pytest.mark.parametrize("maps_on", [0, 1])
def test_focused_transport_with_nlgc_matches_numpy_restatement(maps_on):
    """push_particle_2d_ft fed by calc_spatial_diffusion_coefficients_nlgc(focused_transport = .true.):
    k_perp ~ mu^2, kpp = -k_perp, separate d ln k_para and d ln k_perp."""
    from stochastic_parker_b200 import mhd
    w, P, frames, _ = make_case("c1", grid=48, nptl=300, cli=dict(focused_transport=1, duu_init=5.0, nlgc=1, kperp_kpara=0.05))
    P.deltab_flag = P.correlation_flag = maps_on
    P.rng_mode = RNG_TABLE
    o = Oracle(P, w.nptl_max)
    u = np.random.default_rng(61).uniform(0, 1, (300, 2, 4))
    o.set_rng_table(u)
    maps = [mhd.make_turbulence_maps(P.nx, P.ny, 1, f) for f in (0, 1)]
    for slot in (0, 1):
        o.upload_fields(slot, frames[slot])
        if maps_on:
            o.upload_turbulence(0, slot, maps[slot][0], maps[slot][1])
            o.upload_turbulence(1, slot, maps[slot][2], maps[slot][3])
    o.inject_uniform(300, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    before = o.download_particles()
    assert o.debug_push_n(0.0, w.dt_out, 1) == 300
    after = o.download_particles()
    rt = (before["t"] - 0.0) / w.dt_out
    F = o.interp(before["x"], before["y"], before["z"], rt)
    g = [[np_step.turbulence_grad(m, P.dx, P.dy) for m in maps[s]] for s in (0, 1)]
    aux = np_step.interp_aux(g[0], g[1], P, before["x"], before["y"], rt) if maps_on else None
    x, y, p, v, mu, t, dt = _np_step_2d_ft(P, F, before, u[before["tag_injected"], 0], P.dt_min_rel * w.dt_out,
                                           P.dt_max_rel * w.dt_out, aux=aux)
    for name, ref in (("x", x), ("y", y), ("p", p), ("v", v), ("mu", mu), ("t", t), ("dt", dt)):
        scale = np.maximum(np.abs(ref), 1.0 if name in "xy" else 1e-300)
        assert (np.abs(after[name] - ref) / scale).max() < 1e-13, name


@pytest.mark.parametrize("mode", [1, 2, 4, 5])
def test_targeted_injection_matches_python_restatement(mode):
    """inject_particles_at_large_jz / _absj / _divv / _rho (particle_module.f90:785-905, 919-1061, 1250-1341,
    1356-1468) value for value: per particle a rejection loop over uniform positions in the WHOLE domain (three
    draws per trial; outside the part box counts as a rejection), the criterion interpolated at rt = 0, then mu
    and inject_one_particle."""
    w, P, frames, _ = make_case("c1", grid=48, nptl=8)
    o = Oracle(P, 4000)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    fa1 = np_step.gradients32(frames[0], P.dx, P.dy)
    fa2 = np_step.gradients32(frames[1], P.dx, P.dy)
    g = lambda F, k: F[..., 8 + k - 1]            # fields(nfields + k)

    def criterion(F):                             # the quantity the loop compares with its threshold
        if mode == 1:
            return np.abs(g(F, 16) - g(F, 14))
        if mode == 2:
            return np.sqrt((g(F, 18) - g(F, 20)) ** 2 + (g(F, 19) - g(F, 15)) ** 2 + (g(F, 14) - g(F, 16)) ** 2)
        if mode == 4:
            return -(g(F, 1) + g(F, 5))
        return F[..., 3]
    grid_vals = criterion(fa1[2:-2, 2:-2].astype(np.float64))
    box = box_of(P)
    box[0] += 0.2 * P.lx
    box[4] -= 0.3 * P.ly
    vmin = float(np.quantile(grid_vals, 0.7))
    ninj, ncells = o.inject_targeted(mode, 300, 1e-6, 1, w.particle_v0, 0.2, 0.1, box, 6.2, False, vmin, 2 * 48 * 48)
    a = o.download_particles()
    assert ninj == len(a) and ninj == int(300 * ncells / (2 * 48 * 48)) and ninj > 10
    key = (P.seed & 0xFFFFFFFF, ((P.seed >> 32) + P.mpi_rank) & 0xFFFFFFFF)
    mu_max = float(np.float32(0.99))
    one = lambda v: np.array([v])
    outside = {1: -3.0, 2: -3.0, 4: -3.0, 5: 0.0}[mode]     # divv = 3.0 outside the box, compared as -divv
    trials = 0
    for tag in range(ninj):
        st = dict(k=0, buf=None)

        def u():
            if st["k"] % 4 == 0:
                st["buf"] = philox4x32_10((st["k"] // 4, 0, tag, 0), key)
            v = st["buf"][st["k"] % 4] / 4294967295.0
            st["k"] += 1
            return v
        crit = {1: -2.0, 2: -2.0, 4: -2.0, 5: 0.0}[mode]     # jz = absj = -2, divv = +2, rho = 0 before the loop
        while crit < vmin:
            trials += 1
            x = u() * (P.xmax - P.xmin) + P.xmin
            y = u() * (P.ymax - P.ymin) + P.ymin
            z = u() * (P.zmax - P.zmin) + P.zmin
            if box[0] <= x <= box[3] and box[1] <= y <= box[4] and box[2] <= z <= box[5]:
                crit = float(criterion(np_step.interp32(fa1, fa2, P, one(x), one(y), one(0.0)))[0])
            else:
                crit = outside
        mu = mu_max * (2.0 * u() - 1.0)
        t = 0.2 + u() * 0.1
        r = a[tag]
        assert (r["x"], r["y"], r["z"], r["mu"], r["t"], r["p"]) == (x, y, z, mu, t, P.p0), tag
    assert trials > 2 * ninj            # the rejection loop really rejected


def _py_is_selected(tags, split_times_max, origin, tag_inj, tag_spl, nsplit):
    """is_particle_selected (particle_module.f90:5911-5952); tags: (nptl_tracking, ncols) = Fortran (ncols, n).
    Returns (selected, lo, hi) with 1-based column indices."""
    def findloc(vals, v, back=False):
        idx = np.flatnonzero(vals == v)
        return 0 if len(idx) == 0 else int(idx[-1 if back else 0]) + 1
    if nsplit > split_times_max:
        return False, -1, -1
    i1 = findloc(tags[:, 0], origin)
    if i1 <= 0:
        return False, -1, -1
    i2 = findloc(tags[:, 0], origin, True)
    i3 = findloc(tags[i1 - 1:i2, 1], abs(tag_inj))
    if i3 <= 0:
        return False, -1, -1
    i4 = findloc(tags[i1 - 1:i2, 1], abs(tag_inj), True)
    i3, i4 = i3 + i1 - 1, i4 + i1 - 1
    if nsplit > 0:
        i5 = findloc(tags[i3 - 1:i4, nsplit + 1], abs(tag_spl))
        if i5 <= 0:
            return False, -1, -1
        i6 = findloc(tags[i3 - 1:i4, nsplit + 1], abs(tag_spl), True)
        return True, i5 + i3 - 1, i6 + i3 - 1
    return True, i3, i4


def test_split_of_tracked_particles_matches_python_restatement():
    """split_particle's tracking branch (particle_module.f90:5452-5473): the child of a tracked particle gets
    tag_splitted - 2**(split_times - 1); parent and child each keep their negative tag only while the tag
    table still lists them at their new split level, and are sampled into particles_tracked when
    nsteps_pushed == 0."""
    w, P, _, _ = make_case("c1", grid=16, nptl=8, conf=dict(dt_min_rel=1e-2))
    from helpers import tracked_split_population
    tags, ptl = tracked_split_population(P)
    o = Oracle(P, 32)
    o.init_tracking(tags, 10)
    o.upload_particles(ptl)
    o.split(2.0, 2.0, 10)
    got = o.download_particles()
    rec = o.download_tracked()
    # ---- restatement ----
    split_times_max = tags.shape[1] - 2
    out = [dict((n, ptl[n][i].item()) for n in PARTICLE_DTYPE.names) for i in range(len(ptl))]
    want_rec = {}
    for i in range(len(ptl)):
        q = dict(out[i])
        if not (q["p"] > 2.0 * P.p0 * 2.0 ** q["split_times"] and q["p"] <= P.pmax):
            continue
        q["weight"] = float(np.float32(0.5) ** np.float32(1.0 + q["split_times"]))
        q["split_times"] += 1
        child = dict(q)
        if q["tag_splitted"] < 0:
            child["tag_splitted"] = q["tag_splitted"] - 2 ** (q["split_times"] - 1)
            for who in (child, q):
                sel, lo, hi = _py_is_selected(tags, split_times_max, who["origin"], who["tag_injected"],
                                              who["tag_splitted"], who["split_times"])
                if sel:
                    if q["nsteps_pushed"] == 0:
                        who["nsteps_tracked"] += 1
                        for c in range(lo, hi + 1):
                            want_rec[(c - 1, who["nsteps_tracked"] - 1)] = dict(who)
                else:
                    who["tag_splitted"] = -who["tag_splitted"]
        else:
            child["tag_splitted"] = q["tag_splitted"] + 2 ** (q["split_times"] - 1)
        out.append(child)
        out[i] = q
    assert len(got) == len(out) == 11
    for name in ("origin", "tag_injected", "tag_splitted", "split_times", "weight", "p", "x", "nsteps_tracked"):
        assert [g.item() for g in got[name]] == [q[name] for q in out], name
    # who is still tracked: A and its child, B only (its child is not in the table), D's child only
    assert [q["tag_splitted"] for q in out] == [-1, -1, 1, 1, 1, -1, -2, 2, 2, -3, 2]
    filled = {(r, c) for r in range(rec.shape[0]) for c in range(rec.shape[1]) if rec[r, c]["tag_splitted"] != 0}
    assert filled == set(want_rec)
    for (r, c), q in want_rec.items():
        for name in ("tag_injected", "tag_splitted", "split_times", "nsteps_tracked", "x", "weight"):
            assert rec[r, c][name].item() == q[name], (r, c, name)
