"""GPU parity tests: libgpat_cuda.so (through the C ABI) against the CPU oracle.

Bars (north_star): with identical random increments trajectories agree to 1e-12 relative in
FP64; integer/index work (histogram counts, tags, compaction order, RNG counters) bit-exact.
`strict_math=1` selects the no-contraction build of the push kernel, whose only arithmetic
difference from the oracle is CUDA's libdevice pow/log10 vs glibc's (<= 2 ulp); the fast build
(FMA contraction, fused time blend) is held to the same 1e-12 per step.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_particles_close, assert_particles_identical, box_of, make_case, rel_err, sort_by_key
from oracle.oracle import Oracle
from stochastic_parker_b200 import GpatSim, run_intervals
from stochastic_parker_b200.abi import PARTICLE_DTYPE, rng_steps

pytestmark = pytest.mark.gpu

STEP_RTOL = 1e-12    # one push, FP64 (north_star)
FRAME_RTOL = 1e-9    # a whole MHD interval: ~1e3 dependent steps, errors compound


def pair(P, nptl_max, strict=1):
    Pg = P.copy()
    Pg.strict_math = strict
    return GpatSim(Pg, nptl_max), Oracle(P, nptl_max)


def surfaces_of(P):
    """The synthetic acceleration surfaces of a case, as run_intervals' `surfaces` callable."""
    from stochastic_parker_b200 import mhd
    return lambda which, frame: mhd.make_acc_surface(P, which, frame)


def load_fields(sims, frames, time_interp=True):
    for s in sims:
        s.upload_fields(0, frames[0])
        if time_interp:
            s.upload_fields(1, frames[1])
        if s.P.acc_by_surface:
            for k in range(2 if s.P.surface2_existed else 1):
                s.upload_acc_surface(k, 0, surfaces_of(s.P)(k, 0))
                if time_interp:
                    s.upload_acc_surface(k, 1, surfaces_of(s.P)(k, 1))


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key,grid", [("c1", 48), ("c5", 32)])
def test_gradients_bit_exact(key, grid):
    """calc_fields_gradients (mhd_data_parallel.f90:533-566): all 24 FP32 gradients, every
    grid point including the one-sided ghost edges, bit for bit."""
    w, P, frames, _ = make_case(key, grid=grid, nptl=8)
    g, o = pair(P, w.nptl_max)
    o.upload_fields(0, frames[0])
    ref = o.get_fields(0).reshape(-1, 32)
    got = g.debug_gradients(frames[0]).reshape(-1, 32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    g.close()


@pytest.mark.parametrize("key,grid,cli", [("c1", 48, {}), ("c4", 64, {}), ("c5", 32, {}),
                                          ("c5", 32, dict(dpp_wave=1, dpp_shear=1))])
@pytest.mark.parametrize("strict", [1, 0])
def test_interp_parity(key, grid, cli, strict):
    """get_interp_paramters + interp_fields incl. the time blend, on every slot the pusher
    reads (strict: identical operations -> identical bits)."""
    w, P, frames, _ = make_case(key, grid=grid, nptl=8, cli=cli)
    g, o = pair(P, w.nptl_max, strict)
    load_fields((g, o), frames)
    rng = np.random.default_rng(1)
    n = 4000
    x = rng.uniform(P.xmin - 0.5 * P.dx, P.xmax + 0.5 * P.dx, n)
    y = rng.uniform(P.ymin - 0.5 * P.dy, P.ymax + 0.5 * P.dy, n)
    z = rng.uniform(P.zmin - 0.5 * P.dz, P.zmax + 0.5 * P.dz, n) if P.ndim == 3 else rng.uniform(0, 1, n)
    rt = rng.uniform(0, 1, n)
    ref = o.interp(x, y, z, rt)
    got = g.interp(x, y, z, rt)
    used = np.any(got != 0.0, axis=0)
    assert used.sum() >= 15
    if strict:
        assert np.array_equal(got[:, used], ref[:, used])
    else:
        scale = np.maximum(np.abs(ref[:, used]).max(axis=0), 1e-30)
        assert np.max(np.abs(got[:, used] - ref[:, used]) / scale) < 1e-14
    g.close()


@pytest.mark.parametrize("dist_flag", [1, 0, 2])
def test_inject_parity(dist_flag):
    """inject_particles_spatial_uniform + inject_one_particle: same Philox stream, same
    arithmetic -> bit-exact for the delta distribution; exp/pow ulps for the others."""
    w, P, frames, _ = make_case("c1", grid=32, nptl=3000)
    g, o = pair(P, 4000)
    for s in (g, o):
        s.inject_uniform(3000, 1e-4, dist_flag, w.particle_v0, 0.3, 0.1, box_of(P), 6.2)
        s.inject_uniform(1500, 1e-4, dist_flag, w.particle_v0, 0.4, 0.1, box_of(P), 6.2)  # hits capacity
    a, b = g.download_particles(), o.download_particles()
    assert len(a) == len(b) == 4000
    cg, co = g.counters(), o.counters()
    assert (cg.nptl_current, cg.tag_max) == (co.nptl_current, co.tag_max) == (4000, 4500)
    if dist_flag == 1:
        assert_particles_identical(a, b, 'inject')
    else:
        assert_particles_close(a, b, 1e-13, f"inject dist_flag={dist_flag}", frac_outliers=0.002)
    g.close()


def test_1d_gradients_and_interp_bit_exact():
    """1-D fields (farray(:, -1:nx+2, 1, 1), mhd_data_parallel.f90:78): d/dx gradients incl. the
    one-sided ends, y/z gradients left at the zero fill, and the 2-corner interpolation of
    get_interp_paramters' 1-D branch (particle_module.f90:649-652), bit for bit."""
    w, P, frames, _ = make_case("s1", grid=192, nptl=8, cli=dict(dpp_wave=1))
    g, o = pair(P, w.nptl_max)
    o.upload_fields(0, frames[0])
    ref = o.get_fields(0).reshape(-1, 32)
    got = g.debug_gradients(frames[0]).reshape(-1, 32)
    assert got.shape == (196, 32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    load_fields((g, o), frames)
    rng = np.random.default_rng(2)
    n = 3000
    x = rng.uniform(P.xmin - 0.5 * P.dx, P.xmax + 0.5 * P.dx, n)
    y, z, rt = rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    refi, goti = o.interp(x, y, z, rt), g.interp(x, y, z, rt)
    used = np.any(goti != 0.0, axis=0)
    assert used[[0, 3, 4, 5, 6, 8]].all()  # vx rho bx by bz dvx_dx
    assert np.array_equal(goti[:, used], refi[:, used])
    g.close()


from helpers import CASES  # noqa: E402  (shared with the reference-golden tests)


def _inject(sims, w, P, n, t0=0.0, dist_flag=1):
    for s in sims:
        s.inject_uniform(n, 0.0, dist_flag, w.particle_v0, t0, w.dt_out, box_of(P), w.power_index)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("strict", [1, 0])
def test_step_parity(name, strict):
    """One and then 40 consecutive calls of push_particle_* per particle (BC test, gather,
    kappa, drift, adaptive dt, stochastic step, momentum floor) with identical uniforms."""
    w, P, frames, _ = make_case(**CASES[name], nptl=512)
    g, o = pair(P, w.nptl_max, strict)
    load_fields((g, o), frames, P.time_interp)
    _inject((g, o), w, P, 512, dist_flag=0)
    for nsteps in (1, 40):
        sg = g.debug_push_n(0.0, w.dt_out, nsteps)
        so = o.debug_push_n(0.0, w.dt_out, nsteps)
        assert sg == so
        a, b = g.download_particles(), o.download_particles()
        # a single flipped branch (1-ulp pow difference at a min()/BC edge) is tolerated on
        # at most 1 particle in 500; everything else must sit inside 1e-12
        assert_particles_close(a, b, STEP_RTOL * max(1, nsteps // 4), f"{name} {nsteps} steps strict={strict}",
                               frac_outliers=0.002)
    g.close()


@pytest.mark.parametrize("name", ["c1_2d", "c3_shock_open", "c4_dpp_wave_shear", "c5_3d"])
def test_table_rng_parity(name):
    """north_star wording: 'fed the same pre-generated random increments' -- both sides replay
    one table of uniforms instead of generating them."""
    from stochastic_parker_b200.abi import RNG_TABLE
    w, P, frames, _ = make_case(**CASES[name], nptl=256)
    P.rng_mode = RNG_TABLE
    g, o = pair(P, w.nptl_max, 1)
    u = np.random.default_rng(7).uniform(0, 1, (256, 32, 4))
    g.set_rng_table(u)
    o.set_rng_table(u)
    load_fields((g, o), frames, P.time_interp)
    _inject((g, o), w, P, 256)
    assert g.debug_push_n(0.0, w.dt_out, 30) == o.debug_push_n(0.0, w.dt_out, 30)
    assert_particles_close(g.download_particles(), o.download_particles(), 1e-11, name, frac_outliers=0.004)
    g.close()


@pytest.mark.parametrize("name", sorted(CASES))
def test_interval_parity_strict(name):
    """particle_mover over two MHD intervals (adaptive steps, end-of-interval roll-back +
    fixed-dt re-push, both remove_particles passes, split, field swap), strict build."""
    w, P, frames, ts = make_case(**CASES[name], nptl=400)
    g, o = pair(P, w.nptl_max, 1)
    kw = dict(nptl=400, dist_flag=1, particle_v0=w.particle_v0, inject_new_ptl=True, split_flag=1,
              pmin_split=1.05, split_ratio=1.05, num_fine_steps=2, dump_escaped_dist=True, surfaces=surfaces_of(P))
    rg, sg = run_intervals(g, frames, ts, **kw)
    ro, so = run_intervals(o, frames, ts, **kw)
    a, b = g.download_particles(), o.download_particles()
    assert abs(sg - so) <= 2e-3 * so, (sg, so)
    assert abs(len(a) - len(b)) <= 2
    if len(a) == len(b):
        # same ORDER as the reference's swap-with-tail remove and append-at-tail split
        same_order = all(np.array_equal(a[f], b[f]) for f in ("origin", "tag_injected", "tag_splitted"))
        assert same_order
        assert_particles_close(a, b, FRAME_RTOL, name, int_exact=False, frac_outliers=0.01)
    cg, co = g.counters(), o.counters()
    assert abs(cg.nptl_escaped - co.nptl_escaped) <= 1
    assert rel_err(cg.leak, co.leak) < 1e-2 or abs(cg.leak - co.leak) <= 1.0
    g.close()


SURF_CASE = dict(conf=dict(acc_region_flag=1), cli=dict(acc_by_surface=1, surface_norm1="+z", surface2_existed=1,
                                                        surface_norm2="-y", is_intersection=1))


@pytest.mark.parametrize("key,grid,extra", [("c1", 64, {}),    # L2B, switch-specialised (mag 1, mom 1)
                                            ("c3", 64, {}),    # L2B, specialised (mag 0, mom 1), open x
                                            ("c4", 64, {}),    # L2E (D_pp), specialised
                                            ("c5", 32, {}),    # L3B, generic kernel, L2-sized residency
                                            ("c5", 32, SURF_CASE)])  # L3B + the acceleration-surface gate
def test_interval_parity_fast(key, grid, extra):
    """The production (fast-math) build over two intervals against the oracle: these are the
    kernels bench.py times (particle_mover picks the specialised instantiations; the per-step
    tests above run the generic ones through gpat_debug_push_n)."""
    w, P, frames, ts = make_case(key, grid=grid, nptl=2000, **extra)
    g, o = pair(P, w.nptl_max, 0)
    kw = dict(nptl=2000, dist_flag=1, particle_v0=w.particle_v0, split_flag=1, surfaces=surfaces_of(P))
    rg, sg = run_intervals(g, frames, ts, **kw)
    ro, so = run_intervals(o, frames, ts, **kw)
    assert abs(sg - so) <= 1e-3 * so
    a, b = sort_by_key(g.download_particles()), sort_by_key(o.download_particles())
    if len(a) != len(b):   # an escape decided by the last bits of a position (open boundaries)
        assert abs(len(a) - len(b)) <= 2
        keys = lambda q: set(zip(q["origin"].tolist(), q["tag_injected"].tolist(), q["tag_splitted"].tolist()))
        both = keys(a) & keys(b)
        pick = lambda q: q[[k in both for k in zip(q["origin"].tolist(), q["tag_injected"].tolist(), q["tag_splitted"].tolist())]]
        a, b = pick(a), pick(b)
    assert_particles_close(a, b, FRAME_RTOL, f"fast {key}", int_exact=False, frac_outliers=0.01)
    # spectra: identical up to particles that sit within rounding of a bin edge
    for x, y in zip(rg, ro):
        assert np.abs(x["fglobal"] - y["fglobal"]).sum() <= 4.0
    g.close()


@pytest.mark.parametrize("name,nptl,maps", [
    ("s1_shock_1d", 2000, False),                 # push_particle_1d, 1-D locate / boundary branches
    ("s1_shock_1d_dpp_nlgc", 1000, False),        # 1-D on the extended record (two lanes per particle)
    ("c1_2d_focused_transport", 200, False),      # push_particle_2d_ft: roll-back of v and mu, ~1e4 steps per particle
    ("c5_3d_ft", 300, False),                     # push_particle_3d_ft (two resident CTAs per SM, 255 registers)
    ("c4_2d_focused_transport_dpp", 200, False),  # focused transport + D_pp + NLGC
    ("c1_2d", 1000, True),                        # 2-D Parker + deltab / correlation maps (lane-group map gather)
    ("c5_3d", 400, True),                         # 3-D Parker + maps (d/dz of the maps)
])
def test_interval_parity_general_pushers_production_build(name, nptl, maps):
    """particle_mover over two MHD intervals for the paths outside the five named configs, PRODUCTION build: the
    kSpecAlt / kSpecAltMaps instantiations of push_kernel_coop run their whole state machine here (fine steps,
    roll-back + fixed-dt re-push with the focused-transport v / mu, the 1-D boundary test, split, field and map swap),
    where the per-step tests above drive them through gpat_debug_push_n only.  Tolerance: FRAME_RTOL (1e-9) on 99 % of the
    particles that both sides still hold, like test_interval_parity_fast; focused transport takes ~1e4 dependent steps
    per particle with clamps on mu: 98 % (measured, scripts/r02/ft_dev.py: identical step totals and spectra, 1-3 of 400
    particles at 1e-8 in mu or x)."""
    from stochastic_parker_b200 import mhd
    w, P, frames, ts = make_case(**CASES[name], nptl=nptl)
    if maps:
        P.deltab_flag = 1
        P.correlation_flag = 1
    g, o = pair(P, w.nptl_max, 0)
    mp = (lambda which, f: mhd.make_turbulence_maps(P.nx, P.ny, P.nz, f, ndim=P.ndim)[2 * which:2 * which + 2]) if maps else None
    kw = dict(nptl=nptl, dist_flag=1, particle_v0=w.particle_v0, split_flag=1, num_fine_steps=2, maps=mp)
    rg, sg = run_intervals(g, frames, ts, **kw)
    ro, so = run_intervals(o, frames, ts, **kw)
    ft = bool(P.focused_transport)
    assert abs(sg - so) <= 1e-3 * so, (sg, so)
    a, b = sort_by_key(g.download_particles()), sort_by_key(o.download_particles())
    if len(a) != len(b):   # an escape or a split decided by the last bits of a position / momentum
        assert abs(len(a) - len(b)) <= max(2, len(b) // 100)
        keys = lambda q: set(zip(q["origin"].tolist(), q["tag_injected"].tolist(), q["tag_splitted"].tolist()))
        both = keys(a) & keys(b)
        pick = lambda q: q[[k in both for k in zip(q["origin"].tolist(), q["tag_injected"].tolist(), q["tag_splitted"].tolist())]]
        a, b = pick(a), pick(b)
    assert_particles_close(a, b, FRAME_RTOL, f"production {name} maps={maps}", int_exact=False,
                           frac_outliers=0.02 if ft else 0.01)
    for x, y in zip(rg, ro):   # spectra: identical up to particles within rounding of a bin edge
        assert np.abs(x["fglobal"] - y["fglobal"]).sum() <= 4.0
    g.close()


@pytest.mark.parametrize("name", ["c1_2d", "c3_shock_open", "c5_3d", "s1_shock_1d", "c1_2d_focused_transport"])
def test_histograms_bit_exact(name):
    """calc_particle_distributions + quick_check + get_pmax_global on the SAME particle set:
    every histogram count bit-exact (dyadic weights -> order-independent FP64 sums)."""
    w, P, frames, ts = make_case(**CASES[name], nptl=3000)
    g, o = pair(P, w.nptl_max, 1)
    run_intervals(o, frames, ts, nptl=3000, particle_v0=w.particle_v0, pmin_split=1.02, split_ratio=1.02)
    ptl = o.download_particles()
    assert len(ptl) > 0 and len(np.unique(ptl["weight"])) > 1
    g.upload_particles(ptl)
    c = o.counters()
    g.set_counters(c)
    dg, do = g.diagnostics(True), o.diagnostics(True)
    assert np.array_equal(dg["fglobal"], do["fglobal"])
    assert dg["fglobal"].sum() > 0
    for k in range(4):
        if do["flocal"][k] is None:
            assert dg["flocal"][k] is None
            continue
        assert np.array_equal(dg["flocal"][k], do["flocal"][k]), f"flocal{k + 1}"
    assert np.array_equal(dg["quick"][[0, 1, 2, 3, 4, 6, 7]], do["quick"][[0, 1, 2, 3, 4, 6, 7]])
    assert rel_err(dg["quick"][5], do["quick"][5]) < 1e-12  # sum(dt): not dyadic, order-dependent
    assert dg["pmax"] == do["pmax"]
    pe_g, me_g = g.hist_edges(0)
    pe_o, me_o = o.hist_edges(0)
    assert np.array_equal(pe_g, pe_o) and np.array_equal(me_g, me_o)
    g.close()


def test_split_and_remove_exact():
    """split_particle (append order, tags, weights, capacity stop) and remove_particles
    (swap-with-tail order) reproduce the serial reference bit for bit."""
    w, P, frames, ts = make_case("c1", grid=32, nptl=64)
    nmax = 5000
    g, o = pair(P, nmax, 1)
    rng = np.random.default_rng(3)
    n = 4000
    ptl = np.zeros(n, dtype=PARTICLE_DTYPE)
    ptl["x"] = rng.uniform(P.xmin, P.xmax, n)
    ptl["y"] = rng.uniform(P.ymin, P.ymax, n)
    ptl["z"] = rng.uniform(0, 1, n)
    ptl["p"] = P.p0 * 10 ** rng.uniform(-0.3, 1.5, n)
    ptl["weight"] = 0.5 ** rng.integers(0, 4, n)
    ptl["split_times"] = rng.integers(0, 4, n)
    ptl["count_flag"] = rng.choice([1, 1, 1, 0, -1, -2, -4], n)
    ptl["tag_injected"] = np.arange(n)
    ptl["tag_splitted"] = 1
    ptl["dt"] = 1e-4
    ptl["mu"] = rng.uniform(-0.9, 0.9, n)
    rng_steps(ptl)[:] = rng.integers(0, 1000, n).astype(np.uint64)
    for s in (g, o):
        s.upload_particles(ptl)
    # two consecutive splits: the second one runs into nptl_max
    for it in range(2):
        g.split(2.0, 2.0)
        o.split(2.0, 2.0)
        a, b = g.download_particles(), o.download_particles()
        assert len(a) == len(b)
        assert_particles_identical(a, b, f"split pass {it}")
        assert g.counters().nptl_split == o.counters().nptl_split
    assert g.counters().nptl_current == nmax
    # remove via a mover call over an interval that is already over for every particle
    load_fields((g, o), frames)
    for s in (g, o):
        q = s.download_particles()
        q["t"] = 0.1
        s.upload_particles(q)
    assert g.particle_mover(0.0, 0.1, 100, 1, 1) == o.particle_mover(0.0, 0.1, 100, 1, 1) == 0
    a, b = g.download_particles(), o.download_particles()
    assert len(a) == len(b) and len(a) < nmax
    assert_particles_identical(a, b, 'remove')
    cg, co = g.counters(), o.counters()
    assert (cg.nptl_escaped, cg.leak, cg.leak_negp) == (co.nptl_escaped, co.leak, co.leak_negp)
    ea, eb = sort_by_key(g.download_escaped()), sort_by_key(o.download_escaped())
    assert_particles_identical(ea, eb, 'escaped')
    assert np.array_equal(g.escaped_diagnostics(), o.escaped_diagnostics())
    g.close()


@pytest.mark.parametrize("key,grid,conf", [("c2", 64, {}), ("c3", 64, {}),
                                           ("c5", 32, dict(pbcx=1, pbcy=1, pbcz=1))])
def test_local_escaped_distributions_bit_exact(key, grid, conf):
    """calc_escaped_distributions, local part (diagnostics.f90:956-1170) on the SAME escaped set: the face
    arrays of every enabled local set equal the oracle's bit for bit; an empty escaped list gives zeros."""
    w, P, frames, ts = make_case(key, grid=grid, nptl=4000, conf=conf)
    g, o = pair(P, w.nptl_max, 1)
    load_fields((g, o), frames, P.time_interp)
    empty = g.escaped_local_diagnostics()
    assert all(e is None or all(v is None or not v.any() for v in e.values()) for e in empty)
    _inject((g, o), w, P, 4000, dist_flag=2)
    o.particle_mover(0.0, w.dt_out, 100, 1, 1)
    esc = o.download_escaped()
    assert len(esc) >= 5
    # same escaped particles on both sides: push on the oracle, then hand the GPU the pre-push population
    # and let it produce its own escapees; compare through the particles that both sides lost
    g.particle_mover(0.0, w.dt_out, 100, 1, 1)
    ge = g.download_escaped()
    assert abs(len(ge) - len(esc)) <= 2
    a, b = g.escaped_local_diagnostics(), o.escaped_local_diagnostics()
    if len(ge) == len(esc) and np.array_equal(sort_by_key(ge)["count_flag"], sort_by_key(esc)["count_flag"]):
        for k in range(4):
            assert (a[k] is None) == (b[k] is None)
            if a[k] is None:
                continue
            for f in "xyz":
                assert (a[k][f] is None) == (b[k][f] is None)
                if a[k][f] is not None:
                    # positions agree to 1e-9 after an interval: a bin-edge case may move one count
                    assert np.abs(a[k][f] - b[k][f]).sum() <= 2.0, (k, f)
    # and the binning itself, bit for bit, on the GPU's own escapees re-binned by the oracle's routine
    o2 = Oracle(P, w.nptl_max)
    o2.upload_particles(np.zeros(0, dtype=PARTICLE_DTYPE))
    o2.lib.orc_set_escaped(o2.h, ge.ctypes.data_as(C.c_void_p), C.c_int64(len(ge)))
    c = o2.escaped_local_diagnostics()
    filled = 0.0
    for k in range(4):
        if a[k] is None:
            continue
        for f in "xyz":
            if a[k][f] is not None:
                assert np.array_equal(a[k][f], c[k][f]), (k, f)
                filled += a[k][f].sum()
    assert filled > 0.0     # this (unconditional) comparison is the bit-exact claim; it must not be vacuous
    g.close()


def test_edge_cases():
    """Empty population, zero-particle injection, mover before fields, bad parameters."""
    from stochastic_parker_b200 import GpatError
    w, P, frames, ts = make_case("c1", grid=32, nptl=16)
    g = GpatSim(P, 64)
    with pytest.raises(GpatError):
        g.particle_mover(0.0, 0.1)  # fields not uploaded: GPAT_ERR_STATE
    load_fields((g,), frames)
    assert g.particle_mover(0.0, 0.1) == 0
    g.inject_uniform(0, 0.0, 1, 1.0, 0.0, 0.1, box_of(P), 6.2)
    g.split(2.0, 2.0)
    d = g.diagnostics(True)
    assert d["fglobal"].sum() == 0 and d["quick"][0] == 0 and d["quick"][6] == 1.0 and d["pmax"] == 0.0
    assert len(g.download_particles()) == 0
    with pytest.raises((GpatError, ValueError)):
        g.upload_fields(0, frames[0][:-1])
    with pytest.raises(GpatError):
        g.inject_uniform(4, 0.0, 7, 1.0, 0.0, 0.1, box_of(P), 6.2)
    bad = P.copy()
    bad.focused_transport = 1       # the five-uniform FT pushers cannot replay a four-column table
    bad.include_3rd_dim = 1
    bad.rng_mode = 1
    with pytest.raises(GpatError):
        GpatSim(bad, 64)
    bad = P.copy()
    bad.spherical_coord = 1
    with pytest.raises(GpatError):
        GpatSim(bad, 64)
    bad = P.copy()
    bad.acc_by_surface = 1          # only the 3-D pushers look at the surfaces
    bad.surface_norm1 = 2
    with pytest.raises(GpatError):
        GpatSim(bad, 64)
    w3, P3, frames3, _ = make_case("c5", grid=16, nptl=16, conf=dict(acc_region_flag=1, r1=4, r2=8, r3=8),
                                   cli=dict(acc_by_surface=1, surface_norm1="+z"))
    g3 = GpatSim(P3, 64)
    g3.upload_fields(0, frames3[0])
    g3.upload_fields(1, frames3[1])
    g3.inject_uniform(8, 0.0, 1, 1.0, 0.0, 0.1, box_of(P3), 6.2)
    with pytest.raises(GpatError):
        g3.particle_mover(0.0, 0.1)  # surfaces not uploaded: GPAT_ERR_STATE
    with pytest.raises(GpatError):
        g3.upload_acc_surface(1, 0, np.zeros(g3.surface_shape(0)))  # surface2_existed is false
    g3.close()
    bad = P.copy()
    bad.local[0].rx = 5  # does not divide nx: check_local_dist_configuration
    with pytest.raises(GpatError):
        GpatSim(bad, 64)
    g.close()


# ------------------------------------------------------------------------------------------
# BASELINE.json's full sizes: size-independent properties + a sub-population against the oracle
def test_full_size_c1_properties():
    """C1 at its full shape (1024^2 field, 1e6 particles), production build, two intervals with
    splitting: integer bookkeeping exact, weights conserved, everybody on the frame time, and a
    2000-particle sub-population re-run on the oracle (same Philox streams) agrees."""
    from stochastic_parker_b200 import WORKLOADS, config, mhd
    w = WORKLOADS["c1"].scaled()
    n = 1_000_000
    cfg = mhd.mhd_config(w.nx, w.ny, w.nz, w.lx, w.ly, w.lz, w.dt_out, w.ndim)
    P = config.build_params(w.conf_text(), cfg, w.ndim, nframes=200, cli=w.cli)
    frames = [mhd.make_frame(w.kind, w.nx, w.ny, w.nz, f, w.dt_out) for f in range(3)]
    g = GpatSim(P, 2 * n)
    g.upload_fields(0, frames[0])
    g.upload_fields(1, frames[1])
    g.inject_uniform(n, 0.0, 1, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    start = g.download_particles()
    steps1 = g.particle_mover(0.0, w.dt_out, 100, 1, 0)
    a = g.download_particles()
    assert len(a) == n and steps1 > 100 * n
    assert int(rng_steps(a).sum()) == steps1            # one Philox block per push, nothing else
    assert np.all(a["count_flag"] == 1) and a["weight"].sum() == float(n)
    assert np.max(np.abs(a["t"] - w.dt_out)) < 1e-13
    assert a["x"].min() >= P.xmin and a["x"].max() <= P.xmax and a["y"].min() >= P.ymin and a["y"].max() <= P.ymax
    assert np.array_equal(np.sort(a["tag_injected"]), np.arange(n))
    # sub-population on the oracle
    sel = np.sort(np.random.default_rng(0).choice(n, 2000, replace=False))
    o = Oracle(P, 4096)
    o.upload_fields(0, frames[0])
    o.upload_fields(1, frames[1])
    o.upload_particles(start[sel])
    so = o.particle_mover(0.0, w.dt_out, 100, 1, 0)
    b = sort_by_key(o.download_particles())
    ga = sort_by_key(a[np.isin(a["tag_injected"], start["tag_injected"][sel])])
    assert abs(int(rng_steps(ga).sum()) - so) <= 2e-3 * so
    assert_particles_close(ga, b, FRAME_RTOL, "c1 full size", int_exact=False, frac_outliers=0.01)
    # second interval with a low split threshold: weight is conserved by splitting, histograms add up
    g.swap_fields()
    g.upload_fields(1, frames[2])
    g.particle_mover(w.dt_out, w.dt_out, 100, 1, 0)
    g.split(1.02, 1.02)
    c = g.counters()
    d = g.diagnostics(True)
    assert c.nptl_current == d["quick"][0] > n and c.nptl_split == c.nptl_current - n
    assert d["quick"][2] == float(n)                    # sum of weights, exact (dyadic)
    assert d["fglobal"].sum() == float(n)               # every particle is inside (pmin, pmax]
    for k in range(3):
        fl = d["flocal"][k]
        assert fl[..., -1, :].sum() == 0.0 and 0 < fl.sum() <= float(n)
    g.close()


def test_full_size_c4_and_c5_layouts_run():
    """The two other field-record layouts at a size where the field store no longer fits L2
    (C4: 2-D + D_pp, 24 slots; C5: 3-D, 24 slots): counters consistent, no particle lost."""
    from stochastic_parker_b200 import WORKLOADS, config, mhd
    for key, grid, n in (("c4", 2048, 400_000), ("c5", 192, 400_000)):
        w = WORKLOADS[key].scaled(grid=grid)
        cfg = mhd.mhd_config(w.nx, w.ny, w.nz, w.lx, w.ly, w.lz, w.dt_out, w.ndim)
        P = config.build_params(w.conf_text(), cfg, w.ndim, nframes=200, cli=w.cli)
        g = GpatSim(P, 2 * n)
        for s in (0, 1):
            g.upload_fields(s, mhd.make_frame(w.kind, w.nx, w.ny, w.nz, s, w.dt_out))
        g.inject_uniform(n, 0.0, 1, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
        steps = g.particle_mover(0.0, w.dt_out, 100, 1, 0)
        a = g.download_particles()
        assert len(a) == n and int(rng_steps(a).sum()) == steps > 10 * n
        assert np.max(np.abs(a["t"] - w.dt_out)) < 1e-13 and np.all(np.isfinite(a["p"])) and a["p"].min() >= 0.25 * P.p0
        g.close()


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_gpus_nccl_allreduce(tmp_path):
    """N = 2: two processes, one GPU each, the library's own NCCL all-reduce inside
    gpat_diagnostics.  Reduced histograms == histograms of the union of the shards, bit for bit."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(root, "tests", "multigpu_worker.py"), str(tmp_path), "20001"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    shards = [np.load(tmp_path / f"ptl_{k}.npy") for k in range(2)]
    red = [np.load(tmp_path / f"reduced_{k}.npz") for k in range(2)]
    w, P, frames, ts = make_case("c1", grid=64, nptl=20001)
    o = Oracle(P, 8 * 20001)
    o.upload_particles(np.concatenate(shards))
    d = o.diagnostics(True)
    for k in range(2):  # all-reduce: every rank holds the reduced arrays
        assert np.array_equal(red[k]["fglobal"], d["fglobal"])
        for j in range(3):
            assert np.array_equal(red[k][f"flocal{j}"], d["flocal"][j])
        assert red[k]["quick"][0] == len(shards[0]) + len(shards[1])
        assert red[k]["quick"][2] == d["quick"][2] and float(red[k]["pmax"]) == d["pmax"]
        assert red[k]["quick"][6] == d["quick"][6] and red[k]["quick"][7] == d["quick"][7]


def test_cpp_driver_matches_python_driver(tmp_path):
    """host/gpat_driver (C++ above the C ABI, the reference's switches and files) writes the same
    quick.dat numbers and bit-identical spectra as run_intervals through ctypes."""
    import os
    import subprocess
    from stochastic_parker_b200 import WORKLOADS, config, mhd
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "host")], check=True)
    w = WORKLOADS["c1"].scaled(grid=64, nptl=5000)
    nfr = 4
    cfg = mhd.write_run(str(tmp_path / "mhd"), w.kind, w.nx, w.ny, w.nz, nframes=nfr, lx=w.lx, ly=w.ly, lz=w.lz,
                        dt_out=w.dt_out)
    conf = tmp_path / "conf.dat"
    conf.write_text(w.conf_text())
    out = tmp_path / "out"
    out.mkdir()
    args = [os.path.join(root, "host", "gpat_driver"), "-nl", ".false.", "-pv", repr(w.particle_v0),
            "-dm", str(tmp_path / "mhd") + "/", "-np", "5000", "-ti", "1", "-ts", "0", "-te", str(nfr - 1),
            "-df", "1", "-pi", "6.2", "-sf", "1", "-sr", "1.05", "-ps", "1.05", "-ni", "100", "-dt", "0.0",
            "-dd", str(out) + "/", "-cf", str(conf), "-ld", ".true.", "-nm", "40000", "-in", ".true.",
            "-t0", "7.53877e-5", "-nd", "2", "-dp1", "850964.408", "-dp2", "13575468.975", "-ch", "-1"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    P = config.build_params(w.conf_text(), mhd.read_mhd_config(str(tmp_path / "mhd" / "mhd_config.dat")), 2,
                            nframes=nfr - 1, cli=w.cli)
    g = GpatSim(P, 40000)
    frames = [mhd.read_frame(str(tmp_path / "mhd"), f, dict(cfg, ndim=2)) for f in range(nfr)]
    rec, steps = run_intervals(g, frames, [f * w.dt_out for f in range(nfr)], nptl=5000, dist_flag=1,
                               particle_v0=w.particle_v0, power_index=6.2, split_ratio=1.05, pmin_split=1.05)
    assert f"Total particle steps: {steps} " in r.stdout
    for d in rec:
        raw = open(out / f"fdists_{d['frame']:04d}.bin", "rb").read()
        nmu, npp = np.frombuffer(raw[:8], dtype=np.int32)
        fg = np.frombuffer(raw[8:8 + 8 * nmu * npp], dtype=np.float64).reshape(npp, nmu)
        assert np.array_equal(fg, d["fglobal"])
        loc = open(out / f"fdists_local2_{d['frame']:04d}.bin", "rb").read()
        shp = np.frombuffer(loc[:20], dtype=np.int32)
        assert np.array_equal(np.frombuffer(loc[20:], dtype=np.float64).reshape(tuple(shp[::-1])), d["flocal"][1])
    rows = [l.split() for l in open(out / "quick.dat").read().splitlines()]
    assert rows[0][:3] == ["iframe", "nptl_current", "nptl_split"] and len(rows) == 1 + nfr
    last = open(out / "quick.dat").read().splitlines()[-1]
    assert last[:6] == f"{nfr - 1:06d}" and float(last[6:19]) == float(f"{rec[-1]['quick'][0]:.6E}")
    assert len(open(out / "pmax_global.dat").read().split()) == nfr
    g.close()


def test_cpp_driver_reads_acceleration_surfaces(tmp_path):
    """-as 1 -sn1 +z -s2e .true. -sn2 -y -ii .true. -sf1/-sf2: the C++ driver reads <name>_NNNN.dat like
    read_acc_surface and gives the run run_intervals gives with the same arrays."""
    import os
    import subprocess
    from stochastic_parker_b200 import WORKLOADS, config, mhd
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "host")], check=True)
    w = WORKLOADS["c5"].scaled(grid=32, nptl=3000)
    w.conf = dict(w.conf, acc_region_flag=1, r1=4, r2=8, r3=16)
    w.cli = dict(w.cli, acc_by_surface=1, surface_norm1="+z", surface2_existed=1, surface_norm2="-y", is_intersection=1)
    nfr = 3
    d = tmp_path / "mhd"
    cfg = mhd.write_run(str(d), w.kind, w.nx, w.ny, w.nz, nframes=nfr, lx=w.lx, ly=w.ly, lz=w.lz, dt_out=w.dt_out)
    P = config.build_params(w.conf_text(), mhd.read_mhd_config(str(d / "mhd_config.dat")), 3, nframes=nfr - 1, cli=w.cli)
    for f in range(nfr):
        for k, stem in enumerate(("surf_a", "surf_b")):
            mhd.make_acc_surface(P, k, f).tofile(str(d / f"{stem}_{f:04d}.dat"))
    conf = tmp_path / "conf.dat"
    conf.write_text(w.conf_text())
    out = tmp_path / "out"
    out.mkdir()
    args = [os.path.join(root, "host", "gpat_driver"), "-nl", ".false.", "-pv", repr(w.particle_v0),
            "-dm", str(d) + "/", "-np", "3000", "-ti", "1", "-ts", "0", "-te", str(nfr - 1),
            "-df", "1", "-pi", "6.2", "-sf", "1", "-sr", "1.05", "-ps", "1.05", "-ni", "100", "-dt", "0.0",
            "-dd", str(out) + "/", "-cf", str(conf), "-ld", ".false.", "-nm", "30000", "-in", ".true.",
            "-nd", "3", "-dp1", "850964.408", "-dp2", "13575468.975", "-ch", "-1",
            "-as", "1", "-sn1", "+z", "-s2e", ".true.", "-sn2", "-y", "-ii", ".true.", "-sf1", "surf_a", "-sf2", "surf_b"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    g = GpatSim(P, 30000)
    frames = [mhd.read_frame(str(d), f, dict(cfg, ndim=3)) for f in range(nfr)]
    rec, steps = run_intervals(g, frames, [f * w.dt_out for f in range(nfr)], nptl=3000, dist_flag=1,
                               particle_v0=w.particle_v0, power_index=6.2, split_ratio=1.05, pmin_split=1.05,
                               local_dist=False, surfaces=surfaces_of(P))
    assert f"Total particle steps: {steps} " in r.stdout
    for rd in rec:
        raw = open(out / f"fdists_{rd['frame']:04d}.bin", "rb").read()
        nmu, npp = np.frombuffer(raw[:8], dtype=np.int32)
        assert np.array_equal(np.frombuffer(raw[8:8 + 8 * nmu * npp], dtype=np.float64).reshape(npp, nmu), rd["fglobal"])
    g.close()
    # without the surface files the driver stops with an error instead of running ungated
    os.remove(d / "surf_b_0001.dat")
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0


def test_restart_round_trip_is_bit_exact():
    """dump_particles / read_particles + save/read_particle_module_state (diagnostics.f90:1811-1888,
    particle_module.f90:5532-5664, 5744-5816): a run that is stopped after one interval, downloaded
    to the AoS records of the restart file (Philox counters ride in `padding`), and resumed in a
    NEW handle continues bit-identically to the uninterrupted run."""
    w, P, frames, ts = make_case("c1", grid=64, nptl=3000, nframes=4)
    kw = dict(nptl=3000, dist_flag=2, particle_v0=w.particle_v0, pmin_split=1.05, split_ratio=1.05,
              inject_new_ptl=True)
    for strict in (1, 0):
        Pg = P.copy()
        Pg.strict_math = strict
        a = GpatSim(Pg, w.nptl_max * 4)
        run_intervals(a, frames, ts, **kw)
        ref = a.download_particles()
        ca = a.counters()
        a.close()
        b = GpatSim(Pg, w.nptl_max * 4)
        run_intervals(b, frames[:2], ts[:2], **kw)          # first interval only
        dump, cb = b.download_particles(), b.counters()
        b.close()
        c = GpatSim(Pg, w.nptl_max * 4)                     # "restart"
        c.upload_particles(dump)
        c.set_counters(cb)
        c.upload_fields(0, frames[1])
        for tf in (2, 3):                                    # the loop of run_intervals from frame 2 on
            c.upload_fields(1, frames[tf])
            c.inject_uniform(3000, 0.0, 2, w.particle_v0, ts[tf - 1], ts[tf] - ts[tf - 1], box_of(P), 6.2)
            c.particle_mover(ts[tf - 1], ts[tf] - ts[tf - 1], 100, 1, 0)
            c.split(1.05, 1.05)
            c.swap_fields()
        got, cc = c.download_particles(), c.counters()
        assert_particles_identical(got, ref, f"restart strict={strict}")
        assert (cc.nptl_current, cc.nptl_split, cc.tag_max, cc.leak, cc.leak_negp) == \
               (ca.nptl_current, ca.nptl_split, ca.tag_max, ca.leak, ca.leak_negp)
        c.close()


@pytest.mark.parametrize("key,grid,mode,vmin,same", [
    ("c1", 64, 1, None, True),      # inject_large_jz, every rank injects nptl
    ("c1", 64, 2, None, False),     # inject_large_absj, scaled by ncells / ncells_norm
    ("c3", 64, 4, None, True),      # inject_large_divv (compression at the shock)
    ("c4", 64, 5, None, True),      # inject_large_rho (density slot kept by the D_pp layout)
    ("c5", 32, 1, None, True),      # 3-D
])
def test_targeted_injection_parity(key, grid, mode, vmin, same):
    """inject_particles_at_large_jz/_absj/_divv/_rho + get_ncells_large_* (particle_module.f90:
    785-1468, mhd_data_parallel.f90:2211-2498): same cell count, same number of particles, and
    the same accepted positions from the same per-particle Philox streams, bit for bit."""
    w, P, frames, _ = make_case(key, grid=grid, nptl=8)
    g, o = pair(P, 6000)
    load_fields((g, o), frames, P.time_interp)
    box = box_of(P)
    box[0] += 0.1 * (P.xmax - P.xmin)       # a part_box smaller than the domain: misses are redrawn
    box[4] -= 0.15 * (P.ymax - P.ymin)
    # threshold = a value the interpolated criterion exceeds on roughly a third of the box
    fa = o.get_fields(0).reshape(-1, 32)
    crit = {1: np.abs(fa[:, 8 + 15] - fa[:, 8 + 13]),
            2: np.sqrt((fa[:, 8 + 17] - fa[:, 8 + 19]) ** 2 + (fa[:, 8 + 18] - fa[:, 8 + 14]) ** 2
                       + (fa[:, 8 + 13] - fa[:, 8 + 15]) ** 2),
            4: -(fa[:, 8] + fa[:, 8 + 4] + (fa[:, 8 + 8] if P.ndim == 3 else 0.0)),
            5: fa[:, 3]}[mode]
    vmin = float(np.quantile(crit, 0.7))
    norm = 3 * grid
    rg = g.inject_targeted(mode, 1500, 1e-4, 1, w.particle_v0, 0.0, 0.1, box, 6.2, same, vmin, norm)
    ro = o.inject_targeted(mode, 1500, 1e-4, 1, w.particle_v0, 0.0, 0.1, box, 6.2, same, vmin, norm)
    assert rg == ro and ro[0] > 0 and ro[1] > 0
    if same:
        assert ro[0] == 1500
    a, b = g.download_particles(), o.download_particles()
    assert_particles_identical(a, b, f"targeted injection mode {mode}")
    assert np.all((a["x"] >= box[0]) & (a["x"] <= box[3]) & (a["y"] >= box[1]) & (a["y"] <= box[4]))
    cg, co = g.counters(), o.counters()
    assert (cg.nptl_current, cg.tag_max) == (co.nptl_current, co.tag_max)
    g.close()


def test_targeted_injection_rejects_what_it_cannot_do():
    from stochastic_parker_b200 import GpatError
    w, P, frames, _ = make_case("c1", grid=32, nptl=8)
    g = GpatSim(P, 100)
    g.upload_fields(0, frames[0])
    with pytest.raises(GpatError):   # the base 2-D record has no density slot
        g.inject_targeted(5, 10, 0.0, 1, 1.0, 0.0, 0.1, box_of(P), 6.2, True, 0.5, 1)
    with pytest.raises(GpatError):   # deltab maps are outside the GPU path
        g.inject_targeted(3, 10, 0.0, 1, 1.0, 0.0, 0.1, box_of(P), 6.2, True, 0.5, 1)
    g.close()
    P.keep_rho = 1
    g = GpatSim(P, 100)
    g.upload_fields(0, frames[0])
    n, nc = g.inject_targeted(5, 10, 0.0, 1, 1.0, 0.0, 0.1, box_of(P), 6.2, True, 0.5, 1)
    assert n == 10 and nc > 0
    g.close()


def test_prefetched_frames_give_identical_runs():
    """gpat_prefetch_fields (frame pipeline): copying frame tf+1 on the copy stream while frame
    tf is pushed changes nothing in the results, bit for bit, in either build."""
    w, P, frames, ts = make_case("c1", grid=64, nptl=1500, nframes=4)
    frames = [np.ascontiguousarray(f, dtype=np.float32) for f in frames]
    for strict in (1, 0):
        Pg = P.copy()
        Pg.strict_math = strict
        outs = []
        for pipelined in (False, True):
            g = GpatSim(Pg, w.nptl_max)
            g.upload_fields(0, frames[0])
            for tf in (1, 2, 3):
                g.upload_fields(1, frames[tf])
                if pipelined and tf < 3:
                    g.prefetch_fields(frames[tf + 1])
                if tf == 1:
                    g.inject_uniform(1500, 0.0, 1, w.particle_v0, ts[0], ts[1] - ts[0], box_of(P), 6.2)
                g.particle_mover(ts[tf - 1], ts[tf] - ts[tf - 1], 100, 1, 0)
                g.split(2.0, 2.0)
                g.swap_fields()
            outs.append(g.download_particles())
            g.close()
        assert_particles_identical(outs[1], outs[0], f"prefetch strict={strict}")


@pytest.mark.parametrize("strict", [1, 0])
def test_tracking_run_replays_and_matches_the_oracle(strict):
    """Particle tracking (particle_module.f90:434-440, 1697-1724, 5452-5473, 5825-5990): the second
    run of the two-run workflow on the GPU.  (a) the tracking run reproduces the first GPU run bit
    for bit (Philox keyed by |tag|), in both builds; (b) tags, sample counts and sample slots of
    particles_tracked equal the oracle's, the sampled states agree like any trajectory does."""
    from test_cpu_oracle import _tracking_pair
    first, sel, tags, second, recs = _tracking_pair(lambda P, n: GpatSim(P, n), strict=strict)
    key = lambda q: (q["origin"], np.abs(q["tag_injected"]), np.abs(q["tag_splitted"]))
    oa, ob = np.lexsort(key(first)[::-1]), np.lexsort(key(second)[::-1])
    fa, fb = first[oa], second[ob]
    assert len(fa) == len(fb)
    for f in ("x", "y", "p", "t", "weight", "dt"):
        assert np.array_equal(fa[f], fb[f]), f
    assert np.count_nonzero(fb["tag_splitted"] < 0) == len(sel)
    assert len(recs) == 3 and recs[0].shape[0] == len(sel)
    if strict:
        _, _, tags_o, _, recs_o = _tracking_pair(Oracle)
        # selection from the oracle's own first run: same particles up to trajectories that sit
        # within rounding of a split threshold, so compare run against run where the tables agree
        if np.array_equal(tags, tags_o):
            for rg, ro in zip(recs, recs_o):
                used_g, used_o = rg["tag_splitted"] < 0, ro["tag_splitted"] < 0
                agree = (used_g == used_o).mean()
                assert agree > 0.99
                both = used_g & used_o
                assert np.array_equal(rg["tag_splitted"][both], ro["tag_splitted"][both])
                assert np.array_equal(rg["nsteps_tracked"][both], ro["nsteps_tracked"][both])
                err = rel_err(rg["x"][both], ro["x"][both])
                assert np.quantile(err, 0.99) < 1e-7
    for rec in recs:
        for row in rec:
            used = row["tag_splitted"] < 0
            assert np.all(np.diff(row["t"][used]) > 0)
            assert np.all(row["nsteps_pushed"][used] == 0)


def test_spectra_agree_within_poisson_error_for_independent_streams():
    """north_star's second bar: with on-device Philox, spectra and spatial distributions agree with
    the reference path within Poisson statistical error.  The production build draws from per-particle
    Philox streams, the oracle from one sequential MT19937 seeded with the reference's own seed array
    (the generator family behind mt_stream; nothing shared but the physics); unit weights
    (no splitting) make every bin count a Poisson variate, so chi^2/dof of the difference ~ 1."""
    w, P, frames, ts = make_case("c1", grid=64, nptl=30000, nframes=3)
    kw = dict(nptl=30000, dist_flag=0, particle_v0=w.particle_v0, inject_new_ptl=False, split_flag=0)
    Pg = P.copy()
    Pg.strict_math = 0
    Pg.seed = 0x1234ABCD
    g = GpatSim(Pg, w.nptl_max)
    rg, _ = run_intervals(g, frames, ts, **kw)
    g.close()
    Po = P.copy()
    Po.rng_mode = 2   # oracle-only: one sequential MT19937 seeded like random_number_generator.f90:16,38
    o = Oracle(Po, w.nptl_max)
    ro, _ = run_intervals(o, frames, ts, **kw)

    def chi2(a, b, min_count=25):
        a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
        m = (a + b) >= 2 * min_count
        return float(np.sum((a[m] - b[m]) ** 2 / (a[m] + b[m]))), int(m.sum())

    # momentum spectrum after two intervals
    c, dof = chi2(rg[-1]["fglobal"], ro[-1]["fglobal"])
    assert dof >= 10
    assert c / dof < 1.0 + 4.0 * np.sqrt(2.0 / dof), (c, dof)      # 4 sigma of a chi^2 with dof bins
    # spatial distribution: local set 1 summed over momentum, rebinned to 8 x 8
    # flocal1 arrives as (nrz, nry, nrx, npbins, nmu): sum the last two axes -> the spatial map
    sa = np.asarray(rg[-1]["flocal"][0]).sum(axis=(-1, -2))
    sb = np.asarray(ro[-1]["flocal"][0]).sum(axis=(-1, -2))
    c, dof = chi2(sa, sb)
    assert dof >= 20
    assert c / dof < 1.0 + 4.0 * np.sqrt(2.0 / dof), (c, dof)
    # and the totals: weight is conserved on both sides
    assert abs(rg[-1]["quick"][2] - ro[-1]["quick"][2]) <= 5.0 * np.sqrt(rg[-1]["quick"][2] + 1)


@pytest.mark.parametrize("key,grid,dist_flag", [("c3", 64, 0), ("c1", 48, 1), ("c5", 32, 2)])
def test_shock_injection_parity(key, grid, dist_flag):
    """locate_shock_xpos + inject_particles_at_shock (mhd_data_parallel.f90:1988-2045,
    particle_module.f90:542-633; `-is 1` of config/shock.sh): shock positions, interpolation with
    the reference's weights, momentum envelope and random stream, bit for bit."""
    w, P, frames, ts = make_case(key, grid=grid, nptl=8)
    g, o = pair(P, 5000)
    load_fields((g, o), frames, True)
    for s in (g, o):
        s.inject_at_shock(2000, 1e-5, dist_flag, w.particle_v0, ts[0], 6.2)
        s.inject_at_shock(1500, 1e-5, dist_flag, w.particle_v0, ts[1], 6.2)
    a, b = g.download_particles(), o.download_particles()
    assert len(a) == len(b) == 3500
    if dist_flag == 1:
        assert_particles_identical(a, b, "shock injection")
    else:
        assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"]) and np.array_equal(a["tag_injected"], b["tag_injected"])
        assert_particles_close(a, b, 1e-13, f"shock injection dist_flag={dist_flag}", frac_outliers=0.002)
    g.close()


def _with_maps(sims, P, frames_idx=(0, 1)):
    from stochastic_parker_b200 import mhd
    for slot, f in enumerate(frames_idx):
        m = mhd.make_turbulence_maps(P.nx, P.ny, P.nz, f, ndim=P.ndim)
        for s in sims:
            s.upload_turbulence(0, slot, m[0], m[1])
            s.upload_turbulence(1, slot, m[2], m[3])


@pytest.mark.parametrize("key,grid,cli", [
    ("c1", 64, {}),                                              # 2-D Parker
    ("c1", 64, dict(nlgc=1, kperp_kpara=0.05)),                  # NLGC: sigma2_2d and lc_2d enter kappa_perp
    ("c5", 32, {}),                                              # 3-D: d/dz of the maps
    ("c1", 64, dict(focused_transport=1, duu_init=5.0)),         # D_mumu normalisation
])
def test_turbulence_maps_step_parity(key, grid, cli):
    """deltab_flag + correlation_flag (particle_module.f90:2246-2254, 2314-2321, 2505-2517, 2589-2604,
    3143-3148; maps and gradients mhd_data_parallel.f90:306-497, 771-1604, 1806-1915)."""
    w, P, frames, _ = make_case(key, grid=grid, nptl=512, cli=cli, conf=dict(dt_min_rel=1e-4))
    P.deltab_flag = 1
    P.correlation_flag = 1
    g, o = pair(P, w.nptl_max, 0)   # the library routes map runs to the reference-order build itself
    load_fields((g, o), frames, True)
    with pytest.raises(Exception):   # maps not uploaded yet
        g.debug_push_n(0.0, w.dt_out, 1)
    _with_maps((g, o), P)
    _inject((g, o), w, P, 512, dist_flag=0)
    for nsteps in (1, 40):
        assert g.debug_push_n(0.0, w.dt_out, nsteps) == o.debug_push_n(0.0, w.dt_out, nsteps)
        assert_particles_close(g.download_particles(), o.download_particles(), STEP_RTOL * max(1, nsteps // 4),
                               f"maps {key} {cli} {nsteps} steps", frac_outliers=0.004)
    # and they matter: the same particles without the maps land elsewhere
    P0 = P.copy()
    P0.deltab_flag = P0.correlation_flag = 0
    o0 = Oracle(P0, w.nptl_max)
    load_fields((o0,), frames, True)
    _inject((o0,), w, P0, 512, dist_flag=0)
    o0.debug_push_n(0.0, w.dt_out, 1)
    o1 = Oracle(P, w.nptl_max)
    load_fields((o1,), frames, True)
    _with_maps((o1,), P)
    _inject((o1,), w, P, 512, dist_flag=0)
    o1.debug_push_n(0.0, w.dt_out, 1)
    assert np.max(np.abs(o0.download_particles()["x"] - o1.download_particles()["x"])) > 1e-9
    g.close()


def test_turbulence_maps_swap_and_db2_injection():
    """copy_magnetic_fluctuation via gpat_swap_fields + inject_particles_at_large_db2
    (particle_module.f90:1075-1236, mhd_data_parallel.f90:2343-2377), bit for bit."""
    from stochastic_parker_b200 import mhd
    w, P, frames, ts = make_case("c1", grid=64, nptl=8, nframes=3)
    P.deltab_flag = 1
    g, o = pair(P, 5000, 1)
    load_fields((g, o), frames, True)
    _with_maps((g, o), P)
    for s in (g, o):
        s.swap_fields()
        s.upload_fields(1, frames[2])
    m = mhd.make_turbulence_maps(P.nx, P.ny, 1, 2)
    for s in (g, o):
        s.upload_turbulence(0, 1, m[0], m[1])
    m1 = mhd.make_turbulence_maps(P.nx, P.ny, 1, 1)[0]     # farray1 is now frame 1
    vmin = float(np.quantile(m1, 0.7))
    rg = g.inject_targeted(3, 1200, 0.0, 1, w.particle_v0, ts[1], 0.1, box_of(P), 6.2, True, vmin, 1)
    ro = o.inject_targeted(3, 1200, 0.0, 1, w.particle_v0, ts[1], 0.1, box_of(P), 6.2, True, vmin, 1)
    assert rg == ro and ro[0] == 1200 and ro[1] > 0
    assert_particles_identical(g.download_particles(), o.download_particles(), "inject_large_db2")
    g.close()


@pytest.mark.parametrize("key,grid", [("c1", 64), ("c5", 32)])
def test_cell_sorted_particles_change_nothing_but_the_order(key, grid, monkeypatch):
    """csrc/sort.cu: the production build may sort the particle arrays by grid cell before a push.
    Every particle owns its random stream, so the sorted run must reproduce the unsorted one bit
    for bit, particle by particle (identified by origin / tag_injected / tag_splitted), with the
    same counters and histograms."""
    w, P, frames, ts = make_case(key, grid=grid, nptl=3000, nframes=4)
    Pg = P.copy()
    Pg.strict_math = 0
    kw = dict(nptl=3000, dist_flag=2, particle_v0=w.particle_v0, inject_new_ptl=True, split_flag=1,
              pmin_split=1.05, split_ratio=1.05)
    runs = []
    for sort in ("0", "1"):
        monkeypatch.setenv("GPAT_PUSH_SORT", sort)
        g = GpatSim(Pg, 8 * w.nptl_max)
        res, steps = run_intervals(g, frames, ts, **kw)
        runs.append((g.download_particles(), res, steps, g.counters()))
        g.close()
    (a, ra, sa, ca), (b, rb, sb, cb) = runs
    assert sa == sb and len(a) == len(b)
    assert not np.array_equal(a["tag_injected"], b["tag_injected"])      # the order did change
    assert_particles_identical(sort_by_key(a), sort_by_key(b), f"sorted vs unsorted {key}")
    assert (ca.nptl_current, ca.nptl_split, ca.tag_max, ca.leak) == (cb.nptl_current, cb.nptl_split, cb.tag_max, cb.leak)
    for x, y in zip(ra, rb):
        assert np.array_equal(x["fglobal"], y["fglobal"])
        for fx, fy in zip(x["flocal"], y["flocal"]):
            assert (fx is None and fy is None) or np.array_equal(fx, fy)


# ------------------------------------------------------------------------------------------
# the GPU library against golden vectors computed by the reference's own Fortran
# (tests/golden/ref_f90/, generated by tests/golden/make_ref_f90_golden.py; tests/test_cpu_reference_f90.py
# holds the C oracle to the same files bit for bit)
# ------------------------------------------------------------------------------------------
class _GpuForGolden(GpatSim):
    """GpatSim + get_fields(): the 32-slot store as the device computes it (gpat_debug_gradients)."""

    def upload_fields(self, slot, f, with_grad=0):
        self._last = getattr(self, "_last", {})
        self._last[slot] = np.ascontiguousarray(f, dtype=np.float32)
        super().upload_fields(slot, f, with_grad)

    def get_fields(self, slot):
        return self.debug_gradients(self._last[slot])


@pytest.mark.parametrize("name", __import__("helpers").REF_GOLDEN_CASES)
def test_gpu_against_reference_golden(name):
    """Reference-order build (strict_math = 1) against what the reference's Fortran computes: gradients,
    injection and every integer bit-exact; single pushes at north_star's 1e-12; two whole intervals at the
    interval bar; histogram counts within the few particles a 1-ulp libdevice pow can move across an edge."""
    import os
    from helpers import golden_collect, unpack_sparse
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_f90", name + ".npz"))
    got = golden_collect(lambda P, n: _GpuForGolden(P, n), name)
    assert np.array_equal(got["grad_sha256"], z["grad_sha256"]), "calc_fields_gradients differs from the reference"
    used = np.any(got["interp"] != 0.0, axis=0)
    # (a slot of a synthetic field can be identically zero, e.g. vz of the flare sheet: it then counts as unused)
    assert used.sum() >= (8 if name.startswith("s1") else 12) and np.array_equal(got["interp"][:, used], z["interp"][:, used])
    # injection order: 16 particles each with dist_flag 1 (delta: no transcendental), 0 (Maxwellian: exp), 2 (power law:
    # pow) -- the first group bit for bit, the others to the ulps of libdevice's exp / pow
    assert_particles_identical(got["inject"][:16], z["inject"][:16], name + " inject (delta)")
    assert_particles_close(got["inject"], z["inject"], 1e-13, name + " inject")
    assert int(got["steps1_count"]) == int(z["steps1_count"]) and int(got["steps41_count"]) == int(z["steps41_count"])
    assert_particles_close(got["steps1"], z["steps1"], STEP_RTOL, name + " 1 push")
    assert_particles_close(got["steps41"], z["steps41"], STEP_RTOL * 10, name + " 41 pushes", frac_outliers=0.03)
    sg, so = int(got["run_steps"]), int(z["run_steps"])
    assert abs(sg - so) <= 2e-3 * so + 2, (sg, so)
    a, b = got["run_particles"], z["run_particles"]
    assert abs(len(a) - len(b)) <= 2
    if len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in ("origin", "tag_injected", "tag_splitted")):
        assert_particles_close(a, b, FRAME_RTOL, name + " intervals", int_exact=False, frac_outliers=0.05)
    for k in z.files:
        if k.endswith("fglobal") or k.endswith("fescaped"):
            assert np.abs(got[k] - z[k]).sum() <= 4.0, k
        elif k.endswith(".val") and k[:-4] + ".shape" in got:
            assert np.abs(unpack_sparse(got, k[:-4]) - unpack_sparse(z, k[:-4])).sum() <= 4.0, k
        elif k.endswith("pmax"):
            assert rel_err(got[k], z[k]) < 1e-6, k


@pytest.mark.parametrize("key,grid,env,nslots,absent", [("c4", 64, "GPAT_NO_L2D", 18, 2), ("c5", 32, "GPAT_NO_L3D", 21, 3)])
def test_side_plane_layouts_match_the_one_plane_records(monkeypatch, key, grid, env, nslots, absent):
    """The production layouts with a side plane -- L2D (config C4: the 2-D Parker line with rho in its pad slot +
    dvx_dy / dvy_dx behind it) and L3D (config C5: the 3-D record split at the 128-byte line), both gathered by
    four lanes per particle -- against the one-plane records they replace (L2E / L3B, selected by GPAT_NO_L2D /
    GPAT_NO_L3D): interpolated fields bit for bit, 30 pushes to 1e-12, and whole intervals with histograms."""
    w, P, frames, ts = make_case(key, grid=grid, nptl=512)
    P.strict_math = 0
    n = 300
    rng = np.random.default_rng(5)
    x, y = rng.uniform(P.xmin, P.xmax, n), rng.uniform(P.ymin, P.ymax, n)
    z = rng.uniform(P.zmin, P.zmax, n) if P.ndim == 3 else np.zeros(n)
    rt = rng.uniform(0, 1, n)
    out = {}
    for tag, val in (("side", None), ("one", "1")):
        if val:
            monkeypatch.setenv(env, val)
        else:
            monkeypatch.delenv(env, raising=False)
        g = GpatSim(P, w.nptl_max)
        load_fields((g,), frames, P.time_interp)
        fi = g.interp(x, y, z, rt)
        _inject((g,), w, P, 512, dist_flag=0)
        s = g.debug_push_n(0.0, w.dt_out, 30)
        a = g.download_particles()
        g.close()
        g = GpatSim(P, w.nptl_max)
        rec, steps = run_intervals(g, frames, ts, nptl=512, particle_v0=w.particle_v0, split_flag=0)
        out[tag] = (fi, s, a, rec, steps, sort_by_key(g.download_particles()))
        g.close()
    fs, fe = out["side"][0], out["one"][0]
    used = np.any(fs != 0.0, axis=0)
    assert used.sum() == nslots and not np.any(fs[:, absent] != 0.0)   # C4: no vz; C5: no rho
    assert np.array_equal(fs[:, used], fe[:, used])
    assert out["side"][1] == out["one"][1]
    assert_particles_close(out["side"][2], out["one"][2], STEP_RTOL * 8, "side plane vs one plane, 30 pushes",
                           frac_outliers=0.004)
    assert abs(out["side"][4] - out["one"][4]) <= 2e-3 * out["one"][4]
    assert_particles_close(out["side"][5], out["one"][5], FRAME_RTOL, "side plane vs one plane, intervals",
                           int_exact=False, frac_outliers=0.01)
    assert np.abs(out["side"][3][-1]["fglobal"] - out["one"][3][-1]["fglobal"]).sum() <= 4.0
