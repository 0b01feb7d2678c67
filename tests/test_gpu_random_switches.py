"""48 seeded random combinations of the Parker-transport switches and 36 of the general pushers (focused transport, 1-D,
turbulence maps), both builds: one push and 20 pushes of the GPU library against the oracle (the CPU twin, oracle vs numpy, is tests/test_cpu_oracle.py::
test_random_switch_combinations_match_numpy_restatement)."""
import numpy as np
import pytest

from helpers import assert_particles_close, box_of, make_case
from oracle.oracle import Oracle
from stochastic_parker_b200 import GpatSim

pytestmark = pytest.mark.gpu


def _combo(trial):
    rng = np.random.default_rng(7000 + trial)
    geom = ["2d", "2d3", "3d"][trial % 3]
    conf = dict(mag_dependency=int(rng.integers(0, 2)), momentum_dependency=int(rng.integers(0, 2)),
                kret=float(rng.choice([0.0, 0.01, 0.3])), acc_region_flag=int(rng.integers(0, 2)), dt_min_rel=1e-4)
    cli = dict(nlgc=int(rng.integers(0, 2)), kperp_kpara=0.05, dpp_wave=int(rng.integers(0, 2)),
               dpp_shear=int(rng.integers(0, 2)), weak_scattering=int(rng.integers(0, 2)),
               time_interp=int(rng.integers(0, 2)), check_drift_2d=int(rng.integers(0, 2)) if geom == "2d" else 0)
    if geom == "2d3":
        cli["include_3rd_dim"] = 1
    if geom == "3d":
        conf.update(r1=4, r2=8, r3=16)
    return geom, conf, cli


@pytest.mark.parametrize("trial", range(24))
@pytest.mark.parametrize("strict", [1, 0])
def test_random_switch_combination(trial, strict):
    geom, conf, cli = _combo(trial)
    key, grid = ("c5", 32) if geom == "3d" else ("c1", 64)
    w, P, frames, _ = make_case(key, grid=grid, nptl=512, conf=conf, cli=cli)
    if conf["acc_region_flag"]:
        for i, v in enumerate((0.2, 0.8, 0.1, 0.7, 0.3, 0.9)):
            P.acc_region[i] = v
    Pg = P.copy()
    Pg.strict_math = strict
    g, o = GpatSim(Pg, w.nptl_max), Oracle(P, w.nptl_max)
    for s in (g, o):
        s.upload_fields(0, frames[0])
        if P.time_interp:
            s.upload_fields(1, frames[1])
        s.inject_uniform(512, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    for nsteps in (1, 20):
        assert g.debug_push_n(0.0, w.dt_out, nsteps) == o.debug_push_n(0.0, w.dt_out, nsteps)
        assert_particles_close(g.download_particles(), o.download_particles(), 1e-12 * max(1, nsteps // 4),
                               f"{geom} {conf} {cli} strict={strict} n={nsteps}", frac_outliers=0.004)
    g.close()


def _general_combo(trial):
    """Switch combinations of the pushers outside the five named configs: focused transport (2-D, 2-D + third dimension,
    3-D) and 1-D, each with or without the turbulence maps; Parker 2-D / 3-D with maps."""
    rng = np.random.default_rng(9000 + trial)
    kind = ["ft2d", "ft2d3", "ft3d", "1d", "maps2d", "maps3d"][trial % 6]
    conf = dict(momentum_dependency=int(rng.integers(0, 2)), kret=float(rng.choice([0.0, 0.01, 0.3])),
                acc_region_flag=int(rng.integers(0, 2)), dt_min_rel=1e-4)
    cli = dict(nlgc=int(rng.integers(0, 2)), kperp_kpara=0.05, dpp_wave=int(rng.integers(0, 2)),
               dpp_shear=int(rng.integers(0, 2)), weak_scattering=int(rng.integers(0, 2)),
               time_interp=int(rng.integers(0, 2)))
    if kind != "1d":   # gpat_init rejects mag_dependency in 1-D (the reference multiplies by an unassigned db_dx there)
        conf["mag_dependency"] = int(rng.integers(0, 2))
    else:
        conf["mag_dependency"] = 0
    if kind.startswith("ft"):
        cli.update(focused_transport=1, duu_init=float(rng.choice([1.0, 5.0])))
    if kind == "ft2d3":
        cli["include_3rd_dim"] = 1
    maps = kind.startswith("maps") or bool(rng.integers(0, 2))
    key, grid = {"ft2d": ("c1", 64), "ft2d3": ("c1", 64), "ft3d": ("c5", 32), "1d": ("s1", 256),
                 "maps2d": ("c1", 64), "maps3d": ("c5", 32)}[kind]
    if key == "c5":
        conf.update(r1=4, r2=8, r3=16)
    return kind, key, grid, conf, cli, maps


@pytest.mark.parametrize("trial", range(18))
@pytest.mark.parametrize("strict", [1, 0])
def test_random_general_pusher_combination(trial, strict):
    """strict = 0: the kSpecAlt / kSpecAltMaps instantiations of the production build (lane-group gather of fields and maps
    + the general pushers with the straight-line math); strict = 1: the reference-order kernels.  One and 20 pushes, 1e-12."""
    from stochastic_parker_b200 import mhd
    kind, key, grid, conf, cli, maps = _general_combo(trial)
    w, P, frames, _ = make_case(key, grid=grid, nptl=512, conf=conf, cli=cli)
    if conf["acc_region_flag"]:
        for i, v in enumerate((0.2, 0.8, 0.1, 0.7, 0.3, 0.9)):
            P.acc_region[i] = v
    if maps:
        P.deltab_flag = 1
        P.correlation_flag = 1
    Pg = P.copy()
    Pg.strict_math = strict
    g, o = GpatSim(Pg, w.nptl_max), Oracle(P, w.nptl_max)
    for s in (g, o):
        s.upload_fields(0, frames[0])
        if P.time_interp:
            s.upload_fields(1, frames[1])
        if maps:
            for slot in ((0, 1) if P.time_interp else (0,)):
                m = mhd.make_turbulence_maps(P.nx, P.ny, P.nz, slot, ndim=P.ndim)
                s.upload_turbulence(0, slot, m[0], m[1])
                s.upload_turbulence(1, slot, m[2], m[3])
        s.inject_uniform(512, 0.0, 0, w.particle_v0, 0.0, w.dt_out, box_of(P), w.power_index)
    for nsteps in (1, 20):
        assert g.debug_push_n(0.0, w.dt_out, nsteps) == o.debug_push_n(0.0, w.dt_out, nsteps)
        assert_particles_close(g.download_particles(), o.download_particles(), 1e-12 * max(1, nsteps // 4),
                               f"{kind} {conf} {cli} maps={maps} strict={strict} n={nsteps}", frac_outliers=0.004)
    g.close()


def test_split_of_tracked_particles_hand_made_population():
    """GPU twin of tests/test_cpu_oracle.py::test_split_of_tracked_particles_matches_python_restatement: the same
    six particles and tag table through gpat_init_tracking + gpat_split, against the oracle, field by field
    (tests/test_cpu_reference_f90.py runs the same population through the reference's own split_particle).
    Round 2's first GPU run of this test failed on a population in which two particles shared one
    particles_tracked slot (serial last-writer-wins in the reference, undefined in a parallel split); real runs
    cannot produce that, and helpers.tracked_split_population() no longer does."""
    from stochastic_parker_b200.abi import PARTICLE_DTYPE
    w, P, _, _ = make_case("c1", grid=16, nptl=8, conf=dict(dt_min_rel=1e-2))
    from helpers import tracked_split_population
    tags, ptl = tracked_split_population(P)
    g, o = GpatSim(P, 32), Oracle(P, 32)
    for s in (g, o):
        s.init_tracking(tags, 10)
        s.upload_particles(ptl)
        s.split(2.0, 2.0, 10)
    a, b = g.download_particles(), o.download_particles()
    assert len(a) == len(b) == 11
    for name in PARTICLE_DTYPE.names:
        assert np.array_equal(a[name], b[name]), name
    ra, rb = g.download_tracked(), o.download_tracked()
    for name in PARTICLE_DTYPE.names:
        assert np.array_equal(ra[name], rb[name]), name
    g.close()
