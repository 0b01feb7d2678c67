"""CPU tests of the C++ host driver (host/gpat_driver.cpp): switches, conf.dat, frame / map / surface / tag
files, call order and output files -- the host logic above the C ABI.

The driver binary is the product's; the library underneath it is replaced, for these tests only, by the
TEST DOUBLE in tests/abi_mock/ (LD_PRELOAD), whose entry points forward to the CPU oracle.  The same
sequence driven from Python (run_intervals on oracle.Oracle) must then give bit-identical files.  The
GPU counterparts of these tests (the real library under the same binary) are in test_gpu_parity.py.
"""
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import Oracle
from stochastic_parker_b200 import WORKLOADS, config, mhd, outputs, run_intervals
from stochastic_parker_b200.abi import PARTICLE_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK_DIR = os.path.join(ROOT, "tests", "abi_mock")
MOCK = os.path.join(MOCK_DIR, "_build", "libgpat_testdouble.so")
DRIVER = os.path.join(ROOT, "host", "gpat_driver")


@pytest.fixture(scope="module")
def driver():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    if not os.path.exists(os.path.join(ROOT, "stochastic_parker_b200", "csrc", "libgpat_cuda.so")):
        pytest.skip("the product library is not built (the driver links against it)")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "host")], check=True)
    os.makedirs(os.path.dirname(MOCK), exist_ok=True)
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-Wall", "-o", MOCK, os.path.join(MOCK_DIR, "gpat_mock.c"),
                    "-L" + os.path.join(ROOT, "oracle"), "-lorc", "-Wl,-rpath," + os.path.join(ROOT, "oracle")],
                   check=True)

    def run(args, expect_ok=True):
        env = dict(os.environ, LD_PRELOAD=MOCK, OMP_NUM_THREADS="4")
        r = subprocess.run([DRIVER] + [str(a) for a in args], capture_output=True, text=True, timeout=600, env=env)
        if expect_ok:
            assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        return r
    return run


def _setup(tmp_path, key, grid, nptl, nfr, conf=None, cli=None):
    w = WORKLOADS[key].scaled(grid=grid, nptl=nptl)
    if conf:
        w.conf = dict(w.conf, **conf)
    if cli:
        w.cli = dict(w.cli, **cli)
    d = tmp_path / "mhd"
    cfg = mhd.write_run(str(d), w.kind, w.nx, w.ny, w.nz, nframes=nfr, lx=w.lx, ly=w.ly, lz=w.lz, dt_out=w.dt_out)
    (tmp_path / "conf.dat").write_text(w.conf_text())
    out = tmp_path / "out"
    out.mkdir()
    P = config.build_params(w.conf_text(), mhd.read_mhd_config(str(d / "mhd_config.dat")), w.ndim, nframes=nfr - 1,
                            cli=w.cli)
    frames = [mhd.read_frame(str(d), f, dict(cfg, ndim=w.ndim)) for f in range(nfr)]
    base = ["-nl", ".false.", "-pv", repr(w.particle_v0), "-dm", str(d) + "/", "-np", nptl, "-ti", "1", "-ts", "0",
            "-te", nfr - 1, "-df", "1", "-pi", "6.2", "-sf", "1", "-sr", "1.05", "-ps", "1.05", "-ni", "100",
            "-dt", "0.0", "-dd", str(out) + "/", "-cf", str(tmp_path / "conf.dat"), "-ld", ".true.", "-nm", 12 * nptl,
            "-in", ".true.", "-nd", w.ndim, "-dp1", "850964.408", "-dp2", "13575468.975", "-ch", "-1"]
    return w, P, frames, d, out, base


def _spectra(out, frame):
    return outputs.read_fdists(str(out), frame)["fglobal"]


def _same_run(r, out, rec, steps, nfr, local=True):
    assert f"Total particle steps: {steps} " in r.stdout
    for d in rec:
        assert np.array_equal(_spectra(out, d["frame"]), d["fglobal"]), d["frame"]
        if local and d["flocal"][1] is not None:
            assert np.array_equal(outputs.read_fdists_local(str(out), 2, d["frame"]), d["flocal"][1])
    q = outputs.read_quick(str(out))
    assert list(q)[:3] == ["iframe", "nptl_current", "nptl_split"] and list(q["iframe"]) == list(range(nfr))
    assert q["nptl_current"][-1] == float(f"{rec[-1]['quick'][0]:.5E}")
    assert q["ntot"][-1] == float(f"{rec[-1]['quick'][2]:.5E}") and q["pdt_max"][-1] == float(f"{rec[-1]['quick'][7]:.5E}")
    pm = outputs.read_pmax_global(str(out))
    assert len(pm) == nfr and pm[-1] == float(f"{rec[-1]['pmax']:.5E}")


KW = dict(dist_flag=1, power_index=6.2, split_ratio=1.05, pmin_split=1.05)


def test_basic_2d_run_with_split_and_local_spectra(driver, tmp_path):
    w, P, frames, d, out, base = _setup(tmp_path, "c1", 48, 800, 4)
    r = driver(base)
    rec, steps = run_intervals(Oracle(P, 12 * 800), frames, [f * w.dt_out for f in range(4)], nptl=800,
                               particle_v0=w.particle_v0, **KW)
    _same_run(r, out, rec, steps, 4)


def test_fine_steps_part_box_and_power_law(driver, tmp_path):
    w, P, frames, d, out, base = _setup(tmp_path, "c3", 48, 600, 3)
    box = [P.xmin + 0.8 * P.lx, P.ymin, 0.0, P.xmax, P.ymax, 0.0]   # next to the open high-x boundary
    args = base + ["-nf", "3", "-ip", ".true.", "-xs", box[0], "-ys", box[1], "-zs", box[2], "-xe", box[3],
                   "-ye", box[4], "-ze", box[5], "-df", "2", "-in", ".false.", "-ded", ".true.", "-pd", "1",
                   "-de", ".true."]
    r = driver(args)
    rec, steps = run_intervals(Oracle(P, 12 * 600), frames, [f * w.dt_out for f in range(3)], nptl=600,
                               particle_v0=w.particle_v0, **dict(KW, dist_flag=2), num_fine_steps=3, part_box=box,
                               inject_new_ptl=False, dump_escaped_dist=True, particle_data_dump=True, dump_escaped=True)
    _same_run(r, out, rec, steps, 3)
    # -pd 1 / -de .true.: particles_NNNN and escaped_particles_NNNN hold the same records, in the same order
    for d in rec:
        for stem, key in (("particles", "particles"), ("escaped_particles", "escaped_particles")):
            if key not in d:
                assert not os.path.exists(out / f"{stem}_{d['frame']:04d}.bin")
                continue
            raw = open(out / f"{stem}_{d['frame']:04d}.bin", "rb").read()
            assert np.frombuffer(raw[:8], dtype=np.int64)[0] == len(d[key])
            got = np.frombuffer(raw[8:], dtype=PARTICLE_DTYPE)
            for name in PARTICLE_DTYPE.names:
                assert np.array_equal(got[name], d[key][name]), (stem, d["frame"], name)
    assert sum(len(d.get("escaped_particles", ())) for d in rec) > 0
    # -ded: escaped_dists_NNNN (global) and escaped_dists_localK_NNNN (face arrays) of every interval
    assert sum(d["fescaped"].sum() for d in rec[1:]) > 0
    for d in rec[1:]:
        assert np.array_equal(outputs.read_escaped_dists(str(out), d["frame"]), d["fescaped"])
        for k, loc in enumerate(d["fescaped_local"]):
            if loc is None:
                assert not os.path.exists(out / f"escaped_dists_local{k + 1}_{d['frame']:04d}.bin")
                continue
            got = outputs.read_escaped_dists_local(str(out), k + 1, d["frame"])
            for f in "xyz":
                assert (got[f] is None) == (loc[f] is None)
                if loc[f] is not None:
                    assert np.array_equal(got[f], loc[f]), (d["frame"], k, f)
    # -pd 1 / -de .true.: particles_NNNN and escaped_particles_NNNN hold the same records, in the same order
    for d in rec:
        for stem, key in (("particles", "particles"), ("escaped_particles", "escaped_particles")):
            if key not in d:
                assert not os.path.exists(out / f"{stem}_{d['frame']:04d}.bin")
                continue
            raw = open(out / f"{stem}_{d['frame']:04d}.bin", "rb").read()
            assert np.frombuffer(raw[:8], dtype=np.int64)[0] == len(d[key])
            got = np.frombuffer(raw[8:], dtype=PARTICLE_DTYPE)
            for name in PARTICLE_DTYPE.names:
                assert np.array_equal(got[name], d[key][name]), (stem, d["frame"], name)
    assert sum(len(d.get("escaped_particles", ())) for d in rec) > 0
    # -ded: escaped_dists_NNNN (global) and escaped_dists_localK_NNNN (face arrays) of every interval
    assert sum(d["fescaped"].sum() for d in rec[1:]) > 0
    for d in rec[1:]:
        raw = open(out / f"escaped_dists_{d['frame']:04d}.bin", "rb").read()
        nmu, npp, nface = np.frombuffer(raw[:12], dtype=np.int32)
        assert np.array_equal(np.frombuffer(raw[12:], dtype=np.float64).reshape(nface, npp, nmu), d["fescaped"])
        for k, loc in enumerate(d["fescaped_local"]):
            if loc is None:
                assert not os.path.exists(out / f"escaped_dists_local{k + 1}_{d['frame']:04d}.bin")
                continue
            raw = open(out / f"escaped_dists_local{k + 1}_{d['frame']:04d}.bin", "rb").read()
            body = np.frombuffer(raw[24:], dtype=np.float64)
            want = np.concatenate([loc[f].ravel() for f in "xyz" if loc[f] is not None])
            assert np.array_equal(body, want), (d["frame"], k)


@pytest.mark.parametrize("switch,extra,mode,vmin,norm", [
    ("-ij", ["-jz", "0.5", "-nn", "40"], 1, 0.5, 40),
    ("-iaj", ["-ajm", "0.5", "-naj", "40"], 2, 0.5, 40),
    ("-iv", ["-dv", "0.2", "-nv", "40"], 4, 0.2, 40),
    ("-ir", ["-rm", "1.05", "-nr", "40"], 5, 1.05, 40),
])
def test_targeted_injection_switches(driver, tmp_path, switch, extra, mode, vmin, norm):
    w, P, frames, d, out, base = _setup(tmp_path, "c1", 48, 500, 3)
    r = driver(base + [switch, ".true.", "-sn", ".false."] + extra)
    rec, steps = run_intervals(Oracle(P, 12 * 500), frames, [f * w.dt_out for f in range(3)], nptl=500,
                               particle_v0=w.particle_v0, **KW, inject_mode=mode, inject_same_nptl=False,
                               inject_min=vmin, ncells_norm=norm)
    _same_run(r, out, rec, steps, 3)


def test_shock_injection_switch(driver, tmp_path):
    w, P, frames, d, out, base = _setup(tmp_path, "c3", 48, 500, 3)
    r = driver(base + ["-is", ".true."])
    rec, steps = run_intervals(Oracle(P, 12 * 500), frames, [f * w.dt_out for f in range(3)], nptl=500,
                               particle_v0=w.particle_v0, **KW, inject_mode=6)
    _same_run(r, out, rec, steps, 3)


def test_shock_injection_ignores_the_inject_new_ptl_switch(driver, tmp_path):
    """stochastic-mhd.f90:451-454 calls locate_shock_xpos + inject_particles_at_shock on EVERY frame, before and
    outside the `tf == 1 .or. inject_new_ptl` gate: with -in .false. both host drivers still inject each frame."""
    w, P, frames, d, out, base = _setup(tmp_path, "c3", 48, 300, 3)
    r = driver(base + ["-is", ".true.", "-in", ".false."])
    rec, steps = run_intervals(Oracle(P, 12 * 300), frames, [f * w.dt_out for f in range(3)], nptl=300,
                               particle_v0=w.particle_v0, **KW, inject_mode=6, inject_new_ptl=False)
    _same_run(r, out, rec, steps, 3)
    assert rec[-1]["quick"][0] > 1.5 * 300     # two injections, not one


def test_one_dimensional_run(driver, tmp_path):
    w, P, frames, d, out, base = _setup(tmp_path, "s1", 256, 500, 3)
    r = driver(base)
    rec, steps = run_intervals(Oracle(P, 12 * 500), frames, [f * w.dt_out for f in range(3)], nptl=500,
                               particle_v0=w.particle_v0, **KW)
    _same_run(r, out, rec, steps, 3)


def test_focused_transport_switches(driver, tmp_path):
    w, P, frames, d, out, base = _setup(tmp_path, "c1", 48, 400, 3, conf=dict(dt_min_rel=1e-3),
                                        cli=dict(focused_transport=1, duu_init=5.0))
    r = driver(base + ["-ft", ".true.", "-du", "5.0"])
    rec, steps = run_intervals(Oracle(P, 12 * 400), frames, [f * w.dt_out for f in range(3)], nptl=400,
                               particle_v0=w.particle_v0, **KW)
    assert P.nmu_global > 1
    _same_run(r, out, rec, steps, 3)


def test_turbulence_map_files(driver, tmp_path):
    w, P, frames, d, out, base = _setup(tmp_path, "c1", 48, 400, 3)
    P.deltab_flag = 1
    P.correlation_flag = 1
    maps = [mhd.make_turbulence_maps(w.nx, w.ny, w.nz, f, 2, w.dt_out) for f in range(3)]
    for f, (s2s, s22, lcs, lc2) in enumerate(maps):
        np.stack([s2s, s22]).tofile(str(d / f"deltab_{f:04d}"))
        np.stack([lcs, lc2]).tofile(str(d / f"lc_{f:04d}"))
    r = driver(base + ["-db", "1", "-co", "1"])

    rec, steps = run_intervals(Oracle(P, 12 * 400), frames, [f * w.dt_out for f in range(3)], nptl=400,
                               particle_v0=w.particle_v0, **KW,
                               maps=lambda which, f: (maps[f][2 * which], maps[f][2 * which + 1]))
    _same_run(r, out, rec, steps, 3)
    os.remove(d / "lc_0002")
    assert driver(base + ["-db", "1", "-co", "1"], expect_ok=False).returncode != 0


def test_acceleration_surface_files(driver, tmp_path):
    w, P, frames, d, out, base = _setup(tmp_path, "c5", 24, 400, 3, conf=dict(acc_region_flag=1, r1=4, r2=8, r3=12),
                                        cli=dict(acc_by_surface=1, surface_norm1="+z", surface2_existed=1,
                                                 surface_norm2="-y", is_intersection=1))
    for f in range(3):
        for k, stem in enumerate(("surf_a", "surf_b")):
            mhd.make_acc_surface(P, k, f).tofile(str(d / f"{stem}_{f:04d}.dat"))
    args = base + ["-as", "1", "-sn1", "+z", "-s2e", ".true.", "-sn2", "-y", "-ii", ".true.", "-sf1", "surf_a",
                   "-sf2", "surf_b"]
    r = driver(args)
    rec, steps = run_intervals(Oracle(P, 12 * 400), frames, [f * w.dt_out for f in range(3)], nptl=400,
                               particle_v0=w.particle_v0, **KW, surfaces=lambda k, f: mhd.make_acc_surface(P, k, f))
    _same_run(r, out, rec, steps, 3)
    ungated, _ = run_intervals(Oracle(config.build_params(w.conf_text(), mhd.read_mhd_config(str(d / "mhd_config.dat")),
                                                          3, nframes=2, cli=dict(w.cli, acc_by_surface=0)), 12 * 400),
                               frames, [f * w.dt_out for f in range(3)], nptl=400, particle_v0=w.particle_v0, **KW)
    assert not np.array_equal(ungated[-1]["fglobal"], rec[-1]["fglobal"])   # the gate matters in this case
    os.remove(d / "surf_b_0001.dat")
    assert driver(args, expect_ok=False).returncode != 0


def test_tracking_run(driver, tmp_path):
    from stochastic_parker_b200 import tracking
    w, P, frames, d, out, base = _setup(tmp_path, "c1", 48, 500, 3, conf=dict(dt_min_rel=1e-4))
    ts = [f * w.dt_out for f in range(3)]
    first = Oracle(P, 12 * 500)
    run_intervals(first, frames, ts, nptl=500, particle_v0=w.particle_v0, **KW, inject_new_ptl=False)
    dump = first.download_particles()
    assert dump["split_times"].max() >= 1
    tags = tracking.select_tags(dump, np.argsort(dump["p"])[-12:])
    with open(tmp_path / "tags.bin", "wb") as f:
        np.array(tags.shape, dtype=np.int32).tofile(f)
        np.ascontiguousarray(tags, dtype=np.int32).tofile(f)
    r = driver(base + ["-tf", ".true.", "-ptf", str(tmp_path / "tags.bin"), "-in", ".false."])
    got = {}
    o = Oracle(P, 12 * 500)
    rec, steps = run_intervals(o, frames, ts, nptl=500, particle_v0=w.particle_v0, **KW, inject_new_ptl=False,
                               track_tags=tags, on_tracked=lambda tf, a: got.__setitem__(tf, a.copy()))
    assert f"Total particle steps: {steps} " in r.stdout
    for tf, want in got.items():
        raw = open(out / f"particle_tracking_particles_tracked_{tf:04d}.bin", "rb").read()
        ntrk, nmax = np.frombuffer(raw[:16], dtype=np.int64)
        have = np.frombuffer(raw[16:], dtype=PARTICLE_DTYPE).reshape(ntrk, nmax)
        assert have.shape == want.shape
        for name in PARTICLE_DTYPE.names:   # field by field: the two pad bytes of the record are not data
            assert np.array_equal(have[name], want[name]), (tf, name)
    assert not os.path.exists(out / "fdists_0001.bin")   # no distributions in a tracking run


def test_bad_switches_stop_the_driver(driver, tmp_path):
    w, P, frames, d, out, base = _setup(tmp_path, "c1", 32, 50, 2)
    assert driver(base + ["-nosuch", "1"], expect_ok=False).returncode == 2
    assert driver(base + ["-sc", "1"], expect_ok=False).returncode == 2
    assert driver(base + ["-vdt", ".true."], expect_ok=False).returncode != 0     # no time_stamps.dat
    assert driver(base + ["-rf", ".true."], expect_ok=False).returncode != 0      # no restart/ files to read
    assert driver(base[:4] + ["-dm", str(tmp_path / "nowhere") + "/"] + base[6:], expect_ok=False).returncode == 2


def test_restart_files_and_restart_flag(driver, tmp_path):
    """-te 2, then -rf .true. -te 4: the restart files the driver always writes at the end
    (stochastic-mhd.f90:252-271) resume the run bit-identically; run_intervals' dump_restart / read_restart use
    the same files."""
    from stochastic_parker_b200 import dump_restart, read_restart
    w, P, frames, d, out, base = _setup(tmp_path, "c3", 48, 500, 5)
    ts = [f * w.dt_out for f in range(5)]
    te = base.index("-te") + 1
    first = list(base)
    first[te] = 2
    driver(first + ["-nf", "2"])
    rdir = out / "restart"
    assert np.fromfile(rdir / "latest_restart", dtype=np.int32)[0] == 2
    r = driver(base + ["-nf", "2", "-rf", ".true."])
    assert "This is a restart" in r.stdout and " Starting step 3" in r.stdout and " Starting step 2" not in r.stdout
    full = Oracle(P, 12 * 500)
    rec, steps = run_intervals(full, frames, ts, nptl=500, particle_v0=w.particle_v0, **KW, num_fine_steps=2)
    for rd in rec:
        assert np.array_equal(_spectra(out, rd["frame"]), rd["fglobal"]), rd["frame"]
    rows = open(out / "quick.dat").read().splitlines()
    assert len(rows) == 1 + 5 and [int(x[:6]) for x in rows[1:]] == [0, 1, 2, 3, 4]
    # the files of the second run hold the final population of the uninterrupted run
    assert np.fromfile(rdir / "latest_restart", dtype=np.int32)[0] == 4
    raw = open(rdir / "particles_0004.bin", "rb").read()
    n = np.frombuffer(raw[:8], dtype=np.int64)[0]
    got = np.frombuffer(raw[8:], dtype=PARTICLE_DTYPE)
    want = full.download_particles()
    assert n == len(want) == len(got)
    for name in PARTICLE_DTYPE.names:
        assert np.array_equal(got[name], want[name]), name
    # and Python reads what C++ wrote (first run's files), continuing to the same end state
    o = Oracle(P, 12 * 500)
    np.array([2], dtype=np.int32).tofile(rdir / "latest_restart")
    assert read_restart(o, str(out) + "/") == 2
    run_intervals(o, frames, ts, nptl=500, particle_v0=w.particle_v0, **KW, num_fine_steps=2, tmin=2)
    have = o.download_particles()
    for name in PARTICLE_DTYPE.names:
        assert np.array_equal(have[name], want[name]), name
    dump_restart(o, str(tmp_path / "py") + "/", 4, 4)
    assert open(tmp_path / "py" / "restart" / "particle_module_state_0004.bin", "rb").read() == \
        open(rdir / "particle_module_state_0004.bin", "rb").read()


def test_varying_frame_interval(driver, tmp_path):
    """-vdt .true.: load_tstamps_mhd (mhd_config.f90:221-254) -- frame times from time_stamps.dat, frames past
    -tm continue with the last interval."""
    w, P, frames, d, out, base = _setup(tmp_path, "c1", 48, 500, 5)
    stamps = np.array([0.0, 0.07, 0.19, 0.26, 0.40])
    mhd.write_time_stamps(str(d), 0, 3, stamps[:4])                 # the MHD run has frames 0..3 only
    r = driver(base + ["-vdt", ".true.", "-tm", "3", "-st", "0"], expect_ok=False)
    # frame 4 does not exist as a file: with -tm 3 the driver must not try to read it
    ts = mhd.read_time_stamps(str(d), 0, 4, 3)
    assert np.allclose(ts, [0.0, 0.07, 0.19, 0.26, 0.33]) and ts[4] - ts[3] == ts[3] - ts[2]
    os.remove(d / "mhd_data_0004")
    r = driver(base + ["-vdt", ".true.", "-tm", "3"])

    rec, steps = run_intervals(Oracle(P, 12 * 500), frames[:4], list(ts), nptl=500, particle_v0=w.particle_v0, **KW,
                               tmax_mhd=3)
    _same_run(r, out, rec, steps, 5)


def test_single_time_frame_without_time_interpolation(driver, tmp_path):
    """-st 1 -ti 0: the first frame is the only one read (stochastic-mhd.f90:400); every interval uses it."""
    w, P, frames, d, out, base = _setup(tmp_path, "c1", 48, 400, 4, cli=dict(time_interp=0))
    ti = base.index("-ti") + 1
    args = list(base)
    args[ti] = "0"
    for f in (1, 2, 3):
        os.remove(d / f"mhd_data_{f:04d}")          # must not be needed
    r = driver(args + ["-st", "1"])
    assert P.time_interp == 0
    rec, steps = run_intervals(Oracle(P, 12 * 400), [frames[0]] * 4, [f * w.dt_out for f in range(4)], nptl=400,
                               particle_v0=w.particle_v0, **KW)
    _same_run(r, out, rec, steps, 4)


def test_third_dimension_dpp_and_nlgc_switches(driver, tmp_path):
    """-i3 1 -dw 1 -ds 1 -ws 0 -nl .true. -kk 0.05 -cd 0: the switch-to-parameter mapping of the physics options."""
    cli = dict(include_3rd_dim=1, dpp_wave=1, dpp_shear=1, weak_scattering=0, nlgc=1, kperp_kpara=0.05, tau0=1e-3)
    w, P, frames, d, out, base = _setup(tmp_path, "c1", 48, 400, 3, conf=dict(dt_min_rel=1e-3), cli=cli)
    args = list(base)
    args[args.index("-nl") + 1] = ".true."
    r = driver(args + ["-i3", "1", "-dw", "1", "-ds", "1", "-ws", "0", "-kk", "0.05", "-t0", "1e-3"])
    rec, steps = run_intervals(Oracle(P, 12 * 400), frames, [f * w.dt_out for f in range(3)], nptl=400,
                               particle_v0=w.particle_v0, **KW)
    assert P.nlgc == 1 and P.include_3rd_dim == 1 and P.dpp_wave == 1 and P.weak_scattering == 0 and P.tau0 == 1e-3
    _same_run(r, out, rec, steps, 3)


def test_three_dimensional_run(driver, tmp_path):
    w, P, frames, d, out, base = _setup(tmp_path, "c5", 24, 300, 3, conf=dict(r1=4, r2=8, r3=12))
    r = driver(base)
    rec, steps = run_intervals(Oracle(P, 12 * 300), frames, [f * w.dt_out for f in range(3)], nptl=300,
                               particle_v0=w.particle_v0, **KW)
    _same_run(r, out, rec, steps, 3)


def test_large_db2_injection_with_map_files(driver, tmp_path):
    """-ib .true. -db2 ... -nb ... with -db 1: inject_particles_at_large_db2 reads the uploaded dB^2 map."""
    w, P, frames, d, out, base = _setup(tmp_path, "c1", 48, 400, 3)
    P.deltab_flag = 1
    maps = [mhd.make_turbulence_maps(w.nx, w.ny, w.nz, f, 2, w.dt_out) for f in range(3)]
    for f, (s2s, s22, lcs, lc2) in enumerate(maps):
        np.stack([s2s, s22]).tofile(str(d / f"deltab_{f:04d}"))
    vmin = float(np.median(maps[0][0]))
    r = driver(base + ["-db", "1", "-ib", ".true.", "-db2", repr(vmin), "-nb", "40", "-sn", ".false."])

    rec, steps = run_intervals(Oracle(P, 12 * 400), frames, [f * w.dt_out for f in range(3)], nptl=400,
                               particle_v0=w.particle_v0, **KW, maps=lambda which, f: (maps[f][0], maps[f][1]),
                               inject_mode=3, inject_same_nptl=False, inject_min=vmin,
                               ncells_norm=40)
    _same_run(r, out, rec, steps, 3)
    assert rec[-1]["quick"][0] > 0
