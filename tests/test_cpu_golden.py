"""Golden fixtures generated from the reference's own Python and data files
(tests/golden/make_golden.py, run once in the build container; the tests never read
/root/reference).  They pin the data formats on either side of the particle path and the
run constants of config C1; the particle path itself has no reference golden vectors
(SURVEY.md section 4), see tests/test_cpu_oracle.py for how the oracle is pinned instead.
"""
import json
import os

import numpy as np

from stochastic_parker_b200 import WORKLOADS, config, mhd

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_c1_constants_match_the_reference_normalisation():
    """examples/reconnection_2d/sde.py::reconnection_test() vs the constants its run script
    hard-codes (diffusion_reconnection.sh:182-189), which config C1 uses.  The shipped script
    values differ from today's sde.py by ~0.3 % (the script was generated from slightly
    different physical constants); both are within 0.5 %."""
    g = json.load(open(os.path.join(G, "reconnection_norm.json")))
    w = WORKLOADS["c1"]
    assert abs(w.conf["kpara0"] / g["normed_kappa_parallel"] - 1) < 5e-3
    assert abs(w.cli["drift_param1"] / g["drift_param1"] - 1) < 5e-3
    assert abs(w.cli["drift_param2"] / g["drift_param2"] - 1) < 5e-3
    assert abs(w.cli["tau0"] / g["tau0_scattering"] - 1) < 5e-3


def test_conf_template_equals_the_shipped_conf_file():
    """Every key of examples/reconnection_2d/conf_reconnection.dat that the particle path reads
    has the same value in this repo's C1 conf text (before the run script's kpara0/kret
    overrides), read with the reference's forward-scan semantics."""
    g = json.load(open(os.path.join(G, "conf_reconnection.json")))
    w = WORKLOADS["c1"].scaled()
    w.conf = dict(w.conf, kpara0=g["kpara0"], kret=g["kret"])  # undo diffusion_reconnection.sh:183-185
    r = config.ConfReader(w.conf_text())
    order = ["b0", "p0", "pmin", "pmax", "momentum_dependency", "gamma_turb", "mag_dependency", "kpara0", "kret",
             "dt_min_rel", "dt_max_rel", "npp_global", "nmu_global"]
    for k in range(1, 5):
        order += [f"dump_interval{k}", f"pmin{k}", f"pmax{k}", f"npbins{k}", f"nmu{k}", f"rx{k}", f"ry{k}", f"rz{k}"]
    order += ["acc_region_flag", "pbcx", "pbcy", "pbcz", "mpi_sizex", "mpi_sizey", "mpi_sizez"]
    for key in order:
        assert r.get(key) == g[key], key
    # and the Params the ABI receives
    cfg = mhd.mhd_config(1024, 1024, 1, 2.0, 2.0, 1.0, 0.1, 2)
    P = config.build_params(w.conf_text(), cfg, 2, nframes=200, cli=w.cli)
    assert (P.p0, P.pmin, P.pmax, P.gamma_turb) == (g["p0"], g["pmin"], g["pmax"], g["gamma_turb"])
    assert (P.dt_min_rel, P.dt_max_rel, P.npp_global) == (g["dt_min_rel"], g["dt_max_rel"], int(g["npp_global"]))
    assert [(P.local[k].npbins, P.local[k].rx) for k in range(3)] == [(12, 4), (64, 8), (32, 16)]
    assert P.local[3].enabled == 0  # dump_interval4 = 10000 > number of frames (diagnostics.f90:2129)


def _check_layout(tag, mode):
    z = np.load(os.path.join(G, f"reorganize_{tag}.npz"))
    fdata, ref = z["fdata"], z["mhd_data"]
    nx, ny = int(z["nx"]), int(z["ny"])
    assert ref.shape == (ny + 4, nx + 4, 8) and ref.dtype == np.float32
    # variable order vx vy vz rho bx by bz |B| from the Athena order rho p vx vy vz bx by bz
    # (reorganize_fields.py:75-82), C-order (y, x, var)
    interior = ref[2:ny + 2, 2:nx + 2]
    for dst, src in enumerate((2, 3, 4, 0, 5, 6, 7)):
        assert np.array_equal(interior[..., dst], fdata[..., src].T.astype(np.float32))
    absb = np.sqrt(np.sum(fdata[..., 5:8] ** 2, axis=2))  # f64 sqrt, then the cast (reorganize_fields.py:62)
    assert np.array_equal(interior[..., 7], absb.T.astype(np.float32))
    # this repo's ghost fill reproduces the reference's, corners included
    mine = np.zeros_like(ref)
    mine[2:ny + 2, 2:nx + 2] = interior
    mhd._ghost_fill(mine, 0, mode)
    mhd._ghost_fill(mine, 1, mode)
    assert np.array_equal(mine, ref)
    return z


def test_mhd_data_layout_periodic():
    _check_layout("periodic", "periodic")


def test_mhd_data_layout_reflect():
    _check_layout("reflect", "reflect")


def test_mhd_config_bytes(tmp_path):
    """mhd_config.dat: 13 f64 + 14 i32 as written by reorganize_fields.py:204-259; this repo's
    writer produces the same bytes for the fields the Fortran side reads (mhd_config.f90:139-148:
    13 f64 + 13 i32) and its reader parses the reference's file."""
    z = np.load(os.path.join(G, "reorganize_periodic.npz"))
    raw = z["mhd_config_bytes"].tobytes()
    assert len(raw) == 13 * 8 + 14 * 4
    p = tmp_path / "mhd_config.dat"
    p.write_bytes(raw)
    cfg = mhd.read_mhd_config(str(p))
    nx, ny = int(z["nx"]), int(z["ny"])
    assert (cfg["nx"], cfg["ny"], cfg["nz"], cfg["nvar"]) == (nx, ny, 1, 9)
    assert (cfg["lx"], cfg["ly"], cfg["dt_out"]) == (float(z["lx"]), float(z["ly"]), float(z["dt_out"]))
    assert cfg["dx"] == float(z["lx"]) / nx and cfg["xmax"] == float(z["lx"]) and cfg["xmin"] == 0.0
    mine = mhd.mhd_config(nx, ny, 1, float(z["lx"]), float(z["ly"]), float(z["lz"]), float(z["dt_out"]), 2)
    mine.update(nxs=cfg["nxs"], nys=cfg["nys"], nzs=cfg["nzs"], topox=cfg["topox"], topoy=cfg["topoy"],
                topoz=cfg["topoz"])  # the MPI topology of the reorganising script is not ours to choose
    q = tmp_path / "mine.dat"
    mhd.write_mhd_config(str(q), mine)
    assert q.read_bytes()[:13 * 8 + 13 * 4] == raw[:13 * 8 + 13 * 4]
