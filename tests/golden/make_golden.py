#!/usr/bin/env python
"""Generate tests/golden/*.{json,npz} from the reference's own Python and data files.

Run in the build container only (needs /root/reference, read-only); the fixtures are
committed, the tests never read /root/reference.  What is pinned:

  reconnection_norm.json   examples/reconnection_2d/sde.py::reconnection_test(): the
                           normalisation the reference's run script hard-codes
                           (diffusion_reconnection.sh:182-189): kpara0, drift parameters, tau0
  conf_reconnection.json   examples/reconnection_2d/conf_reconnection.dat read key by key
  reorganize_*.npz         examples/reconnection_2d/mhd_data/reorganize_fields.py::
                           save_mhd_fields_with_ghost + save_mhd_config on a small synthetic
                           Athena-style input: the on-disk layout of mhd_data_NNNN (ghost fill,
                           variable order, |B| slot) and the bytes of mhd_config.dat
"""
import contextlib
import io
import json
import os
import re
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def golden_norm():
    sys.path.insert(0, os.path.join(REF, "examples", "reconnection_2d"))
    import sde  # the reference's script (math only)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        sde.reconnection_test()
    text = buf.getvalue()
    out = {
        "normed_kappa_parallel": float(re.search(r"Normed kappa parallel: (\S+)", text).group(1)),
        "kperp_over_kpara": float(re.search(r"kperp / kpara: (\S+)", text).group(1)),
        "drift_param1": float(re.search(r"Parameters for particle drift: (\S+), (\S+)", text).group(1)),
        "drift_param2": float(re.search(r"Parameters for particle drift: (\S+), (\S+)", text).group(2)),
        "tau0_scattering": float(re.search(r"Scattering time for initial particles: (\S+)", text).group(1)),
        "source": "examples/reconnection_2d/sde.py::reconnection_test() stdout",
    }
    json.dump(out, open(os.path.join(HERE, "reconnection_norm.json"), "w"), indent=1)
    return out


def golden_conf():
    path = os.path.join(REF, "examples", "reconnection_2d", "conf_reconnection.dat")
    vals = {}
    for line in open(path):
        m = re.match(r"\s*([A-Za-z_0-9]+)\s*=\s*([-+0-9.EeDd]+)", line)
        if m and not line.lstrip().startswith("!"):
            vals[m.group(1)] = float(m.group(2).replace("D", "E").replace("d", "e"))
    # the same file through this repo's reader, in the reference's read order: must agree
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from stochastic_parker_b200.config import ConfReader
    r = ConfReader.open(path)
    for k in ("b0", "p0", "pmin", "pmax", "momentum_dependency", "gamma_turb", "mag_dependency", "kpara0", "kret",
              "dt_min_rel", "dt_max_rel"):
        assert r.get(k) == vals[k], k
    vals["_source"] = "examples/reconnection_2d/conf_reconnection.dat"
    json.dump(vals, open(os.path.join(HERE, "conf_reconnection.json"), "w"), indent=1, sort_keys=True)
    return vals


def golden_reorganize():
    """Run the reference's reorganize_fields on fake reader modules (matplotlib and the Athena
    readers are stubbed: they are I/O and plotting, not part of the layout logic)."""
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].rcParams = {}
    sys.modules["matplotlib"].rc = lambda *a, **k: None
    sys.modules["matplotlib.pyplot"].style = types.SimpleNamespace(use=lambda *a, **k: None)
    fake = types.ModuleType("mhd_data")
    sys.modules["mhd_data"] = fake
    sys.path.insert(0, os.path.join(REF, "examples", "reconnection_2d", "mhd_data"))
    rng = np.random.default_rng(2024)
    nx, ny = 12, 10
    fdata = rng.uniform(-1, 1, (nx, ny, 8))           # rho, p, vx, vy, vz, bx, by, bz
    fdata[..., 0] = np.abs(fdata[..., 0]) + 0.5
    fake.read_fields_data = lambda info, tframe: (np.zeros((ny, nx)), None, fdata)
    cfg = types.SimpleNamespace(nx=nx, ny=ny, nz=1, xmin=0.0, xmax=2.0, ymin=0.0, ymax=1.5, zmin=0.0, zmax=1.0,
                                dt_out=0.1)
    fake.read_mhd_config = lambda name, code: cfg
    import reorganize_fields as rf
    for boundary, tag in ((0, "periodic"), (1, "reflect")):
        with tempfile.TemporaryDirectory() as d:
            info = dict(mhd_code="Athena", run_dir=d + "/", config_name="athinput", xmirror=False, ymirror=False,
                        with_z_component=True, boundary=boundary, output_type="reconnection")
            with contextlib.redirect_stdout(io.StringIO()):
                rf.save_mhd_fields_with_ghost(info, [0.0, 1.0, 0.0, 1.0], 0)
            out = np.fromfile(os.path.join(d, "bin_data", "mhd_data_0000"), dtype=np.float32).reshape(ny + 4, nx + 4, 8)
            raw_cfg = np.frombuffer(open(os.path.join(d, "bin_data", "mhd_config.dat"), "rb").read(), dtype=np.uint8)
            np.savez_compressed(os.path.join(HERE, f"reorganize_{tag}.npz"), fdata=fdata, mhd_data=out,
                                mhd_config_bytes=raw_cfg, nx=nx, ny=ny, lx=2.0, ly=1.5, lz=1.0, dt_out=0.1)
            print(tag, out.shape, len(raw_cfg), "bytes of mhd_config.dat")


if __name__ == "__main__":
    print(golden_norm())
    c = golden_conf()
    print(len(c), "conf keys")
    golden_reorganize()
