"""Run configuration: the conf.dat grammar of the reference and the five named workloads.

`ConfReader` mirrors read_config.f90:22-45 (forward-only substring scan per open file,
value after '=', -1.0 when the key is not found).  `build_params` collects what
read_particle_params (particle_module.f90:2778-2878), read_diagnostics_params
(diagnostics.f90:2049-2200), read_particle_boundary_conditions
(simulation_setup.f90:107-123) and the set_* calls (particle_module.f90:260-335) put into
module variables, as one POD `Params` for the C ABI.

WORKLOADS are BASELINE.json's configs C1..C5 on synthetic fields:
physics values from the cited reference files, field shapes from BASELINE.json.
"""
from __future__ import annotations

import dataclasses
import io

from .abi import Params


class ConfReader:
    """get_variable() semantics of read_config.f90 for a conf.dat text."""

    def __init__(self, text: str):
        self._lines = text.splitlines()
        self._pos = 0

    @classmethod
    def open(cls, path: str) -> "ConfReader":
        with io.open(path, "r") as f:
            return cls(f.read())

    def get(self, name: str, delimiter: str = "=") -> float:
        while self._pos < len(self._lines):
            line = self._lines[self._pos]
            self._pos += 1
            if name in line:
                body = line.split(";")[0] if ";" in line else line
                try:
                    return float(body[body.index(delimiter) + 1:].split()[0].replace("D", "E").replace("d", "e"))
                except (ValueError, IndexError):
                    return -1.0
        return -1.0


# command-line defaults of the driver (stochastic-mhd.f90:871-881 for drift, 635-637 nlgc)
CLI_DEFAULTS = dict(drift_param1=4.0e7, drift_param2=2.0e8, charge=-1, nlgc=0, kperp_kpara=0.01,
                    dpp_wave=0, dpp_shear=0, weak_scattering=1, tau0=1.0, check_drift_2d=0,
                    include_3rd_dim=0, time_interp=1, focused_transport=0, duu_init=1.0,
                    # acceleration surfaces (stochastic-mhd.f90:899-938)
                    acc_by_surface=0, surface_norm1="+y", surface_norm2="-y", surface2_existed=0,
                    is_intersection=0)


def surface_norm_code(s: str) -> int:
    """'+x' / '-z' ... -> sign * (axis + 1), the C ABI's encoding of surface_norm1/2.  As in
    acc_region_surface.f90 (44-50, 350-366) anything but 'x' / 'y' in the second character is z and
    anything but '+' in the first is the negative direction."""
    s = (str(s) + "  ")[:2]
    axis = {"x": 1, "y": 2}.get(s[1], 3)
    return axis if s[0] == "+" else -axis


def build_params(conf_text: str, mhd_cfg: dict, ndim: int, nframes: int = 1 << 30,
                 cli: dict | None = None, mpi_rank: int = 0, seed: int = 0x5DE2024) -> Params:
    """conf.dat text + mhd_config + CLI switches -> Params, in the reference's read order."""
    c = dict(CLI_DEFAULTS)
    c.update(cli or {})
    P = Params()
    P.ndim = ndim
    for k in ("nx", "ny", "nz"):
        setattr(P, k, int(mhd_cfg[k]))
    for k in ("dx", "dy", "dz", "xmin", "ymin", "zmin", "xmax", "ymax", "zmax", "lx", "ly", "lz"):
        setattr(P, k, float(mhd_cfg[k]))
    P.time_interp = int(c["time_interp"])

    # read_particle_params: one open, keys in this order (particle_module.f90:2790-2814)
    r = ConfReader(conf_text)
    P.b0 = r.get("b0")
    P.p0 = r.get("p0")
    P.pmin = r.get("pmin")
    P.pmax = r.get("pmax")
    P.momentum_dependency = int(r.get("momentum_dependency"))
    P.gamma_turb = r.get("gamma_turb")
    P.pindex = 3.0 - P.gamma_turb
    P.mag_dependency = int(r.get("mag_dependency"))
    P.kpara0 = r.get("kpara0")
    P.kret = r.get("kret")
    P.dt_min_rel = r.get("dt_min_rel")
    P.dt_max_rel = r.get("dt_max_rel")
    P.acc_region_flag = int(r.get("acc_region_flag"))
    acc = [r.get(k) for k in ("acc_xmin", "acc_xmax", "acc_ymin", "acc_ymax", "acc_zmin", "acc_zmax")]
    if P.acc_region_flag != 1:  # particle_module.f90:2868-2877
        acc = [0.0, 1.0, 0.0, 1.0, 0.0, 1.0]
    for i, v in enumerate(acc):
        P.acc_region[i] = v

    # read_diagnostics_params: a fresh open (diagnostics.f90:2060-2104)
    r = ConfReader(conf_text)
    P.npp_global = int(r.get("npp_global"))
    ft = int(c["focused_transport"])
    nmu_global = int(r.get("nmu_global"))
    P.nmu_global = nmu_global if ft else 1  # 1 for Parker transport, diagnostics.f90:2107-2111
    for k in range(4):
        s = P.local[k]
        dump_interval = int(r.get(f"dump_interval{k + 1}"))
        s.pmin = r.get(f"pmin{k + 1}")
        s.pmax = r.get(f"pmax{k + 1}")
        s.npbins = int(r.get(f"npbins{k + 1}"))
        nmu_k = int(r.get(f"nmu{k + 1}"))
        s.nmu = nmu_k if ft else 1  # diagnostics.f90:2124-2128
        s.rx = int(r.get(f"rx{k + 1}"))
        s.ry = int(r.get(f"ry{k + 1}"))
        s.rz = int(r.get(f"rz{k + 1}"))
        s.enabled = 1 if dump_interval < nframes else 0  # diagnostics.f90:2129

    # read_particle_boundary_conditions (simulation_setup.f90:107-123)
    r = ConfReader(conf_text)
    P.pbc[0] = int(r.get("pbcx"))
    P.pbc[1] = int(r.get("pbcy"))
    P.pbc[2] = int(r.get("pbcz"))

    P.dpp_wave = int(c["dpp_wave"])
    P.dpp_shear = int(c["dpp_shear"])
    P.weak_scattering = int(c["weak_scattering"])
    P.tau0 = float(c["tau0"])
    P.drift1 = float(c["drift_param1"])
    P.drift2 = float(c["drift_param2"])
    P.pcharge = int(c["charge"])
    P.check_drift_2d = int(c["check_drift_2d"])
    P.include_3rd_dim = int(c["include_3rd_dim"])
    P.nlgc = int(c["nlgc"])
    P.focused_transport = ft
    P.duu0 = float(c["duu_init"])  # set_duu_params, particle_module.f90:279-283
    P.kperp_kpara = float(c["kperp_kpara"])
    P.acc_by_surface = int(c["acc_by_surface"])
    P.surface_norm1 = surface_norm_code(c["surface_norm1"])
    P.surface_norm2 = surface_norm_code(c["surface_norm2"])
    P.surface2_existed = int(bool(c["surface2_existed"]))
    P.is_intersection = int(bool(c["is_intersection"]))
    P.seed = seed
    P.rng_mode = 0
    P.mpi_rank = mpi_rank
    P.strict_math = 0
    return P


CONF_TEMPLATE = """\
b0 = 1.0
p0 = {p0}
pmin = {pmin}
pmax = {pmax}
momentum_dependency = {momentum_dependency}
gamma_turb = {gamma_turb}
mag_dependency = {mag_dependency}
kpara0 = {kpara0}
kret = {kret}
dt_min = 1E-8
dt_min_rel = {dt_min_rel}
dt_max_rel = {dt_max_rel}
npp_global = {npp_global}
nmu_global = 32
dump_interval1 = {di1}
pmin1 = {pmin}
pmax1 = {pmax}
npbins1 = 12
nmu1 = 1
rx1 = {r1}
ry1 = {r1}
rz1 = {r1}
dump_interval2 = {di2}
pmin2 = {pmin}
pmax2 = {pmax}
npbins2 = 64
nmu2 = 1
rx2 = {r2}
ry2 = {r2}
rz2 = {r2}
dump_interval3 = {di3}
pmin3 = {pmin}
pmax3 = {pmax}
npbins3 = 32
nmu3 = 16
rx3 = {r3}
ry3 = {r3}
rz3 = {r3}
dump_interval4 = 10000
pmin4 = {pmin}
pmax4 = {pmax}
npbins4 = 32
nmu4 = 16
rx4 = 8
ry4 = 8
rz4 = 8
acc_region_flag = {acc_region_flag}
acc_xmin = 0.0
acc_xmax = 1.0
acc_ymin = 0.0
acc_ymax = 1.0
acc_zmin = 0.0
acc_zmax = 1.0
pbcx = {pbcx}
pbcy = {pbcy}
pbcz = {pbcz}
mpi_sizex = 1
mpi_sizey = 1
mpi_sizez = 1
"""


@dataclasses.dataclass
class Workload:
    """One of BASELINE.json's configs on a synthetic field."""
    name: str
    kind: str                  # mhd.KINDS key
    nx: int
    ny: int
    nz: int
    lx: float = 2.0
    ly: float = 2.0
    lz: float = 1.0
    dt_out: float = 0.1
    nptl: int = 1_000_000      # particles injected at the first interval
    nptl_max: int = 2_000_000
    dist_flag: int = 1
    power_index: float = 6.2
    particle_v0: float = 17.20195
    split_flag: int = 1
    split_ratio: float = 2.0
    pmin_split: float = 2.0
    inject_new_ptl: bool = False
    nsteps_interval: int = 100
    num_fine_steps: int = 1
    local_dist: bool = True
    conf: dict = dataclasses.field(default_factory=dict)
    cli: dict = dataclasses.field(default_factory=dict)
    source: str = ""

    @property
    def ndim(self) -> int:
        return 3 if self.kind.endswith("3d") else (1 if self.kind.endswith("1d") else 2)

    def conf_text(self) -> str:
        d = dict(p0=0.1, pmin=1.0e-2, pmax=1.0e1, momentum_dependency=1, gamma_turb=1.6666667,
                 mag_dependency=1, kpara0=0.01, kret=0.03, dt_min_rel=1e-7, dt_max_rel=1e-2,
                 npp_global=128, di1=1, di2=1, di3=1, r1=4, r2=8, r3=16, acc_region_flag=0,
                 pbcx=0, pbcy=0, pbcz=0)
        d.update(self.conf)
        return CONF_TEMPLATE.format(**d)

    def scaled(self, grid: int | None = None, nptl: int | None = None) -> "Workload":
        """Same physics on a smaller grid / population (parity-test sizes)."""
        w = dataclasses.replace(self)
        if grid is not None:
            w.nx = grid
            if w.ndim >= 2:
                w.ny = grid
            if w.ndim == 3:
                w.nz = grid
        if nptl is not None:
            w.nptl = nptl
            w.nptl_max = max(2 * nptl, 16)
        return w


_DRIFT = dict(drift_param1=850964.408, drift_param2=13575468.975, charge=-1)

WORKLOADS = {
    # 1-D shock (push_particle_1d, particle_module.f90:2993-3111); mag_dependency must be 0 there
    # because the reference's 1-D kappa branch reads an unassigned db_dx otherwise (:2274)
    "s1": Workload("s1_shock_1d", "shock_1d", 2048, 1, 1, lx=8.0, ly=1.0,
                   conf=dict(momentum_dependency=1, gamma_turb=3.0, mag_dependency=0,
                             kpara0=8.0 / 2048 * 10 * 1.0, kret=0.03, pbcx=1, r1=4, r2=8, r3=16),
                   cli=dict(_DRIFT), source="config/shock.sh:52 in 1-D"),
    # C1: examples/reconnection_2d/conf_reconnection.dat + diffusion_reconnection.sh:181-194
    "c1": Workload("c1_reconnection_2d", "reconnection_2d", 1024, 1024, 1,
                   conf=dict(kpara0=0.00743592, kret=0.01), cli=dict(_DRIFT, tau0=7.53877e-5),
                   source="examples/reconnection_2d/conf_reconnection.dat, diffusion_reconnection.sh:181-194"),
    # C2: config/conf_flarecs.dat + config/diffusion_flarecs.sh:48-65 (open x,y)
    "c2": Workload("c2_flare_2d", "flare_2d", 4096, 4096, 1, nptl=100_000_000, nptl_max=150_000_000,
                   conf=dict(kpara0=0.004, kret=0.01, dt_min_rel=5e-5, pbcx=1, pbcy=1, r1=16, r2=32, r3=64),
                   cli=dict(_DRIFT), source="config/conf_flarecs.dat, config/diffusion_flarecs.sh:48-65"),
    # C3: config/shock.sh:52 (pindex 0 -> gamma_turb 3 in today's grammar), open x
    "c3": Workload("c3_shock_2d", "shock_2d", 2048, 256, 1, lx=8.0, ly=1.0,
                   conf=dict(momentum_dependency=1, gamma_turb=3.0, mag_dependency=0,
                             kpara0=8.0 / 2048 * 10 * 1.0, kret=0.03, pbcx=1, r1=4, r2=8, r3=16),
                   cli=dict(_DRIFT), source="config/shock.sh:52 (kpara0 rescaled so kappa/u = 10 dx)"),
    # C4: config/pic_diffusion.sh:56 + momentum diffusion (wave + shear)
    "c4": Workload("c4_pic_2d", "turbulence_2d", 4096, 4096, 1, nptl=10_000_000, nptl_max=12_000_000,
                   split_flag=0,
                   conf=dict(kpara0=0.1 * (1024.0 / 4096.0) ** 2, kret=0.05, r1=16, r2=32, r3=64),
                   cli=dict(_DRIFT, dpp_wave=1, dpp_shear=1, weak_scattering=1, tau0=7.53877e-5),
                   source="config/pic_diffusion.sh:56 (+ -dw 1 -ds 1)"),
    # C5: config/diffusion_fluxrope.sh:52, 3-D
    "c5": Workload("c5_fluxrope_3d", "fluxrope_3d", 512, 512, 512, lx=1.0, ly=1.0, lz=1.0,
                   nptl=125_000_000, nptl_max=150_000_000,
                   conf=dict(momentum_dependency=0, kpara0=0.0043 * (64.0 / 512.0), kret=0.02,
                             r1=8, r2=16, r3=32),
                   cli=dict(_DRIFT), source="config/diffusion_fluxrope.sh:52"),
}
