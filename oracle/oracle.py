"""ctypes wrapper of the CPU oracle (oracle/gpat_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
Pinned bit for bit to golden vectors computed by executing the reference's own Fortran
(oracle/f90/, tests/golden/ref_f90/, tests/test_cpu_reference_f90.py; see gpat_oracle.c header).

The method names mirror stochastic_parker_b200.driver.GpatSim so parity tests
read the same on both sides.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from stochastic_parker_b200.abi import PARTICLE_DTYPE, Counters, Params, ptr  # noqa: E402  (POD layouts only)

NVAR = 32


def build(fast_native: bool = False) -> None:
    """make -C oracle (gcc only). fast_native rebuilds the timing copy for this host's CPU."""
    args = ["make", "-s", "-C", _HERE]
    if fast_native:
        subprocess.run(["rm", "-f", os.path.join(_HERE, "liborc_fast.so")], check=True)
        args.append("FAST_ARCH=native")
    subprocess.run(args, check=True)


def _load(fast: bool) -> C.CDLL:
    path = os.path.join(_HERE, "liborc_fast.so" if fast else "liborc.so")
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.POINTER(Params), C.c_int64]
    lib.orc_get_particles.restype = C.c_int64
    lib.orc_get_escaped.restype = C.c_int64
    lib.orc_total_steps.restype = C.c_uint64
    lib.orc_num_threads.restype = C.c_int
    return lib


class Oracle:
    def __init__(self, params: Params, nptl_max: int, fast: bool = False):
        self.lib = _load(fast)
        self.P = params.copy()
        self.nptl_max = int(nptl_max)
        self.h = C.c_void_p(self.lib.orc_create(C.byref(self.P), C.c_int64(nptl_max)))
        self._table = None

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # ---- shapes -----------------------------------------------------------
    @property
    def grid_shape(self):
        P = self.P
        nxg = P.nx + 4
        nyg = P.ny + 4 if P.ndim > 1 else P.ny
        nzg = P.nz + 4 if P.ndim > 2 else P.nz
        return nzg, nyg, nxg

    def set_params(self, params: Params):
        self.P = params.copy()
        self.lib.orc_set_params(self.h, C.byref(self.P))

    # ---- fields -----------------------------------------------------------
    def upload_fields(self, slot: int, f: np.ndarray, with_grad: int = 0):
        f = np.ascontiguousarray(f, dtype=np.float32)
        nvar = f.shape[-1]
        self.lib.orc_set_fields(self.h, C.c_int(slot), ptr(f), C.c_int(nvar), C.c_int(with_grad))
        if not with_grad:
            self.lib.orc_calc_gradients(self.h, C.c_int(slot))

    def upload_turbulence(self, which: int, slot: int, slab: np.ndarray, two_d: np.ndarray):
        """read_magnetic_fluctuation (which = 0) / read_correlation_length (1) + their gradient passes;
        slab and two_d are float32 arrays over the ghosted grid (the two halves of the reference's file)."""
        data = np.ascontiguousarray(np.stack([slab, two_d]), dtype=np.float32)
        if data.size != 2 * int(np.prod(self.grid_shape)):
            raise ValueError("turbulence map has the wrong size")
        self.lib.orc_set_turbulence(self.h, C.c_int(which), C.c_int(slot), ptr(data))

    def upload_acc_surface(self, which: int, slot: int, heights: np.ndarray):
        """acc_surfaceK1 (slot 0) / acc_surfaceK2 (slot 1) of surface `which`, float64 (n2, n1) C-order."""
        heights = np.ascontiguousarray(heights, dtype=np.float64)
        self.lib.orc_set_acc_surface(self.h, C.c_int(which), C.c_int(slot), ptr(heights))

    def get_fields(self, slot: int) -> np.ndarray:
        out = np.empty(self.grid_shape + (NVAR,), dtype=np.float32)
        self.lib.orc_get_fields(self.h, C.c_int(slot), ptr(out))
        return out

    def swap_fields(self):
        self.lib.orc_copy_fields(self.h)

    def interp(self, x, y, z, rt) -> np.ndarray:
        x, y, z, rt = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, rt))
        out = np.empty((len(x), NVAR), dtype=np.float64)
        self.lib.orc_interp(self.h, C.c_int64(len(x)), ptr(x), ptr(y), ptr(z), ptr(rt), ptr(out))
        return out

    # ---- particles ----------------------------------------------------------
    def inject_uniform(self, nptl, dt, dist_flag, particle_v0, t_frame, dt_mhd, part_box,
                       power_index):
        box = (C.c_double * 6)(*part_box)
        self.lib.orc_inject_uniform(self.h, C.c_int64(nptl), C.c_double(dt), C.c_int(dist_flag),
                                    C.c_double(particle_v0), C.c_double(t_frame),
                                    C.c_double(dt_mhd), box, C.c_double(power_index))

    def inject_targeted(self, mode, nptl, dt, dist_flag, particle_v0, t_frame, dt_mhd, part_box,
                        power_index, inject_same_nptl=True, vmin=0.0, ncells_norm=1):
        box = (C.c_double * 6)(*part_box)
        self.lib.orc_ncells_large.restype = C.c_int64
        self.lib.orc_inject_targeted.restype = C.c_int64
        ncells = self.lib.orc_ncells_large(self.h, C.c_int(mode), C.c_double(vmin), box)
        ninj = self.lib.orc_inject_targeted(self.h, C.c_int(mode), C.c_int64(nptl), C.c_double(dt),
                                            C.c_int(dist_flag), C.c_double(particle_v0),
                                            C.c_double(t_frame), C.c_double(dt_mhd), box,
                                            C.c_double(power_index), C.c_int(int(bool(inject_same_nptl))),
                                            C.c_double(vmin), C.c_int64(ncells_norm))
        return int(ninj), int(ncells)

    def inject_at_shock(self, nptl, dt, dist_flag, particle_v0, t_frame, power_index):
        self.lib.orc_inject_at_shock.restype = C.c_int64
        self.lib.orc_inject_at_shock(self.h, C.c_int64(nptl), C.c_double(dt), C.c_int(dist_flag),
                                     C.c_double(particle_v0), C.c_double(t_frame), C.c_double(power_index))

    def particle_mover(self, t0, dtf, nsteps_interval=100, num_fine_steps=1,
                       dump_escaped_dist=0) -> int:
        steps = C.c_uint64(0)
        self.lib.orc_particle_mover(self.h, C.c_double(t0), C.c_double(dtf),
                                    C.c_int(nsteps_interval), C.c_int(num_fine_steps),
                                    C.c_int(dump_escaped_dist), C.byref(steps))
        return steps.value

    def debug_push_n(self, t0, dtf, nsteps) -> int:
        steps = C.c_uint64(0)
        self.lib.orc_debug_push_n(self.h, C.c_double(t0), C.c_double(dtf), C.c_int(nsteps),
                                  C.byref(steps))
        return steps.value

    def split(self, split_ratio, pmin_split, nsteps_interval=100):
        self.lib.orc_split(self.h, C.c_double(split_ratio), C.c_double(pmin_split),
                           C.c_int(nsteps_interval))

    def download_particles(self) -> np.ndarray:
        out = np.zeros(self.nptl_max, dtype=PARTICLE_DTYPE)
        n = self.lib.orc_get_particles(self.h, ptr(out), C.c_int64(self.nptl_max))
        return out[:n].copy()

    def upload_particles(self, ptl: np.ndarray):
        ptl = np.ascontiguousarray(ptl, dtype=PARTICLE_DTYPE)
        self.lib.orc_set_particles(self.h, ptr(ptl), C.c_int64(len(ptl)))

    def download_escaped(self) -> np.ndarray:
        out = np.zeros(self.nptl_max, dtype=PARTICLE_DTYPE)
        n = self.lib.orc_get_escaped(self.h, ptr(out), C.c_int64(self.nptl_max))
        return out[:min(n, self.nptl_max)].copy()

    def reset_escaped(self):
        self.lib.orc_reset_escaped(self.h)

    # ---- particle tracking (PM:5825-5990) ----
    def init_tracking(self, tags: np.ndarray, nsteps_interval: int):
        tags = np.ascontiguousarray(tags, dtype=np.int32)  # (ntrack, ncols) C == (ncols, ntrack) Fortran
        self._ntrack = tags.shape[0]
        self.lib.orc_init_tracking(self.h, ptr(tags), C.c_int(tags.shape[1]), C.c_int64(tags.shape[0]),
                                   C.c_int(nsteps_interval))

    def download_tracked(self) -> np.ndarray:
        """particles_tracked as (nptl_tracking, nsteps_tracking_max) records."""
        nmax = C.c_int64(0)
        self.lib.orc_get_tracked.restype = C.c_int64
        n = self.lib.orc_get_tracked(self.h, None, C.byref(nmax))
        out = np.zeros((n, nmax.value), dtype=PARTICLE_DTYPE)
        self.lib.orc_get_tracked(self.h, ptr(out), C.byref(nmax))
        return out

    def reset_tracked(self):
        self.lib.orc_reset_tracked(self.h)

    def counters(self) -> Counters:
        c = Counters()
        self.lib.orc_get_counters(self.h, C.byref(c))
        return c

    def set_counters(self, c: Counters):
        self.lib.orc_set_counters(self.h, C.byref(c))

    def set_rng_table(self, u: np.ndarray):
        """u: (nslots, max_steps, 4) uniforms in [0,1]."""
        self._table = np.ascontiguousarray(u, dtype=np.float64)
        self.lib.orc_set_rng_table(self.h, ptr(self._table), C.c_int64(u.shape[0]),
                                   C.c_int64(u.shape[1]))

    # ---- diagnostics --------------------------------------------------------
    def local_shape(self, k: int):
        P, s = self.P, self.P.local[k]
        nrx = (P.nx + s.rx - 1) // s.rx
        nry = (P.ny + s.ry - 1) // s.ry
        nrz = (P.nz + s.rz - 1) // s.rz
        return nrz, nry, nrx, s.npbins, s.nmu  # C-order view of (nmu,np,nrx,nry,nrz)

    def diagnostics(self, local_dist: bool = True):
        P = self.P
        fglobal = np.zeros((P.npp_global, P.nmu_global), dtype=np.float64)
        flocal = [np.zeros(self.local_shape(k), dtype=np.float64) if P.local[k].enabled else None
                  for k in range(4)]
        ptrs = (C.c_void_p * 4)(*[ptr(a) if a is not None else None for a in flocal])
        quick = np.zeros(8, dtype=np.float64)
        pmax = C.c_double(0.0)
        self.lib.orc_diagnostics(self.h, C.c_int(int(local_dist)), ptr(fglobal), ptrs, ptr(quick),
                                 C.byref(pmax))
        return dict(fglobal=fglobal, flocal=flocal, quick=quick, pmax=pmax.value)

    def escaped_diagnostics(self) -> np.ndarray:
        P = self.P
        out = np.zeros((2 * P.ndim, P.npp_global, P.nmu_global), dtype=np.float64)
        self.lib.orc_escaped_diagnostics(self.h, ptr(out))
        return out

    def escaped_local_shapes(self, k: int):
        nrz, nry, nrx, npb, nmu = self.local_shape(k)
        P = self.P
        return ((2, nrz, nry, npb, nmu), (2, nrz, nrx, npb, nmu) if P.ndim > 1 else None,
                (2, nry, nrx, npb, nmu) if P.ndim > 2 else None)

    def escaped_local_diagnostics(self):
        P = self.P
        arrs = [[None] * 4 for _ in range(3)]
        for k in range(4):
            if P.local[k].enabled:
                for f, shp in enumerate(self.escaped_local_shapes(k)):
                    if shp is not None:
                        arrs[f][k] = np.zeros(shp, dtype=np.float64)
        ptrs = [(C.c_void_p * 4)(*[ptr(a) if a is not None else None for a in arrs[f]]) for f in range(3)]
        self.lib.orc_escaped_local_diagnostics(self.h, ptrs[0], ptrs[1], ptrs[2])
        return [dict(x=arrs[0][k], y=arrs[1][k], z=arrs[2][k]) if P.local[k].enabled else None for k in range(4)]

    def hist_edges(self, which: int = 0):
        P = self.P
        np_ = P.local[which - 1].npbins if which else P.npp_global
        nmu = P.local[which - 1].nmu if which else P.nmu_global
        pe = np.zeros(np_ + 1)
        me = np.zeros(nmu + 1)
        self.lib.orc_hist_edges(self.h, C.c_int(which), ptr(pe), ptr(me))
        return pe, me

    def num_threads(self) -> int:
        return int(self.lib.orc_num_threads())

    def set_num_threads(self, n: int):
        """OpenMP team size of the mover (bench.py's reference arm: all host cores, whatever OMP_NUM_THREADS says)"""
        self.lib.orc_set_num_threads(C.c_int(int(n)))


def philox4x32_10(ctr, key):
    lib = _load(False)
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]
