"""Host-side mirror of the reference driver for the GPU path.

`GpatSim` is a one-to-one wrapper of the C ABI (include/gpat_cuda.h): the method names are
the reference procedure names (particle_module / diagnostics public lists,
particle_module.f90:16-36, diagnostics.f90:24-30).  `run_intervals` reproduces the order
of calls of solve_transport_equation (stochastic-mhd.f90:312-567) for one rank.

Everything numerical happens inside libgpat_cuda.so; this file only sequences calls
and owns the host buffers, exactly like the Fortran driver would.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import abi
from .abi import PARTICLE_DTYPE, Counters, Params, Timings, ptr


class GpatError(RuntimeError):
    pass


class GpatSim:
    def __init__(self, params: Params, nptl_max: int, device: int = 0, lib_path: str | None = None):
        self.lib = abi.load_library(lib_path)
        self.P = params.copy()
        self.nptl_max = int(nptl_max)
        self.h = C.c_void_p()
        rc = self.lib.gpat_init(C.byref(self.h), device, self.nptl_max, C.byref(self.P))
        if rc != abi.GPAT_OK:
            msg = self.lib.gpat_last_error(None)
            self.h = None
            raise GpatError(f"gpat_init failed ({rc}): {msg.decode() if msg else ''}")

    # ---- plumbing -----------------------------------------------------------
    def _ck(self, rc: int, what: str):
        if rc != abi.GPAT_OK:
            msg = self.lib.gpat_last_error(self.h)
            raise GpatError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.gpat_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def grid_shape(self):
        P = self.P
        nzg = P.nz + 4 if P.ndim > 2 else 1
        nyg = P.ny + 4 if P.ndim > 1 else 1
        return nzg, nyg, P.nx + 4

    def set_params(self, params: Params):
        self._ck(self.lib.gpat_set_params(self.h, C.byref(params)), "gpat_set_params")
        self.P = params.copy()

    # ---- fields -------------------------------------------------------------
    def upload_turbulence(self, which: int, slot: int, slab: np.ndarray, two_d: np.ndarray):
        """read_magnetic_fluctuation (which = 0) / read_correlation_length (which = 1) and their gradient
        passes (mhd_data_parallel.f90:306-497, 771-1604); slab, two_d: float32 over the ghosted grid."""
        data = np.ascontiguousarray(np.stack([slab, two_d]), dtype=np.float32)
        if data.size != 2 * int(np.prod(self.grid_shape)):
            raise ValueError("turbulence map has the wrong size")
        self._ck(self.lib.gpat_upload_turbulence(self.h, which, slot, ptr(data)), "gpat_upload_turbulence")

    def surface_shape(self, which: int):
        """(n2, n1) C-order shape of acc_surfaceK1/K2 (acc_region_surface.f90:33-39): the two grid axes
        other than the surface's normal, ghosted."""
        P = self.P
        axis = abs(P.surface_norm2 if which else P.surface_norm1) - 1
        n1 = P.ny + 4 if axis == 0 else P.nx + 4
        n2 = P.ny + 4 if axis == 2 else P.nz + 4
        return n2, n1

    def upload_acc_surface(self, which: int, slot: int, heights: np.ndarray):
        """read_acc_surface (acc_region_surface.f90:91-206), in memory: heights of surface `which` (0, 1)
        for frame slot 0 (acc_surfaceK1) or 1 (acc_surfaceK2), float64 over the ghosted plane."""
        heights = np.ascontiguousarray(heights, dtype=np.float64)
        if heights.shape != self.surface_shape(which):
            raise ValueError(f"surface {which} has shape {heights.shape}, {self.surface_shape(which)} expected")
        self._ck(self.lib.gpat_upload_acc_surface(self.h, which, slot, ptr(heights)), "gpat_upload_acc_surface")

    def prefetch_fields(self, f: np.ndarray):
        """Start the H2D copy of a frame that a later upload_fields(slot, f) will pack; `f` must be
        C-contiguous float32 (ideally page-locked) and must not change until then."""
        if f.dtype != np.float32 or not f.flags["C_CONTIGUOUS"]:
            raise ValueError("prefetch_fields needs a C-contiguous float32 array (no temporary copies)")
        nvar = f.shape[-1]
        if f.size != int(np.prod(self.grid_shape)) * nvar:
            raise ValueError(f"field array has {f.size} floats, grid {self.grid_shape} x {nvar} expected")
        self._ck(self.lib.gpat_prefetch_fields(self.h, f.ctypes.data_as(C.c_void_p), nvar), "gpat_prefetch_fields")

    def upload_fields(self, slot: int, f: np.ndarray, with_grad: int = 0):
        f = np.ascontiguousarray(f, dtype=np.float32)
        nvar = f.shape[-1]
        if f.size != int(np.prod(self.grid_shape)) * nvar:
            raise ValueError(f"field array has {f.size} floats, grid {self.grid_shape} x {nvar} expected")
        self._ck(self.lib.gpat_upload_fields(self.h, slot, ptr(f), nvar, with_grad), "gpat_upload_fields")

    def swap_fields(self):
        self._ck(self.lib.gpat_swap_fields(self.h), "gpat_swap_fields")

    def debug_gradients(self, f8: np.ndarray) -> np.ndarray:
        f8 = np.ascontiguousarray(f8, dtype=np.float32)
        out = np.empty(f8.shape[:-1] + (32,), dtype=np.float32)
        self._ck(self.lib.gpat_debug_gradients(self.h, ptr(f8), ptr(out)), "gpat_debug_gradients")
        return out

    def interp(self, x, y, z, rt) -> np.ndarray:
        x, y, z, rt = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, rt))
        out = np.empty((len(x), 32), dtype=np.float64)
        self._ck(self.lib.gpat_debug_interp(self.h, len(x), ptr(x), ptr(y), ptr(z), ptr(rt), ptr(out)),
                 "gpat_debug_interp")
        return out

    # ---- particles ----------------------------------------------------------
    def inject_uniform(self, nptl, dt, dist_flag, particle_v0, t_frame, dt_mhd, part_box, power_index):
        box = (C.c_double * 6)(*part_box)
        self._ck(self.lib.gpat_inject_uniform(self.h, nptl, dt, dist_flag, particle_v0, t_frame, dt_mhd,
                                              box, power_index), "gpat_inject_uniform")

    def inject_targeted(self, mode, nptl, dt, dist_flag, particle_v0, t_frame, dt_mhd, part_box, power_index,
                        inject_same_nptl=True, vmin=0.0, ncells_norm=1):
        """inject_particles_at_large_jz/_absj/_divv/_rho (particle_module.f90:785-1468); returns
        (nptl_inject, ncells)."""
        box = (C.c_double * 6)(*part_box)
        ninj, ncells = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.gpat_inject_targeted(self.h, mode, nptl, dt, dist_flag, particle_v0, t_frame, dt_mhd,
                                               box, power_index, int(bool(inject_same_nptl)), vmin,
                                               ncells_norm, C.byref(ninj), C.byref(ncells)),
                 "gpat_inject_targeted")
        return ninj.value, ncells.value

    def inject_at_shock(self, nptl, dt, dist_flag, particle_v0, t_frame, power_index):
        """locate_shock_xpos + inject_particles_at_shock (mhd_data_parallel.f90:1988, particle_module.f90:542)."""
        self._ck(self.lib.gpat_inject_at_shock(self.h, nptl, dt, dist_flag, particle_v0, t_frame, power_index),
                 "gpat_inject_at_shock")

    # ---- particle tracking ----------------------------------------------------
    def init_tracking(self, tags: np.ndarray, nsteps_interval: int):
        """init_particle_tracking (particle_module.f90:5825-5879); tags: (nptl_tracking, split_times_max+2)
        int32 rows as written by tracking.select_tags (== the Fortran (ncols, nptl) array)."""
        tags = np.ascontiguousarray(tags, dtype=np.int32)
        self._ck(self.lib.gpat_init_tracking(self.h, ptr(tags), tags.shape[1], tags.shape[0], nsteps_interval),
                 "gpat_init_tracking")

    def download_tracked(self) -> np.ndarray:
        """particles_tracked as (nptl_tracking, nsteps_tracking_max) records."""
        nmax, ntrk = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.gpat_tracked_shape(self.h, C.byref(nmax), C.byref(ntrk)), "gpat_tracked_shape")
        out = np.zeros((ntrk.value, nmax.value), dtype=PARTICLE_DTYPE)
        self._ck(self.lib.gpat_download_tracked(self.h, ptr(out)), "gpat_download_tracked")
        return out

    def reset_tracked(self):
        self._ck(self.lib.gpat_reset_tracked(self.h), "gpat_reset_tracked")

    def particle_mover(self, t0, dtf, nsteps_interval=100, num_fine_steps=1, dump_escaped_dist=0) -> int:
        steps = C.c_uint64(0)
        self._ck(self.lib.gpat_particle_mover(self.h, t0, dtf, nsteps_interval, num_fine_steps,
                                              dump_escaped_dist, C.byref(steps)), "gpat_particle_mover")
        return steps.value

    def debug_push_n(self, t0, dtf, nsteps) -> int:
        steps = C.c_uint64(0)
        self._ck(self.lib.gpat_debug_push_n(self.h, t0, dtf, nsteps, C.byref(steps)), "gpat_debug_push_n")
        return steps.value

    def split(self, split_ratio, pmin_split, nsteps_interval=100):
        self._ck(self.lib.gpat_split(self.h, split_ratio, pmin_split, nsteps_interval), "gpat_split")

    def download_particles(self) -> np.ndarray:
        n = C.c_int64(0)
        self._ck(self.lib.gpat_download_particles(self.h, None, 0, C.byref(n)), "gpat_download_particles")
        out = np.zeros(max(n.value, 1), dtype=PARTICLE_DTYPE)
        self._ck(self.lib.gpat_download_particles(self.h, ptr(out), n.value, C.byref(n)),
                 "gpat_download_particles")
        return out[:n.value]

    def upload_particles(self, ptl: np.ndarray):
        ptl = np.ascontiguousarray(ptl, dtype=PARTICLE_DTYPE)
        self._ck(self.lib.gpat_upload_particles(self.h, ptr(ptl) if len(ptl) else None, len(ptl)),
                 "gpat_upload_particles")

    def download_escaped(self) -> np.ndarray:
        n = C.c_int64(0)
        self._ck(self.lib.gpat_download_escaped(self.h, None, 0, C.byref(n)), "gpat_download_escaped")
        out = np.zeros(max(n.value, 1), dtype=PARTICLE_DTYPE)
        self._ck(self.lib.gpat_download_escaped(self.h, ptr(out), n.value, C.byref(n)),
                 "gpat_download_escaped")
        return out[:n.value]

    def reset_escaped(self):
        self._ck(self.lib.gpat_reset_escaped(self.h), "gpat_reset_escaped")

    def counters(self) -> Counters:
        c = Counters()
        self._ck(self.lib.gpat_get_counters(self.h, C.byref(c)), "gpat_get_counters")
        return c

    def set_counters(self, c: Counters):
        self._ck(self.lib.gpat_set_counters(self.h, C.byref(c)), "gpat_set_counters")

    def set_rng_table(self, u: np.ndarray):
        u = np.ascontiguousarray(u, dtype=np.float64)
        self._ck(self.lib.gpat_set_rng_table(self.h, ptr(u), u.shape[0], u.shape[1]), "gpat_set_rng_table")

    # ---- diagnostics --------------------------------------------------------
    def local_shape(self, k: int):
        P, s = self.P, self.P.local[k]
        nrx = (P.nx + s.rx - 1) // s.rx
        nry = (P.ny + s.ry - 1) // s.ry
        nrz = (P.nz + s.rz - 1) // s.rz
        return nrz, nry, nrx, s.npbins, s.nmu  # C-order view of Fortran (nmu,np,nrx,nry,nrz)

    def alloc_diagnostics(self):
        """Host arrays with the shapes of diagnostics.f90:182-191, 235-245."""
        P = self.P
        fglobal = np.zeros((P.npp_global, P.nmu_global), dtype=np.float64)
        flocal = [np.zeros(self.local_shape(k), dtype=np.float64) if P.local[k].enabled else None
                  for k in range(4)]
        return fglobal, flocal

    def diagnostics(self, local_dist: bool = True, out=None):
        fglobal, flocal = out if out is not None else self.alloc_diagnostics()
        ptrs = (C.c_void_p * 4)(*[ptr(a) if a is not None else None for a in flocal])
        quick = np.zeros(8, dtype=np.float64)
        pmax = C.c_double(0.0)
        self._ck(self.lib.gpat_diagnostics(self.h, int(local_dist), ptr(fglobal), ptrs, ptr(quick),
                                           C.byref(pmax)), "gpat_diagnostics")
        return dict(fglobal=fglobal, flocal=flocal, quick=quick, pmax=pmax.value)

    def escaped_local_shapes(self, k: int):
        """C-order shapes of fescaped{k+1}_x, _y, _z (Fortran (nmu, npbins, n1, n2, 2), diagnostics.f90:358-405);
        None for an axis the run does not have."""
        nrz, nry, nrx, npb, nmu = self.local_shape(k)
        P = self.P
        return ((2, nrz, nry, npb, nmu), (2, nrz, nrx, npb, nmu) if P.ndim > 1 else None,
                (2, nry, nrx, npb, nmu) if P.ndim > 2 else None)

    def escaped_local_diagnostics(self):
        """calc_escaped_distributions, local part (diagnostics.f90:956-1170): for each local set a dict
        {"x": ..., "y": ..., "z": ...} of the face spectra (None: disabled set / absent axis)."""
        P = self.P
        arrs = [[None] * 4 for _ in range(3)]
        for k in range(4):
            if P.local[k].enabled:
                for f, shp in enumerate(self.escaped_local_shapes(k)):
                    if shp is not None:
                        arrs[f][k] = np.zeros(shp, dtype=np.float64)
        ptrs = [(C.c_void_p * 4)(*[ptr(a) if a is not None else None for a in arrs[f]]) for f in range(3)]
        self._ck(self.lib.gpat_escaped_local_diagnostics(self.h, ptrs[0], ptrs[1], ptrs[2]),
                 "gpat_escaped_local_diagnostics")
        return [dict(x=arrs[0][k], y=arrs[1][k], z=arrs[2][k]) if P.local[k].enabled else None for k in range(4)]

    def escaped_diagnostics(self) -> np.ndarray:
        P = self.P
        out = np.zeros((2 * P.ndim, P.npp_global, P.nmu_global), dtype=np.float64)
        self._ck(self.lib.gpat_escaped_diagnostics(self.h, ptr(out)), "gpat_escaped_diagnostics")
        return out

    def hist_edges(self, which: int = 0):
        P = self.P
        np_ = P.local[which - 1].npbins if which else P.npp_global
        nmu = P.local[which - 1].nmu if which else P.nmu_global
        pe, me = np.zeros(np_ + 1), np.zeros(nmu + 1)
        self._ck(self.lib.gpat_hist_edges(self.h, which, ptr(pe), ptr(me)), "gpat_hist_edges")
        return pe, me

    # ---- multi-GPU ------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.gpat_comm_unique_id(buf)
        if rc != abi.GPAT_OK:
            msg = self.lib.gpat_last_error(None)
            raise GpatError(f"gpat_comm_unique_id failed ({rc}): {msg.decode() if msg else ''}")
        return buf.raw

    def comm_init(self, uid: bytes, nranks: int, rank: int):
        self._ck(self.lib.gpat_comm_init(self.h, uid, nranks, rank), "gpat_comm_init")

    def comm_destroy(self):
        self._ck(self.lib.gpat_comm_destroy(self.h), "gpat_comm_destroy")

    def timings(self) -> Timings:
        t = Timings()
        self._ck(self.lib.gpat_get_timings(self.h, C.byref(t)), "gpat_get_timings")
        return t


def run_intervals(sim, frames, tstamps, *, nptl, dist_flag=1, particle_v0=1.0, power_index=6.2,
                  part_box=None, inject_new_ptl=True, tmax_to_inject=1 << 30, split_flag=1,
                  split_ratio=2.0, pmin_split=2.0, nsteps_interval=100, num_fine_steps=1,
                  local_dist=True, dump_escaped_dist=False, dt_inject=0.0, on_interval=None,
                  inject_mode=0, inject_same_nptl=True, inject_min=0.0, ncells_norm=1,
                  track_tags=None, on_tracked=None, surfaces=None, maps=None, tmin=0, quota_seconds=None, tmax_mhd=1 << 30,
                  particle_data_dump=False, dump_escaped=False):
    """solve_transport_equation (stochastic-mhd.f90:312-567) for one rank.

    `sim` is a GpatSim (or the test oracle, which has the same methods); `frames` is a
    sequence (or callable frame -> ndarray) of 8-variable MHD frames; `tstamps[i]` is the
    time of frame i (tstamps_mhd, mhd_config.f90:263-271).  Returns the per-interval records
    the reference writes to quick.dat / pmax_global.dat / fdists_NNNN.h5.  With acc_by_surface,
    `surfaces(which, frame)` returns the float64 heights of acceleration surface `which` at that frame
    (the content of <surface_filenameK>_NNNN.dat).  `tmin` > 0 continues a run restored with
    read_restart() (frames are still indexed from the run's t_start = 0); `quota_seconds` ends the loop after
    the first interval that finishes beyond it (reached_quota, :558-565); the last record's "frame" is the
    frame to hand to dump_restart().  Past `tmax_mhd` no new frame is read (:400): the
    last one is sent to slot 1 again, because swap_fields exchanges the two device halves where the
    reference's copy_fields leaves farray2 in place.  `particle_data_dump` / `dump_escaped` add the
    population (dump_particles, :491, 525-527) and the interval's escapees (dump_escaped_particles, :529-531)
    to every record as "particles" / "escaped_particles".
    """
    get = frames if callable(frames) else (lambda i: frames[i])
    P = sim.P
    track = track_tags is not None
    if track:                                                  # :226-229 init_particle_tracking
        sim.init_tracking(track_tags, nsteps_interval)
    if part_box is None:  # stochastic-mhd.f90:384-390
        part_box = [P.xmin, P.ymin, P.zmin, P.xmax, P.ymax, P.zmax]
    nframes = len(tstamps)
    records = []
    start = time.time()
    sim.upload_fields(0, get(tmin))                            # :320-321, :350 (mhd_data_<tmin>)
    nsurf = (2 if P.surface2_existed else 1) if P.acc_by_surface else 0
    if nsurf and surfaces is None:
        raise ValueError("acc_by_surface needs the `surfaces` callable")
    for k in range(nsurf):                                     # :323-334 read_acc_surface(0, ...)
        sim.upload_acc_surface(k, 0, surfaces(k, tmin))
    # turbulence maps (:336-347, 413-420): `maps(which, frame)` -> (slab, two_d) of deltab_NNNN (which = 0) /
    # lc_NNNN (which = 1).  They swap with the fields, so they must be re-sent every frame like them.
    which_maps = [w for w, on in ((0, P.deltab_flag), (1, P.correlation_flag)) if on]
    if which_maps and maps is None:
        raise ValueError("deltab_flag / correlation_flag need the `maps` callable (the maps swap with the fields "
                         "every interval and would go stale)")
    for wm in which_maps:
        sim.upload_turbulence(wm, 0, *maps(wm, tmin))
    total_steps = 0
    for tf in range(tmin + 1, nframes):                        # :397
        # read_field_data_parallel(..., var_flag=time_interp_flag): without time interpolation
        # the new frame REPLACES farray1 (:404-406, :426)
        if tf <= tmax_mhd or not P.time_interp:
            fr = min(tf, tmax_mhd)
        elif tf > tmin + 1:
            fr = tmax_mhd
        else:
            fr = None
        if fr is not None:
            sim.upload_fields(1 if P.time_interp else 0, get(fr))
            for k in range(nsurf):                             # :405-416 read_acc_surface(time_interp_flag, ...)
                sim.upload_acc_surface(k, 1 if P.time_interp else 0, surfaces(k, fr))
            for wm in which_maps:                              # :413-420 read_magnetic_fluctuation / _correlation_length
                sim.upload_turbulence(wm, 1 if P.time_interp else 0, *maps(wm, fr))
        t0, dtf = tstamps[tf - 1], tstamps[tf] - tstamps[tf - 1]
        if inject_mode == 6:                                   # :451-454 inject_at_shock: EVERY frame, whatever
            sim.inject_at_shock(nptl, dt_inject, dist_flag, particle_v0, t0, power_index)   # -in / tmax_to_inject say
        elif (tf == 1 or inject_new_ptl) and tf <= tmax_to_inject:   # :462-485
            if inject_mode:                                    # :464-480 (inject_large_jz ... _rho)
                sim.inject_targeted(inject_mode, nptl, dt_inject, dist_flag, particle_v0, t0, dtf, part_box,
                                    power_index, inject_same_nptl, inject_min, ncells_norm)
            else:
                sim.inject_uniform(nptl, dt_inject, dist_flag, particle_v0, t0, dtf, part_box, power_index)
        if tf == 1 and not track:                              # :488-494
            d0 = sim.diagnostics(local_dist)
            d0["frame"] = 0
            if particle_data_dump:
                d0["particles"] = sim.download_particles()
            records.append(d0)
        # a tracking run moves the particles with num_fine_steps = 1 (:497-503)
        steps = sim.particle_mover(t0, dtf, nsteps_interval, 1 if track else num_fine_steps,
                                   int(dump_escaped_dist))
        total_steps += steps
        if track:                                              # :509-511 dump_tracked_particles
            if on_tracked is not None:
                on_tracked(tf, sim.download_tracked())
            sim.reset_tracked()
        if split_flag == 1:
            sim.split(split_ratio, pmin_split, nsteps_interval)  # :515
        if track:                                              # no distributions in a tracking run (:516)
            d = dict(frame=tf, steps=steps)
            records.append(d)
            if P.time_interp:
                sim.swap_fields()
            if quota_seconds is not None and time.time() - start > quota_seconds:
                break
            continue
        d = sim.diagnostics(local_dist)                        # :518-521
        d["frame"] = tf
        d["steps"] = steps
        if particle_data_dump:
            d["particles"] = sim.download_particles()
        if dump_escaped_dist:
            if dump_escaped:
                d["escaped_particles"] = sim.download_escaped()
            d["fescaped"] = sim.escaped_diagnostics()
            if local_dist:
                d["fescaped_local"] = sim.escaped_local_diagnostics()
            sim.reset_escaped()                                # :533
        records.append(d)
        if on_interval is not None:
            on_interval(tf, d)
        if P.time_interp:
            sim.swap_fields()                                  # :538
        if quota_seconds is not None and time.time() - start > quota_seconds:   # :558-565
            break
    return records, total_steps



def dump_restart(sim, diagnostics_directory: str, t_end: int, tf: int) -> None:
    """The restart files the reference always writes when a run ends (stochastic-mhd.f90:252-271):
    dump_particles(t_end) + save_particle_module_state(t_end) into <dir>/restart/ and latest_restart = tf.
    (The files carry t_end in their names and latest_restart the last finished frame, as in the reference,
    so a run cut short by its quota can only be resumed after renaming -- kept as it is there.)  The
    reference's HDF5 containers are replaced by raw records (no HDF5 in this image):
      particles_NNNN.bin               int64 nptl, then nptl x gpat_particle (104 B; the Philox step counter
                                       rides in `padding`, so no separate save_prng file is needed)
      particle_module_state_NNNN.bin   gpat_counters (nptl_current, nptl_split, nptl_escaped, nptl_max, tag_max,
                                       leak, leak_negp)
      latest_restart                   int32 tf (the reference's own format)"""
    d = os.path.join(diagnostics_directory, "restart")
    os.makedirs(d, exist_ok=True)
    ptl = sim.download_particles()
    with open(os.path.join(d, f"particles_{t_end:04d}.bin"), "wb") as f:
        np.array([len(ptl)], dtype=np.int64).tofile(f)
        ptl.tofile(f)
    with open(os.path.join(d, f"particle_module_state_{t_end:04d}.bin"), "wb") as f:
        f.write(bytes(sim.counters()))
    np.array([tf], dtype=np.int32).tofile(os.path.join(d, "latest_restart"))


def read_restart(sim, diagnostics_directory: str) -> int:
    """restart_flag (stochastic-mhd.f90:224-236): tmin from latest_restart, then read_particles(tmin) and
    read_particle_module_state(tmin).  Returns tmin for run_intervals(..., tmin=tmin)."""
    from . import outputs
    d = os.path.join(diagnostics_directory, "restart")
    tmin = int(np.fromfile(os.path.join(d, "latest_restart"), dtype=np.int32)[0])
    ptl = outputs.read_particles(os.path.join(d, f"particles_{tmin:04d}.bin"))
    c = outputs.read_module_state(os.path.join(d, f"particle_module_state_{tmin:04d}.bin"))
    sim.upload_particles(ptl)
    sim.set_counters(c)
    return tmin
