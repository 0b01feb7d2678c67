"""Readers of the files host/gpat_driver writes (and dump_restart): the raw stand-ins of the reference's
HDF5 outputs (no HDF5 library in this image).  One function per file family; every array comes back in the
C-order view of the reference's Fortran shape.

  quick.dat / pmax_global.dat            text, the reference's own formats (diagnostics.f90:158-168, 1709-1718)
  fdists_NNNN.bin                        int32 nmu, npp; fglobal(nmu, npp); pbins_edges(npp+1); mubins_edges(nmu+1)
  fdists_localK_NNNN.bin                 int32 nmu, npbins, nrx, nry, nrz; flocalK(nmu, npbins, nrx, nry, nrz)
  escaped_dists_NNNN.bin                 int32 nmu, npp, nface; fescaped(nmu, npp, nface)
  escaped_dists_localK_NNNN.bin          int32 nmu, npbins, nrx, nry, nrz, ndim; fescapedK_x, [_y, [_z]]
  particles_NNNN.bin, escaped_particles_NNNN.bin, restart/particles_NNNN.bin
                                         int64 n; n x gpat_particle (abi.PARTICLE_DTYPE)
  restart/particle_module_state_NNNN.bin gpat_counters; restart/latest_restart: int32
  particle_tracking_particles_tracked_NNNN.bin   int64 nptl_tracking, nsteps_tracking_max; the records
"""
from __future__ import annotations

import os

import numpy as np

from .abi import PARTICLE_DTYPE, Counters


def read_quick(directory: str) -> dict:
    """quick.dat -> {column: array}; the E13.6 columns have no separators (Fortran fixed format)."""
    lines = open(os.path.join(directory, "quick.dat")).read().splitlines()
    names = lines[0].split()
    cols = {n: [] for n in names}
    for l in lines[1:]:
        cols[names[0]].append(int(l[:6]))
        for k, n in enumerate(names[1:]):
            cols[n].append(float(l[6 + 13 * k:19 + 13 * k]))
    return {n: np.array(v) for n, v in cols.items()}


def read_pmax_global(directory: str) -> np.ndarray:
    return np.array([float(v) for v in open(os.path.join(directory, "pmax_global.dat")).read().split()])


def read_fdists(directory: str, frame: int) -> dict:
    raw = open(os.path.join(directory, f"fdists_{frame:04d}.bin"), "rb").read()
    nmu, npp = (int(v) for v in np.frombuffer(raw[:8], dtype=np.int32))
    body = np.frombuffer(raw[8:], dtype=np.float64)
    return dict(fglobal=body[:nmu * npp].reshape(npp, nmu), pbins_edges=body[nmu * npp:nmu * npp + npp + 1],
                mubins_edges=body[nmu * npp + npp + 1:])


def read_fdists_local(directory: str, k: int, frame: int) -> np.ndarray:
    """flocalK as (nrz, nry, nrx, npbins, nmu); k = 1..4."""
    raw = open(os.path.join(directory, f"fdists_local{k}_{frame:04d}.bin"), "rb").read()
    shp = tuple(int(v) for v in np.frombuffer(raw[:20], dtype=np.int32))
    return np.frombuffer(raw[20:], dtype=np.float64).reshape(shp[::-1])


def read_escaped_dists(directory: str, frame: int) -> np.ndarray:
    """fescaped as (nface, npp, nmu)."""
    raw = open(os.path.join(directory, f"escaped_dists_{frame:04d}.bin"), "rb").read()
    nmu, npp, nface = (int(v) for v in np.frombuffer(raw[:12], dtype=np.int32))
    return np.frombuffer(raw[12:], dtype=np.float64).reshape(nface, npp, nmu)


def read_escaped_dists_local(directory: str, k: int, frame: int) -> dict:
    """{"x": (2, nrz, nry, npbins, nmu), "y": (2, nrz, nrx, ...) or None, "z": (2, nry, nrx, ...) or None}."""
    raw = open(os.path.join(directory, f"escaped_dists_local{k}_{frame:04d}.bin"), "rb").read()
    nmu, npb, nrx, nry, nrz, ndim = (int(v) for v in np.frombuffer(raw[:24], dtype=np.int32))
    body = np.frombuffer(raw[24:], dtype=np.float64)
    shapes = [(2, nrz, nry, npb, nmu), (2, nrz, nrx, npb, nmu) if ndim > 1 else None,
              (2, nry, nrx, npb, nmu) if ndim > 2 else None]
    out, pos = {}, 0
    for name, shp in zip("xyz", shapes):
        if shp is None:
            out[name] = None
            continue
        n = int(np.prod(shp))
        out[name] = body[pos:pos + n].reshape(shp)
        pos += n
    return out


def read_particles(path: str) -> np.ndarray:
    """particles_NNNN.bin / escaped_particles_NNNN.bin / restart/particles_NNNN.bin."""
    with open(path, "rb") as f:
        n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
        ptl = np.fromfile(f, dtype=PARTICLE_DTYPE, count=n)
    if len(ptl) != n:
        raise IOError(f"{path} is truncated")
    return ptl


def read_module_state(path: str) -> Counters:
    return Counters.from_buffer_copy(open(path, "rb").read())


def read_tracked(directory: str, frame: int) -> np.ndarray:
    """particles_tracked as (nptl_tracking, nsteps_tracking_max) records."""
    raw = open(os.path.join(directory, f"particle_tracking_particles_tracked_{frame:04d}.bin"), "rb").read()
    ntrk, nmax = (int(v) for v in np.frombuffer(raw[:16], dtype=np.int64))
    return np.frombuffer(raw[16:], dtype=PARTICLE_DTYPE).reshape(ntrk, nmax)
