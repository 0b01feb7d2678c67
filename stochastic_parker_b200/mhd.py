"""Synthetic MHD frames in GPAT's on-disk format.

No MHD data ships with the reference, so every workload here is generated from
closed-form, time-dependent expressions evaluated in float64 and cast to float32,
then written exactly as the reference's pre-processing writes its files:

* ``mhd_data_NNNN``: float32, C-order ``(nz+4 [3-D only], ny+4, nx+4, 8)`` ==
  Fortran ``(8, nx+4, ny+4, nz+4|1)``; variables vx, vy, vz, rho, bx, by, bz, |B|;
  two ghost cells per side (examples/reconnection_2d/mhd_data/reorganize_fields.py:57,
  62, 82, 185-193).
* ghost fill per axis: ``periodic`` copies ``ghost_lo = phys[n-3:n-1]``,
  ``ghost_hi = phys[1:3]`` and ``reflect`` mirrors (reorganize_fields.py:134-142 and
  160-167; the reference's periodic fill treats phys[n-1] as a duplicate of phys[0],
  so the synthetic fields are given the period (n-1)*dx to be smooth across it).
* ``mhd_config.dat``: 13 float64 + 14 int32 (reorganize_fields.py:204-259,
  read back by mhd_config.f90:139-148 and python/sde_util.py:48-57).

Grid convention: physical sample i sits at x = xmin + i*dx (the push maps
x = xmin to Fortran index 1, particle_module.f90:1616 + 654).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

NVAR = 8


def _ghost_fill(a: np.ndarray, axis: int, mode: str) -> None:
    """In-place ghost fill of one spatial axis (length n+4) of a (..., 8) array."""
    n = a.shape[axis] - 4
    sl = [slice(None)] * a.ndim

    def S(s):
        q = list(sl)
        q[axis] = s
        return tuple(q)

    if mode == "periodic":  # reorganize_fields.py:134-137
        a[S(slice(0, 2))] = a[S(slice(n - 1, n + 1))]
        a[S(slice(n + 2, n + 4))] = a[S(slice(3, 5))]
    elif mode == "reflect":  # reorganize_fields.py:160-163
        a[S(slice(0, 2))] = a[S(slice(3, 1, -1))]
        a[S(slice(n + 2, n + 4))] = a[S(slice(n + 1, n - 1, -1))]
    else:
        raise ValueError(mode)


def _sech2(x):
    return 1.0 / np.cosh(x) ** 2


def _fields_reconnection_2d(X, Y, t):
    """Double Harris sheet + growing island perturbation + compressive flow.
    X, Y in [0,1] with the duplicate-endpoint convention (see module docstring)."""
    lam = 8.0 / X.shape[-1]
    thx, thy = 2 * np.pi * X, 2 * np.pi * Y
    eps = 0.1 * (1.0 + t)
    v0 = 0.5 * (1.0 + 0.1 * t)
    bx = np.tanh((Y - 0.25) / lam) - np.tanh((Y - 0.75) / lam) - 1.0 + eps * np.cos(thx) * np.cos(thy)
    by = eps * np.sin(thx) * np.sin(thy)
    bz = 0.2 + 0.05 * np.cos(thx) * np.cos(2 * thy)
    vx = -v0 * np.sin(thx) * np.cos(thy)
    vy = 0.5 * v0 * np.cos(thx) * np.sin(thy)
    vz = 0.1 * np.sin(thx + thy)
    rho = 1.0 + 0.5 * _sech2((Y - 0.25) / lam) + 0.5 * _sech2((Y - 0.75) / lam)
    return vx, vy, vz, rho, bx, by, bz


def _fields_flare_2d(X, Y, t):
    """Vertical flare current sheet at X=0.5 with inflow/outflow; open box."""
    lam = 6.0 / X.shape[-1]
    s = np.tanh((X - 0.5) / lam)
    vin = 0.05 * (1.0 + 0.2 * t)
    vout = 0.8 * (1.0 + 0.1 * t)
    by = s + 0.05 * np.sin(2 * np.pi * Y) * (1 - s * s)
    bx = 0.05 * (1.0 + t) * np.sin(np.pi * (X - 0.5)) * np.cos(2 * np.pi * Y) * _sech2((X - 0.5) / (4 * lam))
    bz = 0.1 + 0.3 * _sech2((X - 0.5) / lam)
    vx = -vin * s
    vy = vout * np.tanh((Y - 0.4) / 0.1) * _sech2((X - 0.5) / (3 * lam))
    vz = 0.0 * X
    rho = 1.0 + 1.5 * _sech2((X - 0.5) / lam)
    return vx, vy, vz, rho, bx, by, bz


def _fields_shock_2d(X, Y, t):
    """Planar shock moving along +x with compression ratio 4 (upstream on the left).  A weak
    smooth ripple covers the whole box: exactly uniform regions would make dx/dt, dy/dt or dp/dt
    exactly zero, and the reference then falls back to dt = dt_min (particle_module.f90:3532-3534),
    which no real MHD frame ever triggers."""
    w = 2.0 / X.shape[-1]
    xs = 0.45 + 0.05 * t
    u_u, u_d = 1.0, 0.25
    prof = 0.5 * (1.0 - np.tanh((X - xs) / w))  # 1 upstream, 0 downstream
    ripple = 0.02 * np.sin(2 * np.pi * Y)
    thx, thy = 2 * np.pi * X, 2 * np.pi * Y
    vx = u_d + (u_u - u_d) * prof + ripple * (1 - prof) + 0.004 * np.sin(thx + 0.3) * np.cos(thy)
    vy = 0.02 * np.cos(2 * np.pi * Y) * (1 - prof) + 0.003 * np.cos(thx) * np.sin(thy + 0.2)
    vz = 0.0 * X
    rho = 4.0 - 3.0 * prof
    bx = 0.5 + 0.01 * np.cos(thy) * np.sin(thx)
    by = 0.3 * rho + 0.01 * np.sin(thx + 0.7)
    bz = 0.05 * rho + 0.005 * np.cos(thx) * np.cos(thy)
    return vx, vy, vz, rho, bx, by, bz


def _fields_shock_1d(X, t):
    """1-D planar shock along x (the 1-D runs of push_particle_1d, particle_module.f90:2993):
    compression ratio 4, smooth ripple so that dx/dt and dp/dt are never exactly zero."""
    w = 2.0 / X.shape[-1]
    xs = 0.45 + 0.05 * t
    prof = 0.5 * (1.0 - np.tanh((X - xs) / w))
    thx = 2 * np.pi * X
    vx = 0.25 + 0.75 * prof + 0.004 * np.sin(thx + 0.3)
    rho = 4.0 - 3.0 * prof
    bx = 0.5 + 0.0 * X
    by = 0.3 * rho + 0.01 * np.sin(thx + 0.7)
    bz = 0.05 * rho + 0.005 * np.cos(thx)
    return vx, 0.0 * X, 0.0 * X, rho, bx, by, bz


def _fields_turbulence_2d(X, Y, t, nmodes=6, seed=7):
    """Multi-mode 'PIC-like' fluctuating field on a mean field."""
    rng = np.random.default_rng(seed)
    bx = 1.0 + 0.0 * X
    by = 0.0 * X
    bz = 0.3 + 0.0 * X
    vx = 0.0 * X
    vy = 0.0 * X
    vz = 0.0 * X
    rho = 1.0 + 0.0 * X
    for _ in range(nmodes):
        kx, ky = rng.integers(1, 6, size=2)
        ph = rng.uniform(0, 2 * np.pi, size=4)
        amp = 0.6 / np.hypot(kx, ky) ** (5.0 / 6.0)
        om = 0.7 * np.hypot(kx, ky)
        arg = 2 * np.pi * (kx * X + ky * Y)
        bx += amp * ky / np.hypot(kx, ky) * np.sin(arg + ph[0] + om * t)
        by -= amp * kx / np.hypot(kx, ky) * np.sin(arg + ph[0] + om * t)
        bz += 0.3 * amp * np.cos(arg + ph[1] - om * t)
        vx += 0.5 * amp * np.cos(arg + ph[2] + om * t)
        vy += 0.5 * amp * np.sin(arg + ph[3] - om * t)
        rho += 0.15 * amp * np.cos(arg + ph[1] + om * t)
    return vx, vy, vz, rho, bx, by, bz


def _fields_fluxrope_3d(X, Y, Z, t):
    """Flux rope along z with a kink that grows in time, plus a helical flow."""
    a2 = 0.15 ** 2
    xc = 0.5 + 0.05 * (1 + t) * np.sin(2 * np.pi * Z)
    yc = 0.5 + 0.05 * (1 + t) * np.cos(2 * np.pi * Z)
    rx, ry = X - xc, Y - yc
    den = rx * rx + ry * ry + a2
    bx = -0.15 * ry / den
    by = 0.15 * rx / den
    bz = 0.2 + a2 / den
    v0 = 0.3 * (1.0 + 0.1 * t)
    vx = -v0 * ry * np.exp(-den / (4 * a2)) + 0.05 * np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Z)
    vy = v0 * rx * np.exp(-den / (4 * a2)) + 0.05 * np.sin(2 * np.pi * Y)
    vz = 0.2 * v0 * np.exp(-den / (2 * a2)) + 0.05 * np.sin(2 * np.pi * Z)
    rho = 1.0 + 0.5 * a2 / den
    return vx, vy, vz, rho, bx, by, bz


KINDS = {
    "shock_1d": (_fields_shock_1d, 1, "reflect"),
    "reconnection_2d": (_fields_reconnection_2d, 2, "periodic"),
    "flare_2d": (_fields_flare_2d, 2, "reflect"),
    "shock_2d": (_fields_shock_2d, 2, "reflect"),
    "turbulence_2d": (_fields_turbulence_2d, 2, "periodic"),
    "fluxrope_3d": (_fields_fluxrope_3d, 3, "periodic"),
}


def _pack(out, comps):
    vx, vy, vz, rho, bx, by, bz = comps
    out[..., 0] = vx
    out[..., 1] = vy
    out[..., 2] = vz
    out[..., 3] = rho
    out[..., 4] = bx
    out[..., 5] = by
    out[..., 6] = bz
    out[..., 7] = np.sqrt(bx * bx + by * by + bz * bz)  # f64 sqrt then cast, reorganize_fields.py:62


def make_frame(kind: str, nx: int, ny: int, nz: int, frame: int, dt_out: float = 0.1,
               boundary: str | None = None) -> np.ndarray:
    """One frame with ghost cells, float32, shape (nx+4, 8), (ny+4, nx+4, 8) or
    (nz+4, ny+4, nx+4, 8)."""
    fn, ndim, default_bc = KINDS[kind]
    mode = boundary or default_bc
    t = frame * dt_out
    xs = np.arange(nx, dtype=np.float64) / max(nx - 1, 1)
    ys = np.arange(ny, dtype=np.float64) / max(ny - 1, 1)
    if ndim == 1:  # farray(:, -1:nx+2, 1, 1), mhd_data_parallel.f90:78
        out = np.zeros((nx + 4, NVAR), dtype=np.float32)
        _pack(out[2:nx + 2], fn(xs, t))
        _ghost_fill(out, 0, mode)
        return out
    # elementwise closed forms: row blocks (2-D) / planes (3-D) are evaluated by a thread pool
    # (numpy ufuncs release the GIL); the values do not depend on the blocking
    nthr = max(1, min(32, os.cpu_count() or 1))
    if ndim == 2:
        out = np.zeros((ny + 4, nx + 4, NVAR), dtype=np.float32)
        rows = max(16, -(-ny // (4 * nthr)))

        def block(j0):
            j1 = min(ny, j0 + rows)
            X, Y = np.meshgrid(xs, ys[j0:j1])  # (rows, nx)
            _pack(out[2 + j0:2 + j1, 2:nx + 2], fn(X, Y, t))

        with ThreadPoolExecutor(nthr) as ex:
            list(ex.map(block, range(0, ny, rows)))
        _ghost_fill(out, 0, mode)
        _ghost_fill(out, 1, mode)
        return out
    zs = np.arange(nz, dtype=np.float64) / max(nz - 1, 1)
    out = np.zeros((nz + 4, ny + 4, nx + 4, NVAR), dtype=np.float32)
    X, Y = np.meshgrid(xs, ys)

    def plane(k):  # plane by plane to bound memory at 512^3
        _pack(out[k + 2, 2:ny + 2, 2:nx + 2], fn(X, Y, zs[k] + 0.0 * X, t))

    with ThreadPoolExecutor(nthr) as ex:
        list(ex.map(plane, range(nz)))
    _ghost_fill(out, 0, mode)
    _ghost_fill(out, 1, mode)
    _ghost_fill(out, 2, mode)
    return out


def mhd_config(nx: int, ny: int, nz: int, lx: float, ly: float, lz: float, dt_out: float,
               ndim: int, bc: int = 0) -> dict:
    """The fields of `mhd_configuration` (mhd_config.f90:16-27) for a box at the origin."""
    return dict(dx=lx / nx, dy=ly / ny, dz=lz / nz, xmin=0.0, ymin=0.0, zmin=0.0, xmax=lx,
                ymax=ly, zmax=lz, lx=lx, ly=ly, lz=lz, dt_out=dt_out, nx=nx, ny=ny, nz=nz,
                nxs=nx, nys=ny, nzs=nz, topox=1, topoy=1, topoz=1, nvar=9, bcx=bc, bcy=bc,
                bcz=bc, ndim=ndim)


def write_mhd_config(path: str, cfg: dict) -> None:
    dbl = np.array([cfg[k] for k in ("dx", "dy", "dz", "xmin", "ymin", "zmin", "xmax", "ymax",
                                     "zmax", "lx", "ly", "lz", "dt_out")], dtype=np.float64)
    ints = np.zeros(14, dtype=np.int32)
    ints[:13] = [cfg[k] for k in ("nx", "ny", "nz", "nxs", "nys", "nzs", "topox", "topoy",
                                  "topoz", "nvar", "bcx", "bcy", "bcz")]
    with open(path, "wb") as f:
        dbl.tofile(f)
        ints.tofile(f)


def read_mhd_config(path: str) -> dict:
    """load_mhd_config (mhd_config.f90:139-148): 13 f64 + 13 i32 stream read."""
    raw = open(path, "rb").read()
    dbl = np.frombuffer(raw[:104], dtype=np.float64)
    ints = np.frombuffer(raw[104:104 + 52], dtype=np.int32)
    keys_d = ("dx", "dy", "dz", "xmin", "ymin", "zmin", "xmax", "ymax", "zmax", "lx", "ly", "lz",
              "dt_out")
    keys_i = ("nx", "ny", "nz", "nxs", "nys", "nzs", "topox", "topoy", "topoz", "nvar", "bcx",
              "bcy", "bcz")
    cfg = {k: float(v) for k, v in zip(keys_d, dbl)}
    cfg.update({k: int(v) for k, v in zip(keys_i, ints)})
    return cfg


def write_run(directory: str, kind: str, nx: int, ny: int, nz: int, nframes: int,
              lx: float = 2.0, ly: float = 2.0, lz: float = 1.0, dt_out: float = 0.1) -> dict:
    """Write mhd_config.dat and mhd_data_0000..NNNN like reorganize_fields.py does."""
    os.makedirs(directory, exist_ok=True)
    ndim = KINDS[kind][1]
    cfg = mhd_config(nx, ny, nz, lx, ly, lz, dt_out, ndim)
    write_mhd_config(os.path.join(directory, "mhd_config.dat"), cfg)
    for f in range(nframes):
        make_frame(kind, nx, ny, nz, f, dt_out).tofile(os.path.join(directory, f"mhd_data_{f:04d}"))
    return cfg


def read_frame(directory: str, frame: int, cfg: dict) -> np.ndarray:
    nx, ny, nz = cfg["nx"], cfg["ny"], cfg["nz"]
    a = np.fromfile(os.path.join(directory, f"mhd_data_{frame:04d}"), dtype=np.float32)
    ndim = cfg.get("ndim", 3 if nz > 1 else (2 if ny > 1 else 1))
    if ndim == 1:
        return a.reshape(nx + 4, NVAR)
    if ndim == 2:
        return a.reshape(ny + 4, nx + 4, NVAR)
    return a.reshape(nz + 4, ny + 4, nx + 4, NVAR)


def make_turbulence_maps(nx: int, ny: int, nz: int, frame: int, ndim: int = 2, dt_out: float = 0.1):
    """Synthetic deltab_NNNN / lc_NNNN content (read_magnetic_fluctuation, read_correlation_length,
    mhd_data_parallel.f90:306-497: each file holds the slab array then the 2-D array over the ghosted
    grid, float32).  Smooth, strictly positive, time-dependent: returns
    (sigma2_slab, sigma2_2d, lc_slab, lc_2d), each (nx+4,), (ny+4, nx+4) or (nz+4, ny+4, nx+4)."""
    t = frame * dt_out
    xs = (np.arange(nx + 4, dtype=np.float64) - 2.0) / max(nx - 1, 1)
    ys = (np.arange(ny + 4, dtype=np.float64) - 2.0) / max(ny - 1, 1)
    if ndim == 1:   # one row, no ghost rows in y (the 1-D files hold nx + 4 values per array)
        X = xs
        Y = Z = 0.0 * X
    elif ndim == 2:
        X, Y = np.meshgrid(xs, ys)
        Z = 0.0 * X
    else:
        zs = (np.arange(nz + 4, dtype=np.float64) - 2.0) / max(nz - 1, 1)
        Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    tx, ty, tz = 2 * np.pi * X, 2 * np.pi * Y, 2 * np.pi * Z
    s2s = 0.05 * (1.0 + 0.5 * np.sin(tx + 0.3 + t) * np.cos(ty) + 0.2 * np.cos(tz))
    s22 = 0.08 * (1.0 + 0.4 * np.cos(tx) * np.sin(ty + 0.5 - t) + 0.1 * np.sin(tz))
    lcs = 0.6 * (1.0 + 0.3 * np.sin(tx - 0.2) * np.sin(ty + t) + 0.1 * np.cos(tz + 0.4))
    lc2 = 0.4 * (1.0 + 0.25 * np.cos(tx + t) * np.cos(ty - 0.7) + 0.15 * np.sin(tz))
    return tuple(a.astype(np.float32) for a in (s2s, s22, lcs, lc2))


def make_acc_surface(P, which: int, frame: int) -> np.ndarray:
    """Synthetic <surface_filenameK>_NNNN.dat content (read_acc_surface, acc_region_surface.f90:118-206:
    one float64 per point of the ghosted plane spanned by the two axes other than the surface's normal,
    first axis fastest).  A rippled sheet near the mid-plane of the box that drifts with the frame number;
    the second surface sits a quarter of the box further along its normal.  Returns (n2, n1) C-order."""
    norm = P.surface_norm2 if which else P.surface_norm1
    axis = abs(norm) - 1
    n = (P.nx, P.ny, P.nz)
    lo = (P.xmin, P.ymin, P.zmin)
    ext = (P.lx, P.ly, P.lz)
    a1, a2 = [k for k in range(3) if k != axis]
    u = (np.arange(n[a1] + 4, dtype=np.float64) - 2.0) / max(n[a1], 1)
    v = (np.arange(n[a2] + 4, dtype=np.float64) - 2.0) / max(n[a2], 1)
    V, U = np.meshgrid(v, u, indexing="ij")
    mid = 0.5 + (0.2 if which else 0.0) * (1 if norm < 0 else -1)
    h = mid + 0.12 * np.sin(2 * np.pi * (U + 0.07 * frame)) * np.cos(2 * np.pi * V + 0.3 * which) + 0.01 * frame
    return np.ascontiguousarray(lo[axis] + ext[axis] * h, dtype=np.float64)


def write_time_stamps(directory: str, ts_mhd: int, te_mhd: int, stamps) -> None:
    """time_stamps.dat as load_tstamps_mhd reads it (mhd_config.f90:221-254): the first and last frame of the
    MHD run in two 8-byte slots, then one float64 time per frame ts_mhd .. te_mhd."""
    with open(os.path.join(directory, "time_stamps.dat"), "wb") as f:
        np.array([ts_mhd, te_mhd], dtype=np.int64).tofile(f)
        np.asarray(stamps, dtype=np.float64).tofile(f)


def read_time_stamps(directory: str, t_start: int, t_end: int, tmax_mhd: int) -> np.ndarray:
    """tstamps_mhd(1 : t_end - t_start + 1) with -vdt .true.; frames past tmax_mhd repeat the last interval."""
    with open(os.path.join(directory, "time_stamps.dat"), "rb") as f:
        ts_mhd = int(np.fromfile(f, dtype=np.int32, count=1)[0])   # a default integer read at pos = 1
        f.seek((t_start - ts_mhd + 2) * 8)
        nread = min(tmax_mhd, t_end) - t_start + 1
        head = np.fromfile(f, dtype=np.float64, count=nread)
    if len(head) != nread or nread < 2:
        raise IOError("time_stamps.dat is too short")
    out = np.empty(t_end - t_start + 1)
    out[:nread] = head
    for i in range(nread, len(out)):
        out[i] = out[i - 1] + (head[-1] - head[-2])
    return out
