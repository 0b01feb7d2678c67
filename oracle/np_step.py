"""Second, independent restatement (numpy, vectorised over particles) of ONE adaptive step of
the 2-D Cartesian Parker path, used to cross-check oracle/gpat_oracle.c.

TEST INFRASTRUCTURE ONLY (tests/ imports it; the product never does).  PARITY UNPINNED in the
same sense as gpat_oracle.c: the reference ships no golden vectors and cannot be built here.
Written directly from the cited Fortran, not from the C oracle:

  calc_fields_gradients      mhd_data_parallel.f90:533-566   (FP32 differences x f64 0.5/dx)
  get_interp_paramters       particle_module.f90:642-674
  interp_fields              mhd_data_parallel.f90:1751-1793
  calc_spatial_diffusion_coefficients (ndim_field == 2 branch)  particle_module.f90:2208-2387
  push_particle_2d           particle_module.f90:3358-3606 (Cartesian, uniform grid, no D_pp)

numpy float64 element-wise arithmetic is IEEE without contraction, so following the Fortran
association order gives the same bits as the C oracle except inside pow().
"""
from __future__ import annotations

import numpy as np

NFIELDS = 8
EPS = np.finfo(np.float64).eps  # EPSILON(1.0d0)


def gradients32(f8: np.ndarray, dx: float, dy: float) -> np.ndarray:
    """(ny+4, nx+4, 8) float32 -> (ny+4, nx+4, 32) float32, slots as farray (mhd_data_parallel.f90:77-81)."""
    f8 = np.asarray(f8, dtype=np.float32)
    out = np.zeros(f8.shape[:2] + (32,), dtype=np.float32)
    out[..., :8] = f8
    idxh = np.float64(0.5) / np.float64(dx)
    idyh = np.float64(0.5) / np.float64(dy)
    three, four = np.float32(3.0), np.float32(4.0)

    def d_axis(a, axis, ih):
        a = np.moveaxis(a, axis, 0)
        g = np.empty_like(a)
        # interior: FP32 difference, promoted, times the f64 factor, stored as FP32
        g[1:-1] = ((a[2:] - a[:-2]).astype(np.float64) * ih).astype(np.float32)
        lo = (-three * a[0] + four * a[1]) - a[2]
        hi = (three * a[-1] - four * a[-2]) + a[-3]
        g[0] = (lo.astype(np.float64) * ih).astype(np.float32)
        g[-1] = (hi.astype(np.float64) * ih).astype(np.float32)
        return np.moveaxis(g, 0, axis)

    gx = d_axis(f8, 1, idxh)
    gy = d_axis(f8, 0, idyh)
    for k in range(NFIELDS):
        out[..., NFIELDS + 3 * k + 0] = gx[..., k]
        out[..., NFIELDS + 3 * k + 1] = gy[..., k]
        # d/dz stays 0 in 2-D (unz == lnz)
    return out


def interp32(fa1, fa2, P, x, y, t_rel_over_dtf):
    """fields(1:32) at the particles; fa1/fa2 are the 32-slot frames (fa2 None = no time interp)."""
    px = (x - P.xmin) / P.dx
    py = (y - P.ymin) / P.dy
    ix = np.floor(px).astype(np.int64) + 1  # Fortran pos(1)
    iy = np.floor(py).astype(np.int64) + 1
    rx = px - ix + 1
    ry = py - iy + 1
    rx1 = 1.0 - rx
    ry1 = 1.0 - ry
    w = [rx1 * ry1 * 1.0, rx * ry1 * 1.0, rx1 * ry * 1.0, rx * ry * 1.0]
    f1 = np.zeros((len(x), 32))
    f2 = np.zeros((len(x), 32))
    c = 0
    for j in (0, 1):
        for i in (0, 1):
            # Fortran lower bound is -1: index ix -> storage ix + 1
            f1 = f1 + fa1[iy + j + 1, ix + i + 1, :].astype(np.float64) * w[c][:, None]
            if fa2 is not None:
                f2 = f2 + fa2[iy + j + 1, ix + i + 1, :].astype(np.float64) * w[c][:, None]
            c += 1
    if fa2 is not None:
        rt = t_rel_over_dtf[:, None]
        f1 = f1 * (1.0 - rt) + f2 * rt
    return f1


def push_2d(P, F, p, dt_min, dt_max, u, x, y, t, qdrift, aux=None):
    """One adaptive push_particle_2d.  F: (n, 32) interpolated fields (1-based slot s at F[:, s-1]).
    u: (n, 4) uniforms.  aux: (n, 16) interpolated db2_slab(1:4) db2_2d(1:4) lc_slab(1:4) lc_2d(1:4)
    when deltab_flag / correlation_flag are set (particle_module.f90:2246-2254, 2364-2371).
    Returns x, y, p, t, dt after the step."""
    nf = NFIELDS
    bx, by, bz = F[:, 4], F[:, 5], F[:, 6]
    b = np.sqrt(bx ** 2 + by ** 2 + bz ** 2)
    tiny = b < EPS
    with np.errstate(divide="ignore"):
        ib1 = np.where(tiny, 1.0, 1.0 / b)  # kappa routine
        ib = np.where(tiny, 0.0, 1.0 / b)   # pusher
    ib2k, ib2 = ib1 * ib1, ib * ib
    ib3k, ib3 = ib1 * ib2k, ib * ib2
    knp = np.ones_like(b)
    if P.mag_dependency == 1:
        knp = knp * b ** (P.gamma_turb - 2.0)
    if P.deltab_flag:
        knp = knp / aux[:, 0]
    if P.correlation_flag:
        knp = knp * aux[:, 8] ** (P.gamma_turb - 1.0)
    knorm = knp * (p / P.p0) ** P.pindex if P.momentum_dependency == 1 else knp
    kpara = P.kpara0 * knorm
    kperp = kpara * P.kret
    skpara = np.sqrt(2.0 * kpara)
    skperp = np.sqrt(2.0 * kperp)
    skpp = np.sqrt(2.0 * (kpara - kperp))
    dbx_dx, dbx_dy = F[:, nf + 12], F[:, nf + 13]
    dby_dx, dby_dy = F[:, nf + 15], F[:, nf + 16]
    dbz_dx, dbz_dy = F[:, nf + 18], F[:, nf + 19]
    db_dx, db_dy = F[:, nf + 21], F[:, nf + 22]
    dkdx = np.zeros_like(b)
    dkdy = np.zeros_like(b)
    if P.mag_dependency == 1:
        dkdx = db_dx * ib1 * (P.gamma_turb - 2.0)
        dkdy = db_dy * ib1 * (P.gamma_turb - 2.0)
    if P.deltab_flag:
        dkdx = dkdx - aux[:, 1] / aux[:, 0]
        dkdy = dkdy - aux[:, 2] / aux[:, 0]
    if P.correlation_flag:
        dkdx = dkdx + (P.gamma_turb - 1.0) * aux[:, 9] / aux[:, 8]
        dkdy = dkdy + (P.gamma_turb - 1.0) * aux[:, 10] / aux[:, 8]
    kpp = kpara - kperp
    dkxx_dx = kperp * dkdx + kpp * dkdx * bx ** 2 * ib2k + 2.0 * kpp * bx * (dbx_dx * b - bx * db_dx) * ib3k
    dkyy_dy = kperp * dkdy + kpp * dkdy * by ** 2 * ib2k + 2.0 * kpp * by * (dby_dy * b - by * db_dy) * ib3k
    dkxy_dx = kpp * dkdx * bx * by * ib2k + kpp * ((dbx_dx * by + bx * dby_dx) * ib2k - 2.0 * bx * by * db_dx * ib3k)
    dkxy_dy = kpp * dkdy * bx * by * ib2k + kpp * ((dbx_dy * by + bx * dby_dy) * ib2k - 2.0 * bx * by * db_dy * ib3k)

    vx, vy = F[:, 0], F[:, 1]
    dvx_dx, dvy_dy = F[:, nf + 0], F[:, nf + 4]
    vdp = qdrift / np.sqrt((P.drift1 * P.p0 / p) ** 2 + (P.drift2 * P.p0 ** 2 / p ** 2) ** 2)
    vdx = vdp * (dbz_dy * ib2 - 2.0 * bz * db_dy * ib3)
    vdy = vdp * (-dbz_dx * ib2 + 2.0 * bz * db_dx * ib3)
    dx_dt = vx + vdx + dkxx_dx + dkxy_dy
    dy_dt = vy + vdy + dkxy_dx + dkyy_dy
    divv = dvx_dx + dvy_dy
    dp_dt = -p * divv / 3.0
    s = np.where(skperp > 0.0, skperp, skpara)
    with np.errstate(divide="ignore", invalid="ignore"):
        cand = np.minimum.reduce([(0.5 * P.dx / skpara) ** 2, (0.5 * P.dy / skpara) ** 2,
                                  (s / dx_dt) ** 2, (s / dy_dt) ** 2,
                                  np.float64(np.float32(0.1)) * p / np.abs(dp_dt)])
    ok = (dx_dt != 0.0) & (dy_dt != 0.0) & (dp_dt != 0.0)
    dt = np.where(ok, cand, dt_min)
    dt = np.where(dt < dt_min, dt_min, dt)
    dt = np.where(dt > dt_max, dt_max, dt)
    sdt = np.sqrt(dt)
    sqrt3 = np.sqrt(3.0)
    ran1 = (2.0 * u[:, 0] - 1.0) * sqrt3
    ran2 = (2.0 * u[:, 1] - 1.0) * sqrt3
    ran3 = (2.0 * u[:, 2] - 1.0) * sqrt3
    ddx = dx_dt * dt + ran1 * skperp * sdt + ran3 * skpp * sdt * bx * ib
    ddy = dy_dt * dt + ran2 * skperp * sdt + ran3 * skpp * sdt * by * ib
    ranp = (2.0 * u[:, 3] - 1.0) * sqrt3
    ddp = dp_dt * dt + ranp * np.sqrt(2 * 0.0) * sdt
    pn = p + ddp
    low = pn < 0.25 * P.p0
    pn = np.where(low, 0.25 * P.p0, pn)
    return x + ddx, y + ddy, pn, t + dt, dt


def turbulence_grad(a, dx, dy):
    """value + d/dx, d/dy (+ zero d/dz) of one 2-D turbulence map (ny+4, nx+4) float32, with the
    arithmetic of calc_grad_sigma2_slab (mhd_data_parallel.f90:771-872): FP32 differences, one-sided
    at the array ends, times 0.5/dx in FP64, stored FP32.  Returns (ny+4, nx+4, 4) float32."""
    out = np.zeros(a.shape + (4,), dtype=np.float32)
    out[..., 0] = a
    for axis, h in ((1, dx), (0, dy)):
        g = np.zeros_like(a)
        sl = lambda s: tuple(s if k == axis else slice(None) for k in range(2))
        g[sl(slice(1, -1))] = a[sl(slice(2, None))] - a[sl(slice(None, -2))]
        g[sl(0)] = (np.float32(-3) * a[sl(0)] + np.float32(4) * a[sl(1)]) - a[sl(2)]
        g[sl(-1)] = (np.float32(3) * a[sl(-1)] - np.float32(4) * a[sl(-2)]) + a[sl(-3)]
        out[..., 1 if axis == 1 else 2] = (g.astype(np.float64) * (0.5 / h)).astype(np.float32)
    return out


def interp_aux(maps1, maps2, P, x, y, rt):
    """interp_magnetic_fluctuation / interp_correlation_length (mhd_data_parallel.f90:1806-1915) for
    four maps given as lists of (ny+4, nx+4, 4) arrays: returns (n, 16)."""
    px = (x - P.xmin) / P.dx
    py = (y - P.ymin) / P.dy
    ix = np.floor(px).astype(np.int64) + 1
    iy = np.floor(py).astype(np.int64) + 1
    rx, ry = px - ix + 1, py - iy + 1
    w = [(1.0 - rx) * (1.0 - ry) * 1.0, rx * (1.0 - ry) * 1.0, (1.0 - rx) * ry * 1.0, rx * ry * 1.0]
    out = []
    for m1, m2 in zip(maps1, maps2):
        f1 = np.zeros((len(x), 4))
        f2 = np.zeros((len(x), 4))
        c = 0
        for j in (0, 1):
            for i in (0, 1):
                f1 = f1 + m1[iy + j + 1, ix + i + 1, :].astype(np.float64) * w[c][:, None]
                f2 = f2 + m2[iy + j + 1, ix + i + 1, :].astype(np.float64) * w[c][:, None]
                c += 1
        out.append(f1 * (1.0 - rt[:, None]) + f2 * rt[:, None])
    return np.concatenate(out, axis=1)
