"""Second, independent restatement (numpy, vectorised over particles) of ONE adaptive step of
the Cartesian Parker pushers (2-D, 2-D + third dimension, 3-D; standard and NLGC kappa; D_pp),
used to cross-check oracle/gpat_oracle.c.

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


# ------------------------------------------------------------------------------------------------
# 3-D gather, the general kappa tensor (standard and NLGC), D_pp, and the 3-D-like Parker pushers.
# Written from the Fortran; nothing below calls the C oracle or the 2-D functions above.
#   calc_fields_gradients (d/dz part)                 mhd_data_parallel.f90:556-566
#   interp_fields (8 corners)                         mhd_data_parallel.f90:1751-1793
#   calc_spatial_diffusion_coefficients               particle_module.f90:2208-2450
#   calc_spatial_diffusion_coefficients_nlgc          particle_module.f90:2464-2771
#   calc_dpp_wave_scattering / calc_dpp_flow_shear    particle_module.f90:2918-2979
#   push_particle_2d (D_pp part)                      particle_module.f90:3476-3496
#   push_particle_2d_include_3rd                      particle_module.f90:3979-4245
#   push_particle_3d                                  particle_module.f90:4625-4907
# ------------------------------------------------------------------------------------------------
def gradients32_3d(f8: np.ndarray, dx: float, dy: float, dz: float) -> np.ndarray:
    """(nz+4, ny+4, nx+4, 8) float32 -> (..., 32) float32."""
    f8 = np.asarray(f8, dtype=np.float32)
    out = np.zeros(f8.shape[:3] + (32,), dtype=np.float32)
    out[..., :8] = f8
    three, four = np.float32(3.0), np.float32(4.0)
    for d, (axis, h) in enumerate(((2, dx), (1, dy), (0, dz))):
        ih = np.float64(0.5) / np.float64(h)
        a = np.moveaxis(f8, axis, 0)
        g = np.empty_like(a)
        g[1:-1] = ((a[2:] - a[:-2]).astype(np.float64) * ih).astype(np.float32)
        g[0] = (((-three * a[0] + four * a[1]) - a[2]).astype(np.float64) * ih).astype(np.float32)
        g[-1] = (((three * a[-1] - four * a[-2]) + a[-3]).astype(np.float64) * ih).astype(np.float32)
        g = np.moveaxis(g, 0, axis)
        for k in range(NFIELDS):
            out[..., NFIELDS + 3 * k + d] = g[..., k]
    return out


def interp32_3d(fa1, fa2, P, x, y, z, rt):
    """fields(1:32) from the 8 surrounding grid points, corner order i fastest, then j, then k."""
    pc = [(x - P.xmin) / P.dx, (y - P.ymin) / P.dy, (z - P.zmin) / P.dz]
    idx = [np.floor(c).astype(np.int64) + 1 for c in pc]   # Fortran pos
    r = [c - i + 1 for c, i in zip(pc, idx)]
    r1 = [1.0 - v for v in r]
    f1 = np.zeros((len(x), 32))
    f2 = np.zeros((len(x), 32))
    for k in (0, 1):
        for j in (0, 1):
            for i in (0, 1):
                w = (r[0] if i else r1[0]) * (r[1] if j else r1[1]) * (r[2] if k else r1[2])
                sl = (idx[2] + k + 1, idx[1] + j + 1, idx[0] + i + 1)   # lower bound -1
                f1 = f1 + fa1[sl].astype(np.float64) * w[:, None]
                if fa2 is not None:
                    f2 = f2 + fa2[sl].astype(np.float64) * w[:, None]
    if fa2 is not None:
        f1 = f1 * (1.0 - rt[:, None]) + f2 * rt[:, None]
    return f1


def kappa_tensor(P, F, p, mu, geometry, aux=None, focused=False):
    """kappa_type for geometry '2d' (ndim_field = 2), '2d3' (2-D + include_3rd_dim) or '3d'.
    Returns a dict of arrays named as the Fortran components."""
    nf = NFIELDS
    g2, g1 = P.gamma_turb - 2.0, P.gamma_turb - 1.0
    bx, by, bz = F[:, 4], F[:, 5], F[:, 6]
    b = np.sqrt(bx ** 2 + by ** 2 + bz ** 2)
    with np.errstate(divide="ignore"):
        ib1 = np.where(b < EPS, 1.0, 1.0 / b)
    ib2 = ib1 * ib1
    ib3 = ib1 * ib2
    one = np.ones_like(b)
    k = {}
    kn_para, kn_perp = one.copy(), one.copy()
    if P.mag_dependency == 1:
        kn_para = kn_para * b ** g2
        kn_perp = kn_perp * b ** (g2 / 3.0)
    if P.deltab_flag:
        kn_para = kn_para / aux[:, 0]
        kn_perp = kn_perp * aux[:, 0] ** (-1.0 / 3.0) * aux[:, 4] ** (2.0 / 3.0)
    if P.correlation_flag:
        kn_para = kn_para * aux[:, 8] ** g1
        kn_perp = kn_perp * aux[:, 8] ** (g1 / 3.0) * aux[:, 12] ** (2.0 / 3.0)
    k["knorm_para"] = kn_para
    if P.nlgc:
        if P.momentum_dependency == 1:
            np_para = kn_para * (p / P.p0) ** P.pindex
            np_perp = kn_perp * (p / P.p0) ** ((5.0 - P.gamma_turb) / 3.0)
        else:
            np_para, np_perp = kn_para, kn_perp
        kpara = P.kpara0 * np_para
        kperp = P.kpara0 * P.kperp_kpara * np_perp * mu ** 2
    else:
        knorm = kn_para * (p / P.p0) ** P.pindex if P.momentum_dependency == 1 else kn_para
        kpara = P.kpara0 * knorm
        kperp = kpara * P.kret
    k["kpara"], k["kperp"] = kpara, kperp
    k["skpara"] = np.sqrt(2.0 * kpara)
    k["skperp"] = np.sqrt(2.0 * kperp)
    k["skpara_perp"] = np.sqrt(2.0 * (kpara - kperp))

    zero = np.zeros_like(b)
    three = geometry != "2d"
    full = geometry == "3d"
    dB = {("x", "x"): F[:, nf + 12], ("x", "y"): F[:, nf + 13], ("x", "z"): F[:, nf + 14] if full else zero,
          ("y", "x"): F[:, nf + 15], ("y", "y"): F[:, nf + 16], ("y", "z"): F[:, nf + 17] if full else zero,
          ("z", "x"): F[:, nf + 18], ("z", "y"): F[:, nf + 19], ("z", "z"): F[:, nf + 20] if full else zero}
    db = {"x": F[:, nf + 21], "y": F[:, nf + 22], "z": F[:, nf + 23] if full else zero}
    comp = {"x": bx, "y": by, "z": bz}
    axes = ("x", "y", "z") if three else ("x", "y")
    # d ln(k_para) and d ln(k_perp) along each axis
    dpa, dpe = {}, {}
    for n, d in enumerate(axes):
        live = full or d != "z"
        a = zero
        e = zero
        if P.mag_dependency == 1 and live:
            # particle_module.f90:2406-2408: the standard 3-D branch has no 1/B here; NLGC and 2-D do
            a = db[d] * g2 if (full and not P.nlgc) else db[d] * ib1 * g2
            e = db[d] * ib1 * g2 / 3.0
        if P.deltab_flag and live:
            a = a - aux[:, 1 + n] / aux[:, 0]
            e = e - aux[:, 1 + n] / aux[:, 0] / 3.0 + 2.0 * aux[:, 5 + n] / aux[:, 4] / 3.0
        if P.correlation_flag and live:
            a = a + g1 * aux[:, 9 + n] / aux[:, 8]
            e = e + g1 * aux[:, 9 + n] / aux[:, 8] / 3.0 + 2.0 * aux[:, 13 + n] / aux[:, 12] / 3.0
        dpa[d], dpe[d] = a, e
    # focused transport keeps the perpendicular part only (particle_module.f90:2372-2376, 2606-2610)
    kpp = -kperp if focused else kpara - kperp
    for d in axes:       # dk_dd_dd
        if P.nlgc:
            iso, ani = kperp * dpe[d], kpara * dpa[d] - kperp * dpe[d]
        else:
            iso, ani = kperp * dpa[d], kpp * dpa[d]
        c = comp[d]
        k[f"dk{d}{d}_d{d}"] = iso + ani * c ** 2 * ib2 + 2.0 * kpp * c * (dB[(d, d)] * b - c * db[d]) * ib3
    pairs = [("x", "y", "x"), ("x", "y", "y")]
    if three:
        pairs += [("x", "z", "x"), ("x", "z", "z"), ("y", "z", "y"), ("y", "z", "z")]
    for i, j, d in pairs:
        ani = (kpara * dpa[d] - kperp * dpe[d]) if P.nlgc else kpp * dpa[d]
        ci, cj = comp[i], comp[j]
        k[f"dk{i}{j}_d{d}"] = ani * ci * cj * ib2 + kpp * ((dB[(i, d)] * cj + ci * dB[(j, d)]) * ib2
                                                          - 2.0 * ci * cj * db[d] * ib3)
    return k


def dpp_terms(P, F, p, k, dvx_dx, dvy_dy, dvz_dz, divv, dp_dt, geometry):
    """dp_dt and dpp after calc_dpp_wave_scattering and calc_dpp_flow_shear."""
    nf = NFIELDS
    bx, by, bz = F[:, 4], F[:, 5], F[:, 6]
    b = np.sqrt(bx ** 2 + by ** 2 + bz ** 2)
    dpp = np.zeros_like(b)
    if P.dpp_wave:
        va = b / np.sqrt(F[:, 3])
        if P.momentum_dependency == 1:
            dp_dt = dp_dt + (8 * p / (27 * k["kpara"])) * va ** 2
        else:
            dp_dt = dp_dt + (4 * p / (9 * k["kpara"])) * va ** 2
        dpp = dpp + (p * va) ** 2 / (9 * k["kpara"])
    if P.dpp_shear:
        zero = np.zeros_like(b)
        dvx_dy, dvy_dx = F[:, nf + 1], F[:, nf + 3]
        if geometry == "2d":
            sxz = syz = zero
            szz = -divv / 3
        else:
            full = geometry == "3d"
            dvx_dz = F[:, nf + 2] if full else zero
            dvy_dz = F[:, nf + 5] if full else zero
            dvz_dx, dvz_dy = F[:, nf + 6], F[:, nf + 7]
            szz = dvz_dz - divv / 3
            sxz = (dvx_dz + dvz_dx) / 2
            syz = (dvy_dz + dvz_dy) / 2
        sxx = dvx_dx - divv / 3
        syy = dvy_dy - divv / 3
        sxy = (dvx_dy + dvy_dx) / 2
        if P.weak_scattering:
            with np.errstate(divide="ignore"):
                ib = np.where(b < EPS, 0.0, 1.0 / b)
            bbs = sxx * bx ** 2 + syy * by ** 2 + szz * bz ** 2 + 2.0 * (sxy * bx * by + sxz * bx * bz + syz * by * bz)
            bbs = bbs * ib * ib
            gshear = bbs ** 2 / 5
        else:
            gshear = 2 * (sxx ** 2 + syy ** 2 + szz ** 2 + 2 * (sxy ** 2 + sxz ** 2 + syz ** 2)) / 15
        on = gshear > 0.0
        add_dp = (2 + P.pindex) * gshear * P.tau0 * k["knorm_para"] * p ** (P.pindex - 1) * P.p0 ** (2.0 - P.pindex)
        add_pp = gshear * P.tau0 * k["knorm_para"] * p ** P.pindex * P.p0 ** (2.0 - P.pindex)
        dp_dt = np.where(on, dp_dt + add_dp, dp_dt)
        dpp = np.where(on, dpp + add_pp, dpp)
    return dp_dt, dpp


def _finish_momentum(P, p, dp_dt, dpp, dt, sdt, ranp, inside=None, deltas=None):
    ddp = dp_dt * dt + ranp * np.sqrt(2 * dpp) * sdt
    if inside is not None:  # acc_region_flag == 1
        ddp = np.where(inside, ddp, 0.0)
    pn = p + ddp
    low = pn < 0.25 * P.p0
    if deltas is not None:  # deltap as the mover's roll-back sees it (particle_module.f90:3601-3605)
        deltas["p"] = np.where(low, 0.25 * P.p0 - (pn - ddp), ddp)
    return np.where(low, 0.25 * P.p0, pn)


def _acc_region(P, x, y, z, ndim):
    inside = ((x - P.xmin) / P.lx >= P.acc_region[0]) & ((x - P.xmin) / P.lx <= P.acc_region[1])
    inside &= ((y - P.ymin) / P.ly >= P.acc_region[2]) & ((y - P.ymin) / P.ly <= P.acc_region[3])
    if ndim == 3:
        inside &= ((z - P.zmin) / P.lz >= P.acc_region[4]) & ((z - P.zmin) / P.lz <= P.acc_region[5])
    return inside


def push_2d_general(P, F, p, mu, dt_min, dt_max, u, x, y, t, qdrift, aux=None, dt_fixed=None, deltas=None):
    """push_particle_2d with every Parker-transport switch: NLGC kappa, D_pp (wave + shear), acc region.
    dt_fixed: the fixed_dt = .true. call of the mover's re-push; deltas: dict that receives deltax/y/p."""
    nf = NFIELDS
    k = kappa_tensor(P, F, p, mu, "2d", aux)
    bx, by, bz = F[:, 4], F[:, 5], F[:, 6]
    b = np.sqrt(bx ** 2 + by ** 2 + bz ** 2)
    with np.errstate(divide="ignore"):
        ib = np.where(b < EPS, 0.0, 1.0 / b)
    ib2 = ib * ib
    ib3 = ib * ib2
    vx, vy = F[:, 0], F[:, 1]
    dvx_dx, dvy_dy = F[:, nf + 0], F[:, nf + 4]
    dbz_dx, dbz_dy, db_dx, db_dy = F[:, nf + 18], F[:, nf + 19], F[:, nf + 21], F[:, nf + 22]
    vdp = qdrift / np.sqrt((P.drift1 * P.p0 / p) ** 2 + (P.drift2 * P.p0 ** 2 / p ** 2) ** 2)
    vdx = vdp * (dbz_dy * ib2 - 2.0 * bz * db_dy * ib3)
    vdy = vdp * (-dbz_dx * ib2 + 2.0 * bz * db_dx * ib3)
    dx_dt = vx + vdx + k["dkxx_dx"] + k["dkxy_dy"]
    dy_dt = vy + vdy + k["dkxy_dx"] + k["dkyy_dy"]
    divv = dvx_dx + dvy_dy
    dp_dt = -p * divv / 3.0
    dp_dt, dpp = dpp_terms(P, F, p, k, dvx_dx, dvy_dy, None, divv, dp_dt, "2d")
    s = np.where(k["skperp"] > 0.0, k["skperp"], k["skpara"])
    with np.errstate(divide="ignore", invalid="ignore"):
        cand = np.minimum.reduce([(0.5 * P.dx / k["skpara"]) ** 2, (0.5 * P.dy / k["skpara"]) ** 2,
                                  (s / dx_dt) ** 2, (s / dy_dt) ** 2,
                                  np.float64(np.float32(0.1)) * p / np.abs(dp_dt)])
    ok = (dx_dt != 0.0) & (dy_dt != 0.0) & (dp_dt != 0.0)
    dt = np.where(ok, cand, dt_min)
    dt = np.where(dt < dt_min, dt_min, dt)
    dt = np.where(dt > dt_max, dt_max, dt)
    if dt_fixed is not None:
        dt = dt_fixed
    sdt = np.sqrt(dt)
    sqrt3 = np.sqrt(3.0)
    ran1, ran2, ran3, ranp = [(2.0 * u[:, c] - 1.0) * sqrt3 for c in range(4)]
    ddx = dx_dt * dt + ran1 * k["skperp"] * sdt + ran3 * k["skpara_perp"] * sdt * bx * ib
    ddy = dy_dt * dt + ran2 * k["skperp"] * sdt + ran3 * k["skpara_perp"] * sdt * by * ib
    xn, yn = x + ddx, y + ddy
    if deltas is not None:
        deltas["x"], deltas["y"] = ddx, ddy
    if deltas is not None:
        # check_drift_2d (particle_module.f90:3448-3460, 3582-3586): the out-of-plane drift moves ptl%z inside the
        # pusher; the mover's own deltaz stays 0 in 2-D, so a roll-back does not undo it
        if P.check_drift_2d:
            dbx_dy, dby_dx = F[:, nf + 13], F[:, nf + 15]
            vdz = vdp * ((dby_dx - dbx_dy) * ib2 - 2 * (by * db_dx - bx * db_dy) * ib3)
        else:
            vdz = np.zeros_like(vdp)
        deltas["z_in_pusher"] = vdz * dt
    inside = _acc_region(P, xn, yn, None, 2) if P.acc_region_flag == 1 else None
    return xn, yn, _finish_momentum(P, p, dp_dt, dpp, dt, sdt, ranp, inside, deltas), t + dt, dt


def push_3d_like(P, F, p, mu, dt_min, dt_max, u, x, y, z, t, qdrift, full3d, aux=None, dt_fixed=None, deltas=None):
    """push_particle_3d (full3d) or push_particle_2d_include_3rd, Cartesian uniform grid.
    Returns x, y, z, p, t, dt after the step."""
    nf = NFIELDS
    geometry = "3d" if full3d else "2d3"
    k = kappa_tensor(P, F, p, mu, geometry, aux)
    zero = np.zeros(len(p))
    vx, vy, vz = F[:, 0], F[:, 1], F[:, 2]
    bx, by, bz = F[:, 4], F[:, 5], F[:, 6]
    b = np.sqrt(bx ** 2 + by ** 2 + bz ** 2)
    with np.errstate(divide="ignore"):
        ib = np.where(b < EPS, 0.0, 1.0 / b)
    bxn, byn, bzn = bx * ib, by * ib, bz * ib
    bxyn = np.sqrt(bxn ** 2 + byn ** 2)
    with np.errstate(divide="ignore"):
        ibxyn = np.where(bxyn < EPS, 0.0, 1.0 / bxyn)
    dvx_dx, dvy_dy = F[:, nf + 0], F[:, nf + 4]
    dvz_dz = F[:, nf + 8] if full3d else zero
    dbx_dy, dby_dx = F[:, nf + 13], F[:, nf + 15]
    dbx_dz = F[:, nf + 14] if full3d else zero
    dby_dz = F[:, nf + 17] if full3d else zero
    dbz_dx, dbz_dy = F[:, nf + 18], F[:, nf + 19]
    db_dx, db_dy = F[:, nf + 21], F[:, nf + 22]
    db_dz = F[:, nf + 23] if full3d else zero
    ib2 = ib * ib
    ib3 = ib * ib2
    vdp = qdrift / np.sqrt((P.drift1 * P.p0 / p) ** 2 + (P.drift2 * P.p0 ** 2 / p ** 2) ** 2)
    vdx = vdp * ((dbz_dy - dby_dz) * ib2 - 2 * (bz * db_dy - by * db_dz) * ib3)
    vdy = vdp * ((dbx_dz - dbz_dx) * ib2 - 2 * (bx * db_dz - bz * db_dx) * ib3)
    vdz = vdp * ((dby_dx - dbx_dy) * ib2 - 2 * (by * db_dx - bx * db_dy) * ib3)
    if not full3d:  # push_particle_2d_include_3rd zeroes the d/dz components, particle_module.f90:4096-4098
        k["dkxz_dz"], k["dkyz_dz"], k["dkzz_dz"] = zero, zero, zero
    dx_dt = vx + vdx + k["dkxx_dx"] + k["dkxy_dy"] + k["dkxz_dz"]
    dy_dt = vy + vdy + k["dkxy_dx"] + k["dkyy_dy"] + k["dkyz_dz"]
    dz_dt = vz + vdz + k["dkxz_dx"] + k["dkyz_dy"] + k["dkzz_dz"]
    divv = dvx_dx + dvy_dy + dvz_dz
    dp_dt = -p * divv / 3.0
    dp_dt, dpp = dpp_terms(P, F, p, k, dvx_dx, dvy_dy, dvz_dz, divv, dp_dt, geometry)
    s = np.where(k["skperp"] > 0.0, k["skperp"], k["skpara"])
    with np.errstate(divide="ignore", invalid="ignore"):
        cands = [(0.5 * P.dx / k["skpara"]) ** 2, (0.5 * P.dy / k["skpara"]) ** 2]
        if full3d:
            cands.append((0.5 * P.dz / k["skpara"]) ** 2)
        cands += [(s / dx_dt) ** 2, (s / dy_dt) ** 2]
        if full3d:
            cands.append((s / dz_dt) ** 2)   # dz_dt is not tested against 0 (particle_module.f90:4792-4794)
        cands.append(np.float64(np.float32(0.1)) * p / np.abs(dp_dt))
        cand = np.minimum.reduce(cands)
    ok = (dx_dt != 0.0) & (dy_dt != 0.0) & (dp_dt != 0.0)
    dt = np.where(ok, cand, dt_min)
    dt = np.where(dt < dt_min, dt_min, dt)
    dt = np.where(dt > dt_max, dt_max, dt)
    if dt_fixed is not None:
        dt = dt_fixed
    sdt = np.sqrt(dt)
    sqrt3 = np.sqrt(3.0)
    ran1, ran2, ran3, ranp = [(2.0 * u[:, c] - 1.0) * sqrt3 for c in range(4)]
    skpa, skpe = k["skpara"], k["skperp"]
    ddx = dx_dt * dt + (bxn * skpa * ran1 - bxn * bzn * skpe * ibxyn * ran2 - byn * skpe * ibxyn * ran3) * sdt
    ddy = dy_dt * dt + (byn * skpa * ran1 - byn * bzn * skpe * ibxyn * ran2 + bxn * skpe * ibxyn * ran3) * sdt
    ddz = dz_dt * dt + (bzn * skpa * ran1 + bxyn * skpe * ran2) * sdt
    xn, yn, zn = x + ddx, y + ddy, z + ddz
    if deltas is not None:
        deltas["x"], deltas["y"], deltas["z"] = ddx, ddy, ddz
    inside = _acc_region(P, xn, yn, zn, 3 if full3d else 2) if P.acc_region_flag == 1 else None
    return xn, yn, zn, _finish_momentum(P, p, dp_dt, dpp, dt, sdt, ranp, inside, deltas), t + dt, dt


# ------------------------------------------------------------------------------------------------
# The mover: one MHD interval of ONE particle at a time, in plain Python (small cases only).
#   particle_boundary_condition (single rank)         particle_module.f90:1984-2129
#   particle_mover_one_cycle                          particle_module.f90:1481-1833
#   particle_mover (extended bounds, final BC pass)   particle_module.f90:1846-1974
# ------------------------------------------------------------------------------------------------
INBOX, OTHERS = 1, 0


def boundary_condition(P, s, e, tally):
    """s: dict with x, y, z, weight, count_flag; e = (xmin, xmax, ymin, ymax, zmin, zmax) as passed in
    (the step loop passes the EXTENDED bounds, so a periodic wrap shifts by L + dx)."""
    for axis, (name, lo, hi) in enumerate((("x", e[0], e[1]), ("y", e[2], e[3]), ("z", e[4], e[5]))):
        if axis >= P.ndim:
            break
        if s[name] < lo and s["count_flag"] == INBOX:
            if P.pbc[axis]:                      # open: neighbors < 0
                tally["leak"] += s["weight"]
                s["count_flag"] = -(2 * axis + 1)
            else:                                # periodic, the neighbour is this rank
                s[name] = s[name] - lo + hi
        elif s[name] > hi and s["count_flag"] == INBOX:
            if P.pbc[axis]:
                tally["leak"] += s["weight"]
                s["count_flag"] = -(2 * axis + 2)
            else:
                s[name] = s[name] - hi + lo


def _negp_or_bc(P, s, e, tally):
    if s["p"] < 0.0:
        s["count_flag"] = OTHERS
        tally["leak_negp"] += s["weight"]
    else:
        boundary_condition(P, s, e, tally)


def mover_one_particle(P, s, push, t0, dtf, nsteps_interval, num_fine_steps, tally):
    """particle_mover_one_cycle for one particle.  `push(s, fixed_dt)` performs one push_particle_* call
    on the state dict (position, p, t, dt) and returns (deltax, deltay, deltaz, deltap[, deltav, deltamu])."""
    dt_fine = dtf / num_fine_steps
    e = (P.xmin - P.dx * 0.5, P.xmax + P.dx * 0.5, P.ymin - P.dy * 0.5, P.ymax + P.dy * 0.5,
         P.zmin - P.dz * 0.5, P.zmax + P.dz * 0.5)
    d = (0.0, 0.0, 0.0, 0.0)
    step = int(np.ceil((s["t"] - t0) / dt_fine))
    dt_target = dt_fine if step <= 0 else step * dt_fine
    if dt_target > dtf:
        dt_target = dtf
    if s["p"] < 0.0 and s["count_flag"] == INBOX:
        s["count_flag"] = OTHERS
        tally["leak_negp"] += s["weight"]
    else:
        boundary_condition(P, s, e, tally)
    if s["count_flag"] != INBOX:
        return
    while dt_target < dtf + dt_fine * np.float64(np.float32(0.1)):
        if s["count_flag"] != INBOX:
            break
        while (s["t"] - t0) < dt_target and s["count_flag"] == INBOX:
            _negp_or_bc(P, s, e, tally)
            if s["count_flag"] != INBOX:
                break
            d = push(s, False)
            tally["steps"] += 1
            s["nsteps_pushed"] = (s["nsteps_pushed"] + 1) % nsteps_interval
        if (s["t"] - t0) > dt_target and s["count_flag"] == INBOX:
            s["x"], s["y"], s["z"], s["p"] = s["x"] - d[0], s["y"] - d[1], s["z"] - d[2], s["p"] - d[3]
            if len(d) == 6:   # focused transport: deltav, deltamu (particle_module.f90:1712-1713)
                s["v"], s["mu"] = s["v"] - d[4], s["mu"] - d[5]
            s["t"] = s["t"] - s["dt"]
            dt_old = s["dt"]
            s["dt"] = t0 + dt_target - s["t"]
            if s["dt"] > 0:
                s["nsteps_pushed"] = s["nsteps_pushed"] - 1
                d = push(s, True)
                tally["steps"] += 1
                s["nsteps_pushed"] = (s["nsteps_pushed"] + 1) % nsteps_interval
            s["dt"] = dt_old
            _negp_or_bc(P, s, e, tally)
        dt_target = dt_target + dt_fine


def final_boundary_pass(P, s, tally):
    """The pass after the cycle loop: the TRUE bounds (particle_module.f90:1955-1967)."""
    if s["p"] < 0.0 and s["count_flag"] != INBOX:
        s["count_flag"] = OTHERS
        tally["leak_negp"] += s["weight"]
    else:
        boundary_condition(P, s, (P.xmin, P.xmax, P.ymin, P.ymax, P.zmin, P.zmax), tally)


def remove_particles(flags):
    """remove_particles (particle_module.f90:5365-5403): swap-with-tail compaction.  flags: count_flag per
    slot; returns (order of the surviving slots as original indices, escaped indices in the order they are
    appended to escaped_ptls)."""
    idx = list(range(len(flags)))
    n = len(idx)
    nremoved = 0
    escaped = []
    i = 1
    while i <= n:
        if (n - i) == (nremoved - 1):
            break
        if flags[idx[i - 1]] == INBOX:
            i += 1
        else:
            if flags[idx[i - 1]] < 0:
                escaped.append(idx[i - 1])
            tail = n - nremoved
            idx[tail - 1], idx[i - 1] = idx[i - 1], idx[tail - 1]
            nremoved += 1
    return idx[:n - nremoved], escaped


def turbulence_grad_3d(a, dx, dy, dz):
    """3-D version of turbulence_grad: (nz+4, ny+4, nx+4) float32 -> (..., 4) float32 = value, d/dx, d/dy, d/dz
    (calc_grad_sigma2_slab and its siblings, mhd_data_parallel.f90:771-1604, d/dz part included)."""
    out = np.zeros(a.shape + (4,), dtype=np.float32)
    out[..., 0] = a
    for comp, (axis, h) in enumerate(((2, dx), (1, dy), (0, dz)), start=1):
        b = np.moveaxis(a, axis, 0)
        g = np.zeros_like(b)
        g[1:-1] = b[2:] - b[:-2]
        g[0] = (np.float32(-3) * b[0] + np.float32(4) * b[1]) - b[2]
        g[-1] = (np.float32(3) * b[-1] - np.float32(4) * b[-2]) + b[-3]
        out[..., comp] = np.moveaxis((g.astype(np.float64) * (0.5 / h)).astype(np.float32), 0, axis)
    return out


def interp_aux_3d(maps1, maps2, P, x, y, z, rt):
    """interp_magnetic_fluctuation / interp_correlation_length with the eight trilinear weights: (n, 16)."""
    pc = [(x - P.xmin) / P.dx, (y - P.ymin) / P.dy, (z - P.zmin) / P.dz]
    idx = [np.floor(c).astype(np.int64) + 1 for c in pc]
    r = [c - i + 1 for c, i in zip(pc, idx)]
    r1 = [1.0 - v for v in r]
    out = []
    for m1, m2 in zip(maps1, maps2):
        f1 = np.zeros((len(x), 4))
        f2 = np.zeros((len(x), 4))
        for k in (0, 1):
            for j in (0, 1):
                for i in (0, 1):
                    w = (r[0] if i else r1[0]) * (r[1] if j else r1[1]) * (r[2] if k else r1[2])
                    sl = (idx[2] + k + 1, idx[1] + j + 1, idx[0] + i + 1)
                    f1 = f1 + m1[sl].astype(np.float64) * w[:, None]
                    f2 = f2 + m2[sl].astype(np.float64) * w[:, None]
        out.append(f1 * (1.0 - rt[:, None]) + f2 * rt[:, None])
    return np.concatenate(out, axis=1)
