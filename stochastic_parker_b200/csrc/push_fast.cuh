// push_fast.cuh -- production arithmetic of one push_particle_* call (GPAT_STRICT=0 build).
//
// Same physics as push_once() in push.cu (which keeps the reference's operation order for the
// parity build), restructured for the FP64 pipe of sm_100a:
//   * divisions by run constants (dx, dtf, p0, 3) are multiplications by reciprocals;
//   * b, 1/b come from one rsqrt; sqrt(2 kperp) and sqrt(2(kpara-kperp)) are sqrt(2 kpara)
//     times constants; the three sqrt(.)*sqrt(dt) products are one sqrt(2 kpara dt);
//   * the two pow() of kappa (particle_module.f90:2242,2258) are one exp of a sum of logs;
//   * the five dt candidates (particle_module.f90:3519-3523) are compared as fractions by
//     cross-multiplication and only the smallest one is divided out (1 reciprocal, not 5);
//   * the drift speed needs no 1/p: q/sqrt((d1 p0/p)^2 + (d2 p0^2/p^2)^2) =
//     q p^2 / sqrt((d1 p0)^2 p^2 + (d2 p0^2)^2);
//   * log, exp, reciprocal, rsqrt and sqrt are the straight-line versions of fastmath.cuh;
//   * the uniform -> [-sqrt3, sqrt3] map is a single FMA.
// Each change moves a result by a few ulp; tests hold this build to 1e-12 per step against the
// oracle (tests/test_gpu_parity.py::test_step_parity[0-*]).
#pragma once

// FP64 max/min as one compare + select (fmax/fmin add NaN handling: 4-5 instructions each); a NaN
// first argument yields the second, like fmax/fmin
__device__ __forceinline__ double max2(double a, double b) { return (a > b) ? a : b; }
__device__ __forceinline__ double min2f(double a, double b) { return (a < b) ? a : b; }

// F: the interpolated record of this particle (registers, or a shared-memory row written by
// the lane group that gathered it)
// SPEC != 0: the run's switches are compile-time constants: no NLGC, on-device Philox, no
// out-of-plane drift check, no acceleration region, time interpolation on (launch_one checks
// this), mag_dependency = bit 1 and momentum_dependency = bit 2 of SPEC.  ~15 uniform branches and
// their flag loads disappear and the log/exp/rsqrt chains share one basic block (-10 % executed
// instructions on C1, profiles/r01f_push_coop_spec_ncu.txt).  SPEC = 0 reads every switch at run time.
constexpr int kSpec11 = 1 | 2 | 4;  // mag_dependency = 1, momentum_dependency = 1 (C1, C2, C4)
constexpr int kSpec01 = 1 | 4;      // mag_dependency = 0, momentum_dependency = 1 (C3)
constexpr int kSpec10 = 1 | 2;      // mag_dependency = 1, momentum_dependency = 0 (C5; instantiated for L3D only)
// bit 0 clear: every switch is read at run time.  kSpecSurf adds the acc_by_surface gate of the 3-D
// pusher (particle_module.f90:4887-4892) to that generic code, so runs without surfaces never carry it.
constexpr int kSpecSurf = 8;
template <int L, typename FT, bool TRACK = false, int SPEC = 0>
__device__ __forceinline__ void physics_fast(const DevParams& prm, const PushArgs& a,
                                             const FT& F, Lane& q, bool fixed_dt)
{
    const bool f_mag = (SPEC & 1) ? bool(SPEC & 2) : prm.mag_dependency == 1;
    const bool f_mom = (SPEC & 1) ? bool(SPEC & 4) : prm.momentum_dependency == 1;
    const bool f_nlgc = !(SPEC & 1) && prm.nlgc;
    const bool f_table = !(SPEC & 1) && prm.rng_mode == GPAT_RNG_TABLE;
    const bool f_drift2d = !(SPEC & 1) && prm.check_drift_2d;
    const bool f_acc = !(SPEC & 1) && prm.acc_region_flag == 1;
    // tracked particles carry negated tags; the random streams are keyed by the magnitudes
    const int tag_inj = TRACK ? abs(q.tag_inj) : q.tag_inj, tag_spl = TRACK ? abs(q.tag_spl) : q.tag_spl;
    constexpr bool D3 = (Rec<L>::NDIM == 3);
    constexpr bool EXT = Rec<L>::EXT;

    // ---- uniforms -> ran1, ran2, ran3, ran_p in [-sqrt3, sqrt3] ----
    double ran1, ran2, ran3, ranp;
    {
        const double sqrt3 = 1.7320508075688772;
        if (f_table) {
            double u0 = 0.5, u1 = 0.5, u2 = 0.5, u3 = 0.5;
            long long slot = tag_inj;
            if (a.rng_table && slot >= 0 && slot < a.rng_slots && (long long)q.rng < a.rng_max_steps) {
                const double* tb = a.rng_table + ((size_t)slot * a.rng_max_steps + q.rng) * 4;
                u0 = tb[0]; u1 = tb[1]; u2 = tb[2]; u3 = tb[3];
            }
            ran1 = (2.0 * u0 - 1.0) * sqrt3; ran2 = (2.0 * u1 - 1.0) * sqrt3;
            ran3 = (2.0 * u2 - 1.0) * sqrt3; ranp = (2.0 * u3 - 1.0) * sqrt3;
        } else {
            uint4 r = philox4x32_10(make_uint4((unsigned)q.rng, (unsigned)(q.rng >> 32),
                                               (unsigned)tag_inj, (unsigned)tag_spl),
                                    prm.key0, prm.key1 + (unsigned)q.origin);
            const double c = 2.0 * sqrt3 / 4294967295.0;
            // u32 -> f64 through the 2^52 exponent trick (an FP64 add, not a conversion-unit op)
            const double two52 = 4503599627370496.0;
            ran1 = fma(__hiloint2double(0x43300000, (int)r.x) - two52, c, -sqrt3);
            ran2 = fma(__hiloint2double(0x43300000, (int)r.y) - two52, c, -sqrt3);
            ran3 = fma(__hiloint2double(0x43300000, (int)r.z) - two52, c, -sqrt3);
            ranp = fma(__hiloint2double(0x43300000, (int)r.w) - two52, c, -sqrt3);
        }
        q.rng += 1;
    }

    // ---- unpack the interpolated record ----
    double vx, vy, vz = 0.0, rho = 1.0;
    double bx, by, bz;
    double dbx_dx, dbx_dy, dbx_dz = 0.0, dby_dx, dby_dy, dby_dz = 0.0, dbz_dx, dbz_dy, dbz_dz = 0.0;
    double db_dx, db_dy, db_dz = 0.0;
    double dvx_dx, dvy_dy, dvz_dz = 0.0;
    double dvx_dy = 0.0, dvx_dz = 0.0, dvy_dx = 0.0, dvy_dz = 0.0, dvz_dx = 0.0, dvz_dy = 0.0;
    if constexpr (!D3) {
        vx = F[s2::vx]; vy = F[s2::vy];
        bx = F[s2::bx]; by = F[s2::by]; bz = F[s2::bz];
        dbx_dx = F[s2::dbx_dx]; dbx_dy = F[s2::dbx_dy]; dby_dx = F[s2::dby_dx]; dby_dy = F[s2::dby_dy];
        dbz_dx = F[s2::dbz_dx]; dbz_dy = F[s2::dbz_dy]; db_dx = F[s2::db_dx]; db_dy = F[s2::db_dy];
        dvx_dx = F[s2::dvx_dx]; dvy_dy = F[s2::dvy_dy];
        if constexpr (L == L2D) {  // rho in the pad slot of the base line, the shear gradients from the side plane
            rho = F[s2d::rho]; dvx_dy = F[s2d::dvx_dy]; dvy_dx = F[s2d::dvy_dx];
        } else if constexpr (EXT) {
            vz = F[s2::vz]; rho = F[s2::rho];
            dvx_dy = F[s2::dvx_dy]; dvy_dx = F[s2::dvy_dx]; dvz_dx = F[s2::dvz_dx]; dvz_dy = F[s2::dvz_dy];
        }
    } else {
        vx = F[s3::vx]; vy = F[s3::vy]; vz = F[s3::vz];
        bx = F[s3::bx]; by = F[s3::by]; bz = F[s3::bz];
        dbx_dx = F[s3::dbx_dx]; dbx_dy = F[s3::dbx_dy]; dbx_dz = F[s3::dbx_dz];
        dby_dx = F[s3::dby_dx]; dby_dy = F[s3::dby_dy]; dby_dz = F[s3::dby_dz];
        dbz_dx = F[s3::dbz_dx]; dbz_dy = F[s3::dbz_dy]; dbz_dz = F[s3::dbz_dz];
        db_dx = F[s3::db_dx]; db_dy = F[s3::db_dy]; db_dz = F[s3::db_dz];
        dvx_dx = F[s3::dvx_dx]; dvy_dy = F[s3::dvy_dy]; dvz_dz = F[s3::dvz_dz];
        if constexpr (EXT) {
            rho = F[s3::rho];
            dvx_dy = F[s3::dvx_dy]; dvx_dz = F[s3::dvx_dz]; dvy_dx = F[s3::dvy_dx];
            dvy_dz = F[s3::dvy_dz]; dvz_dx = F[s3::dvz_dx]; dvz_dy = F[s3::dvz_dy];
        }
    }
    const bool third = (Rec<L>::THIRD == 1) || (Rec<L>::THIRD == 2 && prm.include_3rd_dim);

    // ---- |B|, 1/|B| from one rsqrt ----
    const double b2 = bx * bx + by * by + bz * bz;
    const bool tiny = b2 < kEps * kEps;             // b < EPSILON(b)
    const double ibr = tiny ? 0.0 : fm::rsqrt(tiny ? 1.0 : b2);  // push_particle_*: ib = 0 for tiny b
    const double b = b2 * ibr;
    const double ibk = tiny ? 1.0 : ibr;            // kappa routine: ib1 = 1 for tiny b
    const double ibk2 = ibk * ibk, ibk3 = ibk2 * ibk;
    const double ib2 = ibr * ibr, ib3 = ib2 * ibr;

    // ---- kappa_para, kappa_perp (particle_module.f90:2239-2269 / 2497-2533) ----
    // log|B| = log(b2)/2; a vanishing field only has to stay finite here
    const double lb = f_mag ? 0.5 * fm::log_pos(max2(b2, 1e-300)) : 0.0;
    const double lpr = fm::log_pos(q.p * prm.ip0);
    double knp = 1.0, kpara, rk, srk, s1mrk;  // rk = kperp/kpara and its square roots
    if (EXT || f_nlgc) {
        if (f_mag) knp = fm::exp_mid(prm.gm2 * lb);
        const double pp = f_mom ? fm::exp_mid(prm.pindex * lpr) : 1.0;
        kpara = prm.kpara0 * knp * pp;
    } else {
        const double e = (f_mag ? prm.gm2 * lb : 0.0) +
                         (f_mom ? prm.pindex * lpr : 0.0);
        kpara = prm.kpara0 * fm::exp_mid(e);
    }
    if (!f_nlgc) {
        rk = prm.kret; srk = prm.sqrt_kret; s1mrk = prm.sqrt_1mkret;
    } else {
        const double e = (f_mag ? prm.gm2_3 * lb : 0.0) +
                         (f_mom ? prm.pidx_perp * lpr : 0.0);
        const double kperp = prm.kpara0 * prm.kperp_kpara * fm::exp_mid(e) * q.mu * q.mu;
        rk = kperp * fm::rcp(kpara);
        srk = fm::sqrt_pos(rk);
        s1mrk = fm::sqrt_pos(1.0 - rk);
    }
    const double kperp = kpara * rk;
    const double kpp = kpara - kperp;

    // ---- gradients of the kappa tensor (particle_module.f90:2377-2387, 2425-2448) ----
    double ax, ay, az = 0.0, ex, ey, ez = 0.0;  // kpp-like and kperp-like d(kappa)/dx_i factors
    {
        double gx = 0.0, gy = 0.0, gz = 0.0;
        if (f_mag) {
            // 3-D non-NLGC omits 1/B (particle_module.f90:2405-2409)
            const double s = (D3 && !f_nlgc) ? prm.gm2 : prm.gm2 * ibk;
            gx = db_dx * s; gy = db_dy * s; gz = db_dz * s;
        }
        if (!f_nlgc) {
            ex = kperp * gx; ey = kperp * gy; ez = kperp * gz;
            ax = kpp * gx; ay = kpp * gy; az = kpp * gz;
        } else {
            const double third_ = 1.0 / 3.0;
            ex = kperp * gx * third_; ey = kperp * gy * third_; ez = kperp * gz * third_;
            ax = kpara * gx - ex; ay = kpara * gy - ey; az = kpara * gz - ez;
        }
    }
    const double bxn2 = bx * bx * ibk2, byn2 = by * by * ibk2, bxyn2 = bx * by * ibk2;
    const double k2 = 2.0 * kpp * ibk3;
    const double dkxx_dx = ex + ax * bxn2 + k2 * bx * (dbx_dx * b - bx * db_dx);
    const double dkyy_dy = ey + ay * byn2 + k2 * by * (dby_dy * b - by * db_dy);
    const double dkxy_dx = ax * bxyn2 + kpp * ((dbx_dx * by + bx * dby_dx) * ibk2 - 2.0 * bx * by * db_dx * ibk3);
    const double dkxy_dy = ay * bxyn2 + kpp * ((dbx_dy * by + bx * dby_dy) * ibk2 - 2.0 * bx * by * db_dy * ibk3);

    // ---- drift (particle_module.f90:3436-3446 / 4739-4741) ----
    const double p2 = q.p * q.p;
    const double vdp = prm.qdrift * p2 * fm::rsqrt(fma(prm.d1p0sq, p2, prm.d2p02sq));
    double dx_dt, dy_dt, dz_dt, divv;
    if (!third) {
        const double vdx = vdp * (dbz_dy * ib2 - 2.0 * bz * db_dy * ib3);
        const double vdy = vdp * (-dbz_dx * ib2 + 2.0 * bz * db_dx * ib3);
        dz_dt = f_drift2d
                    ? vdp * ((dby_dx - dbx_dy) * ib2 - 2.0 * (by * db_dx - bx * db_dy) * ib3)
                    : 0.0;
        dx_dt = vx + vdx + dkxx_dx + dkxy_dy;
        dy_dt = vy + vdy + dkxy_dx + dkyy_dy;
        divv = dvx_dx + dvy_dy;
    } else {
        const double vdx = vdp * ((dbz_dy - dby_dz) * ib2 - 2.0 * (bz * db_dy - by * db_dz) * ib3);
        const double vdy = vdp * ((dbx_dz - dbz_dx) * ib2 - 2.0 * (bx * db_dz - bz * db_dx) * ib3);
        const double vdz = vdp * ((dby_dx - dbx_dy) * ib2 - 2.0 * (by * db_dx - bx * db_dy) * ib3);
        const double bzn2 = bz * bz * ibk2, bxzn2 = bx * bz * ibk2, byzn2 = by * bz * ibk2;
        const double dkxz_dx = ax * bxzn2 + kpp * ((dbx_dx * bz + bx * dbz_dx) * ibk2 - 2.0 * bx * bz * db_dx * ibk3);
        const double dkyz_dy = ay * byzn2 + kpp * ((dby_dy * bz + by * dbz_dy) * ibk2 - 2.0 * by * bz * db_dy * ibk3);
        double dkzz_dz = 0.0, dkxz_dz = 0.0, dkyz_dz = 0.0;
        if (D3) {  // the 2-D variant zeroes the d/dz terms (particle_module.f90:4094-4096)
            dkzz_dz = ez + az * bzn2 + k2 * bz * (dbz_dz * b - bz * db_dz);
            dkxz_dz = az * bxzn2 + kpp * ((dbx_dz * bz + bx * dbz_dz) * ibk2 - 2.0 * bx * bz * db_dz * ibk3);
            dkyz_dz = az * byzn2 + kpp * ((dby_dz * bz + by * dbz_dz) * ibk2 - 2.0 * by * bz * db_dz * ibk3);
        }
        dx_dt = vx + vdx + dkxx_dx + dkxy_dy + dkxz_dz;
        dy_dt = vy + vdy + dkxy_dx + dkyy_dy + dkyz_dz;
        dz_dt = vz + vdz + dkxz_dx + dkyz_dy + dkzz_dz;
        divv = dvx_dx + dvy_dy + dvz_dz;
    }
    double dp_dt = -q.p * divv * (1.0 / 3.0);

    // ---- momentum diffusion (particle_module.f90:2918-2979) ----
    double dpp = 0.0;
    if constexpr (EXT) {
        if (prm.dpp_wave) {
            const double pv = q.p * b2 * fm::rcp(rho * kpara);  // p va^2 / kpara
            dp_dt += (f_mom ? 8.0 / 27.0 : 4.0 / 9.0) * pv;
            dpp += q.p * pv * (1.0 / 9.0);
        }
        if (prm.dpp_shear) {
            const double d3 = divv * (1.0 / 3.0);
            const double sxx = dvx_dx - d3, syy = dvy_dy - d3, szz = dvz_dz - d3;
            const double sxy = 0.5 * (dvx_dy + dvy_dx);
            const double sxz = third ? 0.5 * (dvx_dz + dvz_dx) : 0.0;
            const double syz = third ? 0.5 * (dvy_dz + dvz_dy) : 0.0;
            double gshear;
            if (prm.weak_scattering) {
                double bbs = sxx * bx * bx + syy * by * by + szz * bz * bz +
                             2.0 * (sxy * bx * by + sxz * bx * bz + syz * by * bz);
                bbs = bbs * ib2;
                gshear = bbs * bbs * 0.2;
            } else {
                gshear = (2.0 / 15.0) * (sxx * sxx + syy * syy + szz * szz +
                                         2.0 * (sxy * sxy + sxz * sxz + syz * syz));
            }
            if (gshear > 0.0) {
                // p**pindex * p0**(2-pindex) = p0^2 (p/p0)**pindex
                const double pw = prm.p0 * prm.p0 * fm::exp_mid(prm.pindex * lpr);
                const double g = gshear * prm.tau0 * knp * pw;
                dp_dt += (2.0 + prm.pindex) * g * fm::rcp(q.p);
                dpp += g;
            }
        }
    }

    // ---- time step (particle_module.f90:3499-3543 / 4791-4843) ----
    if (!fixed_dt) {
        double d;
        if (dx_dt != 0.0 && dy_dt != 0.0 && dp_dt != 0.0) {
            // three positive fractions n/dn; pick the smallest by cross-multiplication
            double m = max2(dx_dt * dx_dt, dy_dt * dy_dt);
            if (D3) m = max2(m, dz_dt * dz_dt);  // (s/0)^2 = +Inf is ignored by min, as in the reference
            double n = (D3 ? prm.hd2min3 : prm.hd2min2) * 0.5, dn = kpara;
            const double n1 = 2.0 * ((rk > 0.0) ? kperp : kpara);
            if (n1 * dn < n * m) { n = n1; dn = m; }
            const double n2 = (double)0.1f * q.p, d2 = fabs(dp_dt);
            if (n2 * dn < n * d2) { n = n2; dn = d2; }
            d = n * fm::rcp(dn);
        } else {
            d = a.dt_min;
        }
        d = max2(d, a.dt_min);
        d = min2f(d, a.dt_max);
        q.dt = d;
    }

    // ---- stochastic step ----
    const double sA = fm::sqrt_pos(2.0 * kpara * q.dt);  // sqrt(2 kpara) sqrt(dt)
    double ddx, ddy, ddz;
    if (!third) {
        const double sp = sA * srk, spp = sA * s1mrk * ran3 * ibr;
        ddx = fma(dx_dt, q.dt, fma(ran1, sp, spp * bx));
        ddy = fma(dy_dt, q.dt, fma(ran2, sp, spp * by));
        ddz = dz_dt * q.dt;
        q.dzl = 0.0;  // the mover's own deltaz stays 0 in plain 2-D (particle_module.f90:1564)
    } else {
        const double bxn = bx * ibr, byn = by * ibr, bzn = bz * ibr;
        const double h2 = bxn * bxn + byn * byn;
        const double ih = (h2 < kEps * kEps) ? 0.0 : fm::rsqrt(h2 < kEps * kEps ? 1.0 : h2);
        const double hxy = h2 * ih;
        const double sp = sA * srk;
        const double t2 = sp * ih * ran2, t3 = sp * ih * ran3, t1 = sA * ran1;
        ddx = fma(dx_dt, q.dt, bxn * t1 - bxn * bzn * t2 - byn * t3);
        ddy = fma(dy_dt, q.dt, byn * t1 - byn * bzn * t2 + bxn * t3);
        ddz = fma(dz_dt, q.dt, bzn * t1 + hxy * sp * ran2);
        q.dzl = ddz;
    }
    double sh1 = 0.0, sh2 = 0.0;
    if constexpr (D3 && bool(SPEC & kSpecSurf)) {  // interp_acc_surface at the OLD position, particle_module.f90:1683-1686
        if (f_acc) surface_heights(prm, a, q.x, q.y, q.z, (q.t - a.t0) * a.idtf, sh1, sh2);
    }
    q.x += ddx;
    q.y += ddy;
    q.z += ddz;
    q.t += q.dt;
    q.dxl = ddx;
    q.dyl = ddy;

    double ddp = dp_dt * q.dt;
    if constexpr (EXT) ddp = fma(ranp, fm::sqrt_pos(2.0 * dpp * q.dt), ddp);
    if (f_acc) {
        bool in = in_acc_region(prm, q);
        if constexpr (D3 && bool(SPEC & kSpecSurf)) in = in && above_surface(prm, q, sh1, sh2);
        if (in) q.p += ddp;
        else ddp = 0.0;
    } else {
        q.p += ddp;
    }
    if (q.p < prm.pfloor) {  // particle_module.f90:3601-3605
        q.p -= ddp;
        ddp = prm.pfloor - q.p;
        q.p = prm.pfloor;
    }
    q.dpl = ddp;
}

template <int L, bool TRACK = false>
__device__ __forceinline__ void push_once_fast(const DevParams& prm, const PushArgs& a,
                                               const float* __restrict__ fld, Lane& q, bool fixed_dt)
{
    double F[Rec<L>::NF];
    const double rt = (q.t - a.t0) * a.idtf;
    gather<L>(prm, fld, a.sel, q.x, q.y, q.z, rt, F);
    if constexpr (Rec<L>::NDIM == 3) {
        if (prm.acc_by_surface) {
            physics_fast<L, double[Rec<L>::NF], TRACK, kSpecSurf>(prm, a, F, q, fixed_dt);
            return;
        }
    }
    physics_fast<L, double[Rec<L>::NF], TRACK>(prm, a, F, q, fixed_dt);
}
