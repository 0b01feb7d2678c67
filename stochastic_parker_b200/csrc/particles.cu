// particles.cu -- injection, splitting, dead-particle compaction, AoS<->SoA conversion.
//
//  * inject_kernel   : inject_particles_spatial_uniform + inject_one_particle
//                      (particle_module.f90:454-530, 385-441), one thread per new particle
//  * split           : split_particle (particle_module.f90:5430-5480) as flag -> scan -> append
//  * remove          : remove_particles (particle_module.f90:5365-5403) as stream compaction
//                      that reproduces the reference's swap-with-tail ORDER exactly
//  * to_aos/from_aos : particle_type records for dumps and restart
//
// The scans are hand-written "warp ballot + block scan" stream compaction: pass A counts
// selected items per 1024-item tile (__syncthreads_count), pass B scans the tile counts in
// one block, pass C recomputes the predicate and writes each selected index at
// tile offset + warp offset + popc(ballot below lane).
#include "gpat_internal.cuh"

namespace gpat {

constexpr int kTile = 1024;

// ---- Philox (same algorithm as push.cu; injection stream) -----------------------------------
__device__ __forceinline__ uint4 philox_inj(uint4 c, unsigned k0, unsigned k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// SoA slot -> the reference's particle_type record (for particles_tracked)
__device__ __forceinline__ gpat_particle soa_record(const PtlSoA& P, long long i)
{
    gpat_particle q;
    q.split_times = P.split_times[i]; q.count_flag = P.count_flag[i]; q.pad_[0] = q.pad_[1] = 0;
    q.origin = P.origin[i]; q.nsteps_tracked = P.nsteps_tracked[i];
    q.nsteps_pushed = P.nsteps_pushed[i]; q.tag_injected = P.tag_injected[i];
    q.tag_splitted = P.tag_splitted[i];
    q.x = P.x[i]; q.y = P.y[i]; q.z = P.z[i]; q.p = P.p[i]; q.v = P.v[i]; q.mu = P.mu[i];
    q.weight = P.weight[i]; q.t = P.t[i]; q.dt = P.dt[i];
    q.padding = __longlong_as_double((long long)P.rng[i]);
    return q;
}

struct InjStream {  // word k of ctr = (k/4, 0, tag_injected, 0)
    unsigned tag, k0, k1, k;
    uint4 buf;
    __device__ double next()
    {
        if ((k & 3u) == 0) buf = philox_inj(make_uint4(k >> 2, 0u, tag, 0u), k0, k1);
        unsigned w = (k & 3u) == 0 ? buf.x : (k & 3u) == 1 ? buf.y : (k & 3u) == 2 ? buf.z : buf.w;
        k++;
        return __ddiv_rn((double)w, 4294967295.0);
    }
};

struct InjectArgs {
    long long n, start, nptl_max;
    long long tag0;
    double dt, particle_v0, t_frame, dt_mhd, power_index;
    double box[6];
    int dist_flag;
    // targeted injection (inject_particles_at_large_jz/_absj/_divv/_rho, particle_module.f90:785-1468)
    int mode;            // 0 uniform in part_box; GPAT_INJECT_LARGE_*
    double vmin;         // jz_min / absj_min / divv_min / rho_min
    const float* fld;    // packed field store
    int nrec, half;      // floats per record half-pair / which half is farray1
    int pos[10];         // packed position of dbx_dy dbx_dz dby_dx dby_dz dbz_dx dbz_dy dvx_dx dvy_dy dvz_dz rho (-1: absent = 0)
    int* fail;           // set when a rejection loop hits kMaxTrials
    TrackDev trk;        // particle tracking (particle_module.f90:434-440)
    const int* shock_x;  // mode GPAT_INJECT_AT_SHOCK: shock_xpos2(nyg, nzg), locate_shock_xpos
    const float* aux;    // mode GPAT_INJECT_LARGE_DB2: turbulence maps (sigma2_slab is chunk 0)
};

constexpr int kMaxTrials = 1 << 22;  // the reference's loop is unbounded; a kernel must end

// one reference slot of farray1 at storage cell `cell` (gpat_internal.cuh record layout)
__device__ __forceinline__ float rec_get(const float* __restrict__ fld, long long cell, int nrec, int half, int pos)
{
    return fld[cell * (2LL * nrec) + (pos >> 2) * 8 + half * 4 + (pos & 3)];
}

// get_interp_paramters + interp_fields at rt = 0 for ONE slot, reference summation order, no
// contraction (particle_module.f90:642-674, mhd_data_parallel.f90:1776-1792)
struct InjInterp {
    long long cell[8];
    double w[8];
    int nc;
    __device__ void locate(const DevParams& prm, double x, double y, double z)
    {
        const double px = __ddiv_rn(__dsub_rn(x, prm.xmin), prm.dx);
        const double py = __ddiv_rn(__dsub_rn(y, prm.ymin), prm.dy);
        const double pz = __ddiv_rn(__dsub_rn(z, prm.zmin), prm.dz);
        const int ix = (int)floor(px) + 1;
        const int iy = (prm.ndim > 1) ? (int)floor(py) + 1 : 1;
        const int iz = (prm.ndim > 2) ? (int)floor(pz) + 1 : 1;
        const double rx = __dadd_rn(__dsub_rn(px, (double)ix), 1.0);
        const double ry = (prm.ndim > 1) ? __dadd_rn(__dsub_rn(py, (double)iy), 1.0) : 0.0;
        const double rz = (prm.ndim > 2) ? __dadd_rn(__dsub_rn(pz, (double)iz), 1.0) : 0.0;
        const double rx1 = __dsub_rn(1.0, rx), ry1 = __dsub_rn(1.0, ry), rz1 = __dsub_rn(1.0, rz);
        const int cx = min(max(ix + 1, 0), prm.nxg - 2);
        const int cy = (prm.ndim > 1) ? min(max(iy + 1, 0), prm.nyg - 2) : 0;
        const int cz = (prm.ndim > 2) ? min(max(iz + 1, 0), prm.nzg - 2) : 0;
        nc = (prm.ndim > 2) ? 8 : (prm.ndim > 1 ? 4 : 2);
        for (int c = 0; c < nc; ++c) {
            const int i = c & 1, j = (c >> 1) & 1, k = c >> 2;
            cell[c] = ((long long)(cz + k) * prm.nyg + (cy + j)) * prm.nxg + (cx + i);
            const double wx = i ? rx : rx1, wy = j ? ry : ry1, wz = k ? rz : rz1;
            w[c] = __dmul_rn(__dmul_rn(wx, wy), wz);
        }
    }
    __device__ double slot(const InjectArgs& a, int pos) const
    {
        if (pos < 0) return 0.0;  // a gradient the layout does not store is identically zero (d/dz in 2-D)
        double f = 0.0;
        for (int c = 0; c < nc; ++c)
            f = __dadd_rn(f, __dmul_rn((double)rec_get(a.fld, cell[c], a.nrec, a.half, pos), w[c]));
        return f;
    }
};

__device__ double inject_criterion(const DevParams& prm, const InjectArgs& a, double x, double y, double z)
{
    InjInterp I;
    I.locate(prm, x, y, z);
    if (a.mode == GPAT_INJECT_LARGE_DB2) {  // db2_slab(1) at rt = 0, particle_module.f90:1173-1174
        double f = 0.0;
        for (int c = 0; c < I.nc; ++c)
            f = __dadd_rn(f, __dmul_rn((double)a.aux[I.cell[c] * 32 + a.half * 4], I.w[c]));
        return f;
    }
    if (a.mode == GPAT_INJECT_LARGE_JZ)
        return fabs(__dsub_rn(I.slot(a, a.pos[2]), I.slot(a, a.pos[0])));
    if (a.mode == GPAT_INJECT_LARGE_ABSJ) {
        const double dbx_dy = I.slot(a, a.pos[0]), dbx_dz = I.slot(a, a.pos[1]), dby_dx = I.slot(a, a.pos[2]);
        const double dby_dz = I.slot(a, a.pos[3]), dbz_dx = I.slot(a, a.pos[4]), dbz_dy = I.slot(a, a.pos[5]);
        const double j1 = __dsub_rn(dby_dz, dbz_dy), j2 = __dsub_rn(dbz_dx, dbx_dz), j3 = __dsub_rn(dbx_dy, dby_dx);
        return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(j1, j1), __dmul_rn(j2, j2)), __dmul_rn(j3, j3)));
    }
    if (a.mode == GPAT_INJECT_LARGE_DIVV) {
        double d = I.slot(a, a.pos[6]);
        if (prm.ndim > 1) {
            d = __dadd_rn(d, I.slot(a, a.pos[7]));
            if (prm.ndim > 2) d = __dadd_rn(d, I.slot(a, a.pos[8]));
        }
        return d;
    }
    return I.slot(a, a.pos[9]);  // rho
}

__global__ void inject_kernel(const __grid_constant__ DevParams prm, const PtlSoA P,
                              const __grid_constant__ InjectArgs a)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    // particle_module.f90:491-492: past capacity every particle lands in slot nptl_max; the
    // serial loop leaves the LAST one there.
    long long slot = a.start + i;
    if (a.start + a.n > a.nptl_max && slot >= a.nptl_max - 1) {
        if (i != a.n - 1) return;  // slot nptl_max is owned by the last injected particle
        slot = a.nptl_max - 1;
    }
    InjStream s{(unsigned)(a.tag0 + i), prm.key0, prm.key1 + (unsigned)prm.mpi_rank, 0u, make_uint4(0, 0, 0, 0)};
    const double mu_max = (double)0.99f;  // particle_module.f90:121
    // no contraction here: these are parity-checked bit for bit against the oracle
    double x, y, z;
    if (a.mode == GPAT_INJECT_AT_SHOCK) {
        // inject_particles_at_shock (particle_module.f90:570-587), kept as written: rz from dpy, time
        // weights swapped (rt = 0 picks shock_xpos2), two ghost cells added to the index, scaled by
        // (xmax - xmin)/nxg.  sx2 starts at zero here (uninitialised in the reference, see the header).
        y = __dadd_rn(__dmul_rn(s.next(), prm.ymax - prm.ymin), prm.ymin);
        const double dpy = __ddiv_rn(y, prm.dy);
        const int iy = (int)floor(dpy);
        z = __dadd_rn(__dmul_rn(s.next(), prm.zmax - prm.zmin), prm.zmin);
        const int iz = (int)floor(__ddiv_rn(z, prm.dz));
        const double ry = __dsub_rn(dpy, (double)iy), rz = ry;
        const double w[4] = {__dmul_rn(__dsub_rn(1.0, ry), __dsub_rn(1.0, rz)), __dmul_rn(ry, __dsub_rn(1.0, rz)),
                             __dmul_rn(__dsub_rn(1.0, ry), rz), __dmul_rn(ry, rz)};
        double sx2 = 0.0;
        const int nyr = (prm.ndim > 1) ? prm.ny + 4 : 1, nzr = (prm.ndim > 2) ? prm.nz + 4 : 1;
        if (prm.ndim == 1) {
            sx2 = (double)a.shock_x[0];
        } else if (prm.ndim == 2) {
            for (int j = 0; j <= 1; ++j) {
                const int fj = min(max(iy + j, 0), nyr - 1);
                sx2 = __dadd_rn(sx2, __dmul_rn((double)a.shock_x[fj], w[j]));
            }
        } else {
            for (int k = 0; k <= 1; ++k)
                for (int j = 0; j <= 1; ++j) {
                    const int fj = min(max(iy + j, 0), nyr - 1), fk = min(max(iz + k, 0), nzr - 1);
                    sx2 = __dadd_rn(sx2, __dmul_rn((double)a.shock_x[fj + (long long)nyr * fk], w[k * 2 + j]));
                }
        }
        const double shock_xpos = __dadd_rn(sx2, 2.0);
        x = __ddiv_rn(__dmul_rn(shock_xpos, prm.xmax - prm.xmin), (double)(prm.nx + 4));
    } else if (a.mode == 0) {
        x = __dadd_rn(__dmul_rn(s.next(), a.box[3] - a.box[0]), a.box[0]);
        y = __dadd_rn(__dmul_rn(s.next(), a.box[4] - a.box[1]), a.box[1]);
        z = __dadd_rn(__dmul_rn(s.next(), a.box[5] - a.box[2]), a.box[2]);
    } else {
        // rejection loop of inject_particles_at_large_* (e.g. particle_module.f90:860-899): positions
        // uniform in the whole domain, criterion = interpolated field at rt = 0, a position outside
        // part_box counts as a miss.  Initial / miss values as in the reference.
        const bool divv = (a.mode == GPAT_INJECT_LARGE_DIVV), rho = (a.mode == GPAT_INJECT_LARGE_RHO);
        double crit = divv ? 2.0 : (rho ? 0.0 : -2.0);
        x = a.box[0]; y = a.box[1]; z = a.box[2];
        int trials = 0;
        while (divv ? (-crit < a.vmin) : (crit < a.vmin)) {
            if (++trials > kMaxTrials) { *a.fail = 1; break; }
            x = __dadd_rn(__dmul_rn(s.next(), prm.xmax - prm.xmin), prm.xmin);
            y = __dadd_rn(__dmul_rn(s.next(), prm.ymax - prm.ymin), prm.ymin);
            z = __dadd_rn(__dmul_rn(s.next(), prm.zmax - prm.zmin), prm.zmin);
            if (x >= a.box[0] && x <= a.box[3] && y >= a.box[1] && y <= a.box[4] && z >= a.box[2] && z <= a.box[5])
                crit = inject_criterion(prm, a, x, y, z);
            else
                crit = divv ? 3.0 : (rho ? 0.0 : -3.0);
        }
    }
    const bool shock = (a.mode == GPAT_INJECT_AT_SHOCK);
    // the shock injector draws mu AFTER the momentum (particle_module.f90:611), the others before
    double mu = 0.0;
    if (!shock) mu = __dmul_rn(mu_max, __dsub_rn(__dmul_rn(2.0, s.next()), 1.0));
    double p;
    if (a.dist_flag == 0) {  // particle_module.f90:399-407 / 588-596 (shock: another envelope)
        double ftest = 1.0, fxp = 0.5, ptmp = 0.0;
        while (ftest > fxp) {
            ptmp = __ddiv_rn(__dadd_rn(__dmul_rn(s.next(), prm.pmax - prm.pmin), prm.pmin), prm.p0);
            double p2 = __dmul_rn(ptmp, ptmp);
            fxp = shock ? __dmul_rn(p2, exp(__dmul_rn(-0.5, p2))) : __dmul_rn(p2, exp(-p2));
            ftest = __dmul_rn(s.next(), shock ? (double)0.75f : (double)0.37f);
        }
        p = __dmul_rn(ptmp, prm.p0);
    } else if (a.dist_flag == 2) {  // particle_module.f90:410-418
        double r01 = s.next();
        if ((int)a.power_index == 1) {
            p = __dmul_rn(pow(prm.pmax / prm.p0, r01), prm.p0);
        } else {
            double e = -a.power_index + 1;
            double norm = pow(prm.pmax, e) - pow(prm.p0, e);
            p = pow(__dadd_rn(__dmul_rn(r01, norm), pow(prm.p0, e)), 1.0 / e);
        }
    } else {
        p = prm.p0;
    }
    P.x[slot] = x; P.y[slot] = y; P.z[slot] = z; P.p[slot] = p;
    P.v[slot] = __ddiv_rn(__dmul_rn(a.particle_v0, p), prm.p0);
    P.weight[slot] = 1.0;
    if (shock) mu = __dmul_rn(mu_max, __dsub_rn(__dmul_rn(2.0, s.next()), 1.0));
    P.mu[slot] = mu;
    P.t[slot] = shock ? a.t_frame : __dadd_rn(a.t_frame, __dmul_rn(s.next(), a.dt_mhd));  // particle_module.f90:613
    P.dt[slot] = a.dt;
    P.rng[slot] = 0ull;
    P.split_times[slot] = 0;
    P.count_flag[slot] = GPAT_COUNT_FLAG_INBOX;
    P.origin[slot] = prm.mpi_rank;
    P.nsteps_tracked[slot] = 0;
    P.nsteps_pushed[slot] = 0;
    P.tag_injected[slot] = (int)(a.tag0 + i);
    P.tag_splitted[slot] = 1;
    if (a.trk.enabled) {  // particle_module.f90:434-440
        long long lo, hi;
        if (trk_selected(a.trk, prm.mpi_rank, (int)(a.tag0 + i), 1, 0, lo, hi)) {
            P.nsteps_tracked[slot] = 1;
            P.tag_injected[slot] = -(int)(a.tag0 + i);
            P.tag_splitted[slot] = -1;
            trk_record(a.trk, soa_record(P, slot), lo, hi);
        }
    }
}

// ---- get_ncells_large_jz/_absj/_divv/_rho (mhd_data_parallel.f90:2211-2498) ---------------------
// One thread per physical cell; the cell positions are xpos_local(ix) = dx*(ix-1) + xmin.
__global__ void ncells_kernel(const __grid_constant__ DevParams prm, const __grid_constant__ InjectArgs a,
                              unsigned long long* count)
{
    const long long ncell = (long long)prm.nx * prm.ny * prm.nz;
    unsigned long long mine = 0;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell;
         c += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(c % prm.nx) + 1, iy = (int)((c / prm.nx) % prm.ny) + 1;
        const int iz = (int)(c / ((long long)prm.nx * prm.ny)) + 1;
        const double xp = __dadd_rn(__dmul_rn(prm.dx, (double)(ix - 1)), prm.xmin);
        bool in = xp > a.box[0] && xp < a.box[3];
        if (prm.ndim > 1) {
            const double yp = __dadd_rn(__dmul_rn(prm.dy, (double)(iy - 1)), prm.ymin);
            in = in && yp > a.box[1] && yp < a.box[4];
        }
        if (prm.ndim > 2) {
            const double zp = __dadd_rn(__dmul_rn(prm.dz, (double)(iz - 1)), prm.zmin);
            in = in && zp > a.box[2] && zp < a.box[5];
        }
        if (!in) continue;
        // Fortran index -> storage index: +1 on resolved axes
        auto cell_of = [&](int fx, int fy, int fz) {
            const int cy = (prm.ndim > 1) ? fy + 1 : 0, cz = (prm.ndim > 2) ? fz + 1 : 0;
            return ((long long)cz * prm.nyg + cy) * prm.nxg + (fx + 1);
        };
        auto get = [&](long long cell, int pos) { return pos < 0 ? 0.0f : rec_get(a.fld, cell, a.nrec, a.half, pos); };
        double v;
        if (a.mode == GPAT_INJECT_LARGE_DB2) {  // sigma2_slab_1(1, ix, iy, iz), mhd_data_parallel.f90:2371
            v = (double)a.aux[cell_of(ix, iy, iz) * 32 + a.half * 4];
        } else if (a.mode == GPAT_INJECT_LARGE_JZ) {  // FP32 difference and abs, mhd_data_parallel.f90:2235
            const long long cc = cell_of(ix, iy, iz);
            v = (double)fabsf(__fsub_rn(get(cc, a.pos[2]), get(cc, a.pos[0])));
        } else if (a.mode == GPAT_INJECT_LARGE_ABSJ) {  // all FP32, mhd_data_parallel.f90:2306-2311
            const long long cc = cell_of(ix, iy, iz);
            const float j1 = __fsub_rn(get(cc, a.pos[3]), get(cc, a.pos[5]));
            const float j2 = __fsub_rn(get(cc, a.pos[4]), get(cc, a.pos[1]));
            const float j3 = __fsub_rn(get(cc, a.pos[0]), get(cc, a.pos[2]));
            const float s2 = __fadd_rn(__fadd_rn(__fmul_rn(j1, j1), __fmul_rn(j2, j2)), __fmul_rn(j3, j3));
            v = (double)__fsqrt_rn(s2);
        } else if (a.mode == GPAT_INJECT_LARGE_DIVV) {
            // mhd_data_parallel.f90:2417-2424 assigns the whole ghosted array to an allocatable of
            // shape (nx,ny,nz); after reallocation-on-assignment divv(ix,iy,iz) is the value two
            // cells to the lower-left on every resolved axis.  Reproduced, not repaired.
            const long long cc = cell_of(ix - 2, (prm.ndim > 1) ? iy - 2 : iy, (prm.ndim > 2) ? iz - 2 : iz);
            double d = (double)get(cc, a.pos[6]);
            if (prm.ndim > 1) {
                d = __dadd_rn(d, (double)get(cc, a.pos[7]));
                if (prm.ndim > 2) d = __dadd_rn(d, (double)get(cc, a.pos[8]));
            }
            v = -d;
        } else {
            v = (double)get(cell_of(ix, iy, iz), a.pos[9]);
        }
        if (v > a.vmin) mine++;
    }
    // warp-aggregated count
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, mine);
}

// locate_shock_xpos (mhd_data_parallel.f90:1988-2006): maxloc(abs(farray(nfields+1, :, j, k)), dim=1)
// for every row of the ghosted extent; one warp per row, first maximum wins.
__global__ void shock_xpos_kernel(const __grid_constant__ DevParams prm, const float* __restrict__ fld, int nrec,
                                  int half, int pos, int nyr, int nzr, int* __restrict__ out)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nyr * nzr) return;
    const int j = row % nyr, k = row / nyr;
    float best = -1.0f;
    int at = 1;
    for (int i = lane; i < prm.nxg; i += 32) {
        const long long cell = ((long long)k * prm.nyg + j) * prm.nxg + i;
        const float v = fabsf(rec_get(fld, cell, nrec, half, pos));
        if (v > best) { best = v; at = i + 1; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oa = __shfl_down_sync(0xffffffffu, at, o);
        if (ob > best || (ob == best && oa < at)) { best = ob; at = oa; }
    }
    if (lane == 0) out[row] = at;
}

void launch_shock_xpos(const DevParams& prm, int layout, const float* fld, int half, int* d_out, cudaStream_t st)
{
    const int nrec = nrec_of(layout);
    int pos = -1;
    for (int k = 0; k < nrec; ++k)
        if (slot_of(layout, k) == 8 + 1) pos = k;  // dvx_dx
    const int nyr = (prm.ndim > 1) ? prm.ny + 4 : 1, nzr = (prm.ndim > 2) ? prm.nz + 4 : 1;
    const int rows = nyr * nzr;
    shock_xpos_kernel<<<(rows + 3) / 4, 128, 0, st>>>(prm, fld, nrec, half, pos, nyr, nzr, d_out);
}

static void fill_target(InjectArgs& a, const DevParams& prm, int layout, const float* fld, int sel,
                        int mode, double vmin, const double box[6], int* fail)
{
    a.mode = mode; a.vmin = vmin; a.fld = fld; a.nrec = nrec_of(layout);
    a.half = prm.time_interp ? sel : 0;
    a.fail = fail;
    for (int i = 0; i < 6; ++i) a.box[i] = box[i];
    const int want[10] = {8 + 14, 8 + 15, 8 + 16, 8 + 18, 8 + 19, 8 + 20, 8 + 1, 8 + 5, 8 + 9, 4};
    for (int i = 0; i < 10; ++i) {
        a.pos[i] = -1;
        for (int k = 0; k < a.nrec; ++k)
            if (slot_of(layout, k) == want[i]) a.pos[i] = k;
    }
}

void launch_ncells(const DevParams& prm, int layout, const float* fld, int sel, int mode, double vmin,
                   const double box[6], unsigned long long* d_count, int sm_count, cudaStream_t st, const float* aux)
{
    InjectArgs a{};
    fill_target(a, prm, layout, fld, sel, mode, vmin, box, nullptr);
    a.aux = aux;
    cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st);
    ncells_kernel<<<sm_count * 4, 256, 0, st>>>(prm, a, d_count);
}

void launch_inject(const DevParams& prm, const PtlSoA& P, long long n, long long start,
                   long long nptl_max, long long tag0, double dt, int dist_flag, double particle_v0,
                   double t_frame, double dt_mhd, const double box[6], double power_index,
                   cudaStream_t st, int mode, double vmin, int layout, const float* fld, int sel, int* fail,
                   const TrackDev* trk, const int* shock_x, const float* aux)
{
    if (n <= 0) return;
    InjectArgs a{};
    if (trk) a.trk = *trk;
    a.shock_x = shock_x;
    if (mode == GPAT_INJECT_AT_SHOCK) a.mode = mode;
    if (mode != 0 && mode != GPAT_INJECT_AT_SHOCK) fill_target(a, prm, layout, fld, sel, mode, vmin, box, fail);
    a.aux = aux;
    a.n = n; a.start = start; a.nptl_max = nptl_max; a.tag0 = tag0; a.dt = dt;
    a.particle_v0 = particle_v0; a.t_frame = t_frame; a.dt_mhd = dt_mhd;
    a.power_index = power_index; a.dist_flag = dist_flag;
    for (int i = 0; i < 6; ++i) a.box[i] = box[i];
    inject_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(prm, P, a);
}

// ---- stream compaction primitives --------------------------------------------------------
enum Pred : int { PRED_ALIVE = 0, PRED_HOLE, PRED_FILLER, PRED_ESCAPED, PRED_SPLIT };

struct PredArgs {
    const signed char* count_flag;
    const signed char* split_times;
    const double* p;
    const long long* m;  // device: number of alive particles (for HOLE / FILLER)
    double thr0, split_ratio, pmax;  // SPLIT: p > thr0 * split_ratio**split_times, p <= pmax
};

// libgcc's __powidf2, which is what gfortran emits for real(dp)**integer
__device__ __forceinline__ double powi(double x, int mexp)
{
    unsigned n = (mexp < 0) ? (unsigned)(-mexp) : (unsigned)mexp;
    double y = (n & 1u) ? x : 1.0;
    while (n >>= 1) {
        x = __dmul_rn(x, x);
        if (n & 1u) y = __dmul_rn(y, x);
    }
    return (mexp < 0) ? 1.0 / y : y;
}

template <int PRED>
__device__ __forceinline__ bool pred(const PredArgs& a, long long i)
{
    if (PRED == PRED_ALIVE) return a.count_flag[i] == GPAT_COUNT_FLAG_INBOX;
    if (PRED == PRED_HOLE) return i < *a.m && a.count_flag[i] != GPAT_COUNT_FLAG_INBOX;
    if (PRED == PRED_FILLER) return i >= *a.m && a.count_flag[i] == GPAT_COUNT_FLAG_INBOX;
    if (PRED == PRED_ESCAPED) return a.count_flag[i] < 0;
    // particle_module.f90:5441-5442
    double thr = __dmul_rn(a.thr0, powi(a.split_ratio, (int)a.split_times[i]));
    return a.p[i] > thr && a.p[i] <= a.pmax;
}

template <int PRED>
__global__ void __launch_bounds__(kTile) tile_count_kernel(PredArgs a, long long n, unsigned* tile_counts)
{
    long long i = blockIdx.x * (long long)kTile + threadIdx.x;
    int f = (i < n) ? (int)pred<PRED>(a, i) : 0;
    int c = __syncthreads_count(f);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = (unsigned)c;
}

// exclusive scan of the tile counts by one block; total -> *total
__global__ void __launch_bounds__(1024) tile_scan_kernel(unsigned* tile_counts, long long ntiles,
                                                          long long* tile_offsets, long long* total)
{
    __shared__ long long warp_sums[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    for (long long base = 0; base < ntiles; base += 1024) {
        long long i = base + threadIdx.x;
        long long v = (i < ntiles) ? (long long)tile_counts[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            long long ws = warp_sums[lane];
            long long wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (unsigned)o) wi += t;
            }
            warp_sums[lane] = wi - ws;  // exclusive warp offsets
        }
        __syncthreads();
        long long excl = carry + warp_sums[wid] + incl - v;
        if (i < ntiles) tile_offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// writes the index of every selected item, in ascending order, to out[]
template <int PRED>
__global__ void __launch_bounds__(kTile) tile_scatter_kernel(PredArgs a, long long n,
                                                             const long long* tile_offsets,
                                                             long long* out)
{
    __shared__ unsigned warp_cnt[32];
    long long i = blockIdx.x * (long long)kTile + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    bool f = (i < n) && pred<PRED>(a, i);
    unsigned b = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_cnt[wid] = __popc(b);
    __syncthreads();
    if (wid == 0) {
        unsigned c = warp_cnt[lane], incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        warp_cnt[lane] = incl - c;
    }
    __syncthreads();
    if (f) out[tile_offsets[blockIdx.x] + warp_cnt[wid] + __popc(b & ((1u << lane) - 1u))] = i;
}

template <int PRED>
static void select_indices(const PredArgs& a, long long n, const ScanWork& w, long long* out,
                           long long* total, cudaStream_t st)
{
    long long ntiles = (n + kTile - 1) / kTile;
    if (ntiles == 0) {
        cudaMemsetAsync(total, 0, sizeof(long long), st);
        return;
    }
    tile_count_kernel<PRED><<<(unsigned)ntiles, kTile, 0, st>>>(a, n, w.tile_counts);
    tile_scan_kernel<<<1, 1024, 0, st>>>(w.tile_counts, ntiles, w.tile_offsets, total);
    if (out) tile_scatter_kernel<PRED><<<(unsigned)ntiles, kTile, 0, st>>>(a, n, w.tile_offsets, out);
}

__device__ __forceinline__ void copy_particle(const PtlSoA& D, long long d, const PtlSoA& S, long long s)
{
    D.x[d] = S.x[s]; D.y[d] = S.y[s]; D.z[d] = S.z[s]; D.p[d] = S.p[s]; D.v[d] = S.v[s];
    D.mu[d] = S.mu[s]; D.weight[d] = S.weight[s]; D.t[d] = S.t[s]; D.dt[d] = S.dt[s];
    D.rng[d] = S.rng[s]; D.origin[d] = S.origin[s]; D.nsteps_tracked[d] = S.nsteps_tracked[s];
    D.nsteps_pushed[d] = S.nsteps_pushed[s]; D.tag_injected[d] = S.tag_injected[s];
    D.tag_splitted[d] = S.tag_splitted[s]; D.split_times[d] = S.split_times[s];
    D.count_flag[d] = S.count_flag[s];
}

// ---- remove_particles --------------------------------------------------------------------
// The serial reference walks i upward and swaps every non-INBOX particle with the current
// tail.  Net effect: the k-th hole (ascending) among the first m = #alive slots receives the
// k-th alive particle counted DOWN from the end.  Reproduced here with two index lists.
__global__ void fill_holes_kernel(PtlSoA P, const long long* holes, const long long* fillers,
                                  const long long* nholes)
{
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long nh = *nholes;
    if (k >= nh) return;
    copy_particle(P, holes[k], P, fillers[nh - 1 - k]);
}

__global__ void save_escaped_kernel(PtlSoA E, long long ebase_cap, const long long* ebase,
                                    PtlSoA P, const long long* idx, const long long* nesc)
{
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= *nesc) return;
    long long d = *ebase + k;
    if (d < ebase_cap) copy_particle(E, d, P, idx[k]);
}

// counters layout (device, long long): [0] nptl_current [1] nptl_escaped [2] scratch alive m
// [3] nholes [4] nfillers [5] nesc_this_pass [6] nsplit_this_pass
__global__ void after_remove_kernel(long long* c)
{
    c[1] += c[5];
    c[0] = c[2];
}

void launch_remove(const PtlSoA& P, const PtlSoA& E, long long ecap, long long n, long long* counters,
                   const ScanWork& w, long long* idx_a, long long* idx_b, int dump_escaped,
                   cudaStream_t st)
{
    if (n <= 0) return;
    PredArgs a{P.count_flag, P.split_times, P.p, counters + 2, 0.0, 0.0, 0.0};
    // escaped particles first (they are overwritten by the hole filling)
    select_indices<PRED_ESCAPED>(a, n, w, dump_escaped ? idx_a : nullptr, counters + 5, st);
    if (dump_escaped)
        save_escaped_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(E, ecap, counters + 1, P, idx_a,
                                                                         counters + 5);
    select_indices<PRED_ALIVE>(a, n, w, nullptr, counters + 2, st);
    select_indices<PRED_HOLE>(a, n, w, idx_a, counters + 3, st);
    select_indices<PRED_FILLER>(a, n, w, idx_b, counters + 4, st);
    fill_holes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, idx_a, idx_b, counters + 3);
    after_remove_kernel<<<1, 1, 0, st>>>(counters);
}

// final BC pass with the un-extended box, particle_module.f90:1959-1970
__global__ void final_bc_kernel(const __grid_constant__ DevParams prm, PtlSoA P, const long long* n,
                                double* leak)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= *n) return;
    int flag = P.count_flag[i];
    double x = P.x[i], y = P.y[i], z = P.z[i], w = P.weight[i];
    if (P.p[i] < 0.0 && flag != GPAT_COUNT_FLAG_INBOX) {
        flag = GPAT_COUNT_FLAG_OTHERS;
        atomicAdd(leak + 1, w);
    } else {
        if (x < prm.xmin && flag == GPAT_COUNT_FLAG_INBOX) {
            if (prm.pbc[0]) { atomicAdd(leak, w); flag = GPAT_COUNT_FLAG_ESCAPE_LX; }
            else x = x - prm.xmin + prm.xmax;
        } else if (x > prm.xmax && flag == GPAT_COUNT_FLAG_INBOX) {
            if (prm.pbc[0]) { atomicAdd(leak, w); flag = GPAT_COUNT_FLAG_ESCAPE_HX; }
            else x = x - prm.xmax + prm.xmin;
        }
        if (prm.ndim == 1) {
            // particle_module.f90:2036: y is not tested in 1-D
        } else if (y < prm.ymin && flag == GPAT_COUNT_FLAG_INBOX) {
            if (prm.pbc[1]) { atomicAdd(leak, w); flag = GPAT_COUNT_FLAG_ESCAPE_LY; }
            else y = y - prm.ymin + prm.ymax;
        } else if (y > prm.ymax && flag == GPAT_COUNT_FLAG_INBOX) {
            if (prm.pbc[1]) { atomicAdd(leak, w); flag = GPAT_COUNT_FLAG_ESCAPE_HY; }
            else y = y - prm.ymax + prm.ymin;
        }
        if (prm.ndim == 3 || prm.include_3rd_dim) {
            if (z < prm.zmin && flag == GPAT_COUNT_FLAG_INBOX) {
                if (prm.pbc[2]) { atomicAdd(leak, w); flag = GPAT_COUNT_FLAG_ESCAPE_LZ; }
                else z = z - prm.zmin + prm.zmax;
            } else if (z > prm.zmax && flag == GPAT_COUNT_FLAG_INBOX) {
                if (prm.pbc[2]) { atomicAdd(leak, w); flag = GPAT_COUNT_FLAG_ESCAPE_HZ; }
                else z = z - prm.zmax + prm.zmin;
            }
        }
    }
    P.x[i] = x; P.y[i] = y; P.z[i] = z;
    P.count_flag[i] = (signed char)flag;
}

void launch_final_bc(const DevParams& prm, const PtlSoA& P, long long nmax, const long long* n_dev,
                     double* leak, cudaStream_t st)
{
    if (nmax <= 0) return;
    final_bc_kernel<<<(unsigned)((nmax + 255) / 256), 256, 0, st>>>(prm, P, n_dev, leak);
}

// ---- split_particle ----------------------------------------------------------------------
__global__ void split_apply_kernel(PtlSoA P, const long long* idx, long long* counters, long long n,
                                   long long nptl_max, TrackDev trk)
{
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = counters[6];
    long long room = nptl_max - n;  // the serial loop stops at the first child that does not fit
    long long nsplit = total < room ? total : room;
    if (k >= nsplit) return;
    long long i = idx[k];
    long long child = n + k;
    int st = P.split_times[i];
    // particle_module.f90:5449: 0.5**(1.0 + split_times) in default real -- exact power of two
    double wgt = ldexp(1.0, -(1 + st));
    P.weight[i] = wgt;
    P.split_times[i] = (signed char)(st + 1);
    copy_particle(P, child, P, i);
    const int tag = P.tag_splitted[i];
    if (tag >= 0) {
        P.tag_splitted[child] = tag + (1 << st);  // 2**(split_times_new - 1)
        return;
    }
    // a tracked particle splits (particle_module.f90:5452-5473): whichever branch is still an
    // ancestor of a selected particle keeps its negative tag (and is sampled if it sits on a
    // sampling step); the other one stops being tracked.
    const int origin = P.origin[i], tinj = P.tag_injected[i], ns = st + 1;
    const bool sample = (P.nsteps_pushed[i] == 0);
    long long lo, hi;
    const int ctag = tag - (1 << st);
    P.tag_splitted[child] = ctag;
    if (trk.enabled && trk_selected(trk, origin, tinj, ctag, ns, lo, hi)) {
        if (sample) {
            P.nsteps_tracked[child] = P.nsteps_tracked[child] + 1;
            trk_record(trk, soa_record(P, child), lo, hi);
        }
    } else {
        P.tag_splitted[child] = -ctag;
    }
    if (trk.enabled && trk_selected(trk, origin, tinj, tag, ns, lo, hi)) {
        if (sample) {
            P.nsteps_tracked[i] = P.nsteps_tracked[i] + 1;
            trk_record(trk, soa_record(P, i), lo, hi);
        }
    } else {
        P.tag_splitted[i] = -tag;  // stop tracking
    }
}

__global__ void after_split_kernel(long long* c, long long n, long long nptl_max, long long* nptl_split)
{
    long long total = c[6];
    long long room = nptl_max - n;
    long long nsplit = total < room ? total : room;
    // overflow: nptl_current = nptl_max (particle_module.f90:5444-5447)
    c[0] = (total > room) ? nptl_max : n + nsplit;
    *nptl_split += nsplit;
}

void launch_split(const DevParams& prm, const PtlSoA& P, long long n, long long nptl_max,
                  double split_ratio, double pmin_split, long long* counters, long long* nptl_split,
                  const ScanWork& w, long long* idx_a, cudaStream_t st, const TrackDev* trk)
{
    if (n <= 0) return;
    PredArgs a{P.count_flag, P.split_times, P.p, counters + 2, pmin_split * prm.p0, split_ratio, prm.pmax};
    select_indices<PRED_SPLIT>(a, n, w, idx_a, counters + 6, st);
    split_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, idx_a, counters, n, nptl_max,
                                                                      trk ? *trk : TrackDev{});
    after_split_kernel<<<1, 1, 0, st>>>(counters, n, nptl_max, nptl_split);
}

// ---- AoS <-> SoA ----------------------------------------------------------------------------
__global__ void to_aos_kernel(PtlSoA P, gpat_particle* out, long long n)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    gpat_particle q;
    q.split_times = P.split_times[i]; q.count_flag = P.count_flag[i]; q.pad_[0] = q.pad_[1] = 0;
    q.origin = P.origin[i]; q.nsteps_tracked = P.nsteps_tracked[i];
    q.nsteps_pushed = P.nsteps_pushed[i]; q.tag_injected = P.tag_injected[i];
    q.tag_splitted = P.tag_splitted[i];
    q.x = P.x[i]; q.y = P.y[i]; q.z = P.z[i]; q.p = P.p[i]; q.v = P.v[i]; q.mu = P.mu[i];
    q.weight = P.weight[i]; q.t = P.t[i]; q.dt = P.dt[i];
    q.padding = __longlong_as_double((long long)P.rng[i]);
    out[i] = q;
}

__global__ void from_aos_kernel(PtlSoA P, const gpat_particle* in, long long n)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    gpat_particle q = in[i];
    P.split_times[i] = q.split_times; P.count_flag[i] = q.count_flag;
    P.origin[i] = q.origin; P.nsteps_tracked[i] = q.nsteps_tracked;
    P.nsteps_pushed[i] = q.nsteps_pushed; P.tag_injected[i] = q.tag_injected;
    P.tag_splitted[i] = q.tag_splitted;
    P.x[i] = q.x; P.y[i] = q.y; P.z[i] = q.z; P.p[i] = q.p; P.v[i] = q.v; P.mu[i] = q.mu;
    P.weight[i] = q.weight; P.t[i] = q.t; P.dt[i] = q.dt;
    P.rng[i] = (unsigned long long)__double_as_longlong(q.padding);
}

void launch_to_aos(const PtlSoA& P, gpat_particle* out, long long n, cudaStream_t st)
{
    if (n > 0) to_aos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, out, n);
}
void launch_from_aos(const PtlSoA& P, const gpat_particle* in, long long n, cudaStream_t st)
{
    if (n > 0) from_aos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, in, n);
}

}  // namespace gpat
