// diag.cu -- diagnostics: energy spectrum, spatial distributions, counters.
//
// One pass over the particles replaces calc_particle_distributions (diagnostics.f90:738-879),
// the particle loop of quick_check (diagnostics.f90:138-142) and get_pmax_global
// (diagnostics.f90:1701-1705).  The global p(-mu) spectrum is accumulated per block in shared
// memory with warp-aggregated atomics (lanes that hit the same bin are merged with
// __match_any_sync before one shared atomic); the large local (mu,p,x,y,z) histograms go
// straight to global memory with no-return FP64 reductions (REDG), their bins being spread
// over space.  Weights are dyadic (0.5**k, particle_module.f90:5449) so FP64 sums are
// exact and independent of the accumulation order: counts are bit-exact vs the reference.
#include "gpat_internal.cuh"

namespace gpat {

__device__ __forceinline__ bool ifloor_ok(double v, long long& out)
{
    if (!(v > -2.0e9 && v < 2.0e9)) return false;  // NaN/Inf never index a bin
    out = (long long)floor(v);
    return true;
}

// momentum bin index f(p) in [-1, K] from the host-libm thresholds T[0..K]; lp = device log10(p)
// only provides the starting guess.
__device__ __forceinline__ int pbin(const double* __restrict__ T, int K, double p, double lp,
                                    double pmin_log, double dp_log)
{
    double v = (lp - pmin_log) / dp_log;
    int c = !(v == v) ? -1 : (v < -1.0) ? -1 : (v > (double)K) ? K : (int)floor(v);
    while (c < K && p >= T[c + 1]) ++c;
    while (c >= 0 && p < T[c]) --c;
    return c;
}

constexpr int kMaxSharedBins = 4096;

__global__ void __launch_bounds__(256) diag_kernel(const PtlSoA P, const __grid_constant__ DiagArgs a)
{
    __shared__ double sh[kMaxSharedBins];
    __shared__ double red_w[8], red_dt[8];
    __shared__ unsigned long long red_mn[8], red_mx[8], red_pm[8];
    const int nglob = a.nmu_g * a.npp_g;
    const bool use_sh = (a.fglobal != nullptr) && (nglob <= kMaxSharedBins);
    if (use_sh)
        for (int b = threadIdx.x; b < nglob; b += blockDim.x) sh[b] = 0.0;
    __syncthreads();

    double sw = 0.0, sdt = 0.0;
    unsigned long long mn = 0x7ff0000000000000ull, mx = 0ull, pm = 0ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // all lanes of a warp run the same number of iterations (match_any needs them converged)
    const long long nround = (a.n + stride - 1) / stride;
    for (long long r = 0; r < nround; ++r) {
        const long long i = r * stride + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        const bool live = i < a.n;
        double x = 0, y = 0, z = 0, p = 0, mu = 0, w = 0, dt = 0;
        if (live) {
            x = P.x[i]; y = P.y[i]; z = P.z[i]; p = P.p[i]; mu = P.mu[i]; w = P.weight[i];
            dt = P.dt[i];
            sw += w;
            sdt += dt;
            unsigned long long db = (unsigned long long)__double_as_longlong(dt);
            unsigned long long pb = (unsigned long long)__double_as_longlong(p);
            // dt and p are non-negative: IEEE order == unsigned integer order
            if (dt < 1.0 && db < mn) mn = db;  // pdt_min starts at 1.0 (diagnostics.f90:131)
            if (db > mx && dt > 0.0) mx = db;
            if (pb > pm && p > 0.0) pm = pb;
        }
        const double lp = live ? log10(p) : 0.0;
        // global spectrum (diagnostics.f90:773-780)
        int gbin = -1;
        if (live && a.fglobal && p > a.pmin && p <= a.pmax && mu >= -1.0 && mu <= 1.0) {
            long long ip = pbin(a.gthr, a.npp_g, p, lp, a.pmin_log, a.dp_log), imu;
            if (ifloor_ok((mu + 1.0) / a.dmu, imu)) {
                ip += 1; imu += 1;
                long long lin = (imu - 1) + (ip - 1) * (long long)a.nmu_g;
                if (ip >= 1 && imu >= 1 && lin >= 0 && lin < nglob) gbin = (int)lin;
            }
        }
        // warp aggregation: one atomic per distinct bin in the warp
        {
            unsigned peers = __match_any_sync(0xffffffffu, gbin);
            double tot = w;
            // sum the weights of the peers (leader = lowest lane of the group)
            const unsigned lane = threadIdx.x & 31u;
            const int leader = __ffs(peers) - 1;
            double acc = 0.0;
            for (unsigned mm = peers; mm; mm &= mm - 1) {
                int src = __ffs(mm) - 1;
                double v = __shfl_sync(peers, tot, src);
                acc += v;
            }
            if (gbin >= 0 && (int)lane == leader) {
                if (use_sh) atomicAdd(&sh[gbin], acc);
                else atomicAdd(&a.fglobal[gbin], acc);
            }
        }
        // local distributions (diagnostics.f90:782-870).  Same warp aggregation as the global spectrum: the lanes of
        // a warp that fall into the same (mu, p, x, y, z) bin are merged with __match_any_sync and their leader
        // issues ONE no-return FP64 reduction (REDG.F64).  The particle arrays are cell-ordered in the production
        // build, so neighbouring lanes share bins (at C1's r = 4 set a warp of 32 particles lands in a handful of
        // bins); the 5-D arrays themselves are MBs (C1: 6.3 + 4.2 + 0.5 MB) and cannot be staged in shared memory.
        if (a.local_dist) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const HistDev& h = a.loc[k];
                if (!h.enabled || !h.data) continue;   // uniform: kernel argument
                long long lin = -1;
                if (live) {
                    long long ix, iy, iz, ip, imu;
                    const bool ok = ifloor_ok((x - a.xmin) / h.dx_diag, ix) &&
                                    ifloor_ok((y - a.ymin) / h.dy_diag, iy) &&
                                    ifloor_ok((z - a.zmin) / h.dz_diag, iz) &&
                                    ifloor_ok((mu + 1.0) / h.dmu, imu);
                    if (ok) {
                        ip = pbin(h.pthr, h.npbins, p, lp, h.pmin_log, h.dp_log);
                        ix += 1; iy += 1; iz += 1; ip += 1; imu += 1;
                        if (ix >= 1 && ix <= h.nrx && iy >= 1 && iy <= h.nry && iz >= 1 && iz <= h.nrz &&
                            ip > 0 && ip < h.npbins /* top bin never filled, diagnostics.f90:799 */ &&
                            imu >= 1 && imu <= h.nmu)
                            lin = (long long)((size_t)(imu - 1) + (size_t)h.nmu * ((size_t)(ip - 1) +
                                  (size_t)h.npbins * ((size_t)(ix - 1) + (size_t)h.nrx *
                                  ((size_t)(iy - 1) + (size_t)h.nry * (size_t)(iz - 1)))));
                    }
                }
                const unsigned peers = __match_any_sync(0xffffffffu, lin);
                const int leader = __ffs(peers) - 1;
                double acc = 0.0;
                for (unsigned mm = peers; mm; mm &= mm - 1)   // dyadic weights: the sum is exact in any order
                    acc += __shfl_sync(peers, w, __ffs(mm) - 1);
                if (lin >= 0 && (int)(threadIdx.x & 31u) == leader) atomicAdd(&h.data[lin], acc);
            }
        }
    }
    // block reduction of the counters
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sw += __shfl_down_sync(0xffffffffu, sw, o);
        sdt += __shfl_down_sync(0xffffffffu, sdt, o);
        unsigned long long t;
        t = __shfl_down_sync(0xffffffffu, mn, o); if (t < mn) mn = t;
        t = __shfl_down_sync(0xffffffffu, mx, o); if (t > mx) mx = t;
        t = __shfl_down_sync(0xffffffffu, pm, o); if (t > pm) pm = t;
    }
    if (lane == 0) { red_w[wid] = sw; red_dt[wid] = sdt; red_mn[wid] = mn; red_mx[wid] = mx; red_pm[wid] = pm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            sw += red_w[k]; sdt += red_dt[k];
            if (red_mn[k] < mn) mn = red_mn[k];
            if (red_mx[k] > mx) mx = red_mx[k];
            if (red_pm[k] > pm) pm = red_pm[k];
        }
        atomicAdd(&a.sums[0], sw);
        atomicAdd(&a.sums[1], sdt);
        atomicMin(&a.minmax[0], mn);
        atomicMax(&a.minmax[1], mx);
        atomicMax(&a.minmax[2], pm);
    }
    if (use_sh) {
        for (int b = threadIdx.x; b < nglob; b += blockDim.x)
            if (sh[b] != 0.0) atomicAdd(&a.fglobal[b], sh[b]);
    }
}

__global__ void finalize_quick_kernel(const double* sums, const unsigned long long* minmax,
                                      const double* leak, double nptl_current, double nptl_split,
                                      double* q)
{
    // var_local(1:6) of quick_check (diagnostics.f90:134-142), then pdt_min, pdt_max, pmax
    q[0] = nptl_current; q[1] = nptl_split; q[2] = sums[0]; q[3] = leak[0]; q[4] = leak[1];
    q[5] = sums[1];
    q[6] = (minmax[0] == 0x7ff0000000000000ull) ? 1.0 : __longlong_as_double((long long)minmax[0]);
    q[7] = __longlong_as_double((long long)minmax[1]);
    q[8] = __longlong_as_double((long long)minmax[2]);
}

void launch_finalize_quick(const double* sums, const unsigned long long* minmax, const double* leak,
                           double nptl_current, double nptl_split, double* q9, cudaStream_t st)
{
    finalize_quick_kernel<<<1, 1, 0, st>>>(sums, minmax, leak, nptl_current, nptl_split, q9);
}

void launch_diag(const PtlSoA& P, const DiagArgs& a, int sm_count, cudaStream_t st)
{
    long long blocks = (a.n + 255) / 256;
    long long cap = (long long)sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    diag_kernel<<<(unsigned)blocks, 256, 0, st>>>(P, a);
}

// escaped particles: fescaped(nmu, npp, 2*ndim), diagnostics.f90:913-1000
__global__ void escaped_diag_kernel(const PtlSoA E, long long n, DiagArgs a, int nface, double* fesc)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p = E.p[i], mu = E.mu[i];
    int face = -(int)E.count_flag[i];
    if (face < 1 || face > nface) return;
    if (p > a.pmin && p <= a.pmax && mu >= -1.0 && mu <= 1.0) {
        long long ip = pbin(a.gthr, a.npp_g, p, log10(p), a.pmin_log, a.dp_log), imu;
        if (ifloor_ok((mu + 1.0) / a.dmu, imu)) {
            ip += 1; imu += 1;
            long long lin = (imu - 1) + (ip - 1) * (long long)a.nmu_g;
            long long nglob = (long long)a.nmu_g * a.npp_g;
            if (ip >= 1 && imu >= 1 && lin >= 0 && lin < nglob)
                atomicAdd(&fesc[lin + (size_t)(face - 1) * nglob], E.weight[i]);
        }
    }
}

void launch_escaped_diag(const PtlSoA& E, long long n, const DiagArgs& a, int nface, double* fesc,
                         cudaStream_t st)
{
    if (n > 0) escaped_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(E, n, a, nface, fesc);
}

// local part of calc_escaped_distributions (diagnostics.f90:956-1170): every escaped particle is binned
// on the face it left through, in the two coordinates of that face, with the local sets' own p and mu bins
__global__ void escaped_local_kernel(const PtlSoA E, long long n, DiagArgs a, EscLocalDev o)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int face = -(int)E.count_flag[i];  // 1 lx, 2 hx, 3 ly, 4 hy, 5 lz, 6 hz
    if (face < 1 || face > 6) return;
    const double x = E.x[i], y = E.y[i], z = E.z[i], p = E.p[i], mu = E.mu[i], w = E.weight[i];
    const double lp = log10(p);
    const size_t side = (size_t)((face - 1) & 1);
    for (int k = 0; k < 4; ++k) {
        const HistDev& h = a.loc[k];
        if (!h.enabled) continue;
        long long ix = 0, iy = 0, iz = 0, imu = 0;
        const bool okx = ifloor_ok((x - a.xmin) / h.dx_diag, ix), oky = ifloor_ok((y - a.ymin) / h.dy_diag, iy);
        const bool okz = ifloor_ok((z - a.zmin) / h.dz_diag, iz), okm = ifloor_ok((mu + 1.0) / h.dmu, imu);
        const long long ip = (long long)pbin(h.pthr, h.npbins, p, lp, h.pmin_log, h.dp_log) + 1;
        ix += 1; iy += 1; iz += 1; imu += 1;
        const bool condx = okx && ix >= 1 && ix <= h.nrx, condy = oky && iy >= 1 && iy <= h.nry;
        const bool condz = okz && iz >= 1 && iz <= h.nrz;
        const bool condp = (p == p) && ip > 0 && ip < h.npbins, condmu = okm && imu >= 1 && imu <= h.nmu;
        if (!(condp && condmu)) continue;
        const size_t nb = (size_t)h.nmu * h.npbins;
        const size_t b = (size_t)(imu - 1) + (size_t)h.nmu * (size_t)(ip - 1);
        if (face <= 2) {
            if (condy && condz && o.fx[k])
                atomicAdd(&o.fx[k][b + nb * ((size_t)(iy - 1) + (size_t)h.nry * ((size_t)(iz - 1) + (size_t)h.nrz * side))], w);
        } else if (face <= 4) {
            if (condx && condz && o.fy[k])
                atomicAdd(&o.fy[k][b + nb * ((size_t)(ix - 1) + (size_t)h.nrx * ((size_t)(iz - 1) + (size_t)h.nrz * side))], w);
        } else {
            if (condx && condy && o.fz[k])
                atomicAdd(&o.fz[k][b + nb * ((size_t)(ix - 1) + (size_t)h.nrx * ((size_t)(iy - 1) + (size_t)h.nry * side))], w);
        }
    }
}

void launch_escaped_local(const PtlSoA& E, long long n, const DiagArgs& a, const EscLocalDev& o, cudaStream_t st)
{
    if (n > 0) escaped_local_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(E, n, a, o);
}

}  // namespace gpat
