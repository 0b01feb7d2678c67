// push.cu -- the pseudo-particle SDE push (the hot path).
//
// One launch moves every particle through one MHD interval: it replaces the particle loop
// of particle_mover_one_cycle (particle_module.f90:1560-1831) with get_interp_paramters
// (particle_module.f90:642-674), interp_fields (mhd_data_parallel.f90:1751-1793),
// calc_spatial_diffusion_coefficients[_nlgc] (particle_module.f90:2208-2771),
// push_particle_2d / _2d_include_3rd / _3d (particle_module.f90:3358-3606, 3979-4245,
// 4625-4907), particle_boundary_condition (particle_module.f90:1984-2129) and unif_01
// (random_number_generator.f90:97-101 -> Philox4x32-10 per particle) fused.
//
// Execution model: persistent grid, one lane per particle, lanes refill from a global
// work counter with a warp-aggregated atomic when their particle reaches the end of the
// interval.  The reference's nested while loops are flattened into a per-lane state
// machine so that every loop iteration of a warp executes the expensive body (gather +
// kappa + step) exactly once, converged, whatever the per-particle step counts are.
//
// This translation unit is compiled twice:
//   GPAT_STRICT=1, -fmad=false : reference operation order, no contraction (parity build)
//   GPAT_STRICT=0              : FMA contraction, time-blend folded into the weights
#include <cstdlib>
#include <type_traits>

#include "gpat_internal.cuh"
#include "fastmath.cuh"

#ifndef GPAT_STRICT
#define GPAT_STRICT 0
#endif
// The production build is three translation units of this file (compile time): GPAT_TU_PART 0 = the kernels of the named
// configs + every generic instantiation, 1 / 2 = the 2-D / 3-D kSpecAlt instantiations (1-D, focused transport, maps).
#ifndef GPAT_TU_PART
#define GPAT_TU_PART 0
#endif

namespace gpat {
namespace {

constexpr double kEps = 2.220446049250313e-16;  // EPSILON(1.0d0)
constexpr int kBlock = 128;

__device__ __forceinline__ double sq(double v) { return v * v; }
__device__ __forceinline__ double min2(double a, double b) { return (b < a) ? b : a; }

// Elementary operations of the general pushers (push_physics and what it calls).  Reference-order build: the plain
// IEEE operations in the reference's order.  Production build: the straight-line versions of fastmath.cuh (<= 2 ulp each;
// pow as exp(y log x)), which is what turns the ~2000 FP64-pipe instructions of a focused-transport step's libdevice
// pow / sqrt / division sequences into ~500.  Arguments are what the pushers guarantee: normal, positive where a
// root or a logarithm is taken; pow(0, y) keeps its IEEE value.
constexpr double kThird = 1.0 / 3.0, kTwoThirds = 2.0 / 3.0;
// PowBase: a base of several powers.  Production build: its logarithm is taken once (a focused-transport step raises
// b, p/p0, |mu| and p to ten exponents; as separate pow calls that was ten logarithms, the compiler cannot merge them
// across the run-time switches).  Reference-order build: just the value, every power is its own pow().
struct PowBase { double x, lx; };
#if GPAT_STRICT
__device__ __forceinline__ double pm_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ PowBase pm_base(double x) { return PowBase{x, 0.0}; }
__device__ __forceinline__ double pm_powb(const PowBase& b, double y) { return pow(b.x, y); }
__device__ __forceinline__ double pm_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ double pm_div(double a, double b) { return a / b; }
__device__ __forceinline__ double pm_divc(double a, double c) { return a / c; }
#else
__device__ __forceinline__ double pm_sqrt(double x) { return fm::sqrt_pos(x); }
__device__ __forceinline__ PowBase pm_base(double x) { return PowBase{x, fm::log_pos(x > 0.0 ? x : 1.0)}; }
__device__ __forceinline__ double pm_powb(const PowBase& b, double y)
{
    const double r = fm::exp_mid(y * b.lx);
    return (b.x > 0.0) ? r : ((y > 0.0) ? 0.0 : (y == 0.0 ? 1.0 : __longlong_as_double(0x7ff0000000000000LL)));
}
__device__ __forceinline__ double pm_pow(double x, double y) { return pm_powb(pm_base(x), y); }
__device__ __forceinline__ double pm_div(double a, double b) { return a * fm::rcp(b); }
__device__ __forceinline__ double pm_divc(double a, double c) { return a * (1.0 / c); }  // c: a literal
#endif

// ---- Philox4x32-10 (Salmon et al. 2011) -------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, unsigned k0, unsigned k1)
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

__device__ __forceinline__ double u01(unsigned w)
{
#if GPAT_STRICT
    return (double)w / 4294967295.0;
#else
    return (double)w * (1.0 / 4294967295.0);
#endif
}

// ---- per-lane particle state ---------------------------------------------------------
struct Lane {
    double x, y, z, p, t, dt, weight, mu;
    double dxl, dyl, dzl, dpl;  // last step's deltas (roll-back, particle_module.f90:1708-1714)
    double dt_target, dt_old;
    unsigned long long rng;
    int tag_inj, tag_spl, origin;
    int nsteps_pushed;
    int count_flag;
    int nsteps_tracked;  // used by the tracking instantiations only
    // Touched only by the reference-order build and by the ALT instantiations of the production build (1-D, focused
    // transport, turbulence maps); every other instantiation never reads or writes them, so they cost no registers there.
    double v, dvl, dmul;  // focused transport: particle speed, last step's delta v / delta mu
    double sh1, sh2;      // acc_by_surface: surface heights at this step's starting position
};

constexpr bool kStrict = (GPAT_STRICT != 0);
// SPEC value of the production-build instantiations that run the general pushers (push_physics: 1-D, focused transport,
// turbulence maps, every switch at run time) behind the lane-group gather
constexpr int kSpecAlt = 16;
// ... and the same with the lane-group gather of the turbulence-map record compiled in (deltab_flag / correlation_flag):
// a separate instantiation, because carrying the map code costs runs without maps 10 % (profiles/README.md, calls Y, Z)
constexpr int kSpecAltMaps = 16 | 32;
__host__ __device__ constexpr bool spec_is_alt(int spec) { return (spec & 16) != 0; }
__host__ __device__ constexpr bool spec_has_maps(int spec) { return (spec & 32) != 0; }
// ... and one instantiation per pusher (bits 6-7 of SPEC), so that a focused-transport kernel does not carry push_1d and
// the Parker pushers and vice versa (the kernel that held all of them was 12 900 instructions, "no instruction" 0.4-0.7
// stalls per issue, profiles/r02n2_push_coop_*_ncu.txt).  PATH_ANY reads ndim / focused_transport at run time: the
// reference-order build.
enum : int { PATH_ANY = 0, PATH_1D = 1, PATH_FT = 2, PATH_PARKER = 3 };
__host__ __device__ constexpr int spec_path(int spec) { return (spec >> 6) & 3; }
__host__ __device__ constexpr int alt_spec(int path, bool maps) { return 16 | (maps ? 32 : 0) | (path << 6); }

// particle_boundary_condition for a single rank (neighbours are self or -1)
// ALT: the run may be 1-D (reference-order build: always checked)
template <bool ALT = false>
__device__ __forceinline__ void boundary(const DevParams& prm, Lane& q, const double* e,
                                         double* leak)
{
    if (q.x < e[0] && q.count_flag == GPAT_COUNT_FLAG_INBOX) {
        if (prm.pbc[0]) { atomicAdd(leak, q.weight); q.count_flag = GPAT_COUNT_FLAG_ESCAPE_LX; }
        else q.x = q.x - e[0] + e[1];
    } else if (q.x > e[1] && q.count_flag == GPAT_COUNT_FLAG_INBOX) {
        if (prm.pbc[0]) { atomicAdd(leak, q.weight); q.count_flag = GPAT_COUNT_FLAG_ESCAPE_HX; }
        else q.x = q.x - e[1] + e[0];
    }
    if ((kStrict || ALT) && prm.ndim == 1) {  // particle_module.f90:2036
    } else if (q.y < e[2] && q.count_flag == GPAT_COUNT_FLAG_INBOX) {
        if (prm.pbc[1]) { atomicAdd(leak, q.weight); q.count_flag = GPAT_COUNT_FLAG_ESCAPE_LY; }
        else q.y = q.y - e[2] + e[3];
    } else if (q.y > e[3] && q.count_flag == GPAT_COUNT_FLAG_INBOX) {
        if (prm.pbc[1]) { atomicAdd(leak, q.weight); q.count_flag = GPAT_COUNT_FLAG_ESCAPE_HY; }
        else q.y = q.y - e[3] + e[2];
    }
    if (prm.ndim == 3 || prm.include_3rd_dim) {
        if (q.z < e[4] && q.count_flag == GPAT_COUNT_FLAG_INBOX) {
            if (prm.pbc[2]) { atomicAdd(leak, q.weight); q.count_flag = GPAT_COUNT_FLAG_ESCAPE_LZ; }
            else q.z = q.z - e[4] + e[5];
        } else if (q.z > e[5] && q.count_flag == GPAT_COUNT_FLAG_INBOX) {
            if (prm.pbc[2]) { atomicAdd(leak, q.weight); q.count_flag = GPAT_COUNT_FLAG_ESCAPE_HZ; }
            else q.z = q.z - e[5] + e[4];
        }
    }
}

// true when negp_or_bc() would change anything
// MAYZ = false: the layout rules out a resolved z axis (base 2-D record: ndim = 2 and no
// include_3rd_dim, which needs the extended record), so the z test is not even compiled
template <bool MAYZ = true, bool ALT = false>
__device__ __forceinline__ bool outside_or_negp(const DevParams& prm, const Lane& q)
{
    bool o = (q.p < 0.0) | (q.x < prm.ext[0]) | (q.x > prm.ext[1]) | (q.y < prm.ext[2]) | (q.y > prm.ext[3]);
    if ((kStrict || ALT) && prm.ndim == 1) o = (q.p < 0.0) | (q.x < prm.ext[0]) | (q.x > prm.ext[1]);
    if (MAYZ && (prm.ndim == 3 || prm.include_3rd_dim)) o = o | (q.z < prm.ext[4]) | (q.z > prm.ext[5]);
    return o;
}

// particle_module.f90:1602-1609
template <bool ALT = false>
__device__ __forceinline__ void negp_or_bc(const DevParams& prm, Lane& q, double* leak)
{
    if (q.p < 0.0) {
        q.count_flag = GPAT_COUNT_FLAG_OTHERS;
        atomicAdd(leak + 1, q.weight);
    } else {
        boundary<ALT>(prm, q, prm.ext, leak);
    }
}

// ---- gather: get_interp_paramters + interp_fields -------------------------------------
// 256-bit read-only load (LDG.E.256 on sm_100a): one 32-byte chunk = four slots of both frames
__device__ __forceinline__ void ldg256(const float* __restrict__ p, float4& lo, float4& hi)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z),
                   "=f"(hi.w)
                 : "l"(p));
}

// cell index + in-cell offsets of a position (get_interp_paramters, particle_module.f90:642-674);
// the cell is clamped so that a runaway particle cannot fault.
template <int NDIM, bool ALT = false>
__device__ __forceinline__ long long locate(const DevParams& prm, double x, double y, double z,
                                            double& rx, double& ry, double& rz)
{
#if GPAT_STRICT
    double px = (x - prm.xmin) / prm.dx;
    double py = (y - prm.ymin) / prm.dy;
#else
    double px = (x - prm.xmin) * prm.idx;
    double py = (y - prm.ymin) * prm.idy;
#endif
    int ix = (int)floor(px) + 1;  // Fortran pos(1)
    int iy = (int)floor(py) + 1;
    rx = px - (double)ix + 1.0;
    ry = py - (double)iy + 1.0;
    // Fortran index -> storage index is +1 (lower bound -1, mhd_data_parallel.f90:82)
    int cx = min(max(ix + 1, 0), prm.nxg - 2);
    int cy = min(max(iy + 1, 0), prm.nyg - 2);
    if ((kStrict || ALT) && prm.ndim == 1) { cy = 0; ry = 0.0; }  // pos(2) = 1, ry = 0 (particle_module.f90:649-652)
    long long cell = (long long)cy * prm.nxg + cx;
    rz = 0.0;
    if (NDIM == 3) {
#if GPAT_STRICT
        double pz = (z - prm.zmin) / prm.dz;
#else
        double pz = (z - prm.zmin) * prm.idz;
#endif
        int iz = (int)floor(pz) + 1;
        rz = pz - (double)iz + 1.0;
        int cz = min(max(iz + 1, 0), prm.nzg - 2);
        cell += (long long)cz * prm.nxg * prm.nyg;
    }
    return cell;
}

template <int L>
__device__ __forceinline__ void gather(const DevParams& prm, const float* __restrict__ fld,
                                       int sel, double x, double y, double z, double rt,
                                       double (&F)[Rec<L>::NF])
{
    constexpr int NREC = Rec<L>::NREC;
    constexpr int NQ = (Rec<L>::NUSED + 3) / 4;
    constexpr int NC = (Rec<L>::NDIM == 3) ? 8 : 4;
    constexpr long long stride = 2LL * NREC;  // floats per grid point (both halves)

    double rx, ry, rz;
    const long long cell = locate<Rec<L>::NDIM>(prm, x, y, z, rx, ry, rz);
    const double rx1 = 1.0 - rx, ry1 = 1.0 - ry;
    double w[NC];
    long long off[NC];
    if (NC == 4) {
        w[0] = rx1 * ry1; w[1] = rx * ry1; w[2] = rx1 * ry; w[3] = rx * ry;
    } else {
        const double rz1 = 1.0 - rz;
        w[0] = rx1 * ry1 * rz1; w[1] = rx * ry1 * rz1; w[2] = rx1 * ry * rz1; w[3] = rx * ry * rz1;
        w[4] = rx1 * ry1 * rz;  w[5] = rx * ry1 * rz;  w[6] = rx1 * ry * rz;  w[7] = rx * ry * rz;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
        off[c] = (cell + (c & 1) + (long long)((c >> 1) & 1) * prm.nxg +
                  (long long)(c >> 2) * prm.nxg * prm.nyg) * stride;

#if GPAT_STRICT
    // reference order: per frame, sum over corners starting from 0, then blend
    const int hA = (prm.time_interp ? sel : 0) * 4;
    const int hB = (sel ^ 1) * 4;
    const double rt1 = 1.0 - rt;
#pragma unroll
    for (int qd = 0; qd < NQ; ++qd) {
        double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float4 fa = __ldg(reinterpret_cast<const float4*>(fld + off[c] + 8 * qd + hA));
            a[0] = a[0] + (double)fa.x * w[c]; a[1] = a[1] + (double)fa.y * w[c];
            a[2] = a[2] + (double)fa.z * w[c]; a[3] = a[3] + (double)fa.w * w[c];
            if (prm.time_interp) {
                float4 fb = __ldg(reinterpret_cast<const float4*>(fld + off[c] + 8 * qd + hB));
                b[0] = b[0] + (double)fb.x * w[c]; b[1] = b[1] + (double)fb.y * w[c];
                b[2] = b[2] + (double)fb.z * w[c]; b[3] = b[3] + (double)fb.w * w[c];
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
            F[4 * qd + e] = prm.time_interp ? (a[e] * rt1 + b[e] * rt) : a[e];
    }
    if constexpr (Rec<L>::SIDE_CHUNKS > 0) {  // side plane in chunk format (L3D): same sums over the side chunks
        const float* side = fld + (long long)prm.nxg * prm.nyg * prm.nzg * stride;
        constexpr int SS = 8 * Rec<L>::SIDE_CHUNKS;
#pragma unroll
        for (int qd = 0; qd < Rec<L>::SIDE_CHUNKS; ++qd) {
            double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const float* pc_ = side + (off[c] / stride) * SS + 8 * qd;
                float4 fa = __ldg(reinterpret_cast<const float4*>(pc_ + hA));
                a[0] = a[0] + (double)fa.x * w[c]; a[1] = a[1] + (double)fa.y * w[c];
                a[2] = a[2] + (double)fa.z * w[c]; a[3] = a[3] + (double)fa.w * w[c];
                if (prm.time_interp) {
                    float4 fb = __ldg(reinterpret_cast<const float4*>(pc_ + hB));
                    b[0] = b[0] + (double)fb.x * w[c]; b[1] = b[1] + (double)fb.y * w[c];
                    b[2] = b[2] + (double)fb.z * w[c]; b[3] = b[3] + (double)fb.w * w[c];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
                F[NREC + 4 * qd + e] = prm.time_interp ? (a[e] * rt1 + b[e] * rt) : a[e];
        }
    } else if constexpr (Rec<L>::NSIDE > 0) {  // side plane (L2D): [s0 s1 of half 0 | s0 s1 of half 1] per grid point
        const float* side = fld + (long long)prm.nxg * prm.nyg * prm.nzg * stride;
        const int sA = prm.time_interp ? sel : 0, sB = sel ^ 1;
        double a[2] = {0.0, 0.0}, b[2] = {0.0, 0.0};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(side + (off[c] / stride) * 4));
            const float ea[2] = {sA ? e.z : e.x, sA ? e.w : e.y}, eb[2] = {sB ? e.z : e.x, sB ? e.w : e.y};
            a[0] = a[0] + (double)ea[0] * w[c]; a[1] = a[1] + (double)ea[1] * w[c];
            if (prm.time_interp) { b[0] = b[0] + (double)eb[0] * w[c]; b[1] = b[1] + (double)eb[1] * w[c]; }
        }
        F[NREC] = prm.time_interp ? (a[0] * rt1 + b[0] * rt) : a[0];
        F[NREC + 1] = prm.time_interp ? (a[1] * rt1 + b[1] * rt) : a[1];
        F[NREC + 2] = F[NREC + 3] = 0.0;
    }
#else
    // fast: fold the time blend into the corner weights, one FMA per loaded value; w0/w1 are the
    // weights of half 0 / half 1 of each chunk
    double w0[NC], w1[NC];
    {
        const double rt1 = 1.0 - rt;
        const double tA = prm.time_interp ? rt1 : 1.0, tB = prm.time_interp ? rt : 0.0;
        // w0 / w1: weights of farray1 / farray2 (the loads below put them in f0 / f1)
#pragma unroll
        for (int c = 0; c < NC; ++c) { w0[c] = w[c] * tA; w1[c] = w[c] * tB; }
    }
#pragma unroll
    for (int qd = 0; qd < NQ; ++qd) {
        double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float4 f0, f1;
            if (sel == 0) ldg256(fld + off[c] + 8 * qd, f0, f1);  // f0 = farray1, f1 = farray2:
            else ldg256(fld + off[c] + 8 * qd, f1, f0);           // same summation order for both sel
            a[0] = fma((double)f0.x, w0[c], a[0]); a[1] = fma((double)f0.y, w0[c], a[1]);
            a[2] = fma((double)f0.z, w0[c], a[2]); a[3] = fma((double)f0.w, w0[c], a[3]);
            a[0] = fma((double)f1.x, w1[c], a[0]); a[1] = fma((double)f1.y, w1[c], a[1]);
            a[2] = fma((double)f1.z, w1[c], a[2]); a[3] = fma((double)f1.w, w1[c], a[3]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) F[4 * qd + e] = a[e];
    }
    if constexpr (Rec<L>::SIDE_CHUNKS > 0) {
        const float* side = fld + (long long)prm.nxg * prm.nyg * prm.nzg * stride;
        constexpr int SS = 8 * Rec<L>::SIDE_CHUNKS;
#pragma unroll
        for (int qd = 0; qd < Rec<L>::SIDE_CHUNKS; ++qd) {
            double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float4 f0, f1;
                if (sel == 0) ldg256(side + (off[c] / stride) * SS + 8 * qd, f0, f1);
                else ldg256(side + (off[c] / stride) * SS + 8 * qd, f1, f0);
                a[0] = fma((double)f0.x, w0[c], a[0]); a[1] = fma((double)f0.y, w0[c], a[1]);
                a[2] = fma((double)f0.z, w0[c], a[2]); a[3] = fma((double)f0.w, w0[c], a[3]);
                a[0] = fma((double)f1.x, w1[c], a[0]); a[1] = fma((double)f1.y, w1[c], a[1]);
                a[2] = fma((double)f1.z, w1[c], a[2]); a[3] = fma((double)f1.w, w1[c], a[3]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) F[NREC + 4 * qd + e] = a[e];
        }
    } else if constexpr (Rec<L>::NSIDE > 0) {
        const float* side = fld + (long long)prm.nxg * prm.nyg * prm.nzg * stride;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(side + (off[c] / stride) * 4));
            const float a0f = sel == 0 ? e.x : e.z, a1f = sel == 0 ? e.y : e.w;  // farray1 lives in half `sel`
            const float b0f = sel == 0 ? e.z : e.x, b1f = sel == 0 ? e.w : e.y;  // farray2
            a0 = fma((double)a0f, w0[c], a0); a1 = fma((double)a1f, w0[c], a1);
            a0 = fma((double)b0f, w1[c], a0); a1 = fma((double)b1f, w1[c], a1);
        }
        F[NREC] = a0; F[NREC + 1] = a1; F[NREC + 2] = F[NREC + 3] = 0.0;
    }
#endif
}

// interp_magnetic_fluctuation + interp_correlation_length (mhd_data_parallel.f90:1806-1915): the
// sixteen values db2_slab(1:4) db2_2d(1:4) lc_slab(1:4) lc_2d(1:4), reference summation order
template <int NDIM>
__device__ __forceinline__ void gather_aux(const DevParams& prm, const float* __restrict__ aux, int sel,
                                           double x, double y, double z, double rt, double (&A)[16])
{
    constexpr int NC = (NDIM == 3) ? 8 : 4;
    double rx, ry, rz;
    const long long cell = locate<NDIM, true>(prm, x, y, z, rx, ry, rz);
    const double rx1 = 1.0 - rx, ry1 = 1.0 - ry, rz1 = 1.0 - rz;
    double w[NC];
    if (NC == 4) {
        w[0] = rx1 * ry1; w[1] = rx * ry1; w[2] = rx1 * ry; w[3] = rx * ry;
    } else {
        w[0] = rx1 * ry1 * rz1; w[1] = rx * ry1 * rz1; w[2] = rx1 * ry * rz1; w[3] = rx * ry * rz1;
        w[4] = rx1 * ry1 * rz;  w[5] = rx * ry1 * rz;  w[6] = rx1 * ry * rz;  w[7] = rx * ry * rz;
    }
    const int hA = (prm.time_interp ? sel : 0) * 4, hB = (sel ^ 1) * 4;
    const double rt1 = 1.0 - rt;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const long long off = (cell + (c & 1) + (long long)((c >> 1) & 1) * prm.nxg +
                                   (long long)(c >> 2) * prm.nxg * prm.nyg) * 32 + 8 * q;
            const float4 fa = __ldg(reinterpret_cast<const float4*>(aux + off + hA));
            a[0] = a[0] + (double)fa.x * w[c]; a[1] = a[1] + (double)fa.y * w[c];
            a[2] = a[2] + (double)fa.z * w[c]; a[3] = a[3] + (double)fa.w * w[c];
            if (prm.time_interp) {
                const float4 fb = __ldg(reinterpret_cast<const float4*>(aux + off + hB));
                b[0] = b[0] + (double)fb.x * w[c]; b[1] = b[1] + (double)fb.y * w[c];
                b[2] = b[2] + (double)fb.z * w[c]; b[3] = b[3] + (double)fb.w * w[c];
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) A[4 * q + e] = prm.time_interp ? (a[e] * rt1 + b[e] * rt) : a[e];
    }
}

// ---- kappa (particle_module.f90:93-104) --------------------------------------------
struct Kappa {
    double knorm_para, kpara, kperp, skpara, skperp, skpara_perp;
    PowBase pb_b, pb_p, pb_lc;  // bases b, p/p0, lc_slab of calc_kappa's powers (set under the switches that use them)
    double dkxx_dx, dkyy_dy, dkzz_dz, dkxy_dx, dkxy_dy, dkxz_dx, dkxz_dz, dkyz_dy, dkyz_dz;
};

struct BField {  // B and its gradients at the particle
    double bx, by, bz, b;
    double dbx_dx, dbx_dy, dbx_dz, dby_dx, dby_dy, dby_dz, dbz_dx, dbz_dy, dbz_dz;
    double db_dx, db_dy, db_dz;
};

// THREE: z-components needed (3-D or 2-D with include_3rd_dim); FULL3D: ndim == 3
// aux (reference-order build): the 16 interpolated turbulence values, or nullptr
template <bool THREE, bool FULL3D>
__device__ __forceinline__ void calc_kappa(const DevParams& prm, const BField& B, double p,
                                           double mu, Kappa& k, const double* aux = nullptr)
{
    const double bx = B.bx, by = B.by, bz = B.bz, b = B.b;
    const double ib1 = (b < kEps) ? 1.0 : pm_div(1.0, b);  // particle_module.f90:2230-2234
    const double ib2 = ib1 * ib1;
    const double ib3 = ib1 * ib2;
    double knp = 1.0, knperp = 1.0;
    if (prm.mag_dependency == 1) {
        k.pb_b = pm_base(b);
        knp = knp * pm_powb(k.pb_b, prm.gm2);
        if (prm.nlgc) knperp = knperp * pm_powb(k.pb_b, prm.gm2_3);
    }
    const bool dbf = aux && prm.deltab_flag, cof = aux && prm.correlation_flag;
    if (dbf) {  // particle_module.f90:2246-2248 / 2505-2509
        knp = pm_div(knp, aux[0]);
        if (prm.nlgc) knperp = knperp * pm_pow(aux[0], -kThird) * pm_pow(aux[4], kTwoThirds);
    }
    if (cof) {  // particle_module.f90:2252-2254 / 2513-2517
        k.pb_lc = pm_base(aux[8]);
        knp = knp * pm_powb(k.pb_lc, prm.gamma_turb - 1.0);
        if (prm.nlgc) knperp = knperp * pm_powb(k.pb_lc, pm_divc(prm.gamma_turb - 1.0, 3.0)) * pm_pow(aux[12], kTwoThirds);
    }
    k.knorm_para = knp;
    double ax = 0.0, ay = 0.0, az = 0.0;  // coefficients multiplying the b_i b_j terms
    if (prm.momentum_dependency == 1) k.pb_p = pm_base(pm_div(p, prm.p0));
    if (!prm.nlgc) {
        double knorm = (prm.momentum_dependency == 1) ? knp * pm_powb(k.pb_p, prm.pindex) : knp;
        k.kpara = prm.kpara0 * knorm;
        k.kperp = k.kpara * prm.kret;
    } else {
        double kn_para = knp, kn_perp = knperp;
        if (prm.momentum_dependency == 1) {
            kn_para = knp * pm_powb(k.pb_p, prm.pindex);
            kn_perp = knperp * pm_powb(k.pb_p, prm.pidx_perp);
        }
        k.kpara = prm.kpara0 * kn_para;
        k.kperp = prm.kpara0 * prm.kperp_kpara * kn_perp * sq(mu);
    }
    k.skpara = pm_sqrt(2.0 * k.kpara);
    k.skperp = pm_sqrt(2.0 * k.kperp);
    k.skpara_perp = pm_sqrt(2.0 * (k.kpara - k.kperp));
    // particle_module.f90:2372-2376: focused transport carries the parallel streaming itself
    const double kpp = prm.focused_transport ? -k.kperp : k.kpara - k.kperp;
    double px_, py_, pz_ = 0.0;  // the "kperp*dkdx" leading terms
    if (!prm.nlgc) {
        double dkdx = 0.0, dkdy = 0.0, dkdz = 0.0;
        if (prm.mag_dependency == 1) {
            if (FULL3D) {  // particle_module.f90:2405-2409: no 1/B in 3-D
                dkdx = B.db_dx * prm.gm2; dkdy = B.db_dy * prm.gm2; dkdz = B.db_dz * prm.gm2;
            } else {
                dkdx = B.db_dx * ib1 * prm.gm2; dkdy = B.db_dy * ib1 * prm.gm2;
            }
        }
        if (dbf) {  // particle_module.f90:2314-2317, 2364-2367, 2410-2414
            dkdx = dkdx - pm_div(aux[1], aux[0]); dkdy = dkdy - pm_div(aux[2], aux[0]);
            if (FULL3D) dkdz = dkdz - pm_div(aux[3], aux[0]);
        }
        if (cof) {  // particle_module.f90:2318-2321, 2368-2371, 2415-2419
            const double g1 = prm.gamma_turb - 1.0;
            dkdx = dkdx + pm_div(g1 * aux[9], aux[8]); dkdy = dkdy + pm_div(g1 * aux[10], aux[8]);
            if (FULL3D) dkdz = dkdz + pm_div(g1 * aux[11], aux[8]);
        }
        px_ = k.kperp * dkdx; py_ = k.kperp * dkdy; pz_ = k.kperp * dkdz;
        ax = kpp * dkdx; ay = kpp * dkdy; az = kpp * dkdz;
    } else {
        double dpa_x = 0.0, dpa_y = 0.0, dpa_z = 0.0, dpe_x = 0.0, dpe_y = 0.0, dpe_z = 0.0;
        if (prm.mag_dependency == 1) {
            dpa_x = B.db_dx * ib1 * prm.gm2; dpa_y = B.db_dy * ib1 * prm.gm2;
            dpe_x = pm_divc(B.db_dx * ib1 * prm.gm2, 3.0); dpe_y = pm_divc(B.db_dy * ib1 * prm.gm2, 3.0);
            if (FULL3D) { dpa_z = B.db_dz * ib1 * prm.gm2; dpe_z = pm_divc(B.db_dz * ib1 * prm.gm2, 3.0); }
        }
        if (dbf) {  // particle_module.f90:2589-2596 and the 2-D+3rd / 3-D twins
            dpa_x = dpa_x - pm_div(aux[1], aux[0]); dpa_y = dpa_y - pm_div(aux[2], aux[0]);
            dpe_x = dpe_x - pm_divc(pm_div(aux[1], aux[0]), 3.0) + pm_divc(pm_div(2.0 * aux[5], aux[4]), 3.0);
            dpe_y = dpe_y - pm_divc(pm_div(aux[2], aux[0]), 3.0) + pm_divc(pm_div(2.0 * aux[6], aux[4]), 3.0);
            if (FULL3D) {
                dpa_z = dpa_z - pm_div(aux[3], aux[0]);
                dpe_z = dpe_z - pm_divc(pm_div(aux[3], aux[0]), 3.0) + pm_divc(pm_div(2.0 * aux[7], aux[4]), 3.0);
            }
        }
        if (cof) {  // particle_module.f90:2597-2604
            const double g1 = prm.gamma_turb - 1.0;
            dpa_x = dpa_x + pm_div(g1 * aux[9], aux[8]); dpa_y = dpa_y + pm_div(g1 * aux[10], aux[8]);
            dpe_x = dpe_x + pm_divc(pm_div(g1 * aux[9], aux[8]), 3.0) + pm_divc(pm_div(2.0 * aux[13], aux[12]), 3.0);
            dpe_y = dpe_y + pm_divc(pm_div(g1 * aux[10], aux[8]), 3.0) + pm_divc(pm_div(2.0 * aux[14], aux[12]), 3.0);
            if (FULL3D) {
                dpa_z = dpa_z + pm_div(g1 * aux[11], aux[8]);
                dpe_z = dpe_z + pm_divc(pm_div(g1 * aux[11], aux[8]), 3.0) + pm_divc(pm_div(2.0 * aux[15], aux[12]), 3.0);
            }
        }
        px_ = k.kperp * dpe_x; py_ = k.kperp * dpe_y; pz_ = k.kperp * dpe_z;
        ax = k.kpara * dpa_x - k.kperp * dpe_x;
        ay = k.kpara * dpa_y - k.kperp * dpe_y;
        az = k.kpara * dpa_z - k.kperp * dpe_z;
    }
    k.dkxx_dx = px_ + ax * sq(bx) * ib2 + 2.0 * kpp * bx * (B.dbx_dx * b - bx * B.db_dx) * ib3;
    k.dkyy_dy = py_ + ay * sq(by) * ib2 + 2.0 * kpp * by * (B.dby_dy * b - by * B.db_dy) * ib3;
    k.dkxy_dx = ax * bx * by * ib2 +
                kpp * ((B.dbx_dx * by + bx * B.dby_dx) * ib2 - 2.0 * bx * by * B.db_dx * ib3);
    k.dkxy_dy = ay * bx * by * ib2 +
                kpp * ((B.dbx_dy * by + bx * B.dby_dy) * ib2 - 2.0 * bx * by * B.db_dy * ib3);
    if (THREE) {
        k.dkzz_dz = pz_ + az * sq(bz) * ib2 + 2.0 * kpp * bz * (B.dbz_dz * b - bz * B.db_dz) * ib3;
        k.dkxz_dx = ax * bx * bz * ib2 +
                    kpp * ((B.dbx_dx * bz + bx * B.dbz_dx) * ib2 - 2.0 * bx * bz * B.db_dx * ib3);
        k.dkxz_dz = az * bx * bz * ib2 +
                    kpp * ((B.dbx_dz * bz + bx * B.dbz_dz) * ib2 - 2.0 * bx * bz * B.db_dz * ib3);
        k.dkyz_dy = ay * by * bz * ib2 +
                    kpp * ((B.dby_dy * bz + by * B.dbz_dy) * ib2 - 2.0 * by * bz * B.db_dy * ib3);
        k.dkyz_dz = az * by * bz * ib2 +
                    kpp * ((B.dby_dz * bz + by * B.dbz_dz) * ib2 - 2.0 * by * bz * B.db_dz * ib3);
    } else {
        k.dkzz_dz = k.dkxz_dx = k.dkxz_dz = k.dkyz_dy = k.dkyz_dz = 0.0;
    }
}

// velocity gradients needed by D_pp
struct VGrad { double dvx_dx, dvy_dy, dvz_dz, dvx_dy, dvx_dz, dvy_dx, dvy_dz, dvz_dx, dvz_dy; };

// calc_dpp_wave_scattering + calc_dpp_flow_shear (particle_module.f90:2918-2979)
__device__ __forceinline__ void momentum_diffusion(const DevParams& prm, const BField& B,
                                                   const VGrad& V, double rho, double divv,
                                                   const Kappa& k, double p, double& dp_dt,
                                                   double& dpp)
{
    if (prm.dpp_wave) {
        double va = pm_div(B.b, pm_sqrt(rho));
        if (prm.momentum_dependency == 1) dp_dt = dp_dt + pm_div(8.0 * p, 27.0 * k.kpara) * sq(va);
        else dp_dt = dp_dt + pm_div(4.0 * p, 9.0 * k.kpara) * sq(va);
        dpp = dpp + pm_div(sq(p * va), 9.0 * k.kpara);
    }
    if (prm.dpp_shear) {
        double sxx = V.dvx_dx - pm_divc(divv, 3.0), syy = V.dvy_dy - pm_divc(divv, 3.0), szz = V.dvz_dz - pm_divc(divv, 3.0);
        double sxy = (V.dvx_dy + V.dvy_dx) / 2.0;
        double sxz = (V.dvx_dz + V.dvz_dx) / 2.0;
        double syz = (V.dvy_dz + V.dvz_dy) / 2.0;
        double gshear;
        if (prm.weak_scattering) {
            double ib = (B.b < kEps) ? 0.0 : pm_div(1.0, B.b);
            double bbs = sxx * sq(B.bx) + syy * sq(B.by) + szz * sq(B.bz) +
                         2.0 * (sxy * B.bx * B.by + sxz * B.bx * B.bz + syz * B.by * B.bz);
            bbs = bbs * ib * ib;
            gshear = pm_divc(sq(bbs), 5.0);
        } else {
            gshear = pm_divc(2.0 * (sq(sxx) + sq(syy) + sq(szz) + 2.0 * (sq(sxy) + sq(sxz) + sq(syz))), 15.0);
        }
        if (gshear > 0.0) {
            const PowBase pbp = pm_base(p);
            dp_dt = dp_dt + (2.0 + prm.pindex) * gshear * prm.tau0 * k.knorm_para *
                                pm_powb(pbp, prm.pindex - 1.0) * prm.p0_pow;
            dpp = dpp + gshear * prm.tau0 * k.knorm_para * pm_powb(pbp, prm.pindex) * prm.p0_pow;
        }
    }
}

__device__ __forceinline__ bool in_acc_region(const DevParams& prm, const Lane& q)
{
    double xn = (q.x - prm.xmin) / prm.lx;
    bool in = (xn >= prm.acc_region[0]) && (xn <= prm.acc_region[1]);
    if (prm.ndim > 1) {  // particle_module.f90:2897
        double yn = (q.y - prm.ymin) / prm.ly;
        in = in && (yn >= prm.acc_region[2]) && (yn <= prm.acc_region[3]);
    }
    if (prm.ndim == 3) {
        double zn = (q.z - prm.zmin) / prm.lz;
        in = in && (zn >= prm.acc_region[4]) && (zn <= prm.acc_region[5]);
    }
    return in;
}

// interp_acc_surface (acc_region_surface.f90:255-334): the two surface heights at the particle's
// pre-step position, from the eight trilinear weights collapsed along each surface's normal
__device__ __forceinline__ void surface_heights(const DevParams& prm, const PushArgs& a, double x, double y,
                                                double z, double rt, double& h1, double& h2)
{
    double rx, ry, rz;
    const double px = (x - prm.xmin) / prm.dx, py = (y - prm.ymin) / prm.dy, pz = (z - prm.zmin) / prm.dz;
    const int pos[3] = {(int)floor(px) + 1, (int)floor(py) + 1, (int)floor(pz) + 1};
    rx = px - (double)pos[0] + 1.0; ry = py - (double)pos[1] + 1.0; rz = pz - (double)pos[2] + 1.0;
    const double rx1 = 1.0 - rx, ry1 = 1.0 - ry, rz1 = 1.0 - rz;
    const double w[8] = {rx1 * ry1 * rz1, rx * ry1 * rz1, rx1 * ry * rz1, rx * ry * rz1,
                         rx1 * ry1 * rz,  rx * ry1 * rz,  rx1 * ry * rz,  rx * ry * rz};
    double out[2] = {0.0, 0.0};
    int i1 = 0, j1 = 0, cur_axis = -1;
    double w2[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < 2; ++k) {
        if (k == 1 && !prm.surface2_existed) break;
        const int norm = k ? prm.surface_norm2 : prm.surface_norm1;
        const int axis = abs(norm) - 1;
        if (axis != cur_axis) {  // the second surface reuses i1, j1, weights_2d when the axes agree
            if (axis == 0) {
                i1 = pos[1]; j1 = pos[2];
                w2[0] = w[0] + w[1]; w2[1] = w[2] + w[3]; w2[2] = w[4] + w[5]; w2[3] = w[6] + w[7];
            } else if (axis == 1) {
                i1 = pos[0]; j1 = pos[2];
                w2[0] = w[0] + w[2]; w2[1] = w[1] + w[3]; w2[2] = w[4] + w[6]; w2[3] = w[5] + w[7];
            } else {
                i1 = pos[0]; j1 = pos[1];
                w2[0] = w[0] + w[4]; w2[1] = w[1] + w[5]; w2[2] = w[2] + w[6]; w2[3] = w[3] + w[7];
            }
            cur_axis = axis;
        }
        const int n1 = a.surf_n1[k], n2 = a.surf_n2[k];
        const int ia = min(max(i1 + 1, 0), n1 - 2), ib = min(max(j1 + 1, 0), n2 - 2);
        const size_t n = (size_t)n1 * n2;
        double hh[2] = {0.0, 0.0};
        const int nslot = prm.time_interp ? 2 : 1;
        for (int sl = 0; sl < nslot; ++sl) {
            const int half = prm.time_interp ? (sl == 0 ? a.sel : (a.sel ^ 1)) : 0;
            const double* sf = a.surf[k] + (size_t)half * n;
            double acc = 0.0;  // sum() over the 2 x 2 section in array-element order
            acc = acc + sf[ia + (size_t)n1 * ib] * w2[0];
            acc = acc + sf[(ia + 1) + (size_t)n1 * ib] * w2[1];
            acc = acc + sf[ia + (size_t)n1 * (ib + 1)] * w2[2];
            acc = acc + sf[(ia + 1) + (size_t)n1 * (ib + 1)] * w2[3];
            hh[sl] = acc;
        }
        out[k] = prm.time_interp ? hh[0] * (1.0 - rt) + hh[1] * rt : hh[0];
    }
    h1 = out[0];
    h2 = out[1];
}

// check_above_acc_surface (acc_region_surface.f90:342-388) at the particle's NEW position
__device__ __forceinline__ bool above_surface(const DevParams& prm, const Lane& q, double h1, double h2)
{
    const double c[3] = {q.x, q.y, q.z};
    double ph = c[abs(prm.surface_norm1) - 1];
    bool in = (prm.surface_norm1 > 0) ? (ph > h1) : (ph < h1);
    if (prm.surface2_existed) {
        ph = c[abs(prm.surface_norm2) - 1];
        const bool in2 = (prm.surface_norm2 > 0) ? (ph > h2) : (ph < h2);
        in = prm.is_intersection ? (in && in2) : (in || in2);
    }
    return in;
}

// push_particle_1d (particle_module.f90:2993-3111) on a 2-D record whose second row is zero.
// Two uniforms per step: ran1 for x, then one for p (particle_module.f90:3085-3088).
template <int L>
__device__ __forceinline__ void push_1d(const DevParams& prm, const PushArgs& a,
                                        const double (&F)[Rec<L>::NREC], double u0, double u1,
                                        Lane& q, bool fixed_dt, const double* aux)
{
    BField B;
    VGrad V;
    B.bx = F[s2::bx]; B.by = F[s2::by]; B.bz = F[s2::bz];
    B.b = pm_sqrt(sq(B.bx) + sq(B.by) + sq(B.bz));
    B.dbx_dx = B.dbx_dy = B.dbx_dz = B.dby_dx = B.dby_dy = B.dby_dz = 0.0;
    B.dbz_dx = B.dbz_dy = B.dbz_dz = B.db_dx = B.db_dy = B.db_dz = 0.0;
    V.dvx_dx = F[s2::dvx_dx];
    V.dvy_dy = V.dvz_dz = V.dvx_dy = V.dvx_dz = V.dvy_dx = V.dvy_dz = V.dvz_dx = V.dvz_dy = 0.0;
    double rho = 1.0;
    if constexpr (Rec<L>::EXT) rho = F[s2::rho];
    Kappa k;
    calc_kappa<false, false>(prm, B, q.p, q.mu, k, aux);
    // particle_module.f90:2271-2292: dkdx has no magnetic term here (mag_dependency = 1 is rejected by
    // gpat_init because the reference would multiply by an unassigned db_dx)
    double dkdx = 0.0;
    if (aux && prm.deltab_flag) dkdx = dkdx - pm_div(aux[1], aux[0]);
    if (aux && prm.correlation_flag) dkdx = dkdx + pm_div((prm.gamma_turb - 1.0) * aux[9], aux[8]);
    k.dkxx_dx = k.kpara * dkdx;
    const double dx_dt = F[s2::vx] + k.dkxx_dx;
    const double divv = V.dvx_dx;
    double dp_dt = pm_divc(-q.p * divv, 3.0);
    double dpp = 0.0;
    if (Rec<L>::EXT) momentum_diffusion(prm, B, V, rho, divv, k, q.p, dp_dt, dpp);
    if (!fixed_dt) {
        double d;
        if (dx_dt != 0.0 && dp_dt != 0.0) {  // particle_module.f90:3060-3072
            const double s = (k.skperp > 0.0) ? k.skperp : k.skpara;
            d = sq(pm_div(0.5 * prm.dx, k.skpara));
            d = min2(d, sq(pm_div(s, dx_dt)));
            d = min2(d, pm_div((double)0.1f * q.p, fabs(dp_dt)));
        } else {
            d = a.dt_min;
        }
        if (d < a.dt_min) d = a.dt_min;
        if (d > a.dt_max) d = a.dt_max;
        q.dt = d;
    }
    const double sdt = pm_sqrt(q.dt);
    const double sqrt3 = 1.7320508075688772;
    const double ran1 = (2.0 * u0 - 1.0) * sqrt3;
    const double ddx = dx_dt * q.dt + ran1 * k.skpara * sdt;
    q.x = q.x + ddx;
    q.t = q.t + q.dt;
    q.dxl = ddx;
    q.dyl = 0.0;
    q.dzl = 0.0;
    const double ranp = (2.0 * u1 - 1.0) * sqrt3;
    double ddp = dp_dt * q.dt + ranp * pm_sqrt(2.0 * dpp) * sdt;
    if (prm.acc_region_flag == 1) {
        if (in_acc_region(prm, q)) q.p = q.p + ddp;
        else ddp = 0.0;
    } else {
        q.p = q.p + ddp;
    }
    const double pfloor = 0.25 * prm.p0;
    if (q.p < pfloor) {
        q.p = q.p - ddp;
        ddp = pfloor - q.p;
        q.p = pfloor;
    }
    q.dpl = ddp;
}

// calc_duu (particle_module.f90:3116-3155) + push_particle_2d_ft (particle_module.f90:3626-3977),
// Cartesian uniform grid.  Uniforms: u0, u1 perpendicular displacement, u2 momentum, u3 pitch angle.
template <int L>
__device__ __forceinline__ void push_2d_ft(const DevParams& prm, const PushArgs& a,
                                           const double (&F)[Rec<L>::NREC], double u0, double u1, double u2,
                                           double u3, Lane& q, bool fixed_dt, const double* aux)
{
    if constexpr (Rec<L>::EXT && Rec<L>::NDIM == 2) {
        const double mu_max = (double)0.99f;
        BField B;
        VGrad V;
        const double vx = F[s2::vx], vy = F[s2::vy], vz = F[s2::vz];
        B.bx = F[s2::bx]; B.by = F[s2::by]; B.bz = F[s2::bz];
        B.dbx_dx = F[s2::dbx_dx]; B.dbx_dy = F[s2::dbx_dy]; B.dby_dx = F[s2::dby_dx];
        B.dby_dy = F[s2::dby_dy]; B.dbz_dx = F[s2::dbz_dx]; B.dbz_dy = F[s2::dbz_dy];
        B.db_dx = F[s2::db_dx]; B.db_dy = F[s2::db_dy];
        B.dbx_dz = B.dby_dz = B.dbz_dz = B.db_dz = 0.0;
        V.dvx_dx = F[s2::dvx_dx]; V.dvy_dy = F[s2::dvy_dy]; V.dvz_dz = 0.0;
        V.dvx_dy = F[s2::dvx_dy]; V.dvy_dx = F[s2::dvy_dx]; V.dvz_dx = 0.0; V.dvz_dy = 0.0;
        V.dvx_dz = V.dvy_dz = 0.0;
        const double dvz_dx = F[s2::dvz_dx], dvz_dy = F[s2::dvz_dy];
        const double rho = F[s2::rho];
        const double bx = B.bx, by = B.by, bz = B.bz;
        B.b = pm_sqrt(sq(bx) + sq(by) + sq(bz));
        const double b = B.b;
        Kappa k;
        calc_kappa<false, false>(prm, B, q.p, q.mu, k, aux);
        const double ib = (b < kEps) ? 0.0 : pm_div(1.0, b);
        const double ib2 = ib * ib, ib3 = ib * ib2;
        // `1.0 / pcharge` is a default-real quotient (particle_module.f90:3716); qdrift holds 1/(3 q)
        const double vdp = pm_div((double)(1.0f / (float)prm.pcharge),
                           pm_sqrt(sq(pm_div(prm.drift1 * prm.p0, q.p)) + sq(pm_div(prm.drift2 * sq(prm.p0), sq(q.p)))));
        const double mu2 = sq(q.mu);
        const double muf1 = 0.5 * (1.0 - mu2), muf2 = 0.5 * (3.0 * mu2 - 1.0);
        const double kx = bx * B.dbx_dx + by * B.dbx_dy;
        const double ky = bx * B.dby_dx + by * B.dby_dy;
        const double kz = bx * B.dbz_dx + by * B.dbz_dy;
        const double bdot_curvb = bx * B.dbz_dy - by * B.dbz_dx + bz * (B.dby_dx - B.dbx_dy);
        const double vdx = vdp * (muf1 * (-bz * B.db_dy) * ib2 + mu2 * (by * kz - bz * ky) * ib3 +
                                  muf1 * bx * bdot_curvb * ib3);
        const double vdy = vdp * (muf1 * (bz * B.db_dx) * ib2 + mu2 * (bz * kx - bx * kz) * ib3 +
                                  muf1 * by * bdot_curvb * ib3);
        double vdz = 0.0;
        if (prm.check_drift_2d)
            vdz = vdp * (muf1 * (bx * B.db_dy - by * B.db_dx) * ib2 + mu2 * (bx * ky - by * kx) * ib3 +
                         muf1 * bz * bdot_curvb * ib3);
        double vbx = q.v * q.mu * ib;
        const double vby = vbx * by;
        vbx = vbx * bx;
        const double dx_dt = vx + vdx + vbx + k.dkxx_dx + k.dkxy_dy;
        const double dy_dt = vy + vdy + vby + k.dkxy_dx + k.dkyy_dy;
        const double dz_dt = vdz;
        const double divv = V.dvx_dx + V.dvy_dy;
        const double bb_gradv = (bx * (bx * V.dvx_dx + by * V.dvx_dy) + by * (bx * V.dvy_dx + by * V.dvy_dy) +
                                 bz * (bx * dvz_dx + by * dvz_dy)) * ib2;
        const double bv_gradv = (bx * (vx * V.dvx_dx + vy * V.dvx_dy) + by * (vx * V.dvy_dx + vy * V.dvy_dy) +
                                 bz * (vx * dvz_dx + vy * dvz_dy)) * ib;
        const double acc_rate = -(muf1 * divv + muf2 * bb_gradv + pm_div(q.mu * bv_gradv, q.v));
        double dp_dt = q.p * acc_rate;
        double dpp = 0.0;
        momentum_diffusion(prm, B, V, rho, divv, k, q.p, dp_dt, dpp);  // sigma_xz = sigma_yz = 0 (V.dvz_* = 0)
        // calc_duu
        const double div_bnorm = -(bx * B.db_dx + by * B.db_dy) * ib2;
        double dmu_dt = q.v * div_bnorm + q.mu * divv - 3 * q.mu * bb_gradv - pm_div(2 * bv_gradv, q.v);
        dmu_dt = dmu_dt * (1 - mu2) * 0.5;
        const double h0 = (double)0.2f;
        const PowBase pbm = pm_base(fabs(q.mu));
        const double dtmp = pm_powb(pbm, prm.gamma_turb - 1) + h0;
        double duu = prm.duu0 * (1 - mu2) * dtmp;
        double duu_du;
        if (q.mu > 0.0) duu_du = prm.duu0 * (-2 * q.mu * dtmp + (1 - mu2) * pm_powb(pbm, prm.gamma_turb - 2));
        else if (q.mu < 0.0) duu_du = prm.duu0 * (-2 * q.mu * dtmp - (1 - mu2) * pm_powb(pbm, prm.gamma_turb - 2));
        else duu_du = 0.0;
        double duu_norm = 1.0;
        if (prm.mag_dependency == 1) duu_norm = duu_norm * pm_powb(k.pb_b, 2.0 - prm.gamma_turb);
        if (aux && prm.deltab_flag) duu_norm = duu_norm * aux[0];                                    // :3143-3145
        if (aux && prm.correlation_flag) duu_norm = duu_norm * pm_powb(k.pb_lc, 1.0 - prm.gamma_turb);   // :3146-3148
        if (prm.momentum_dependency == 1) duu_norm = duu_norm * pm_powb(k.pb_p, prm.gamma_turb - 1);
        duu_du = duu_du * duu_norm;
        duu = duu * duu_norm;
        dmu_dt = dmu_dt + duu_du;

        if (!fixed_dt) {
            double d;
            if (dx_dt != 0.0 && dy_dt != 0.0 && dp_dt != 0.0 && dmu_dt != 0.0) {
                const double s = (k.skperp > 0.0) ? k.skperp : k.skpara;  // particle_module.f90:3847-3864
                d = sq(pm_div(0.5 * prm.dx, s));
                d = min2(d, sq(pm_div(0.5 * prm.dy, s)));
                d = min2(d, sq(pm_div(s, dx_dt)));
                d = min2(d, sq(pm_div(s, dy_dt)));
                d = min2(d, pm_div((double)0.1f * q.p, fabs(dp_dt)));
                d = min2(d, pm_div((double)0.1f, fabs(dmu_dt)));
                d = min2(d, pm_div(2.0 * duu, sq(dmu_dt)));
            } else {
                d = a.dt_min;
            }
            if (d < a.dt_min) d = a.dt_min;
            if (d > a.dt_max) d = a.dt_max;
            q.dt = d;
        }
        const double sdt = pm_sqrt(q.dt);
        const double sqrt3 = 1.7320508075688772;
        double ran1 = (2.0 * u0 - 1.0) * sqrt3;
        const double ran2 = (2.0 * u1 - 1.0) * sqrt3;
        const double bxn = bx * ib, byn = by * ib, bzn = bz * ib;
        const double ibxyn = pm_div(1.0, pm_sqrt(sq(bxn) + sq(byn)));
        // the second term uses the UN-normalised by / bx (particle_module.f90:3912-3913): kept
        const double ddx = dx_dt * q.dt + k.skperp * ibxyn * sdt * (-bxn * bzn * ran1 - by * ran2);
        const double ddy = dy_dt * q.dt + k.skperp * ibxyn * sdt * (-byn * bzn * ran1 + bx * ran2);
        const double ddz = dz_dt * q.dt;
        ran1 = (2.0 * u2 - 1.0) * sqrt3;
        double ddp = dp_dt * q.dt + ran1 * pm_sqrt(2 * dpp) * sdt;
        double ddv = pm_div(q.v * ddp, q.p);
        ran1 = (2.0 * u3 - 1.0) * sqrt3;
        double ddmu = dmu_dt * q.dt + ran1 * pm_sqrt(2 * duu) * sdt;
        q.x = q.x + ddx;
        q.y = q.y + ddy;
        q.z = q.z + ddz;
        q.mu = q.mu + ddmu;
        q.t = q.t + q.dt;
        if (q.mu > mu_max) { ddmu = mu_max - (q.mu - ddmu); q.mu = mu_max; }
        else if (q.mu < -mu_max) { ddmu = -mu_max - (q.mu - ddmu); q.mu = -mu_max; }
        if (prm.acc_region_flag == 1) {
            if (in_acc_region(prm, q)) { q.p = q.p + ddp; q.v = q.v + ddv; }
            else { ddp = 0.0; ddv = 0.0; }
        } else {
            q.p = q.p + ddp;
            q.v = q.v + ddv;
        }
        const double pfloor = 0.25 * prm.p0;
        if (q.p < pfloor) {  // particle_module.f90:3967-3974
            q.v = q.v - ddv;
            ddv = pm_div(q.v * 0.25 * prm.p0, q.p) - q.v;
            q.v = q.v + ddv;
            q.p = q.p - ddp;
            ddp = pfloor - q.p;
            q.p = pfloor;
        }
        q.dxl = ddx; q.dyl = ddy; q.dzl = 0.0;  // the mover's own deltaz stays 0 (particle_module.f90:1564)
        q.dpl = ddp; q.dvl = ddv; q.dmul = ddmu;
    }
}

// push_particle_2d_include_3rd_ft (particle_module.f90:4267-4623) and push_particle_3d_ft (:4930-5320),
// Cartesian, no acc_by_surface: one body, the 2-D variant with every d/dz equal to zero (x - 0, x + 0
// and 0 * x are exact).  Uniforms: u0 u1 (u2 = ran3 is drawn and unused in Cartesian runs), u3 for p,
// u4 for mu -- the first word of a second Philox block of the same step.
template <int L>
__device__ __forceinline__ void push_ft_3d_like(const DevParams& prm, const PushArgs& a,
                                                const double (&F)[Rec<L>::NREC], double u0, double u1, double u3,
                                                double u4, Lane& q, bool fixed_dt, const double* aux)
{
    if constexpr (Rec<L>::EXT) {
        constexpr bool D3 = (Rec<L>::NDIM == 3);
        const double mu_max = (double)0.99f;
        BField B;
        VGrad V;
        double vx, vy, vz, rho;
        if constexpr (!D3) {
            vx = F[s2::vx]; vy = F[s2::vy]; vz = F[s2::vz]; rho = F[s2::rho];
            B.bx = F[s2::bx]; B.by = F[s2::by]; B.bz = F[s2::bz];
            B.dbx_dx = F[s2::dbx_dx]; B.dbx_dy = F[s2::dbx_dy]; B.dby_dx = F[s2::dby_dx];
            B.dby_dy = F[s2::dby_dy]; B.dbz_dx = F[s2::dbz_dx]; B.dbz_dy = F[s2::dbz_dy];
            B.db_dx = F[s2::db_dx]; B.db_dy = F[s2::db_dy];
            B.dbx_dz = B.dby_dz = B.dbz_dz = B.db_dz = 0.0;
            V.dvx_dx = F[s2::dvx_dx]; V.dvy_dy = F[s2::dvy_dy]; V.dvz_dz = 0.0;
            V.dvx_dy = F[s2::dvx_dy]; V.dvy_dx = F[s2::dvy_dx]; V.dvz_dx = F[s2::dvz_dx]; V.dvz_dy = F[s2::dvz_dy];
            V.dvx_dz = V.dvy_dz = 0.0;
        } else {
            vx = F[s3::vx]; vy = F[s3::vy]; vz = F[s3::vz]; rho = F[s3::rho];
            B.bx = F[s3::bx]; B.by = F[s3::by]; B.bz = F[s3::bz];
            B.dbx_dx = F[s3::dbx_dx]; B.dbx_dy = F[s3::dbx_dy]; B.dbx_dz = F[s3::dbx_dz];
            B.dby_dx = F[s3::dby_dx]; B.dby_dy = F[s3::dby_dy]; B.dby_dz = F[s3::dby_dz];
            B.dbz_dx = F[s3::dbz_dx]; B.dbz_dy = F[s3::dbz_dy]; B.dbz_dz = F[s3::dbz_dz];
            B.db_dx = F[s3::db_dx]; B.db_dy = F[s3::db_dy]; B.db_dz = F[s3::db_dz];
            V.dvx_dx = F[s3::dvx_dx]; V.dvy_dy = F[s3::dvy_dy]; V.dvz_dz = F[s3::dvz_dz];
            V.dvx_dy = F[s3::dvx_dy]; V.dvx_dz = F[s3::dvx_dz]; V.dvy_dx = F[s3::dvy_dx];
            V.dvy_dz = F[s3::dvy_dz]; V.dvz_dx = F[s3::dvz_dx]; V.dvz_dy = F[s3::dvz_dy];
        }
        const double bx = B.bx, by = B.by, bz = B.bz;
        B.b = pm_sqrt(sq(bx) + sq(by) + sq(bz));
        const double b = B.b;
        Kappa k;
        if (D3) calc_kappa<true, true>(prm, B, q.p, q.mu, k, aux);
        else calc_kappa<true, false>(prm, B, q.p, q.mu, k, aux);
        const double ib = (b < kEps) ? 0.0 : pm_div(1.0, b);
        const double bxn = bx * ib, byn = by * ib, bzn = bz * ib;
        const double bxyn = pm_sqrt(sq(bxn) + sq(byn));
        const double ibxyn = (bxyn < kEps) ? 0.0 : pm_div(1.0, bxyn);
        const double ib2 = ib * ib, ib3 = ib * ib2;
        const double vdp = pm_div((double)(1.0f / (float)prm.pcharge),
                           pm_sqrt(sq(pm_div(prm.drift1 * prm.p0, q.p)) + sq(pm_div(prm.drift2 * sq(prm.p0), sq(q.p)))));
        const double mu2 = sq(q.mu);
        const double muf1 = 0.5 * (1.0 - mu2), muf2 = 0.5 * (3.0 * mu2 - 1.0);
        double kx, ky, kz, bdot_curvb, vdx, vdy;
        if (D3) {  // particle_module.f90:5049-5063
            kx = bx * B.dbx_dx + by * B.dbx_dy + bz * B.dbx_dz;
            ky = bx * B.dby_dx + by * B.dby_dy + bz * B.dby_dz;
            kz = bx * B.dbz_dx + by * B.dbz_dy + bz * B.dbz_dz;
            bdot_curvb = bx * (B.dbz_dy - B.dby_dz) + by * (B.dbx_dz - B.dbz_dx) + bz * (B.dby_dx - B.dbx_dy);
            vdx = vdp * (muf1 * (by * B.db_dz - bz * B.db_dy) * ib2 + mu2 * (by * kz - bz * ky) * ib3 +
                         muf1 * bx * bdot_curvb * ib3);
            vdy = vdp * (muf1 * (bz * B.db_dx - bx * B.db_dz) * ib2 + mu2 * (bz * kx - bx * kz) * ib3 +
                         muf1 * by * bdot_curvb * ib3);
        } else {  // particle_module.f90:4388-4400
            kx = bx * B.dbx_dx + by * B.dbx_dy;
            ky = bx * B.dby_dx + by * B.dby_dy;
            kz = bx * B.dbz_dx + by * B.dbz_dy;
            bdot_curvb = bx * (B.dbz_dy) + by * (-B.dbz_dx) + bz * (B.dby_dx - B.dbx_dy);
            vdx = vdp * (muf1 * (-bz * B.db_dy) * ib2 + mu2 * (by * kz - bz * ky) * ib3 +
                         muf1 * bx * bdot_curvb * ib3);
            vdy = vdp * (muf1 * (bz * B.db_dx) * ib2 + mu2 * (bz * kx - bx * kz) * ib3 +
                         muf1 * by * bdot_curvb * ib3);
        }
        const double vdz = vdp * (muf1 * (bx * B.db_dy - by * B.db_dx) * ib2 + mu2 * (bx * ky - by * kx) * ib3 +
                                  muf1 * bz * bdot_curvb * ib3);
        double vbx = q.v * q.mu * ib;
        const double vby = vbx * by;
        const double vbz = vbx * bz;
        vbx = vbx * bx;
        double dx_dt, dy_dt, dz_dt, divv, bb_gradv, bv_gradv;
        if (D3) {  // particle_module.f90:5104-5115
            dx_dt = vx + vbx + vdx + k.dkxx_dx + k.dkxy_dy + k.dkxz_dz;
            dy_dt = vy + vby + vdy + k.dkxy_dx + k.dkyy_dy + k.dkyz_dz;
            dz_dt = vz + vbz + vdz + k.dkxz_dx + k.dkyz_dy + k.dkzz_dz;
            divv = V.dvx_dx + V.dvy_dy + V.dvz_dz;
            bb_gradv = (bx * (bx * V.dvx_dx + by * V.dvx_dy + bz * V.dvx_dz) +
                        by * (bx * V.dvy_dx + by * V.dvy_dy + bz * V.dvy_dz) +
                        bz * (bx * V.dvz_dx + by * V.dvz_dy + bz * V.dvz_dz)) * ib2;
            bv_gradv = (bx * (vx * V.dvx_dx + vy * V.dvx_dy + vz * V.dvx_dz) +
                        by * (vx * V.dvy_dx + vy * V.dvy_dy + vz * V.dvy_dz) +
                        bz * (vx * V.dvz_dx + vy * V.dvz_dy + vz * V.dvz_dz)) * ib;
        } else {  // particle_module.f90:4432-4442
            dx_dt = vx + vbx + vdx + k.dkxx_dx + k.dkxy_dy;
            dy_dt = vy + vby + vdy + k.dkxy_dx + k.dkyy_dy;
            dz_dt = vz + vbz + vdz + k.dkxz_dx + k.dkyz_dy;
            divv = V.dvx_dx + V.dvy_dy;
            bb_gradv = (bx * (bx * V.dvx_dx + by * V.dvx_dy) + by * (bx * V.dvy_dx + by * V.dvy_dy) +
                        bz * (bx * V.dvz_dx + by * V.dvz_dy)) * ib2;
            bv_gradv = (bx * (vx * V.dvx_dx + vy * V.dvx_dy) + by * (vx * V.dvy_dx + vy * V.dvy_dy) +
                        bz * (vx * V.dvz_dx + vy * V.dvz_dy)) * ib;
        }
        const double acc_rate = -(muf1 * divv + muf2 * bb_gradv + pm_div(q.mu * bv_gradv, q.v));
        double dp_dt = q.p * acc_rate;
        double dpp = 0.0;
        momentum_diffusion(prm, B, V, rho, divv, k, q.p, dp_dt, dpp);
        const double div_bnorm = D3 ? -(bx * B.db_dx + by * B.db_dy + bz * B.db_dz) * ib2
                                    : -(bx * B.db_dx + by * B.db_dy) * ib2;
        double dmu_dt = q.v * div_bnorm + q.mu * divv - 3 * q.mu * bb_gradv - pm_div(2 * bv_gradv, q.v);
        dmu_dt = dmu_dt * (1 - mu2) * 0.5;
        const double h0 = (double)0.2f;
        const PowBase pbm = pm_base(fabs(q.mu));
        const double dtmp = pm_powb(pbm, prm.gamma_turb - 1) + h0;
        double duu = prm.duu0 * (1 - mu2) * dtmp;
        double duu_du;
        if (q.mu > 0.0) duu_du = prm.duu0 * (-2 * q.mu * dtmp + (1 - mu2) * pm_powb(pbm, prm.gamma_turb - 2));
        else if (q.mu < 0.0) duu_du = prm.duu0 * (-2 * q.mu * dtmp - (1 - mu2) * pm_powb(pbm, prm.gamma_turb - 2));
        else duu_du = 0.0;
        double duu_norm = 1.0;
        if (prm.mag_dependency == 1) duu_norm = duu_norm * pm_powb(k.pb_b, 2.0 - prm.gamma_turb);
        if (aux && prm.deltab_flag) duu_norm = duu_norm * aux[0];
        if (aux && prm.correlation_flag) duu_norm = duu_norm * pm_powb(k.pb_lc, 1.0 - prm.gamma_turb);
        if (prm.momentum_dependency == 1) duu_norm = duu_norm * pm_powb(k.pb_p, prm.gamma_turb - 1);
        duu_du = duu_du * duu_norm;
        duu = duu * duu_norm;
        dmu_dt = dmu_dt + duu_du;
        if (!fixed_dt) {
            double d;
            bool ok = dx_dt != 0.0 && dy_dt != 0.0 && dp_dt != 0.0 && dmu_dt != 0.0;
            if (D3) ok = ok && dz_dt != 0.0;  // particle_module.f90:5156-5160; the 2-D variant does not test dz_dt
            if (ok) {
                const double s = (k.skperp > 0.0) ? k.skperp : k.skpara;
                d = sq(pm_div(0.5 * prm.dx, s));
                d = min2(d, sq(pm_div(0.5 * prm.dy, s)));
                if (D3) d = min2(d, sq(pm_div(0.5 * prm.dz, s)));
                d = min2(d, sq(pm_div(s, dx_dt)));
                d = min2(d, sq(pm_div(s, dy_dt)));
                if (D3) d = min2(d, sq(pm_div(s, dz_dt)));
                d = min2(d, pm_div((double)0.1f * q.p, fabs(dp_dt)));
                d = min2(d, pm_div((double)0.1f, fabs(dmu_dt)));
                d = min2(d, pm_div(2.0 * duu, sq(dmu_dt)));
            } else {
                d = a.dt_min;
            }
            if (d < a.dt_min) d = a.dt_min;
            if (d > a.dt_max) d = a.dt_max;
            q.dt = d;
        }
        const double sdt = pm_sqrt(q.dt);
        const double sqrt3 = 1.7320508075688772;
        double ran1 = (2.0 * u0 - 1.0) * sqrt3;
        const double ran2 = (2.0 * u1 - 1.0) * sqrt3;
        const double ddx = dx_dt * q.dt + (-bxn * bzn * k.skperp * ibxyn * ran1 - byn * k.skperp * ibxyn * ran2) * sdt;
        const double ddy = dy_dt * q.dt + (-byn * bzn * k.skperp * ibxyn * ran1 + bxn * k.skperp * ibxyn * ran2) * sdt;
        const double ddz = dz_dt * q.dt + bxyn * k.skperp * ran1 * sdt;
        ran1 = (2.0 * u3 - 1.0) * sqrt3;
        double ddp = dp_dt * q.dt + ran1 * pm_sqrt(2 * dpp) * sdt;
        double ddv = pm_div(q.v * ddp, q.p);
        ran1 = (2.0 * u4 - 1.0) * sqrt3;
        double ddmu = dmu_dt * q.dt + ran1 * pm_sqrt(2 * duu) * sdt;
        q.x = q.x + ddx;
        q.y = q.y + ddy;
        q.z = q.z + ddz;
        q.mu = q.mu + ddmu;
        q.t = q.t + q.dt;
        if (q.mu > mu_max) { ddmu = mu_max - (q.mu - ddmu); q.mu = mu_max; }
        else if (q.mu < -mu_max) { ddmu = -mu_max - (q.mu - ddmu); q.mu = -mu_max; }
        if (prm.acc_region_flag == 1) {
            bool in = in_acc_region(prm, q);
            if (prm.acc_by_surface) in = in && above_surface(prm, q, q.sh1, q.sh2);  // particle_module.f90:5297-5303
            if (in) { q.p = q.p + ddp; q.v = q.v + ddv; }
            else { ddp = 0.0; ddv = 0.0; }
        } else {
            q.p = q.p + ddp;
            q.v = q.v + ddv;
        }
        const double pfloor = 0.25 * prm.p0;
        if (q.p < pfloor) {
            q.v = q.v - ddv;
            ddv = pm_div(q.v * 0.25 * prm.p0, q.p) - q.v;
            q.v = q.v + ddv;
            q.p = q.p - ddp;
            ddp = pfloor - q.p;
            q.p = pfloor;
        }
        q.dxl = ddx; q.dyl = ddy; q.dzl = ddz;
        q.dpl = ddp; q.dvl = ddv; q.dmul = ddmu;
    }
}

// One call of push_particle_*: everything between the BC test and the step counter.
// One push_particle_* call on the interpolated record F, every model switch read at run time, reference operation
// order.  Reference-order build: called by push_once() below.  Production build: phase C of the kSpecAlt
// instantiations of push_kernel_coop (1-D, focused transport, turbulence maps) -- same statements, compiled with FMA
// contraction and the production locate() / u01(); held to 1e-12 per step like the other production kernels.
// PRE: the caller gathered the map record already (aux_pre, or nullptr in a run without maps)
template <int L, bool TRACK = false, bool PRE = false, int PATH = PATH_ANY>
__device__ __forceinline__ void push_physics(const DevParams& prm, const PushArgs& a,
                                             const double (&F)[Rec<L>::NREC], Lane& q, bool fixed_dt, double rt,
                                             const double* aux_pre = nullptr)
{
    // tracked particles carry negated tags; the streams are keyed by the magnitudes so that a
    // tracking run replays the run its particles were selected from
    const int tag_inj = TRACK ? abs(q.tag_inj) : q.tag_inj, tag_spl = TRACK ? abs(q.tag_spl) : q.tag_spl;
    constexpr bool D3 = (Rec<L>::NDIM == 3);
    constexpr bool EXT = Rec<L>::EXT;

    // uniforms of this step: ran1, ran2, ran3, ran_p
    double u0, u1, u2, u3;
    if (prm.rng_mode == GPAT_RNG_TABLE) {
        long long slot = tag_inj;
        if (a.rng_table && slot >= 0 && slot < a.rng_slots && (long long)q.rng < a.rng_max_steps) {
            const double* tb = a.rng_table + ((size_t)slot * a.rng_max_steps + q.rng) * 4;
            u0 = tb[0]; u1 = tb[1]; u2 = tb[2]; u3 = tb[3];
        } else {
            u0 = u1 = u2 = u3 = 0.5;
        }
    } else {
        uint4 r = philox4x32_10(make_uint4((unsigned)q.rng, (unsigned)(q.rng >> 32),
                                           (unsigned)tag_inj, (unsigned)tag_spl),
                                prm.key0, prm.key1 + (unsigned)q.origin);
        u0 = u01(r.x); u1 = u01(r.y); u2 = u01(r.z); u3 = u01(r.w);
    }
    q.rng += 1;

    const double* auxp = aux_pre;  // production build: the lane group gathered the map record with the fields
    double A[PRE ? 1 : 16];
    if constexpr (!PRE) {
        if (a.aux && (prm.deltab_flag || prm.correlation_flag)) {  // particle_module.f90:1634-1639
            gather_aux<Rec<L>::NDIM>(prm, a.aux, a.sel, q.x, q.y, q.z, rt, A);
            auxp = A;
        }
    }
    if constexpr (D3) {  // particle_module.f90:1662-1665, 1683-1686
        if (prm.acc_by_surface) surface_heights(prm, a, q.x, q.y, q.z, rt, q.sh1, q.sh2);
    }
    const bool is_1d = (PATH == PATH_ANY) ? (prm.ndim == 1) : (PATH == PATH_1D);
    const bool is_ft = (PATH == PATH_ANY) ? (prm.focused_transport != 0) : (PATH == PATH_FT);
    if constexpr (!D3) {
        if (is_1d) {
            push_1d<L>(prm, a, F, u0, u1, q, fixed_dt, auxp);
            return;
        }
    }
    if (is_ft) {
        if (D3 || (EXT && prm.include_3rd_dim)) {
            // fifth uniform: first word of a second Philox block of this step (counter word 1, top bit flipped)
            const unsigned long long step = q.rng - 1;
            const uint4 r5 = philox4x32_10(make_uint4((unsigned)step, (unsigned)(step >> 32) ^ 0x80000000u,
                                                       (unsigned)tag_inj, (unsigned)tag_spl),
                                           prm.key0, prm.key1 + (unsigned)q.origin);
            push_ft_3d_like<L>(prm, a, F, u0, u1, u3, u01(r5.x), q, fixed_dt, auxp);
        } else if constexpr (!D3) {
            push_2d_ft<L>(prm, a, F, u0, u1, u2, u3, q, fixed_dt, auxp);
        }
        return;
    }

    BField B;
    VGrad V;
    double vx, vy, vz = 0.0, rho = 1.0;
    if constexpr (!D3) {
        vx = F[s2::vx]; vy = F[s2::vy];
        B.bx = F[s2::bx]; B.by = F[s2::by]; B.bz = F[s2::bz];
        B.dbx_dx = F[s2::dbx_dx]; B.dbx_dy = F[s2::dbx_dy]; B.dby_dx = F[s2::dby_dx];
        B.dby_dy = F[s2::dby_dy]; B.dbz_dx = F[s2::dbz_dx]; B.dbz_dy = F[s2::dbz_dy];
        B.db_dx = F[s2::db_dx]; B.db_dy = F[s2::db_dy];
        B.dbx_dz = B.dby_dz = B.dbz_dz = B.db_dz = 0.0;
        V.dvx_dx = F[s2::dvx_dx]; V.dvy_dy = F[s2::dvy_dy]; V.dvz_dz = 0.0;
        V.dvx_dz = V.dvy_dz = 0.0;
        if constexpr (EXT) {
            vz = F[s2::vz]; rho = F[s2::rho];
            V.dvx_dy = F[s2::dvx_dy]; V.dvy_dx = F[s2::dvy_dx];
            V.dvz_dx = F[s2::dvz_dx]; V.dvz_dy = F[s2::dvz_dy];
        } else {
            V.dvx_dy = V.dvy_dx = V.dvz_dx = V.dvz_dy = 0.0;
        }
    } else {
        vx = F[s3::vx]; vy = F[s3::vy]; vz = F[s3::vz];
        B.bx = F[s3::bx]; B.by = F[s3::by]; B.bz = F[s3::bz];
        B.dbx_dx = F[s3::dbx_dx]; B.dbx_dy = F[s3::dbx_dy]; B.dbx_dz = F[s3::dbx_dz];
        B.dby_dx = F[s3::dby_dx]; B.dby_dy = F[s3::dby_dy]; B.dby_dz = F[s3::dby_dz];
        B.dbz_dx = F[s3::dbz_dx]; B.dbz_dy = F[s3::dbz_dy]; B.dbz_dz = F[s3::dbz_dz];
        B.db_dx = F[s3::db_dx]; B.db_dy = F[s3::db_dy]; B.db_dz = F[s3::db_dz];
        V.dvx_dx = F[s3::dvx_dx]; V.dvy_dy = F[s3::dvy_dy]; V.dvz_dz = F[s3::dvz_dz];
        if constexpr (EXT) {
            rho = F[s3::rho];
            V.dvx_dy = F[s3::dvx_dy]; V.dvx_dz = F[s3::dvx_dz]; V.dvy_dx = F[s3::dvy_dx];
            V.dvy_dz = F[s3::dvy_dz]; V.dvz_dx = F[s3::dvz_dx]; V.dvz_dy = F[s3::dvz_dy];
        } else {
            V.dvx_dy = V.dvx_dz = V.dvy_dx = V.dvy_dz = V.dvz_dx = V.dvz_dy = 0.0;
        }
    }
    B.b = pm_sqrt(sq(B.bx) + sq(B.by) + sq(B.bz));
    const bool third = D3 || (EXT && prm.include_3rd_dim);  // push_particle_3d-like path

    Kappa k;
    if (D3) calc_kappa<true, true>(prm, B, q.p, q.mu, k, auxp);
    else if (EXT && prm.include_3rd_dim) calc_kappa<true, false>(prm, B, q.p, q.mu, k, auxp);
    else calc_kappa<false, false>(prm, B, q.p, q.mu, k, auxp);

    const double ib = (B.b < kEps) ? 0.0 : pm_div(1.0, B.b);  // particle_module.f90:3414-3418
    const double ib2 = ib * ib;
    const double ib3 = ib * ib2;
    // particle_module.f90:3436: 1.0/(3*pcharge) is an FP32 quotient
    const double vdp = pm_div(prm.qdrift, pm_sqrt(sq(pm_div(prm.drift1 * prm.p0, q.p)) +
                                         sq(pm_div(prm.drift2 * sq(prm.p0), sq(q.p)))));
    double vdx, vdy, vdz;
    if (!third) {
        vdx = vdp * (B.dbz_dy * ib2 - 2.0 * B.bz * B.db_dy * ib3);
        vdy = vdp * (-B.dbz_dx * ib2 + 2.0 * B.bz * B.db_dx * ib3);
        vdz = prm.check_drift_2d
                  ? vdp * ((B.dby_dx - B.dbx_dy) * ib2 - 2.0 * (B.by * B.db_dx - B.bx * B.db_dy) * ib3)
                  : 0.0;
    } else {
        vdx = vdp * ((B.dbz_dy - B.dby_dz) * ib2 - 2.0 * (B.bz * B.db_dy - B.by * B.db_dz) * ib3);
        vdy = vdp * ((B.dbx_dz - B.dbz_dx) * ib2 - 2.0 * (B.bx * B.db_dz - B.bz * B.db_dx) * ib3);
        vdz = vdp * ((B.dby_dx - B.dbx_dy) * ib2 - 2.0 * (B.by * B.db_dx - B.bx * B.db_dy) * ib3);
    }
    double dx_dt, dy_dt, dz_dt, divv;
    if (!third) {
        dx_dt = vx + vdx + k.dkxx_dx + k.dkxy_dy;
        dy_dt = vy + vdy + k.dkxy_dx + k.dkyy_dy;
        dz_dt = vdz;
        divv = V.dvx_dx + V.dvy_dy;
    } else {
        if (!D3) { k.dkxz_dz = 0.0; k.dkyz_dz = 0.0; k.dkzz_dz = 0.0; }  // particle_module.f90:4094-4096
        dx_dt = vx + vdx + k.dkxx_dx + k.dkxy_dy + k.dkxz_dz;
        dy_dt = vy + vdy + k.dkxy_dx + k.dkyy_dy + k.dkyz_dz;
        dz_dt = vz + vdz + k.dkxz_dx + k.dkyz_dy + k.dkzz_dz;
        divv = V.dvx_dx + V.dvy_dy + V.dvz_dz;
    }
    double dp_dt = pm_divc(-q.p * divv, 3.0);
    double dpp = 0.0;
    if (EXT) {
        if (!third) { V.dvz_dx = 0.0; V.dvz_dy = 0.0; }  // push_particle_2d: sigma_xz = sigma_yz = 0
        momentum_diffusion(prm, B, V, rho, divv, k, q.p, dp_dt, dpp);
    }

    if (!fixed_dt) {
        double d;
        if (dx_dt != 0.0 && dy_dt != 0.0 && dp_dt != 0.0) {
            const double s = (k.skperp > 0.0) ? k.skperp : k.skpara;
            d = sq(pm_div(0.5 * prm.dx, k.skpara));
            d = min2(d, sq(pm_div(0.5 * prm.dy, k.skpara)));
            if (D3) d = min2(d, sq(pm_div(0.5 * prm.dz, k.skpara)));
            d = min2(d, sq(pm_div(s, dx_dt)));
            d = min2(d, sq(pm_div(s, dy_dt)));
            if (D3) d = min2(d, sq(pm_div(s, dz_dt)));
            d = min2(d, pm_div((double)0.1f * q.p, fabs(dp_dt)));
        } else {
            d = a.dt_min;
        }
        if (d < a.dt_min) d = a.dt_min;
        if (d > a.dt_max) d = a.dt_max;
        q.dt = d;
    }
    const double sdt = pm_sqrt(q.dt);
    const double sqrt3 = 1.7320508075688772;  // dsqrt(3.0d0), correctly rounded
    const double ran1 = (2.0 * u0 - 1.0) * sqrt3;
    const double ran2 = (2.0 * u1 - 1.0) * sqrt3;
    const double ran3 = (2.0 * u2 - 1.0) * sqrt3;
    double ddx, ddy, ddz;
    if (!third) {
        ddx = dx_dt * q.dt + ran1 * k.skperp * sdt + ran3 * k.skpara_perp * sdt * B.bx * ib;
        ddy = dy_dt * q.dt + ran2 * k.skperp * sdt + ran3 * k.skpara_perp * sdt * B.by * ib;
        ddz = dz_dt * q.dt;
        q.dzl = 0.0;  // the mover's own deltaz stays 0 in plain 2-D (particle_module.f90:1564)
    } else {
        const double bxn = B.bx * ib, byn = B.by * ib, bzn = B.bz * ib;
        const double bxyn = pm_sqrt(sq(bxn) + sq(byn));
        const double ibxyn = (bxyn < kEps) ? 0.0 : pm_div(1.0, bxyn);
        ddx = dx_dt * q.dt + (bxn * k.skpara * ran1 - bxn * bzn * k.skperp * ibxyn * ran2 -
                              byn * k.skperp * ibxyn * ran3) * sdt;
        ddy = dy_dt * q.dt + (byn * k.skpara * ran1 - byn * bzn * k.skperp * ibxyn * ran2 +
                              bxn * k.skperp * ibxyn * ran3) * sdt;
        ddz = dz_dt * q.dt + (bzn * k.skpara * ran1 + bxyn * k.skperp * ran2) * sdt;
        q.dzl = ddz;
    }
    q.x = q.x + ddx;
    q.y = q.y + ddy;
    q.z = q.z + ddz;
    q.t = q.t + q.dt;
    q.dxl = ddx;
    q.dyl = ddy;

    const double ranp = (2.0 * u3 - 1.0) * sqrt3;
    double ddp = dp_dt * q.dt + ranp * pm_sqrt(2.0 * dpp) * sdt;
    if (prm.acc_region_flag == 1) {
        bool in = in_acc_region(prm, q);
        if (prm.acc_by_surface) in = in && above_surface(prm, q, q.sh1, q.sh2);  // particle_module.f90:4887-4892
        if (in) q.p = q.p + ddp;
        else ddp = 0.0;
    } else {
        q.p = q.p + ddp;
    }
    const double pfloor = 0.25 * prm.p0;
    if (q.p < pfloor) {  // particle_module.f90:3601-3605
        q.p = q.p - ddp;
        ddp = pfloor - q.p;
        q.p = pfloor;
    }
    q.dpl = ddp;
}

template <int L, bool TRACK = false>
__device__ __forceinline__ void push_once(const DevParams& prm, const PushArgs& a,
                                          const float* __restrict__ fld, Lane& q, bool fixed_dt)
{
    double F[Rec<L>::NREC];
    const double rt = (q.t - a.t0) / a.dtf;
    gather<L>(prm, fld, a.sel, q.x, q.y, q.z, rt, F);
    push_physics<L, TRACK>(prm, a, F, q, fixed_dt, rt);
}

#if !GPAT_STRICT
#include "push_fast.cuh"
#endif

// ---- the kernel -------------------------------------------------------------------------
enum : int { ST_IDLE = 0, ST_ADAPT = 1, ST_FIX = 2 };
enum : int { AT_OUTER_HEAD = 0, AT_INNER_HEAD = 1, AFTER_FIXED_PUSH = 2 };

// The reference's loop nest (particle_module.f90:1596-1829) as a resumable state machine:
//   do while (dt_target < dtf + 0.1 dt_fine)            <- AT_OUTER_HEAD
//     if (.not. inbox) exit
//     do while (t - t0 < dt_target .and. inbox)          <- AT_INNER_HEAD
//        [BC; adaptive push]                              -> returns ST_ADAPT
//     if (t - t0 > dt_target .and. inbox) roll back, [fixed push] -> returns ST_FIX,
//        then dt = dt_old; BC                             <- AFTER_FIXED_PUSH
//     dt_target += dt_fine
// Returns ST_IDLE when the particle is done for this interval.
template <bool TRACK = false, bool ALT = false>
__device__ __forceinline__ int next_state(const DevParams& prm, const PushArgs& a, Lane& q,
                                          int entry)
{
    for (;;) {
        if (entry == AT_INNER_HEAD) {
            const bool inbox = (q.count_flag == GPAT_COUNT_FLAG_INBOX);
            if ((q.t - a.t0) < q.dt_target && inbox) return ST_ADAPT;
            if ((q.t - a.t0) > q.dt_target && inbox) {  // particle_module.f90:1707-1716
                q.x = q.x - q.dxl; q.y = q.y - q.dyl; q.z = q.z - q.dzl;
                q.p = q.p - q.dpl;
                if constexpr (kStrict || ALT) {
                    if (prm.focused_transport) { q.v = q.v - q.dvl; q.mu = q.mu - q.dmul; }  // particle_module.f90:1712-1713
                }
                q.t = q.t - q.dt;
                q.dt_old = q.dt;
                q.dt = a.t0 + q.dt_target - q.t;
                if (q.dt > 0) {
                    if (TRACK && q.tag_spl < 0 && q.nsteps_pushed == 0) {  // particle_module.f90:1717-1721
                        q.nsteps_tracked = q.nsteps_tracked - 1;             // back one sample
                        q.nsteps_pushed = a.nsteps_interval - 2;
                    } else {
                        q.nsteps_pushed = q.nsteps_pushed - 1;  // particle_module.f90:1723
                    }
                    return ST_FIX;
                }
                q.dt = q.dt_old;  // particle_module.f90:1816-1825
                negp_or_bc<ALT>(prm, q, a.leak);
            }
            q.dt_target = q.dt_target + a.dt_fine;
        } else if (entry == AFTER_FIXED_PUSH) {
            q.dt = q.dt_old;
            negp_or_bc<ALT>(prm, q, a.leak);
            q.dt_target = q.dt_target + a.dt_fine;
        }
        if (!(q.dt_target < a.dt_target_limit) || q.count_flag != GPAT_COUNT_FLAG_INBOX)
            return ST_IDLE;
        entry = AT_INNER_HEAD;
    }
}

// particle record -> lane registers, and the initial state of its loop nest
// (particle_module.f90:1561-1592)
template <bool TRACK = false, bool ALT = false>
__device__ __forceinline__ int load_lane(const DevParams& prm, const PushArgs& a, const PtlSoA& P,
                                         long long idx, Lane& q, int& remaining)
{
    q.nsteps_tracked = 1;  // reset at the start of the MHD interval, particle_module.f90:1910-1913
    q.x = P.x[idx]; q.y = P.y[idx]; q.z = P.z[idx]; q.p = P.p[idx];
    q.t = P.t[idx]; q.dt = P.dt[idx]; q.weight = P.weight[idx]; q.mu = P.mu[idx];
    q.rng = P.rng[idx]; q.tag_inj = P.tag_injected[idx];
    q.tag_spl = P.tag_splitted[idx]; q.origin = P.origin[idx];
    q.nsteps_pushed = P.nsteps_pushed[idx]; q.count_flag = P.count_flag[idx];
    q.dxl = q.dyl = q.dzl = q.dpl = 0.0;
    if constexpr (kStrict || ALT) {
        q.v = P.v[idx]; q.dvl = 0.0; q.dmul = 0.0;
        q.sh1 = 0.0; q.sh2 = 0.0;
    }
    q.dt_old = q.dt;
    if (a.debug_nsteps > 0) {
        remaining = a.debug_nsteps;
        q.dt_target = a.dtf;
        return (q.count_flag == GPAT_COUNT_FLAG_INBOX) ? ST_ADAPT : ST_IDLE;
    }
    // target time of the first fine step, particle_module.f90:1570-1578
    int step = (int)ceil((q.t - a.t0) / a.dt_fine);
    q.dt_target = (step <= 0) ? a.dt_fine : step * a.dt_fine;
    if (q.dt_target > a.dtf) q.dt_target = a.dtf;
    // safe check, particle_module.f90:1581-1592
    if (q.p < 0.0 && q.count_flag == GPAT_COUNT_FLAG_INBOX) {
        q.count_flag = GPAT_COUNT_FLAG_OTHERS;
        atomicAdd(a.leak + 1, q.weight);
    } else {
        boundary<ALT>(prm, q, prm.ext, a.leak);
    }
    return (q.count_flag == GPAT_COUNT_FLAG_INBOX) ? next_state<TRACK, ALT>(prm, a, q, AT_OUTER_HEAD) : ST_IDLE;
}

template <bool ALT = false>
__device__ __forceinline__ void store_lane(const PushArgs& a, const PtlSoA& P, long long idx,
                                           const Lane& q)
{
    P.x[idx] = q.x; P.y[idx] = q.y; P.z[idx] = q.z; P.p[idx] = q.p;
    P.t[idx] = q.t; P.dt[idx] = q.dt; P.rng[idx] = q.rng;
    if constexpr (kStrict || ALT) { P.v[idx] = q.v; P.mu[idx] = q.mu; }  // changed by focused transport only
    P.nsteps_pushed[idx] = q.nsteps_pushed;
    P.count_flag[idx] = (signed char)q.count_flag;
    // particle_module.f90:1913 sets 1 at the start of the interval; tracked particles count up from it
    if (a.debug_nsteps == 0) P.nsteps_tracked[idx] = q.nsteps_tracked;
}

// idle lanes take the next particles of the work counter (one warp-aggregated atomic)
template <bool TRACK = false, bool PERM4 = false, bool ALT = false>
__device__ __forceinline__ void refill(const DevParams& prm, const PushArgs& a, const PtlSoA& P,
                                       unsigned lane, Lane& q, int& state, long long& idx,
                                       bool& exhausted, int& remaining)
{
    __syncwarp();
    const bool want = (state == ST_IDLE) && !exhausted;
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    unsigned long long base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) {
        unsigned rank;
        if (PERM4) {
            // Lane-group kernels with four lanes per particle serve, in round r, the particles of lanes r, 4 + r, 8 + r ...
            // Hand consecutive queue entries to exactly those lanes (virtual position (lane & 3) * 8 + lane / 4): with
            // cell-sorted particles the eight gathers of one load instruction then touch NEIGHBOURING grid points, and
            // the x + 1 corner of one particle is the x corner of the next -- an L1 hit a few instructions later.
            unsigned mp = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned t = (m >> k) & 0x11111111u;         // lanes 4g + k -> bit 4g
                t = (t | (t >> 3)) & 0x03030303u;
                t = (t | (t >> 6)) & 0x000F000Fu;
                t = (t | (t >> 12)) & 0x000000FFu;           // -> bit g
                mp |= t << (8 * k);
            }
            const unsigned v = (lane & 3u) * 8u + (lane >> 2);
            rank = __popc(mp & ((1u << v) - 1u));
        } else {
            rank = __popc(m & ((1u << lane) - 1u));
        }
        idx = (long long)(base + rank);
        if (idx >= a.nptl) {
            exhausted = true;
        } else {
            state = load_lane<TRACK, ALT>(prm, a, P, idx, q, remaining);
            if (state == ST_IDLE) store_lane<ALT>(a, P, idx, q);
        }
    }
}

// after one push_particle_* call: step counter, next state (particle_module.f90:1694-1704)
// the tracking block after every push (particle_module.f90:1697-1703, 1806-1812): a tracked
// particle (negative tag_splitted) is sampled every nsteps_interval pushes into the rows of all
// the selected particles it is an ancestor of
__device__ __noinline__ void track_sample(const PushArgs& a, const PtlSoA& P, long long idx, Lane& q)
{
    long long lo, hi;
    const int nsplit = P.split_times[idx];
    trk_selected(a.trk, q.origin, q.tag_inj, q.tag_spl, nsplit, lo, hi);  // locate_particle
    q.nsteps_tracked = q.nsteps_tracked + 1;
    gpat_particle r;
    r.split_times = (int8_t)nsplit; r.count_flag = (int8_t)q.count_flag; r.pad_[0] = r.pad_[1] = 0;
    r.origin = q.origin; r.nsteps_tracked = q.nsteps_tracked; r.nsteps_pushed = q.nsteps_pushed;
    r.tag_injected = q.tag_inj; r.tag_splitted = q.tag_spl;
    r.x = q.x; r.y = q.y; r.z = q.z; r.p = q.p; r.v = P.v[idx]; r.mu = q.mu;
    r.weight = q.weight; r.t = q.t; r.dt = q.dt;
    r.padding = __longlong_as_double((long long)q.rng);
    trk_record(a.trk, r, lo, hi);
}

template <bool TRACK = false, int SPEC = 0>
__device__ __forceinline__ int after_push(const DevParams& prm, const PushArgs& a, Lane& q, int state,
                                          int& remaining, const PtlSoA& P, long long idx)
{
    // mod(nsteps_pushed + 1, nsteps_interval), particle_module.f90:1694; the counter is already
    // inside [0, interval) except right after a restart with a smaller interval
    {
        const int n1 = q.nsteps_pushed + 1;
        q.nsteps_pushed = (n1 < a.nsteps_interval) ? n1 : (n1 == a.nsteps_interval ? 0 : n1 % a.nsteps_interval);
    }
    if (TRACK && q.tag_spl < 0 && q.nsteps_pushed == 0) track_sample(a, P, idx, q);
    // gpat_debug_push_n: a uniform branch on a kernel argument, kept in the switch-specialised instantiations too so
    // that the per-step parity tests run the very kernels bench.py times
    if (a.debug_nsteps > 0) return (--remaining == 0) ? ST_IDLE : ST_ADAPT;
    return next_state<TRACK, spec_is_alt(SPEC)>(prm, a, q, state == ST_FIX ? AFTER_FIXED_PUSH : AT_INNER_HEAD);
}

template <int L, bool TRACK = false>
__global__ void __launch_bounds__(kBlock)
push_kernel(const __grid_constant__ DevParams prm, const PtlSoA P, const float* __restrict__ fld,
            const __grid_constant__ PushArgs a)
{
    const unsigned lane = threadIdx.x & 31u;
    Lane q;
    int state = ST_IDLE;
    long long idx = -1;
    bool exhausted = false;
    int remaining = 0;
    unsigned long long nsteps = 0;

    for (;;) {
        refill<TRACK>(prm, a, P, lane, q, state, idx, exhausted, remaining);
        if (__all_sync(0xffffffffu, state == ST_IDLE)) {
            if (__all_sync(0xffffffffu, exhausted)) break;
            continue;
        }

        // ---- one push for every busy lane ----
        if (state == ST_ADAPT) {  // top of the inner while body, particle_module.f90:1602-1612
            negp_or_bc(prm, q, a.leak);
            if (q.count_flag != GPAT_COUNT_FLAG_INBOX) {
                store_lane(a, P, idx, q);
                state = ST_IDLE;
            }
        }
        if (state != ST_IDLE) {
#if GPAT_STRICT
            push_once<L, TRACK>(prm, a, fld, q, state == ST_FIX);
#else
            push_once_fast<L, TRACK>(prm, a, fld, q, state == ST_FIX);
#endif
            nsteps++;
            state = after_push<TRACK>(prm, a, q, state, remaining, P, idx);
            if (state == ST_IDLE) store_lane(a, P, idx, q);
        }
    }
    // warp-aggregated step count
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsteps += __shfl_down_sync(0xffffffffu, nsteps, o);
    if (lane == 0 && nsteps) atomicAdd(a.steps, nsteps);
}

#if !GPAT_STRICT
// ---- lane-group gather ---------------------------------------------------------------------
// A lane that gathers its own particle touches 32 different 128-byte lines per warp load
// instruction, and the L1 data pipe spends one wavefront per line: that pipe, not HBM or L2,
// bounds the one-lane-per-particle kernel (profiles/r01a_push_kernel.md).  Here G lanes read
// ONE particle's line together (G x 32 contiguous bytes per corner), so a warp-wide load touches
// 32/G lines; the G particles of a group are served in G rounds and every lane ends a round
// with final values of the slots of its chunks, parked in shared memory for the owner lane.
#ifndef GPAT_COOP_DEPTH
#define GPAT_COOP_DEPTH 1
#endif
#ifndef GPAT_PUB_FACTORS
#define GPAT_PUB_FACTORS 0   // 1: A/B build publishing y x time factors instead of finished weights (measured: -1 %)
#endif
#ifndef GPAT_UNROLL_3D
#define GPAT_UNROLL_3D 0   // 1: A/B build with the four rounds of the 3-D kernels unrolled like the 2-D ones
#endif
#define GPAT_ROUNDS_UNROLL(NC, G) ((GPAT_UNROLL_3D || (NC) != 8) ? (G) : 1)
#ifndef GPAT_NO_PERM
#define GPAT_NO_PERM 0   // 1: A/B build without the neighbour-preserving lane assignment of refill()
#endif
// FP32 -> FP64 on the integer pipe.  F2F.F64.F32 runs on the XU at 16 lanes/clk/SM (2 cycles per
// warp instruction, scripts/micro/pipes.cu) and the 128 conversions of a 2-D step were the busiest
// pipe of the kernel.  Placing the float's sign | 8-bit exponent | 23-bit mantissa into the
// sign | 11-bit exponent | 52-bit mantissa fields of a double WITHOUT re-biasing the exponent gives
// exactly f * 2^-896 (zero, denormals and signs included), and the missing 2^896 is folded into the
// interpolation weight, so the FMA sees the same exact product f*w as with a real conversion.
// GPAT_CVT_ALU_MASK picks which (row cy, frame half h) quarter of the corner values goes this
// way: bit 2*cy + h.
#ifndef GPAT_ALT_MINBLOCKS
#define GPAT_ALT_MINBLOCKS 3   // resident CTAs per SM the 2-D kSpecAlt instantiations are compiled for (168 registers)
#endif
#ifndef GPAT_ALT_MINBLOCKS_3D
#define GPAT_ALT_MINBLOCKS_3D 2  // 3-D: 255 registers, no spills: focused transport +9 %, maps +21 % over 3 CTAs (call Z)
#endif
#ifndef GPAT_SKIP_ROUNDS
#define GPAT_SKIP_ROUNDS 0
#endif
#ifndef GPAT_CVT_WIDE
#define GPAT_CVT_WIDE 0
#endif
#ifndef GPAT_CVT_ALU_MASK
#define GPAT_CVT_ALU_MASK 10  // the second frame half on the integer pipe: XU 60 % -> 33 %, +1.8 % steps/s on C1
#endif
constexpr double kTwo896 = 5.2829453113566525e+269;  // 2^896
__device__ __forceinline__ double cvt(float f) { return (double)f; }
__device__ __forceinline__ double cvt_scaled(float f)  // f * 2^-896
{
    const int b = __float_as_int(f);
#if !GPAT_CVT_WIDE
    return __hiloint2double((b >> 3) & 0x8fffffff, b << 29);  // SHF + LOP3 + IMAD.SHL: three ALU instructions
#else
    // A/B build, MEASURED AND REJECTED (profiles/README.md, call X): one signed 32 x 32 -> 64 multiply by 2^29 produces both
    // words (high word = b >> 3 arithmetic, low word = b << 29), i.e. IMAD.WIDE + one LOP3 instead of three ALU
    // instructions (these conversions are 18.9 % of the executed instructions of the C1 kernel).  64 fewer instructions per
    // warp-step and 1.3 % SLOWER on C1, 1.6 % on C5: IMAD.WIDE is a multi-cycle instruction of the FMA pipe, the three
    // single-cycle ALU instructions were not the limit.  (PTX with an immediate: the compiler turns the C++
    // multiplication back into the two shifts.)
    long long p;
    asm("mul.wide.s32 %0, %1, 0x20000000;" : "=l"(p) : "r"(b));
    return __hiloint2double(__double2hiint(__longlong_as_double(p)) & 0x8fffffff, __double2loint(__longlong_as_double(p)));
#endif
}
template <int CY, int H> __device__ __forceinline__ double cvt_sel(float f)
{
    if constexpr ((GPAT_CVT_ALU_MASK >> (2 * CY + H)) & 1) return cvt_scaled(f);
    else return cvt(f);
}
__host__ __device__ constexpr double cvt_weight_scale(int cy, int h) { return ((GPAT_CVT_ALU_MASK >> (2 * cy + h)) & 1) ? kTwo896 : 1.0; }

// acc[0..3] += four slots of one frame half at corner c (row cy = bit 1 of c) times its weight
template <int CYDUMMY, int H>
__device__ __forceinline__ void fma_chunk(int c, const float4& f, double w, double (&acc)[4])
{
    if ((c >> 1) & 1) {
        acc[0] = fma(cvt_sel<1, H>(f.x), w, acc[0]); acc[1] = fma(cvt_sel<1, H>(f.y), w, acc[1]);
        acc[2] = fma(cvt_sel<1, H>(f.z), w, acc[2]); acc[3] = fma(cvt_sel<1, H>(f.w), w, acc[3]);
    } else {
        acc[0] = fma(cvt_sel<0, H>(f.x), w, acc[0]); acc[1] = fma(cvt_sel<0, H>(f.y), w, acc[1]);
        acc[2] = fma(cvt_sel<0, H>(f.z), w, acc[2]); acc[3] = fma(cvt_sel<0, H>(f.w), w, acc[3]);
    }
}

template <int L, bool AUX = false> struct Coop {
    static constexpr int NREC = Rec<L>::NREC;
    static constexpr int NCH = NREC / 4;                          // 32-byte chunks per grid point
#ifdef GPAT_COOP_G
    static constexpr int G = GPAT_COOP_G;
#else
    static constexpr int G = (NCH == 4 || NCH == 8) ? 4 : 2;      // lanes per particle
#endif
    static constexpr int CPL = NCH / G;                           // chunks per lane
    static constexpr int NC = (Rec<L>::NDIM == 3) ? 8 : 4;
    // result rows (G = 4): NREC doubles + 32 B.  Rows of ODD lane groups start 16 B later (row_off), so that
    // the two groups of a quarter-warp store phase hit disjoint banks, and 8 consecutive owner
    // rows (160 B apart in 2-D) still tile the 32 banks on the read side
    // (profiles/r01e_push_coop_ncu.txt: 3.2e9 shared-memory bank conflicts before this).
    static constexpr bool SKEW = (G == 4);                        // G = 2 rows (NREC = 24) are conflict-free at +16 B
    // side plane (L2D): in round r lane gq of the group loads corner gq's float4 of particle r and parks its two
    // weighted partial sums behind the record part of the owner's row; the owner adds the four corners up
    // Chunk-format side plane (L3D, two chunks per grid point): one load instruction brings the 128 contiguous
    // bytes of an x-adjacent corner PAIR to the four lanes (lane gq: chunk gq & 1 of the corner with x-bit gq >> 1);
    // over the four (y, z) pairs every lane ends with four partial sums, parked behind the record part.
    static constexpr int NSIDE = Rec<L>::NSIDE;
    static constexpr int SCH = Rec<L>::SIDE_CHUNKS;
    static constexpr int SIDEROW = SCH ? 4 * G : (NSIDE ? 2 * G : 0);   // doubles parked per owner row
    // AUX (kSpecAlt instantiations): the 16 interpolated values of the turbulence-map record (128 B per grid point, four
    // chunks in the fields' own chunk format) are gathered by the same lanes with the same weights -- 4 / G map chunks
    // per lane -- and parked behind the record part of the owner's row.
    // They live in rows of their own (AROW doubles per owner) so that the record rows keep the pitch and the skew
    // measured above whether a run has maps or not.
    static constexpr int ACPL = AUX ? 4 / G : 0;
    static constexpr int AROW = AUX ? 18 : 0;
    static constexpr int ROW = (SKEW ? NREC + 4 : NREC + 2) + SIDEROW;
    // parameter rows.  2-D: the owner publishes its eight finished corner weights (time blend and
    // conversion scale folded in) + the cell, 80 B; the other lanes of the group load them instead
    // of recomputing 16 products per round.  3-D: rx ry t0 t1 cell rz, 48 B, weights per round.
    static constexpr bool PUBW = (NC == 4);
    // PUBF (2-D): publish the four y x time factors + rx + the cell (48 B, three LDS.128 per round) and let every lane
    // form the eight weights with the owner's own products (bit-identical: rx1 * a0 ...), instead of the eight finished
    // weights (80 B, four LDS.128 + one LDS.64 per round): 40 fewer shared-memory wavefronts per warp-step for 36 more
    // FP64 instructions.  MEASURED AND REJECTED (profiles/README.md, call R): 1 % slower on C1, C2, C3 and C4 -- trading
    // L1-pipe wavefronts for issue slots does not pay although the L1 data pipe is the busiest unit (80 % vs 65 %).
    static constexpr bool PUBF = PUBW && (GPAT_PUB_FACTORS != 0);
    static constexpr int PAR = PUBF ? 6 : (PUBW ? 10 : 6);
    static constexpr int CELL_AT = PUBF ? 5 : (PUBW ? 8 : 4);
    __device__ static __forceinline__ int row_off(int owner) { return owner * ROW + (SKEW ? ((owner / G) & 1) * 2 : 0); }
};

// resident CTAs per SM the register allocation must allow: the kernel is latency-bound, and
// 2-D Parker sits right at the 128-register edge between 4 and 3 CTAs (16 vs 12 warps: 14 %)
#ifdef GPAT_MINBLOCKS
template <int L> struct MinBlocks { static constexpr int V = GPAT_MINBLOCKS; };
#else
// 3-D Parker at 4 CTAs spills 32 bytes and is still 6 % faster than 3 CTAs on C5 (profiles/README.md)
template <int L> struct MinBlocks { static constexpr int V = (L == L2B || L == L3B || L == L2D || L == L3D) ? 4 : 3; };
#endif
// SEL = which half of the store is farray1 (PushArgs::sel).  It is a template parameter because
// the two frames must enter every sum in the order (farray1, farray2) whatever half they live in:
// a run restarted from a dump starts with sel = 0 again and has to continue bit-identically
// (tests/test_gpu_parity.py::test_restart_round_trip_is_bit_exact).
template <int L, int SEL, bool TRACK = false, int SPEC = 0>
__global__ void __launch_bounds__(kBlock, (spec_is_alt(SPEC) ? (Rec<L>::NDIM == 3 ? GPAT_ALT_MINBLOCKS_3D : GPAT_ALT_MINBLOCKS)
                                                               : MinBlocks<L>::V))
push_kernel_coop(const __grid_constant__ DevParams prm, const PtlSoA P,
                 const float* __restrict__ fld, const __grid_constant__ PushArgs a)
{
    using C = Coop<L, spec_has_maps(SPEC)>;
    constexpr int NW = kBlock / 32;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wid = threadIdx.x >> 5;
    // rows of the kSpecAlt instantiations carry the 16 map values too and exceed the 48 KB of static shared memory with the
    // wide records: those instantiations take their rows from dynamic shared memory (coop_smem_bytes, launch_one)
    extern __shared__ __align__(16) double coop_dyn[];
    __shared__ __align__(16) double sm_par[spec_is_alt(SPEC) ? 1 : NW][spec_is_alt(SPEC) ? 2 : 32 * C::PAR];
    __shared__ __align__(16) double sm_res[spec_is_alt(SPEC) ? 1 : NW][spec_is_alt(SPEC) ? 2 : 32 * C::ROW];
    double* const par = spec_is_alt(SPEC) ? coop_dyn + wid * (32 * C::PAR) : sm_par[wid];
    double* const res = spec_is_alt(SPEC) ? coop_dyn + NW * (32 * C::PAR) + wid * (32 * C::ROW) : sm_res[wid];
    double* const ares = coop_dyn + NW * (32 * (C::PAR + C::ROW)) + wid * (32 * C::AROW);  // kSpecAltMaps only
    const int gq = (int)(lane & (C::G - 1));
    const int gbase = (int)(lane & ~(unsigned)(C::G - 1));
    constexpr long long stride = 2LL * C::NREC;

    Lane q;
    int state = ST_IDLE;
    long long idx = -1;
    bool exhausted = false;
    int remaining = 0;
    unsigned long long nsteps = 0;

    constexpr bool ALT = spec_is_alt(SPEC);
    constexpr bool has_aux = spec_has_maps(SPEC);  // launch_one picks kSpecAltMaps exactly when the run has maps
    for (;;) {
        refill<TRACK, (C::G == 4) && !GPAT_NO_PERM, ALT>(prm, a, P, lane, q, state, idx, exhausted, remaining);
        if (__all_sync(0xffffffffu, state == ST_IDLE)) {
            if (__all_sync(0xffffffffu, exhausted)) break;
            continue;
        }
        // top of the inner while body, particle_module.f90:1602-1612.  One combined test keeps the
        // common case (inside the extended box, p >= 0) to a single untaken branch.
        if (state == ST_ADAPT && outside_or_negp<(Rec<L>::THIRD != 0), ALT>(prm, q)) {
            negp_or_bc<ALT>(prm, q, a.leak);
            if (q.count_flag != GPAT_COUNT_FLAG_INBOX) {
                store_lane<ALT>(a, P, idx, q);
                state = ST_IDLE;
            }
        }

        // ---- phase A: every lane publishes where its particle is ----
        {
            // t0/t1: time-blend factors of half 0 / half 1 of every chunk
            double rx = 0.0, ry = 0.0, rz = 0.0, t0 = 0.0, t1 = 0.0;
            long long cell = 0;
            if (state != ST_IDLE) {
                cell = locate<Rec<L>::NDIM, spec_is_alt(SPEC)>(prm, q.x, q.y, q.z, rx, ry, rz);
                const double rt = (q.t - a.t0) * a.idtf;
                const bool ti = (SPEC & 1) || prm.time_interp;  // SPEC: time interpolation on
                const double tA = ti ? 1.0 - rt : 1.0, tB = ti ? rt : 0.0;
                t0 = (SEL == 0) ? tA : tB;
                t1 = (SEL == 0) ? tB : tA;
            }
            double2* row = reinterpret_cast<double2*>(par + lane * C::PAR);
            if constexpr (C::PUBW) {
                const double rx1 = 1.0 - rx, ry1 = 1.0 - ry;
                const double a0 = ry1 * t0 * cvt_weight_scale(0, 0), b0 = ry * t0 * cvt_weight_scale(1, 0);
                const double a1 = ry1 * t1 * cvt_weight_scale(0, 1), b1 = ry * t1 * cvt_weight_scale(1, 1);
                if constexpr (C::PUBF) {
                    row[0] = make_double2(a0, b0);
                    row[1] = make_double2(a1, b1);
                    row[2] = make_double2(rx, __longlong_as_double(cell));
                } else {
                    row[0] = make_double2(rx1 * a0, rx * a0);  // w0[0..3]: half 0 at the four corners
                    row[1] = make_double2(rx1 * b0, rx * b0);
                    row[2] = make_double2(rx1 * a1, rx * a1);  // w1[0..3]: half 1
                    row[3] = make_double2(rx1 * b1, rx * b1);
                    row[4] = make_double2(__longlong_as_double(cell), 0.0);
                }
            } else {
                row[0] = make_double2(rx, ry);
                row[1] = make_double2(t0, t1);
                row[2] = make_double2(__longlong_as_double(cell), rz);
            }
        }
        __syncwarp();

        // ---- phase B: G rounds, one particle of the group per round; the loads of DEPTH rounds
        // are in flight at once (the L2 round trip is the longest stall of the whole step) ----
        {
            constexpr int NLD = C::NC * C::CPL;  // 256-bit loads per lane per round
            constexpr int DEPTH = (GPAT_COOP_DEPTH < C::G) ? GPAT_COOP_DEPTH : C::G;
            float4 lo[DEPTH][NLD], hi[DEPTH][NLD];
            float4 sd[DEPTH];  // side plane (L2D): corner gq of the round's particle
            float4 slo[DEPTH][C::SCH ? 4 : 1], shi[DEPTH][C::SCH ? 4 : 1];  // side plane (L3D): the four (y, z) corner pairs
            auto issue = [&](int r, int slot) {
                const long long cell = __double_as_longlong(par[(gbase + r) * C::PAR + C::CELL_AT]);
                const float* base = fld + cell * stride + (gq * C::CPL) * 8;
                if constexpr (C::SCH > 0) {
                    static_assert(C::G == 4 && C::NC == 8 && C::SCH == 2, "chunk-format side plane: 3-D, two chunks, four lanes");
                    const float* side = fld + (long long)prm.nxg * prm.nyg * prm.nzg * stride;
                    const float* sb = side + (cell + (gq >> 1)) * 16 + (gq & 1) * 8;
#pragma unroll
                    for (int pr = 0; pr < 4; ++pr)
                        ldg256(sb + ((long long)(pr & 1) * prm.nxg + (long long)(pr >> 1) * prm.nxg * prm.nyg) * 16,
                               slo[slot][pr], shi[slot][pr]);
                } else if constexpr (C::NSIDE > 0) {
                    static_assert(C::G == 4 && C::NC == 4 && C::PUBW, "side plane: one corner per lane of the group");
                    const float* side = fld + (long long)prm.nxg * prm.nyg * prm.nzg * stride;
                    sd[slot] = __ldg(reinterpret_cast<const float4*>(
                        side + (cell + (gq & 1) + (long long)(gq >> 1) * prm.nxg) * 4));
                }
#pragma unroll
                for (int c = 0; c < C::NC; ++c) {
                    const float* pc_ = base + ((c & 1) + (long long)((c >> 1) & 1) * prm.nxg +
                                               (long long)(c >> 2) * prm.nxg * prm.nyg) * stride;
#pragma unroll
                    for (int j = 0; j < C::CPL; ++j)
                        ldg256(pc_ + 8 * j, lo[slot][c * C::CPL + j], hi[slot][c * C::CPL + j]);
                }
            };
            // GPAT_SKIP_ROUNDS=1 (A/B build, MEASURED AND REJECTED, profiles/README.md call T3): skip the rounds whose owner
            // lanes are all idle (warp-uniform test), so that at the end of a launch, when the queue is empty and a warp holds
            // one or two unfinished particles, an iteration costs one or two L2 round trips instead of four.  The uniform
            // branches around the loads stop the compiler from overlapping the rounds: C1 -8 % (2.23e10 -> 2.06e10), and even
            // the tail-dominated 125 000-particle run loses 6 %.  Four-lane groups only: with two lanes per particle the
            // conditional prefetch of 96 load registers spills 400 bytes.
            constexpr bool kSkip = (GPAT_SKIP_ROUNDS != 0) && (C::G == 4);
            const unsigned act = kSkip ? __ballot_sync(0xffffffffu, state != ST_IDLE) : 0xffffffffu;
            auto need = [&](int r) { return !kSkip || (act & (0x11111111u << r)) != 0u; };
            constexpr bool ROLLED = (GPAT_ROUNDS_UNROLL(C::NC, C::G) == 1);  // rolled: loads at the top of every round,
            if constexpr (!ROLLED) {                                         // no buffers carried across the back-edge
#pragma unroll
                for (int r = 0; r < DEPTH; ++r)
                    if (need(r)) issue(r, r);
            }
            // 2-D: the four rounds are unrolled (2700-instruction body, no instruction-fetch stalls).  3-D: a rolled loop --
            // unrolled, the body is 4096 instructions = 64 KB and "no instruction" is the second largest stall reason
            // (2.0 per issue at 512^3, 1.15 at 256^3: profiles/r02n_push_coop_c5_512_ncu.txt)
            constexpr int kRoundsUnroll = GPAT_ROUNDS_UNROLL(C::NC, C::G);
#pragma unroll kRoundsUnroll
            for (int r = 0; r < C::G; ++r) {
                const int owner = gbase + r;
                const int slot = ROLLED ? 0 : r % DEPTH;
                if (!need(r)) {
                    if constexpr (!ROLLED) {
                        if (r + DEPTH < C::G && need(r + DEPTH)) issue(r + DEPTH, slot);
                    }
                    continue;
                }
                if constexpr (ROLLED) issue(r, 0);
                const double2* row = reinterpret_cast<const double2*>(par + owner * C::PAR);
                // weights of half 0 / half 1 at each corner (time blend folded in)
                double w0[C::NC], w1[C::NC];
                if constexpr (C::PUBF) {
                    const double2 f0 = row[0], f1 = row[1];
                    const double rx = row[2].x, rx1 = 1.0 - rx;
                    w0[0] = rx1 * f0.x; w0[1] = rx * f0.x; w0[2] = rx1 * f0.y; w0[3] = rx * f0.y;
                    w1[0] = rx1 * f1.x; w1[1] = rx * f1.x; w1[2] = rx1 * f1.y; w1[3] = rx * f1.y;
                } else if constexpr (C::PUBW) {
                    const double2 q0 = row[0], q1 = row[1], q2 = row[2], q3 = row[3];
                    w0[0] = q0.x; w0[1] = q0.y; w0[2] = q1.x; w0[3] = q1.y;
                    w1[0] = q2.x; w1[1] = q2.y; w1[2] = q3.x; w1[3] = q3.y;
                } else {
                    const double2 pa = row[0], pb = row[1];
                    const double rx = pa.x, ry = pa.y, t0 = pb.x, t1 = pb.y;
                    const double rx1 = 1.0 - rx, ry1 = 1.0 - ry;
                    {
                        const double rz = par[owner * C::PAR + 5];
                        const double rz1 = 1.0 - rz;
                        const double wxy[4] = {rx1 * ry1, rx * ry1, rx1 * ry, rx * ry};
                        if constexpr (cvt_weight_scale(0, 0) == cvt_weight_scale(1, 0) &&
                                      cvt_weight_scale(0, 1) == cvt_weight_scale(1, 1)) {
                            // the conversion scale does not depend on the row: fold it into the four z x time factors
                            // (a power of two: bit-identical to scaling each of the sixteen weights)
                            const double s0 = cvt_weight_scale(0, 0), s1 = cvt_weight_scale(0, 1);
                            const double z00 = rz1 * t0 * s0, z10 = rz * t0 * s0, z01 = rz1 * t1 * s1, z11 = rz * t1 * s1;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                w0[c] = wxy[c] * z00; w0[c + 4] = wxy[c] * z10;
                                w1[c] = wxy[c] * z01; w1[c + 4] = wxy[c] * z11;
                            }
                        } else {
                            const double z00 = rz1 * t0, z10 = rz * t0, z01 = rz1 * t1, z11 = rz * t1;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const double s0 = cvt_weight_scale(c >> 1, 0), s1 = cvt_weight_scale(c >> 1, 1);
                                w0[c] = wxy[c] * z00 * s0; w0[c + 4] = wxy[c] * z10 * s0;
                                w1[c] = wxy[c] * z01 * s1; w1[c + 4] = wxy[c] * z11 * s1;
                            }
                        }
                    }
                }
                double acc[C::CPL][4];
#pragma unroll
                for (int j = 0; j < C::CPL; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;
#pragma unroll
                for (int c = 0; c < C::NC; ++c) {
#pragma unroll
                    for (int j = 0; j < C::CPL; ++j) {
                        const float4 f0 = lo[slot][c * C::CPL + j], f1 = hi[slot][c * C::CPL + j];
                        if constexpr (SEL == 0) {
                            fma_chunk<0, 0>(c, f0, w0[c], acc[j]);
                            fma_chunk<0, 1>(c, f1, w1[c], acc[j]);
                        } else {
                            fma_chunk<0, 1>(c, f1, w1[c], acc[j]);
                            fma_chunk<0, 0>(c, f0, w0[c], acc[j]);
                        }
                    }
                }
                double2 sidep = make_double2(0.0, 0.0);
                double sacc[4] = {0.0, 0.0, 0.0, 0.0};
                if constexpr (C::SCH > 0) {
                    const bool xb = (gq >> 1) != 0;
#pragma unroll
                    for (int pr = 0; pr < 4; ++pr) {  // corner = x-bit | pr << 1: same weights as the record part
                        const double wa = xb ? w0[2 * pr + 1] : w0[2 * pr], wb = xb ? w1[2 * pr + 1] : w1[2 * pr];
                        if constexpr (SEL == 0) {
                            fma_chunk<0, 0>(2 * pr, slo[slot][pr], wa, sacc);
                            fma_chunk<0, 1>(2 * pr, shi[slot][pr], wb, sacc);
                        } else {
                            fma_chunk<0, 1>(2 * pr, shi[slot][pr], wb, sacc);
                            fma_chunk<0, 0>(2 * pr, slo[slot][pr], wa, sacc);
                        }
                    }
                } else if constexpr (C::NSIDE > 0) {
                    // this lane's corner: weights of half 0 / half 1 from the owner's published row (w0[gq], w1[gq])
                    static_assert(cvt_weight_scale(0, 0) == cvt_weight_scale(1, 0) && cvt_weight_scale(0, 1) == cvt_weight_scale(1, 1),
                                  "the side gather assumes a conversion mask that does not depend on the row");
                    // (selected from the registers the record part already loaded: a second trip to the parameter
                    // row costs two 4-wavefront LDS.64 per round, profiles/r02c_push_coop_c4_l2d_ncu.txt)
                    const bool g1 = (gq & 1) != 0, g2 = (gq & 2) != 0;
                    const double wa = g2 ? (g1 ? w0[3] : w0[2]) : (g1 ? w0[1] : w0[0]);
                    const double wb = g2 ? (g1 ? w1[3] : w1[2]) : (g1 ? w1[1] : w1[0]);
                    const float4 e = sd[slot];
                    if constexpr (SEL == 0) {
                        sidep.x = fma(cvt_sel<0, 1>(e.z), wb, cvt_sel<0, 0>(e.x) * wa);
                        sidep.y = fma(cvt_sel<0, 1>(e.w), wb, cvt_sel<0, 0>(e.y) * wa);
                    } else {
                        sidep.x = fma(cvt_sel<0, 0>(e.x), wa, cvt_sel<0, 1>(e.z) * wb);
                        sidep.y = fma(cvt_sel<0, 0>(e.y), wa, cvt_sel<0, 1>(e.w) * wb);
                    }
                }
                if constexpr (!ROLLED) {
                    if (r + DEPTH < C::G && need(r + DEPTH)) issue(r + DEPTH, slot);
                }
                double2* out = reinterpret_cast<double2*>(res + C::row_off(owner) + (gq * C::CPL) * 4);
                if constexpr (C::SCH > 0) {
                    double2* so = reinterpret_cast<double2*>(res + C::row_off(owner) + C::NREC + 4 * gq);
                    so[0] = make_double2(sacc[0], sacc[1]);
                    so[1] = make_double2(sacc[2], sacc[3]);
                } else if constexpr (C::NSIDE > 0)
                    *reinterpret_cast<double2*>(res + C::row_off(owner) + C::NREC + 2 * gq) = sidep;
#pragma unroll
                for (int j = 0; j < C::CPL; ++j) {
                    out[2 * j] = make_double2(acc[j][0], acc[j][1]);
                    out[2 * j + 1] = make_double2(acc[j][2], acc[j][3]);
                }
                if constexpr (C::ACPL > 0) {
                    // turbulence maps (interp_magnetic_fluctuation + interp_correlation_length, mhd_data_parallel.f90:
                    // 1806-1915): a second, smaller gather of this round's particle with the weights already in registers
                    {
                        const long long cell = __double_as_longlong(par[owner * C::PAR + C::CELL_AT]);
                        const float* ab = a.aux + cell * 32 + (gq * C::ACPL) * 8;
                        float4 alo[C::NC * C::ACPL], ahi[C::NC * C::ACPL];
#pragma unroll
                        for (int c = 0; c < C::NC; ++c) {
                            const float* pc_ = ab + ((c & 1) + (long long)((c >> 1) & 1) * prm.nxg +
                                                     (long long)(c >> 2) * prm.nxg * prm.nyg) * 32;
#pragma unroll
                            for (int j = 0; j < C::ACPL; ++j) ldg256(pc_ + 8 * j, alo[c * C::ACPL + j], ahi[c * C::ACPL + j]);
                        }
                        double aacc[C::ACPL][4];
#pragma unroll
                        for (int j = 0; j < C::ACPL; ++j) aacc[j][0] = aacc[j][1] = aacc[j][2] = aacc[j][3] = 0.0;
#pragma unroll
                        for (int c = 0; c < C::NC; ++c) {
#pragma unroll
                            for (int j = 0; j < C::ACPL; ++j) {
                                if constexpr (SEL == 0) {
                                    fma_chunk<0, 0>(c, alo[c * C::ACPL + j], w0[c], aacc[j]);
                                    fma_chunk<0, 1>(c, ahi[c * C::ACPL + j], w1[c], aacc[j]);
                                } else {
                                    fma_chunk<0, 1>(c, ahi[c * C::ACPL + j], w1[c], aacc[j]);
                                    fma_chunk<0, 0>(c, alo[c * C::ACPL + j], w0[c], aacc[j]);
                                }
                            }
                        }
                        double2* ao = reinterpret_cast<double2*>(ares + owner * C::AROW + (gq * C::ACPL) * 4);
#pragma unroll
                        for (int j = 0; j < C::ACPL; ++j) {
                            ao[2 * j] = make_double2(aacc[j][0], aacc[j][1]);
                            ao[2 * j + 1] = make_double2(aacc[j][2], aacc[j][3]);
                        }
                    }
                }
            }
        }
        __syncwarp();

        // ---- phase C: the owner lane finishes its push ----
        if (state != ST_IDLE) {
            double F[Rec<L>::NF];
            const double2* row = reinterpret_cast<const double2*>(res + C::row_off((int)lane));
#pragma unroll
            for (int k = 0; k < C::NREC / 2; ++k) {
                const double2 v = row[k];
                F[2 * k] = v.x;
                F[2 * k + 1] = v.y;
            }
            if constexpr (C::SCH > 0) {  // lanes (0, 2) hold chunk 0 at x-bit 0 / 1, lanes (1, 3) chunk 1
                const double2* sr = row + C::NREC / 2;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const double2 a0 = sr[2 * ch], a1 = sr[2 * ch + 1], b0 = sr[2 * (ch + 2)], b1 = sr[2 * (ch + 2) + 1];
                    F[C::NREC + 4 * ch] = a0.x + b0.x; F[C::NREC + 4 * ch + 1] = a0.y + b0.y;
                    F[C::NREC + 4 * ch + 2] = a1.x + b1.x; F[C::NREC + 4 * ch + 3] = a1.y + b1.y;
                }
            } else if constexpr (C::NSIDE > 0) {  // the four corners' partial sums, always in corner order
                const double2 c0 = row[C::NREC / 2], c1 = row[C::NREC / 2 + 1], c2 = row[C::NREC / 2 + 2], c3 = row[C::NREC / 2 + 3];
                F[C::NREC] = ((c0.x + c1.x) + c2.x) + c3.x;
                F[C::NREC + 1] = ((c0.y + c1.y) + c2.y) + c3.y;
                F[C::NREC + 2] = F[C::NREC + 3] = 0.0;
            }
            if constexpr (ALT) {
                static_assert(!ALT || Rec<L>::NF == Rec<L>::NREC, "the general pushers read one-plane records");
                // Always a valid array (registers, constant indices): the pushers read it only under deltab_flag /
                // correlation_flag, and gpat_particle_mover refuses to run with a flag set and no maps uploaded.
                double A[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) A[k] = 1.0;
                if constexpr (has_aux) {
                    const double2* ar = reinterpret_cast<const double2*>(ares + (int)lane * C::AROW);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const double2 v = ar[k];
                        A[2 * k] = v.x;
                        A[2 * k + 1] = v.y;
                    }
                }
                push_physics<L, TRACK, true, spec_path(SPEC)>(prm, a, F, q, state == ST_FIX, (q.t - a.t0) * a.idtf, A);
            } else {
                physics_fast<L, double[Rec<L>::NF], TRACK, SPEC>(prm, a, F, q, state == ST_FIX);
            }
            nsteps++;
            state = after_push<TRACK, SPEC>(prm, a, q, state, remaining, P, idx);
            if (state == ST_IDLE) store_lane<ALT>(a, P, idx, q);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsteps += __shfl_down_sync(0xffffffffu, nsteps, o);
    if (lane == 0 && nsteps) atomicAdd(a.steps, nsteps);
}
#endif

// debug: interpolated fields in the reference's 32-slot order
template <int L>
__global__ void interp_kernel(const __grid_constant__ DevParams prm, const float* __restrict__ fld,
                              int sel, long long n, const double* x, const double* y,
                              const double* z, const double* rt, double* out32)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double F[Rec<L>::NF];
    gather<L>(prm, fld, sel, x[i], y[i], z[i], rt[i], F);
    for (int s = 0; s < 32; ++s) out32[i * 32 + s] = 0.0;
#pragma unroll
    for (int k = 0; k < Rec<L>::NUSED; ++k) out32[i * 32 + slot_of(L, k) - 1] = F[k];
#pragma unroll
    for (int k = 0; k < Rec<L>::NSIDE; ++k) out32[i * 32 + side_slot_of(L, k) - 1] = F[Rec<L>::NREC + k];
}

template <int L> constexpr int C_NC() { return (Rec<L>::NDIM == 3) ? 8 : 4; }

#if !GPAT_STRICT
// Size residency to the L2, not to the register file.  A lane re-reads the same NC records step after step, so the
// live working set is resident lanes x NC x record bytes; it is served by L2 only while it fits about half of it (the
// 126 MB L2 of a B200 is two partitions).  2-D records: 39-58 MB at full occupancy, no cap.  3-D Parker: 29 MB per
// resident CTA per SM -> 2 CTAs; measured on C5: 1/2/3/4 CTAs per SM = 2.61/3.45/3.17/2.84e9 steps/s
// (profiles/README.md).  GPAT_PUSH_MAXCTAS overrides.
template <int L>
int l2_cap(int sm_count, int per_sm, const PushArgs& a)
{
    static int l2_bytes = 0;
    if (!l2_bytes) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, dev);
        if (l2_bytes <= 0) l2_bytes = 64 << 20;
    }
    const double per_cta = (double)sm_count * kBlock * C_NC<L>() * ((2.0 * Rec<L>::NREC + side_floats_of(L)) * 4.0);
    int cap = (int)(0.5 * (double)l2_bytes / per_cta + 0.5);
    if (cap < 1) cap = 1;
    // Cell-sorted particles (sort.cu) share their records across the lanes of a warp: the working
    // set per lane shrinks and full occupancy wins again (C5 sorted: 2/3/4 CTAs = 4.8/5.6/5.7e9).
    if (a.sorted) cap = per_sm;
    if (const char* e = getenv("GPAT_PUSH_MAXCTAS")) cap = atoi(e) > 0 ? atoi(e) : cap;
    return per_sm > cap ? cap : per_sm;
}
#endif

#if !GPAT_STRICT && GPAT_TU_PART != 0
// The kSpecAlt instantiations (1-D, focused transport, turbulence maps; one per pusher, with or without the map gather):
// rows in dynamic shared memory (the map rows only when the run has maps), grid sized from the occupancy of the
// instantiation that runs.
template <int L, int SPEC>
void launch_alt_spec(const DevParams& prm, const PtlSoA& P, const float* fld, const PushArgs& a,
                     int sm_count, cudaStream_t st)
{
    using CA = Coop<L, spec_has_maps(SPEC)>;
    const size_t smem = (size_t)(kBlock / 32) * 32 * (CA::PAR + CA::ROW + CA::AROW) * sizeof(double);
    int per_sm = 0;
    auto set = [&](auto k) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); };
    set(push_kernel_coop<L, 0, false, SPEC>); set(push_kernel_coop<L, 1, false, SPEC>);
    set(push_kernel_coop<L, 0, true, SPEC>); set(push_kernel_coop<L, 1, true, SPEC>);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, push_kernel_coop<L, 0, false, SPEC>, kBlock, smem);
    if (per_sm < 1) per_sm = 1;
    per_sm = l2_cap<L>(sm_count, per_sm, a);
    const long long want = (a.nptl + kBlock - 1) / kBlock;
    long long grid = (long long)sm_count * per_sm;  // persistent: a multiple of the SM count
    if (want < grid) grid = want > 0 ? want : 1;
    auto go = [&](auto sel_c, auto trk_c) {
        push_kernel_coop<L, decltype(sel_c)::value, decltype(trk_c)::value, SPEC>
            <<<(unsigned)grid, kBlock, smem, st>>>(prm, P, fld, a);
    };
    auto by_trk = [&](auto sel_c) {
        if (a.trk.enabled) go(sel_c, std::true_type{});
        else go(sel_c, std::false_type{});
    };
    if (a.sel == 0) by_trk(std::integral_constant<int, 0>{});
    else by_trk(std::integral_constant<int, 1>{});
}

template <int L>
void launch_alt(const DevParams& prm, const PtlSoA& P, const float* fld, const PushArgs& a,
                int sm_count, cudaStream_t st)
{
    static_assert(Rec<L>::NF == Rec<L>::NREC, "the general pushers read one-plane records");
    const bool maps = a.aux && (prm.deltab_flag || prm.correlation_flag);
    auto with_maps = [&](auto path_c) {
        constexpr int PATH = decltype(path_c)::value;
        if (maps) launch_alt_spec<L, alt_spec(PATH, true)>(prm, P, fld, a, sm_count, st);
        else launch_alt_spec<L, alt_spec(PATH, false)>(prm, P, fld, a, sm_count, st);
    };
    if constexpr (Rec<L>::NDIM == 2) {   // 1-D runs live in the 2-D records
        if (prm.ndim == 1) { with_maps(std::integral_constant<int, PATH_1D>{}); return; }
    }
    if constexpr (Rec<L>::EXT) {         // focused transport reads the extended records only (abi.cu: pick_layout)
        if (prm.focused_transport) { with_maps(std::integral_constant<int, PATH_FT>{}); return; }
    }
    // Parker pushers: only runs with maps come here (without maps they are the named-config kernels of part 0)
    launch_alt_spec<L, alt_spec(PATH_PARKER, true)>(prm, P, fld, a, sm_count, st);
}
#endif

#if GPAT_STRICT || GPAT_TU_PART == 0
template <int L>
void launch_one(const DevParams& prm, const PtlSoA& P, const float* fld, const PushArgs& a,
                int sm_count, cudaStream_t st)
{
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, push_kernel<L>, kBlock, 0);
    if (per_sm < 1) per_sm = 1;
    long long want = (a.nptl + kBlock - 1) / kBlock;
    long long grid = (long long)sm_count * per_sm;  // persistent: a multiple of the SM count
    if (want < grid) grid = want > 0 ? want : 1;
#if !GPAT_STRICT
    if (a.variant == 1) {
        // GPAT_PUSH_SMEM_PAD: occupancy experiments (unused dynamic shared memory per CTA)
        size_t pad = 0;
        // 1-D, focused transport, turbulence maps: the general pushers behind the lane-group gather (one-plane records);
        // their instantiations live in translation units of their own (GPAT_TU_PART 1 and 2, see the Makefile)
        if constexpr (Rec<L>::NF == Rec<L>::NREC) {
            if (prm.ndim == 1 || prm.focused_transport || prm.deltab_flag || prm.correlation_flag) {
                if (Rec<L>::NDIM == 2) launch_push_alt2d(L, prm, P, fld, a, sm_count, st);
                else launch_push_alt3d(L, prm, P, fld, a, sm_count, st);
                return;
            }
        }
        if (const char* e = getenv("GPAT_PUSH_SMEM_PAD")) {
            pad = (size_t)atol(e);
            cudaFuncSetAttribute(push_kernel_coop<L, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
            cudaFuncSetAttribute(push_kernel_coop<L, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
        }
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, push_kernel_coop<L, 0>, kBlock, pad);
        if (per_sm < 1) per_sm = 1;
        per_sm = l2_cap<L>(sm_count, per_sm, a);
        grid = (long long)sm_count * per_sm;
        if (want < grid) grid = want > 0 ? want : 1;
        // tracking runs use their own instantiations: the production kernels carry no tracking code
        // switch-specialised instantiations (physics_fast) of the 2-D layouts: mag = mom = 1 (C1, C2,
        // C4) and, on the base record, mag = 0 (C3).  The 3-D kernels stay generic: at their 128-
        // register cap the specialised code spills more and measured 7 % SLOWER on C5
        // (profiles/README.md).  Tracking runs pick the same SPEC so that they replay the run their
        // particles were selected from with identical arithmetic.
        int spec = 0;
        if (!prm.nlgc && prm.rng_mode != GPAT_RNG_TABLE && !prm.check_drift_2d &&
            prm.acc_region_flag != 1 && prm.time_interp && !a.generic) {
            const int want = 1 | (prm.mag_dependency == 1 ? 2 : 0) | (prm.momentum_dependency == 1 ? 4 : 0);
            if (Rec<L>::NDIM == 2 && (want == kSpec11 || (L == L2B && want == kSpec01))) spec = want;
            if (L == L3D && want == kSpec10 && !prm.acc_by_surface && !getenv("GPAT_NO_SPEC3D")) spec = want;
        }
        auto go = [&](auto sel_c, auto trk_c, auto spec_c) {
            push_kernel_coop<L, decltype(sel_c)::value, decltype(trk_c)::value, decltype(spec_c)::value>
                <<<(unsigned)grid, kBlock, pad, st>>>(prm, P, fld, a);
        };
        auto by_spec = [&](auto sel_c, auto trk_c) {
            using I = std::integral_constant<int, 0>;
            if constexpr (Rec<L>::NDIM == 2) {
                if (spec == kSpec11) { go(sel_c, trk_c, std::integral_constant<int, kSpec11>{}); return; }
                if constexpr (L == L2B) {
                    if (spec == kSpec01) { go(sel_c, trk_c, std::integral_constant<int, kSpec01>{}); return; }
                }
            }
            if constexpr (Rec<L>::NDIM == 3) {  // run-time switches + the acceleration-surface gate
                if (prm.acc_by_surface) { go(sel_c, trk_c, std::integral_constant<int, kSpecSurf>{}); return; }
            }
            if constexpr (L == L3D) {
                if (spec == kSpec10) { go(sel_c, trk_c, std::integral_constant<int, kSpec10>{}); return; }
            }
            go(sel_c, trk_c, I{});
        };
        auto by_trk = [&](auto sel_c) {
            if (a.trk.enabled) by_spec(sel_c, std::true_type{});
            else by_spec(sel_c, std::false_type{});
        };
        if (a.sel == 0) by_trk(std::integral_constant<int, 0>{});
        else by_trk(std::integral_constant<int, 1>{});
        return;
    }
#endif
    if (a.trk.enabled) push_kernel<L, true><<<(unsigned)grid, kBlock, 0, st>>>(prm, P, fld, a);
    else push_kernel<L><<<(unsigned)grid, kBlock, 0, st>>>(prm, P, fld, a);
}
#endif

}  // namespace

#if !GPAT_STRICT && GPAT_TU_PART == 1
void launch_push_alt2d(int layout, const DevParams& prm, const PtlSoA& P, const float* fld,
                       const PushArgs& a, int sm_count, cudaStream_t st)
{
    if (layout == L2B) launch_alt<L2B>(prm, P, fld, a, sm_count, st);
    else launch_alt<L2E>(prm, P, fld, a, sm_count, st);
}
#elif !GPAT_STRICT && GPAT_TU_PART == 2
void launch_push_alt3d(int layout, const DevParams& prm, const PtlSoA& P, const float* fld,
                       const PushArgs& a, int sm_count, cudaStream_t st)
{
    if (layout == L3B) launch_alt<L3B>(prm, P, fld, a, sm_count, st);
    else launch_alt<L3E>(prm, P, fld, a, sm_count, st);
}
#else

#if GPAT_STRICT
#define GPAT_LAUNCH launch_push_strict
#else
#define GPAT_LAUNCH launch_push_fast
#endif

void GPAT_LAUNCH(int layout, const DevParams& prm, const PtlSoA& P, const float* fld,
                 const PushArgs& a, int sm_count, cudaStream_t st)
{
    switch (layout) {
        case L2B: launch_one<L2B>(prm, P, fld, a, sm_count, st); break;
        case L2E: launch_one<L2E>(prm, P, fld, a, sm_count, st); break;
        case L3B: launch_one<L3B>(prm, P, fld, a, sm_count, st); break;
#if !GPAT_STRICT
        case L2D: launch_one<L2D>(prm, P, fld, a, sm_count, st); break;  // production build only (abi.cu: pick_layout)
        case L3D: launch_one<L3D>(prm, P, fld, a, sm_count, st); break;
#endif
        default: launch_one<L3E>(prm, P, fld, a, sm_count, st); break;
    }
}

#if GPAT_STRICT
void launch_interp_debug(int layout, const DevParams& prm, const float* fld, int sel, long long n,
                         const double* x, const double* y, const double* z, const double* rt,
                         double* out32, cudaStream_t st)
{
    unsigned grid = (unsigned)((n + 127) / 128);
    if (grid == 0) return;
    switch (layout) {
        case L2B: interp_kernel<L2B><<<grid, 128, 0, st>>>(prm, fld, sel, n, x, y, z, rt, out32); break;
        case L2E: interp_kernel<L2E><<<grid, 128, 0, st>>>(prm, fld, sel, n, x, y, z, rt, out32); break;
        case L3B: interp_kernel<L3B><<<grid, 128, 0, st>>>(prm, fld, sel, n, x, y, z, rt, out32); break;
        case L2D: interp_kernel<L2D><<<grid, 128, 0, st>>>(prm, fld, sel, n, x, y, z, rt, out32); break;
        case L3D: interp_kernel<L3D><<<grid, 128, 0, st>>>(prm, fld, sel, n, x, y, z, rt, out32); break;
        default: interp_kernel<L3E><<<grid, 128, 0, st>>>(prm, fld, sel, n, x, y, z, rt, out32); break;
    }
}
#endif

#endif  // GPAT_TU_PART

}  // namespace gpat
