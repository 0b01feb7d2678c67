// gpat_internal.cuh -- device-side data layout shared by the kernels of libgpat_cuda.so.
//
// HBM layout (DESIGN.md "Data layout"):
//  * particles: structure-of-arrays, one contiguous device allocation, capacity nptl_max
//    (the reference's AoS particle_type, particle_module.f90:38-50, exists only at the
//    upload/download boundary);
//  * fields: one packed record per grid point holding ONLY the slots the configured
//    pusher reads (15 of the reference's 32 for 2-D Parker), FP32 exactly as the
//    reference stores them (mhd_data_parallel.f90:35).  A grid point owns 2*NREC floats
//    made of 32-byte chunks: chunk c = [slots 4c..4c+3 of half 0 | the same slots of half 1],
//    the two halves being the two MHD time frames (which half is farray1 flips at every
//    gpat_swap_fields).  One 128-byte line therefore holds both frames of a 2-D Parker
//    cell, and one 256-bit load brings both frames of four slots.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/gpat_cuda.h"

namespace gpat {

// ---- packed field record layouts ---------------------------------------------------
// Reference slot numbers (1-based, mhd_data_parallel.f90:77-81): 1 vx 2 vy 3 vz 4 rho 5 bx
// 6 by 7 bz 8 |B|; gradient of primary k along d (1..3) is slot 8 + 3(k-1) + d.
enum Layout : int { L2B = 0, L2E = 1, L3B = 2, L3E = 3, L2D = 4, L3D = 5 };

// NREC/NUSED: floats per frame in the record / slots in use; EXT: momentum-diffusion slots present;
// NSIDE: slots kept in the SIDE PLANE (see L2D); THIRD: 0 = no resolved z axis, 1 = always (3-D),
// 2 = decided at run time by include_3rd_dim; NF = doubles of an interpolated record (record + side slots).
template <int L> struct Rec;
// 2-D Parker without momentum diffusion: 15 slots (particle_module.f90:3392-3399, 3430-3433,
// 2352-2357) + 1 pad = 64 B per frame
template <> struct Rec<L2B> {
    static constexpr int NREC = 16, NUSED = 15, NDIM = 2, NSIDE = 0, SIDE_CHUNKS = 0, THIRD = 0, NF = NREC;
    static constexpr bool EXT = false;
};
// + vz (include_3rd_dim), rho (D_pp wave), dvx_dy dvy_dx dvz_dx dvz_dy (D_pp shear)
template <> struct Rec<L2E> {
    static constexpr int NREC = 24, NUSED = 21, NDIM = 2, NSIDE = 0, SIDE_CHUNKS = 0, THIRD = 2, NF = NREC;
    static constexpr bool EXT = true;
};
// 3-D Parker: 21 slots (particle_module.f90:4665-4670, 4686-4688, 2390-2401)
template <> struct Rec<L3B> {
    static constexpr int NREC = 24, NUSED = 21, NDIM = 3, NSIDE = 0, SIDE_CHUNKS = 0, THIRD = 1, NF = NREC;
    static constexpr bool EXT = false;
};
template <> struct Rec<L3E> {
    static constexpr int NREC = 32, NUSED = 28, NDIM = 3, NSIDE = 0, SIDE_CHUNKS = 0, THIRD = 1, NF = NREC;
    static constexpr bool EXT = true;
};
// 2-D Parker + momentum diffusion WITHOUT the third dimension (BASELINE config C4, production build):
// 18 slots.  The record is the L2B line with rho in its pad slot -- one 128-byte line per grid point,
// both frames, gathered by four lanes exactly like L2B -- and the two shear-only gradients dvx_dy,
// dvy_dx live in a side plane behind the record plane: one float4 per grid point
// [dvx_dy, dvy_dx of half 0 | the same of half 1].  144 B per grid point instead of L2E's 192 B, and
// the gather keeps L2B's four lanes per particle instead of L2E's two.
template <> struct Rec<L2D> {
    static constexpr int NREC = 16, NUSED = 16, NDIM = 2, NSIDE = 2, SIDE_CHUNKS = 0, THIRD = 0, NF = NREC + 4;
    static constexpr bool EXT = true;
};

// 3-D Parker, production build (BASELINE config C5): the 24-slot record split at the 128-byte line.  Record
// plane = the first 16 slots of L3B (one line per grid point, both frames, four lanes per particle like L2B);
// side plane = the remaining 5 slots (+3 pad) as two 32-byte chunks per grid point, same chunk format.  The
// 192-byte L3B record straddles lines and only splits two ways (six chunks), which costs twice the L1
// wavefronts per load instruction (profiles/r02b_push_coop_c5_ncu.txt: L1 data pipe 75 % busy).
// SIDE_CHUNKS: side plane in chunk format (0: L2D's float4 format).
template <> struct Rec<L3D> {
    static constexpr int NREC = 16, NUSED = 16, NDIM = 3, NSIDE = 5, SIDE_CHUNKS = 2, THIRD = 1, NF = NREC + 8;
    static constexpr bool EXT = false;
};

// packed position -> reference slot (1-based); 0 marks padding
__host__ __device__ constexpr int slot_of(int layout, int k)
{
    constexpr int l2[24] = {1, 2, 5, 6, 7, 9, 13, 21, 22, 24, 25, 27, 28, 30, 31, /*15*/ 3,
                            4, 10, 12, 15, 16, 0, 0, 0};
    constexpr int l3[32] = {1, 2, 3, 5, 6, 7, 9, 13, 17, 21, 22, 23, 24, 25, 26, 27, 28, 29,
                            30, 31, 32, /*21*/ 4, 10, 11, 12, 14, 15, 16, 0, 0, 0, 0};
    return (layout == L2B) ? (k < 15 ? l2[k] : 0)
         : (layout == L2D) ? (k < 15 ? l2[k] : 4)
         : (layout == L2E) ? l2[k]
         : (layout == L3B) ? (k < 21 ? l3[k] : 0)
         : (layout == L3D) ? (k < 16 ? l3[k] : 0)
                           : l3[k];
}
__host__ __device__ constexpr int nrec_of(int layout)
{
    return (layout == L2B || layout == L2D || layout == L3D) ? 16 : layout == L3E ? 32 : 24;
}
// side plane: floats per grid point (both frames) behind the record plane, and the reference slots it holds
__host__ __device__ constexpr int side_floats_of(int layout) { return layout == L2D ? 4 : layout == L3D ? 16 : 0; }
__host__ __device__ constexpr int side_slot_of(int layout, int i)
{
    constexpr int l3s[8] = {28, 29, 30, 31, 32, 0, 0, 0};  // dbz_dy dbz_dz db_dx db_dy db_dz (positions 16..20 of L3B)
    return layout == L2D ? (i == 0 ? 10 : i == 1 ? 12 : 0)  // dvx_dy, dvy_dx
         : layout == L3D ? (i < 8 ? l3s[i] : 0) : 0;
}

// named positions inside a record
namespace s2 {  // 2-D layouts
enum { vx = 0, vy, bx, by, bz, dvx_dx, dvy_dy, dbx_dx, dbx_dy, dby_dx, dby_dy, dbz_dx, dbz_dy,
       db_dx, db_dy, vz, rho, dvx_dy, dvy_dx, dvz_dx, dvz_dy };
}
namespace s2d {  // L2D: interpolated record F[NF]: the 15 base slots, rho, then the side slots
enum { rho = 15, dvx_dy = 16, dvy_dx = 17 };
}
namespace s3 {  // 3-D layouts
enum { vx = 0, vy, vz, bx, by, bz, dvx_dx, dvy_dy, dvz_dz, dbx_dx, dbx_dy, dbx_dz, dby_dx, dby_dy,
       dby_dz, dbz_dx, dbz_dy, dbz_dz, db_dx, db_dy, db_dz, rho, dvx_dy, dvx_dz, dvy_dx, dvy_dz,
       dvz_dx, dvz_dy };
}

// ---- particles (SoA) ---------------------------------------------------------------
struct PtlSoA {
    double *x, *y, *z, *p, *v, *mu, *weight, *t, *dt;
    unsigned long long* rng;  // per-particle Philox step counter
    int *origin, *nsteps_tracked, *nsteps_pushed, *tag_injected, *tag_splitted;
    signed char *split_times, *count_flag;
};

// ---- parameters as the kernels see them ----------------------------------------------
struct DevParams {
    // grid
    int ndim, nx, ny, nz, nxg, nyg, nzg;
    int nyg_src;  // rows of a host frame (1 in 1-D, where the device store has a second zero row)
    int time_interp;
    int pbc[3];
    double dx, dy, dz, xmin, ymin, zmin, xmax, ymax, zmax, lx, ly, lz;
    double ext[6];  // extended box xmin1,xmax1,ymin1,ymax1,zmin1,zmax1 (particle_module.f90:1528-1533)
    // physics
    double p0, pmin, pmax, gamma_turb, pindex, kpara0, kret, kperp_kpara;
    double gm2;      // gamma_turb - 2
    double gm2_3;    // (gamma_turb - 2) / 3
    double pidx_perp;  // (5 - gamma_turb) / 3
    double qdrift;   // dble(1.0 / (3*pcharge)) evaluated in FP32 (particle_module.f90:3436)
    double drift1, drift2, tau0;
    double p0_pow;   // p0**(2 - pindex)
    // reciprocals and products used by the production (non-strict) arithmetic
    double idx, idy, idz, ip0;
    double sqrt_kret, sqrt_1mkret;  // sqrt(kret), sqrt(1 - kret)
    double d1p0, d2p02;             // drift1*p0, drift2*p0**2
    double d1p0sq, d2p02sq;         // their squares
    double hd2min2, hd2min3;        // min((dx/2)^2,(dy/2)^2[,(dz/2)^2])
    double pfloor;                  // 0.25*p0
    double acc_region[6];
    int momentum_dependency, mag_dependency, acc_region_flag;
    int dpp_wave, dpp_shear, weak_scattering, check_drift_2d, include_3rd_dim, nlgc;
    int focused_transport;  // Cartesian push_particle_*_ft (production build: kSpecAlt instantiations)
    int deltab_flag, correlation_flag;  // turbulence maps (production build: kSpecAlt instantiations)
    // acceleration surfaces (3-D, reference-order build only): normal = sign * (axis + 1)
    int acc_by_surface, surface_norm1, surface_norm2, surface2_existed, is_intersection;
    int pcharge;
    double duu0;
    // rng
    unsigned int key0, key1;  // Philox key = (seed_lo, seed_hi + origin)
    int rng_mode;
    int mpi_rank;
};

// ---- particle tracking (particle_module.f90:144-150, 5825-5990) -----------------------------
struct TrackDev {
    int enabled;
    int ncols;                 // split_times_max + 2
    long long ntrack;          // nptl_tracking
    long long nsteps_max;      // nsteps_tracking_max
    const int* tags;           // tags_tracking(ncols, ntrack), column-major
    gpat_particle* rec;        // particles_tracked(nsteps_max, ntrack), column-major
};

#ifdef __CUDACC__
// findloc(tags(row, lo:hi), v, dim=1[, back]) relative to lo (1-based), 0 if absent
__device__ __forceinline__ long long trk_findloc(const TrackDev& t, int row, long long lo, long long hi,
                                                 int v, bool back)
{
    if (!back) { for (long long c = lo; c <= hi; ++c) if (t.tags[(row - 1) + (long long)t.ncols * (c - 1)] == v) return c - lo + 1; }
    else { for (long long c = hi; c >= lo; --c) if (t.tags[(row - 1) + (long long)t.ncols * (c - 1)] == v) return c - lo + 1; }
    return 0;
}

// is_particle_selected / locate_particle (particle_module.f90:5920-5990)
__device__ __forceinline__ bool trk_selected(const TrackDev& t, int origin, int tag_inj, int tag_spl,
                                             int nsplit, long long& lo, long long& hi)
{
    lo = hi = -1;
    if (nsplit > t.ncols - 2) return false;
    const long long i1 = trk_findloc(t, 1, 1, t.ntrack, origin, false);
    if (i1 <= 0) return false;
    const long long i2 = trk_findloc(t, 1, 1, t.ntrack, origin, true);
    long long i3 = trk_findloc(t, 2, i1, i2, abs(tag_inj), false);
    if (i3 <= 0) return false;
    long long i4 = trk_findloc(t, 2, i1, i2, abs(tag_inj), true);
    i3 += i1 - 1;
    i4 += i1 - 1;
    if (nsplit > 0) {
        const long long i5 = trk_findloc(t, nsplit + 2, i3, i4, abs(tag_spl), false);
        if (i5 <= 0) return false;
        const long long i6 = trk_findloc(t, nsplit + 2, i3, i4, abs(tag_spl), true);
        lo = i5 + i3 - 1;
        hi = i6 + i3 - 1;
    } else {
        lo = i3;
        hi = i4;
    }
    return true;
}

// particles_tracked(n, lo:hi) = ptl
__device__ __forceinline__ void trk_record(const TrackDev& t, const gpat_particle& q, long long lo, long long hi)
{
    const long long n = q.nsteps_tracked;
    if (n < 1 || n > t.nsteps_max || lo < 1) return;
    for (long long c = lo; c <= hi; ++c) t.rec[(n - 1) + t.nsteps_max * (c - 1)] = q;
}
#endif

struct PushArgs {
    double t0, dtf, dt_fine, dt_min, dt_max;
    double idtf;             // 1/dtf
    double dt_target_limit;  // dtf + dt_fine*0.1 (particle_module.f90:1596)
    int nsteps_interval;
    int debug_nsteps;        // >0: gpat_debug_push_n mode
    int sel;                 // which half of a record pair holds farray1
    int variant;             // fast build: 0 = one lane gathers its own particle, 1 = lane groups
    int generic;             // 1: never pick the switch-specialised instantiation (GPAT_PUSH_GENERIC=1)
    int sorted;              // 1: the particle arrays were cell-sorted for this push (sort.cu)
    long long nptl;
    unsigned long long* queue;   // work counter
    unsigned long long* steps;   // push_particle_* calls
    double* leak;                // [0] leak, [1] leak_negp
    const double* rng_table;
    long long rng_slots, rng_max_steps;
    TrackDev trk;                // particle tracking (enabled = 0: the production kernels)
    // turbulence maps sigma2_slab, sigma2_2d, lc_slab, lc_2d (mhd_data_parallel.f90:36-41): per grid
    // point four 32-byte chunks [value d/dx d/dy d/dz of half 0 | the same of half 1]
    const float* aux;
    // acc_surfaceK1/K2 (acc_region_surface.f90:12-13): surf[k] -> two halves of n1 x n2 doubles each
    const double* surf[2];
    int surf_n1[2], surf_n2[2];
};

// ---- stream-compaction scratch (particles.cu) ---------------------------------------------
struct ScanWork {
    unsigned* tile_counts;
    long long* tile_offsets;
};

// ---- histogram pass (diag.cu) ---------------------------------------------------------------
struct HistDev {
    int enabled, npbins, nmu, nrx, nry, nrz;
    double pmin_log, dp_log, dmu, dx_diag, dy_diag, dz_diag;
    double* data;
    const double* pthr;  // npbins+1 momentum thresholds (see DiagArgs::gthr)
};

struct DiagArgs {
    long long n;
    int local_dist;
    int nmu_g, npp_g;
    double pmin, pmax, pmin_log, dp_log, dmu;
    double xmin, ymin, zmin;
    double* fglobal;  // (nmu_g, npp_g) column-major
    // gthr[k], k = 0..npp_g: the smallest double p for which the HOST libm evaluates
    // floor((log10(p) - pmin_log)/dp_log) >= k.  Binning against these thresholds gives the
    // bin the reference's own expression (diagnostics.f90:777) yields on this host, bit for
    // bit, whatever the last-ulp behaviour of the device log10.
    const double* gthr;
    HistDev loc[4];
    double* sums;                // [0] sum weight [1] sum dt
    unsigned long long* minmax;  // [0] min dt bits [1] max dt bits [2] max p bits
};

// fescapedK_x/_y/_z of the four local sets (diagnostics.f90:358-405); null = set disabled / axis absent
struct EscLocalDev {
    double* fx[4];
    double* fy[4];
    double* fz[4];
};

// device counters (long long): [0] nptl_current [1] nptl_escaped [2] alive m [3] nholes
// [4] nfillers [5] escaped in this pass [6] split candidates
constexpr int kNumCounters = 8;

void launch_pack(const float* src, int nvar, int with_grad, const DevParams& prm, int layout,
                 float* dst, int half, int sm_count, cudaStream_t st);
void launch_fill(float* p, long long n, float v, int sm_count, cudaStream_t st);
void launch_pack_aux(const float* src2, const DevParams& prm, int which, float* aux, int half, int sm_count,
                     cudaStream_t st);
void launch_grad32(const float* src8, const DevParams& prm, float* out32, int sm_count,
                   cudaStream_t st);
void launch_inject(const DevParams& prm, const PtlSoA& P, long long n, long long start,
                   long long nptl_max, long long tag0, double dt, int dist_flag, double particle_v0,
                   double t_frame, double dt_mhd, const double box[6], double power_index,
                   cudaStream_t st, int mode = 0, double vmin = 0.0, int layout = 0,
                   const float* fld = nullptr, int sel = 0, int* fail = nullptr,
                   const TrackDev* trk = nullptr, const int* shock_x = nullptr, const float* aux = nullptr);
void launch_shock_xpos(const DevParams& prm, int layout, const float* fld, int half, int* d_out, cudaStream_t st);
void launch_ncells(const DevParams& prm, int layout, const float* fld, int sel, int mode, double vmin,
                   const double box[6], unsigned long long* d_count, int sm_count, cudaStream_t st,
                   const float* aux = nullptr);
void launch_remove(const PtlSoA& P, const PtlSoA& E, long long ecap, long long n, long long* counters,
                   const ScanWork& w, long long* idx_a, long long* idx_b, int dump_escaped,
                   cudaStream_t st);
void launch_final_bc(const DevParams& prm, const PtlSoA& P, long long nmax, const long long* n_dev,
                     double* leak, cudaStream_t st);
void launch_split(const DevParams& prm, const PtlSoA& P, long long n, long long nptl_max,
                  double split_ratio, double pmin_split, long long* counters, long long* nptl_split,
                  const ScanWork& w, long long* idx_a, cudaStream_t st, const TrackDev* trk = nullptr);
size_t sort_scratch_bytes(long long n);
cudaError_t launch_cell_sort(const DevParams& prm, const PtlSoA& S, const PtlSoA& D, long long n, unsigned* keys,
                             unsigned* idx, void* tmp, size_t tmp_bytes, cudaStream_t st);
void launch_to_aos(const PtlSoA& P, gpat_particle* out, long long n, cudaStream_t st);
void launch_from_aos(const PtlSoA& P, const gpat_particle* in, long long n, cudaStream_t st);
void launch_diag(const PtlSoA& P, const DiagArgs& a, int sm_count, cudaStream_t st);
void launch_escaped_local(const PtlSoA& E, long long n, const DiagArgs& a, const EscLocalDev& o, cudaStream_t st);
void launch_escaped_diag(const PtlSoA& E, long long n, const DiagArgs& a, int nface, double* fesc,
                         cudaStream_t st);
void launch_finalize_quick(const double* sums, const unsigned long long* minmax, const double* leak,
                           double nptl_current, double nptl_split, double* q9, cudaStream_t st);

void launch_push_fast(int layout, const DevParams& prm, const PtlSoA& P, const float* fld,
                      const PushArgs& a, int sm_count, cudaStream_t st);
void launch_push_strict(int layout, const DevParams& prm, const PtlSoA& P, const float* fld,
                        const PushArgs& a, int sm_count, cudaStream_t st);
// production build, translation units of their own: the kSpecAlt instantiations of the 2-D / 3-D one-plane records
void launch_push_alt2d(int layout, const DevParams& prm, const PtlSoA& P, const float* fld,
                       const PushArgs& a, int sm_count, cudaStream_t st);
void launch_push_alt3d(int layout, const DevParams& prm, const PtlSoA& P, const float* fld,
                       const PushArgs& a, int sm_count, cudaStream_t st);
void launch_interp_debug(int layout, const DevParams& prm, const float* fld, int sel, long long n,
                         const double* x, const double* y, const double* z, const double* rt,
                         double* out32, cudaStream_t st);

}  // namespace gpat
