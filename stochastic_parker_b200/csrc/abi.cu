// abi.cu -- the C ABI of libgpat_cuda.so (include/gpat_cuda.h): handle, memory, call sequencing.
//
// Host-side mirror of what the reference keeps in module state for this path
// (particle_module.f90:64-135, diagnostics.f90:32-75, mhd_data_parallel.f90:31-35).  One
// handle == one MPI rank of the reference == one GPU; all work is issued on the handle's own
// stream and every entry point is blocking on return, like the Fortran procedures it replaces.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gpat_internal.cuh"

using namespace gpat;

namespace {

// ---- NCCL through dlopen: no link-time dependency, and the process (torch, an MPI build)
// may already carry its own libnccl -----------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process?
        if (api.lib) break;
    }
    if (!api.lib) {
        const char* env = getenv("GPAT_NCCL_LIB");
        if (env) api.lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    for (const char* n : names) {
        if (api.lib) break;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!api.lib) return api;
#define GPAT_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name))
    GPAT_SYM(GetUniqueId, "ncclGetUniqueId");
    GPAT_SYM(CommInitRank, "ncclCommInitRank");
    GPAT_SYM(CommDestroy, "ncclCommDestroy");
    GPAT_SYM(AllReduce, "ncclAllReduce");
    GPAT_SYM(GroupStart, "ncclGroupStart");
    GPAT_SYM(GroupEnd, "ncclGroupEnd");
    GPAT_SYM(GetErrorString, "ncclGetErrorString");
#undef GPAT_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce &&
             api.GroupStart && api.GroupEnd;
    return api;
}

thread_local std::string g_init_error;

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct gpat_sim {
    gpat_params hp{};
    DevParams dp{};
    int device = 0, sm_count = 148;
    cudaStream_t st = nullptr;
    int layout = L2B;
    int sel = 0;
    int push_generic = 0;  // GPAT_PUSH_GENERIC=1: A/B switch for the specialised kernel
    int push_variant = 1;  // production kernel: lane-group gather (GPAT_PUSH_VARIANT=0: one lane per particle)
    bool have_field[2] = {false, false};
    long long nptl_max = 0;
    // particles
    void* ptl_mem = nullptr;
    PtlSoA P{};
    void* esc_mem = nullptr;
    PtlSoA E{};
    long long ecap = 0;
    // host mirrors of the module counters
    long long nptl_current = 0, nptl_split = 0, nptl_escaped = 0, tag_max = 0;
    double leak = 0.0, leak_negp = 0.0;
    // device scalars
    long long* d_counters = nullptr;      // kNumCounters
    long long* d_nptl_split = nullptr;
    double* d_leak = nullptr;             // [2]
    unsigned long long* d_queue = nullptr;  // [0] queue [1] steps
    // scan scratch
    ScanWork w{};
    long long *idx_a = nullptr, *idx_b = nullptr;
    // fields
    float* fld = nullptr;
    float* stage = nullptr;
    size_t stage_bytes = 0;
    // frame pipeline (gpat_prefetch_fields): the next frame's H2D copy runs on its own stream into a
    // second staging buffer while the push kernel owns the compute stream
    cudaStream_t st_copy = nullptr;
    cudaEvent_t ev_copy = nullptr;
    float* stage2 = nullptr;
    size_t stage2_bytes = 0;
    const float* pf_ptr = nullptr;  // host pointer of the frame in flight / landed in stage2
    int pf_nvar = 0;
    // particle tracking (gpat_init_tracking)
    TrackDev trk{};
    int* d_shock = nullptr;  // shock_xpos2 (gpat_inject_at_shock)
    // spatial ordering before a push (sort.cu): second particle buffer + sort scratch, lazily allocated
    int sorted_now = 0;      // the current particle order is the cell order of this interval's start
    int sort_mode = -1;      // GPAT_PUSH_SORT: 0 off, 1 on, unset: on when the field store exceeds the L2
    int l2_bytes = 0;
    void* ptl_mem2 = nullptr;
    PtlSoA P2{};
    unsigned* sort_keys = nullptr;  // 2n keys + 2n indices
    void* sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    long long sort_cap = 0;
    double* d_fesc_loc = nullptr;  // local escaped distributions (all sets, all faces)
    size_t fesc_loc_n = 0;
    double* d_surf[2] = {nullptr, nullptr};  // acceleration surfaces: two halves each
    int surf_n1[2] = {0, 0}, surf_n2[2] = {0, 0};
    bool have_surf[2][2] = {{false, false}, {false, false}};
    float* aux = nullptr;    // turbulence maps (gpat_upload_turbulence), 32 floats per grid point
    bool have_aux[2] = {false, false};
    int* d_tags = nullptr;
    gpat_particle* d_tracked = nullptr;
    const void* registered_host[2] = {nullptr, nullptr};
    size_t registered_bytes[2] = {0, 0};
    // histograms
    double* d_fglobal = nullptr;
    double* d_flocal[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t local_bins[4] = {0, 0, 0, 0};
    int nrx[4] = {0}, nry[4] = {0}, nrz[4] = {0};
    double* d_fesc = nullptr;
    double* d_pthr = nullptr;             // momentum-bin thresholds: global, then local 1..4
    size_t pthr_off[5] = {0, 0, 0, 0, 0};
    double* d_sums = nullptr;             // [2]
    unsigned long long* d_minmax = nullptr;  // [3]
    double* d_quick = nullptr;            // [9]
    // rng table
    double* d_table = nullptr;
    long long table_slots = 0, table_steps = 0;
    // aos staging
    gpat_particle* d_aos = nullptr;
    long long aos_cap = 0;
    // nccl
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // instrumentation
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    gpat_timings tm{};
    std::string err;
};

namespace {

int fail(gpat_sim* h, int code, const std::string& msg)
{
    if (h) h->err = msg;
    else g_init_error = msg;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(h, GPAT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

#define NC(call)                                                                               \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess)                                                                 \
            return fail(h, GPAT_ERR_NCCL, std::string(#call) + ": " +                           \
                        (nccl().GetErrorString ? nccl().GetErrorString(r_) : "nccl error"));    \
    } while (0)

int validate(const gpat_params* p, std::string& why)
{
    if (p->ndim < 1 || p->ndim > 3) { why = "ndim must be 1, 2 or 3"; return 1; }
    if (p->nx < 1 || p->ny < 1 || p->nz < 1) { why = "bad grid"; return 1; }
    if (p->ndim == 2 && p->nz != 1) { why = "2-D runs need nz = 1"; return 1; }
    if (p->ndim == 1 && (p->ny != 1 || p->nz != 1)) { why = "1-D runs need ny = nz = 1"; return 1; }
    if (p->ndim == 1 && p->mag_dependency) {
        // particle_module.f90:2274 / 2539 multiply by a db_dx the 1-D branch never assigns
        why = "1-D with mag_dependency = 1 reads an uninitialised db_dx in the reference (undefined there)";
        return 1;
    }
    if (p->focused_transport && p->ndim == 1) {
        // particle_module.f90:3225 assigns dv_dt, :3270 and :3304 read dx_dt
        why = "focused transport in 1-D: push_particle_1d_ft advances x with a dx_dt it never assigns (undefined in the reference)";
        return 1;
    }
    if (p->focused_transport && (p->ndim == 3 || p->include_3rd_dim) && p->rng_mode == GPAT_RNG_TABLE) {
        why = "push_particle_2d_include_3rd_ft / _3d_ft draw five uniforms per step; the uniform table holds four";
        return 1;
    }
    if (p->spherical_coord) { why = "spherical coordinates are outside the GPU path"; return 1; }
    if (p->nonuniform_grid) { why = "non-uniform grids are outside the GPU path"; return 1; }
    if (p->acc_by_surface) {
        auto bad = [](int n) { return n == 0 || n < -3 || n > 3; };
        if (p->ndim != 3) { why = "acc_by_surface needs ndim = 3 (only the 3-D pushers use the surfaces)"; return 1; }
        if (bad(p->surface_norm1) || (p->surface2_existed && bad(p->surface_norm2))) {
            why = "surface_norm must be +-1 (x), +-2 (y) or +-3 (z)";
            return 1;
        }
    }
    if (p->include_3rd_dim && p->ndim != 2) { why = "include_3rd_dim needs ndim = 2"; return 1; }
    if (p->npp_global < 1 || p->nmu_global < 1) { why = "npp_global/nmu_global must be >= 1"; return 1; }
    if (p->pcharge == 0) { why = "pcharge must be non-zero"; return 1; }
    for (int k = 0; k < 4; ++k) {
        const gpat_hist_spec& s = p->local[k];
        if (!s.enabled) continue;
        if (s.rx < 1 || s.ry < 1 || s.rz < 1 || s.npbins < 1 || s.nmu < 1) { why = "bad local histogram spec"; return 1; }
        int nrx = (p->nx + s.rx - 1) / s.rx, nry = (p->ny + s.ry - 1) / s.ry, nrz = (p->nz + s.rz - 1) / s.rz;
        // check_local_dist_configuration (diagnostics.f90:1976-2016)
        if (nrx * s.rx != p->nx || (p->ndim >= 2 && nry * s.ry != p->ny) || (p->ndim == 3 && nrz * s.rz != p->nz)) {
            why = "Wrong factor 'rx/ry/rz' for particle distribution";
            return 1;
        }
    }
    return 0;
}

bool Rec_has_rho(int layout) { return layout == L2E || layout == L3E || layout == L2D; }

int pick_layout(const gpat_params& p)
{
    // focused transport reads vz and all six in-plane velocity gradients (particle_module.f90:3764-3789)
    bool ext = p.dpp_wave || p.dpp_shear || (p.ndim == 2 && p.include_3rd_dim) || p.keep_rho || p.focused_transport;
    // production build, 2-D momentum diffusion without the third dimension (config C4): the L2B line with rho in
    // its pad slot + a side plane for the two shear-only gradients (gpat_internal.cuh, Rec<L2D>).  Everything the
    // general pushers serve (strict_math, 1-D, focused transport, turbulence maps) keeps the one-plane record L2E.
    if (p.ndim == 2 && (p.dpp_wave || p.dpp_shear) && !p.include_3rd_dim && !p.focused_transport && !p.strict_math &&
        !p.deltab_flag && !p.correlation_flag && !getenv("GPAT_NO_L2D"))
        return L2D;
    // production build, 3-D Parker without momentum diffusion (config C5): the L3B record split at the 128-byte line
    if (p.ndim == 3 && !ext && !p.strict_math && !p.deltab_flag && !p.correlation_flag && !getenv("GPAT_NO_L3D"))
        return L3D;
    // 1-D runs live in the 2-D record layouts (one physical row + one zero row, fill_dev_params)
    if (p.ndim <= 2) return ext ? L2E : L2B;
    return ext ? L3E : L3B;
}

void fill_dev_params(gpat_sim* h)
{
    const gpat_params& p = h->hp;
    DevParams& d = h->dp;
    d.ndim = p.ndim; d.nx = p.nx; d.ny = p.ny; d.nz = p.nz;
    d.nxg = p.nx + 4;
    // 1-D (farray(:, -1:nx+2, 1, 1), mhd_data_parallel.f90:78): the store gets a second, all-zero
    // row so that the bilinear gather with ry = 0 reads defined memory and adds exact zeros
    d.nyg = (p.ndim == 1) ? 2 : p.ny + 4;
    d.nyg_src = (p.ndim == 1) ? 1 : p.ny + 4;
    d.nzg = (p.ndim == 3) ? p.nz + 4 : 1;
    d.time_interp = p.time_interp ? 1 : 0;
    for (int i = 0; i < 3; ++i) d.pbc[i] = p.pbc[i];
    d.dx = p.dx; d.dy = p.dy; d.dz = p.dz;
    d.xmin = p.xmin; d.ymin = p.ymin; d.zmin = p.zmin;
    d.xmax = p.xmax; d.ymax = p.ymax; d.zmax = p.zmax;
    d.lx = p.lx; d.ly = p.ly; d.lz = p.lz;
    // particle_module.f90:1528-1533
    d.ext[0] = p.xmin - p.dx * 0.5; d.ext[1] = p.xmax + p.dx * 0.5;
    d.ext[2] = p.ymin - p.dy * 0.5; d.ext[3] = p.ymax + p.dy * 0.5;
    d.ext[4] = p.zmin - p.dz * 0.5; d.ext[5] = p.zmax + p.dz * 0.5;
    d.p0 = p.p0; d.pmin = p.pmin; d.pmax = p.pmax; d.gamma_turb = p.gamma_turb; d.pindex = p.pindex;
    d.kpara0 = p.kpara0; d.kret = p.kret; d.kperp_kpara = p.kperp_kpara;
    d.gm2 = p.gamma_turb - 2.0;
    d.gm2_3 = (p.gamma_turb - 2.0) / 3.0;
    d.pidx_perp = (5.0 - p.gamma_turb) / 3.0;
    d.qdrift = (double)(1.0f / (float)(3 * p.pcharge));  // FP32 quotient, particle_module.f90:3436
    d.drift1 = p.drift1; d.drift2 = p.drift2; d.tau0 = p.tau0;
    d.p0_pow = std::pow(p.p0, 2.0 - p.pindex);
    d.idx = 1.0 / p.dx; d.idy = 1.0 / p.dy; d.idz = 1.0 / p.dz; d.ip0 = 1.0 / p.p0;
    d.sqrt_kret = std::sqrt(p.kret);
    d.sqrt_1mkret = std::sqrt(1.0 - p.kret);
    d.d1p0 = p.drift1 * p.p0;
    d.d2p02 = p.drift2 * p.p0 * p.p0;
    d.d1p0sq = d.d1p0 * d.d1p0;
    d.d2p02sq = d.d2p02 * d.d2p02;
    {
        double hx = 0.5 * p.dx, hy = 0.5 * p.dy, hz = 0.5 * p.dz;
        d.hd2min2 = std::fmin(hx * hx, hy * hy);
        d.hd2min3 = std::fmin(d.hd2min2, hz * hz);
    }
    d.pfloor = 0.25 * p.p0;
    for (int i = 0; i < 6; ++i) d.acc_region[i] = p.acc_region[i];
    d.momentum_dependency = p.momentum_dependency; d.mag_dependency = p.mag_dependency;
    d.acc_region_flag = p.acc_region_flag;
    d.dpp_wave = p.dpp_wave; d.dpp_shear = p.dpp_shear; d.weak_scattering = p.weak_scattering;
    d.check_drift_2d = p.check_drift_2d; d.include_3rd_dim = p.include_3rd_dim; d.nlgc = p.nlgc;
    d.focused_transport = p.focused_transport; d.duu0 = p.duu0; d.pcharge = p.pcharge;
    d.deltab_flag = p.deltab_flag; d.correlation_flag = p.correlation_flag;
    d.acc_by_surface = p.acc_by_surface; d.surface_norm1 = p.surface_norm1; d.surface_norm2 = p.surface_norm2;
    d.surface2_existed = p.surface2_existed; d.is_intersection = p.is_intersection;
    d.key0 = (unsigned)p.seed;
    d.key1 = (unsigned)(p.seed >> 32);
    d.rng_mode = p.rng_mode;
    d.mpi_rank = p.mpi_rank;
}

size_t soa_bytes(long long n)
{
    size_t b = 0;
    b += 10 * align_up((size_t)n * 8, 256);
    b += 5 * align_up((size_t)n * 4, 256);
    b += 2 * align_up((size_t)n, 256);
    return b;
}

void carve_soa(void* mem, long long n, PtlSoA& P)
{
    char* c = static_cast<char*>(mem);
    auto take = [&](size_t bytes) { char* r = c; c += align_up(bytes, 256); return r; };
    P.x = (double*)take(n * 8); P.y = (double*)take(n * 8); P.z = (double*)take(n * 8);
    P.p = (double*)take(n * 8); P.v = (double*)take(n * 8); P.mu = (double*)take(n * 8);
    P.weight = (double*)take(n * 8); P.t = (double*)take(n * 8); P.dt = (double*)take(n * 8);
    P.rng = (unsigned long long*)take(n * 8);
    P.origin = (int*)take(n * 4); P.nsteps_tracked = (int*)take(n * 4);
    P.nsteps_pushed = (int*)take(n * 4); P.tag_injected = (int*)take(n * 4);
    P.tag_splitted = (int*)take(n * 4);
    P.split_times = (signed char*)take(n); P.count_flag = (signed char*)take(n);
}

size_t field_floats(const gpat_sim* h)
{
    // both halves, always; + the side plane of L2D
    return (size_t)h->dp.nxg * h->dp.nyg * h->dp.nzg * (nrec_of(h->layout) * 2 + side_floats_of(h->layout));
}

// The smallest positive double p with floor((log10(p) - pmin_log)/dp_log) >= k, found by
// bisection over the bit patterns with the host's libm (the libm the reference itself runs on).
double bin_threshold(double pmin_log, double dp_log, int k)
{
    auto f_ge = [&](double p) { return std::floor((std::log10(p) - pmin_log) / dp_log) >= (double)k; };
    uint64_t lo = 0x0010000000000000ull;  // smallest normal: f very negative
    uint64_t hi = 0x7fe0000000000000ull;  // ~9e307
    auto as_d = [](uint64_t b) { double d; memcpy(&d, &b, 8); return d; };
    if (f_ge(as_d(lo))) return as_d(lo);
    if (!f_ge(as_d(hi))) return HUGE_VAL;
    while (hi - lo > 1) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (f_ge(as_d(mid))) hi = mid; else lo = mid;
    }
    return as_d(hi);
}

int alloc_hists(gpat_sim* h)
{
    const gpat_params& p = h->hp;
    {
        std::vector<double> thr;
        auto add = [&](double pmin, double pmax, int nb) {
            double pmin_log = std::log10(pmin);
            double dp_log = (std::log10(pmax) - pmin_log) / nb;
            for (int k = 0; k <= nb; ++k) thr.push_back(bin_threshold(pmin_log, dp_log, k));
        };
        h->pthr_off[0] = 0;
        add(p.pmin, p.pmax, p.npp_global);
        for (int k = 0; k < 4; ++k) {
            h->pthr_off[k + 1] = thr.size();
            if (p.local[k].enabled) add(p.local[k].pmin, p.local[k].pmax, p.local[k].npbins);
        }
        if (h->d_pthr) cudaFree(h->d_pthr);
        h->d_pthr = nullptr;
        CU(cudaMalloc(&h->d_pthr, thr.size() * sizeof(double)));
        CU(cudaMemcpy(h->d_pthr, thr.data(), thr.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    for (int k = 0; k < 4; ++k) {
        if (h->d_flocal[k]) { cudaFree(h->d_flocal[k]); h->d_flocal[k] = nullptr; }
        h->local_bins[k] = 0;
        const gpat_hist_spec& s = p.local[k];
        if (!s.enabled) continue;
        h->nrx[k] = (p.nx + s.rx - 1) / s.rx;
        h->nry[k] = (p.ny + s.ry - 1) / s.ry;
        h->nrz[k] = (p.nz + s.rz - 1) / s.rz;
        h->local_bins[k] = (size_t)s.nmu * s.npbins * h->nrx[k] * h->nry[k] * h->nrz[k];
        CU(cudaMalloc(&h->d_flocal[k], h->local_bins[k] * sizeof(double)));
    }
    if (h->d_fglobal) cudaFree(h->d_fglobal);
    if (h->d_fesc) cudaFree(h->d_fesc);
    CU(cudaMalloc(&h->d_fglobal, (size_t)p.nmu_global * p.npp_global * sizeof(double)));
    CU(cudaMalloc(&h->d_fesc, (size_t)p.nmu_global * p.npp_global * 2 * p.ndim * sizeof(double)));
    return GPAT_OK;
}

void fill_diag_args(const gpat_sim* h, DiagArgs& a, int local_dist)
{
    const gpat_params& p = h->hp;
    a.n = h->nptl_current;
    a.local_dist = local_dist;
    a.nmu_g = p.nmu_global; a.npp_g = p.npp_global;
    a.pmin = p.pmin; a.pmax = p.pmax;
    // init_particle_distributions, diagnostics.f90:199-206
    a.pmin_log = std::log10(p.pmin);
    a.dp_log = (std::log10(p.pmax) - a.pmin_log) / p.npp_global;
    a.dmu = (double)(2.0f / (float)p.nmu_global);
    a.xmin = p.xmin; a.ymin = p.ymin; a.zmin = p.zmin;
    a.fglobal = h->d_fglobal;
    a.gthr = h->d_pthr + h->pthr_off[0];
    for (int k = 0; k < 4; ++k) {
        HistDev& d = a.loc[k];
        d.pthr = h->d_pthr + h->pthr_off[k + 1];
        const gpat_hist_spec& s = p.local[k];
        d.enabled = s.enabled && h->d_flocal[k];
        d.npbins = s.npbins; d.nmu = s.nmu; d.nrx = h->nrx[k]; d.nry = h->nry[k]; d.nrz = h->nrz[k];
        d.data = h->d_flocal[k];
        if (!d.enabled) continue;
        // init_local_particle_distributions, diagnostics.f90:268-278
        d.dx_diag = p.lx / d.nrx; d.dy_diag = p.ly / d.nry; d.dz_diag = p.lz / d.nrz;
        d.pmin_log = std::log10(s.pmin);
        d.dp_log = (std::log10(s.pmax) - d.pmin_log) / s.npbins;
        d.dmu = (double)(2.0f / (float)s.nmu);
    }
    a.sums = h->d_sums;
    a.minmax = h->d_minmax;
}

float elapsed(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int ensure_aos(gpat_sim* h, long long n)
{
    if (n <= h->aos_cap) return GPAT_OK;
    if (h->d_aos) cudaFree(h->d_aos);
    h->d_aos = nullptr;
    h->aos_cap = 0;
    CU(cudaMalloc(&h->d_aos, (size_t)n * sizeof(gpat_particle)));
    h->aos_cap = n;
    return GPAT_OK;
}

// escaped_ptls with the reference's growth rule (resize_escaped_particles, particle_module.f90:5329-5358,
// called from remove_particles :5372-5376): when the escapees so far plus every current particle would not
// fit, the array grows to max(int(1.25 * cap), cap + nptl_current) and keeps its content.
int ensure_escaped(gpat_sim* h)
{
    if (!h->esc_mem) {
        CU(cudaMalloc(&h->esc_mem, soa_bytes(h->nptl_max)));
        carve_soa(h->esc_mem, h->nptl_max, h->E);
        h->ecap = h->nptl_max;
    }
    if (h->nptl_escaped + h->nptl_current <= h->ecap) return GPAT_OK;
    long long ncap = (long long)(1.25 * (double)h->ecap);
    if (ncap < h->ecap + h->nptl_current) ncap = h->ecap + h->nptl_current;
    void* mem = nullptr;
    CU(cudaMalloc(&mem, soa_bytes(ncap)));
    PtlSoA N;
    carve_soa(mem, ncap, N);
    const long long keep = h->nptl_escaped < h->ecap ? h->nptl_escaped : h->ecap;
    if (keep > 0) {
        const size_t n8 = (size_t)keep * 8, n4 = (size_t)keep * 4, n1 = (size_t)keep;
        double* const src8[10] = {h->E.x, h->E.y, h->E.z, h->E.p, h->E.v, h->E.mu, h->E.weight, h->E.t, h->E.dt, (double*)h->E.rng};
        double* const dst8[10] = {N.x, N.y, N.z, N.p, N.v, N.mu, N.weight, N.t, N.dt, (double*)N.rng};
        for (int k = 0; k < 10; ++k) CU(cudaMemcpyAsync(dst8[k], src8[k], n8, cudaMemcpyDeviceToDevice, h->st));
        int* const src4[5] = {h->E.origin, h->E.nsteps_tracked, h->E.nsteps_pushed, h->E.tag_injected, h->E.tag_splitted};
        int* const dst4[5] = {N.origin, N.nsteps_tracked, N.nsteps_pushed, N.tag_injected, N.tag_splitted};
        for (int k = 0; k < 5; ++k) CU(cudaMemcpyAsync(dst4[k], src4[k], n4, cudaMemcpyDeviceToDevice, h->st));
        CU(cudaMemcpyAsync(N.split_times, h->E.split_times, n1, cudaMemcpyDeviceToDevice, h->st));
        CU(cudaMemcpyAsync(N.count_flag, h->E.count_flag, n1, cudaMemcpyDeviceToDevice, h->st));
    }
    CU(cudaStreamSynchronize(h->st));
    cudaFree(h->esc_mem);
    h->esc_mem = mem;
    h->E = N;
    h->ecap = ncap;
    return GPAT_OK;
}

int sync_counters(gpat_sim* h)
{
    long long c[kNumCounters];
    double lk[2];
    long long ns;
    CU(cudaMemcpyAsync(c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost, h->st));
    CU(cudaMemcpyAsync(lk, h->d_leak, sizeof(lk), cudaMemcpyDeviceToHost, h->st));
    CU(cudaMemcpyAsync(&ns, h->d_nptl_split, sizeof(ns), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    h->nptl_current = c[0];
    h->nptl_escaped = c[1];
    h->leak = lk[0];
    h->leak_negp = lk[1];
    h->nptl_split = ns;
    return GPAT_OK;
}

int push_counters(gpat_sim* h)
{
    long long c[kNumCounters] = {h->nptl_current, h->nptl_escaped, 0, 0, 0, 0, 0, 0};
    double lk[2] = {h->leak, h->leak_negp};
    CU(cudaMemcpyAsync(h->d_counters, c, sizeof(c), cudaMemcpyHostToDevice, h->st));
    CU(cudaMemcpyAsync(h->d_leak, lk, sizeof(lk), cudaMemcpyHostToDevice, h->st));
    CU(cudaMemcpyAsync(h->d_nptl_split, &h->nptl_split, sizeof(long long), cudaMemcpyHostToDevice, h->st));
    CU(cudaStreamSynchronize(h->st));
    return GPAT_OK;
}

// Spatial ordering of the particle arrays (sort.cu): production build only (GPAT_PUSH_SORT=0 turns it
// off).  Never in the reference-order build, whose tests check the reference's own particle order.
int sort_before_push(gpat_sim* h)
{
    const bool strict = h->hp.strict_math;
    // Default: sort when the packed field store is larger than the L2 (+4 % on C1/C2, +10 % on C4, 2x
    // on C5).  A store that is L2-resident as a whole has no locality left to gain, and clustering
    // particles with similar step counts into the same warps costs load balance (C3: -4.5 %).
    if (h->l2_bytes <= 0) {
        cudaDeviceGetAttribute(&h->l2_bytes, cudaDevAttrL2CacheSize, h->device);
        if (h->l2_bytes <= 0) h->l2_bytes = 64 << 20;
    }
    const bool want = (h->sort_mode == 1) ||
                      (h->sort_mode < 0 && field_floats(h) * sizeof(float) > (size_t)h->l2_bytes);
    const long long n = h->nptl_current;
    h->sorted_now = 0;
    if (strict || !want || n < 2 || n > 0x7fffffffLL) return GPAT_OK;
    if (!h->ptl_mem2) {
        CU(cudaMalloc(&h->ptl_mem2, soa_bytes(h->nptl_max)));
        CU(cudaMemsetAsync(h->ptl_mem2, 0, soa_bytes(h->nptl_max), h->st));
        carve_soa(h->ptl_mem2, h->nptl_max, h->P2);
    }
    if (n > h->sort_cap) {
        if (h->sort_keys) cudaFree(h->sort_keys);
        if (h->sort_tmp) cudaFree(h->sort_tmp);
        h->sort_keys = nullptr; h->sort_tmp = nullptr; h->sort_cap = 0;
        const long long cap = h->nptl_max;
        CU(cudaMalloc(&h->sort_keys, (size_t)cap * 4 * sizeof(unsigned)));
        h->sort_tmp_bytes = sort_scratch_bytes(cap);
        CU(cudaMalloc(&h->sort_tmp, h->sort_tmp_bytes));
        h->sort_cap = cap;
    }
    cudaError_t e = launch_cell_sort(h->dp, h->P, h->P2, n, h->sort_keys, h->sort_keys + 2 * h->sort_cap,
                                     h->sort_tmp, h->sort_tmp_bytes, h->st);
    if (e != cudaSuccess) return fail(h, GPAT_ERR_CUDA, std::string("particle sort: ") + cudaGetErrorString(e));
    h->tm.total_launches += 3;
    std::swap(h->ptl_mem, h->ptl_mem2);
    std::swap(h->P, h->P2);
    h->sorted_now = 1;
    return GPAT_OK;
}

int run_push(gpat_sim* h, double t0, double dtf, int nsteps_interval, int num_fine_steps,
             int debug_nsteps, uint64_t* steps_done)
{
    PushArgs a{};
    a.t0 = t0; a.dtf = dtf;
    a.idtf = 1.0 / dtf;
    a.dt_fine = dtf / num_fine_steps;
    a.dt_min = h->hp.dt_min_rel * dtf;  // set_dt_min_max, particle_module.f90:5519-5524
    a.dt_max = h->hp.dt_max_rel * dtf;
    a.dt_target_limit = dtf + a.dt_fine * (double)0.1f;  // particle_module.f90:1596
    a.nsteps_interval = nsteps_interval > 0 ? nsteps_interval : 1;
    a.debug_nsteps = debug_nsteps;
    a.sel = h->sel;
    a.variant = h->push_variant;
    a.generic = h->push_generic;
    a.sorted = h->sorted_now;
    a.nptl = h->nptl_current;
    a.queue = h->d_queue;
    a.steps = h->d_queue + 1;
    a.leak = h->d_leak;
    a.rng_table = h->d_table;
    a.rng_slots = h->table_slots;
    a.rng_max_steps = h->table_steps;
    a.trk = h->trk;
    a.aux = h->aux;
    for (int k = 0; k < 2; ++k) {
        a.surf[k] = h->d_surf[k];
        a.surf_n1[k] = h->surf_n1[k];
        a.surf_n2[k] = h->surf_n2[k];
    }
    if (h->hp.acc_by_surface) {
        const int nslot = h->dp.time_interp ? 2 : 1;
        for (int k = 0; k < (h->hp.surface2_existed ? 2 : 1); ++k)
            for (int sl = 0; sl < nslot; ++sl)
                if (!h->have_surf[k][sl])
                    return fail(h, GPAT_ERR_STATE, "the acceleration surfaces have not been uploaded (gpat_upload_acc_surface)");
    }
    if ((h->hp.deltab_flag && !h->have_aux[0]) || (h->hp.correlation_flag && !h->have_aux[1]))
        return fail(h, GPAT_ERR_STATE, "the deltab / correlation maps have not been uploaded (gpat_upload_turbulence)");
    CU(cudaMemsetAsync(h->d_queue, 0, 2 * sizeof(unsigned long long), h->st));
    CU(cudaEventRecord(h->ev[0], h->st));
    if (a.nptl > 0) {
        // 1-D (push_particle_1d), focused transport (push_particle_*_ft) and the turbulence maps (deltab /
        // correlation) run the general pushers behind the lane-group gather in the production build
        // (push.cu: kSpecAlt); GPAT_ALT_STRICT=1 sends them back to the reference-order kernels (A/B switch)
        const bool alt_strict = getenv("GPAT_ALT_STRICT") && atoi(getenv("GPAT_ALT_STRICT")) != 0;
        const bool alt = h->hp.ndim == 1 || h->hp.focused_transport || h->hp.deltab_flag || h->hp.correlation_flag;
        if (h->hp.strict_math || (alt && (alt_strict || h->push_variant != 1)))
            launch_push_strict(h->layout, h->dp, h->P, h->fld, a, h->sm_count, h->st);
        else launch_push_fast(h->layout, h->dp, h->P, h->fld, a, h->sm_count, h->st);
        h->tm.total_launches++;
    }
    CU(cudaEventRecord(h->ev[1], h->st));
    CU(cudaGetLastError());
    unsigned long long q[2];
    CU(cudaMemcpyAsync(q, h->d_queue, sizeof(q), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    h->tm.push_ms = elapsed(h->ev[0], h->ev[1]);
    h->tm.push_steps = q[1];
    h->tm.push_launches = a.nptl > 0 ? 1 : 0;
    if (steps_done) *steps_done = q[1];
    return GPAT_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* gpat_last_error(gpat_handle h) { return h ? h->err.c_str() : g_init_error.c_str(); }

int gpat_init(gpat_handle* out, int device, int64_t nptl_max, const gpat_params* params)
{
    gpat_sim* h = nullptr;
    if (!out || !params || nptl_max < 1) return fail(nullptr, GPAT_ERR_INVALID, "gpat_init: bad arguments");
    *out = nullptr;
    std::string why;
    if (validate(params, why)) return fail(nullptr, GPAT_ERR_INVALID, "gpat_init: " + why);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, GPAT_ERR_CUDA,
                    std::string("gpat_init: no CUDA device (") + cudaGetErrorString(e) +
                        "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, GPAT_ERR_INVALID, "gpat_init: bad device index");
    h = new gpat_sim();
    auto bail = [&](int code) { std::string m = h->err; gpat_finalize(h); g_init_error = m; return code; };
#define CUI(call)                                                                       \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                \
            return bail(GPAT_ERR_CUDA);                                                 \
        }                                                                               \
    } while (0)
    h->device = device;
    CUI(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUI(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    CUI(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    for (auto& ev : h->ev) CUI(cudaEventCreate(&ev));
    h->hp = *params;
    h->nptl_max = nptl_max;
    h->layout = pick_layout(h->hp);
    if (const char* v = getenv("GPAT_PUSH_VARIANT")) h->push_variant = atoi(v);
    if (const char* v = getenv("GPAT_PUSH_GENERIC")) h->push_generic = atoi(v);
    if (const char* v = getenv("GPAT_PUSH_SORT")) h->sort_mode = atoi(v);
    fill_dev_params(h);
    CUI(cudaMalloc(&h->ptl_mem, soa_bytes(nptl_max)));
    CUI(cudaMemsetAsync(h->ptl_mem, 0, soa_bytes(nptl_max), h->st));  // init_particles zero fill
    carve_soa(h->ptl_mem, nptl_max, h->P);
    CUI(cudaMalloc(&h->d_counters, kNumCounters * sizeof(long long)));
    CUI(cudaMalloc(&h->d_nptl_split, sizeof(long long)));
    CUI(cudaMalloc(&h->d_leak, 2 * sizeof(double)));
    CUI(cudaMalloc(&h->d_queue, 2 * sizeof(unsigned long long)));
    CUI(cudaMalloc(&h->d_sums, 2 * sizeof(double)));
    CUI(cudaMalloc(&h->d_minmax, 3 * sizeof(unsigned long long)));
    CUI(cudaMalloc(&h->d_quick, 9 * sizeof(double)));
    CUI(cudaMemsetAsync(h->d_counters, 0, kNumCounters * sizeof(long long), h->st));
    CUI(cudaMemsetAsync(h->d_nptl_split, 0, sizeof(long long), h->st));
    CUI(cudaMemsetAsync(h->d_leak, 0, 2 * sizeof(double), h->st));
    long long ntiles = (nptl_max + 1023) / 1024;
    CUI(cudaMalloc(&h->w.tile_counts, ntiles * sizeof(unsigned)));
    CUI(cudaMalloc(&h->w.tile_offsets, ntiles * sizeof(long long)));
    CUI(cudaMalloc(&h->idx_a, nptl_max * sizeof(long long)));
    CUI(cudaMalloc(&h->idx_b, nptl_max * sizeof(long long)));
    CUI(cudaMalloc(&h->fld, field_floats(h) * sizeof(float)));
    CUI(cudaMemsetAsync(h->fld, 0, field_floats(h) * sizeof(float), h->st));  // farray = 0.0
    if (alloc_hists(h) != GPAT_OK) return bail(GPAT_ERR_CUDA);
    CUI(cudaStreamSynchronize(h->st));
#undef CUI
    *out = h;
    return GPAT_OK;
}

int gpat_set_params(gpat_handle h, const gpat_params* params)
{
    if (!h || !params) return fail(h, GPAT_ERR_INVALID, "gpat_set_params: bad arguments");
    std::string why;
    if (validate(params, why)) return fail(h, GPAT_ERR_INVALID, "gpat_set_params: " + why);
    const gpat_params& o = h->hp;
    if (params->ndim != o.ndim || params->nx != o.nx || params->ny != o.ny || params->nz != o.nz ||
        (params->time_interp != 0) != (o.time_interp != 0))
        return fail(h, GPAT_ERR_INVALID, "gpat_set_params: the grid shape cannot change after gpat_init");
    if (pick_layout(*params) != h->layout)
        return fail(h, GPAT_ERR_INVALID,
                    "gpat_set_params: dpp/include_3rd switches change the field record layout; re-init");
    bool hist_changed = params->npp_global != o.npp_global || params->nmu_global != o.nmu_global ||
                        memcmp(params->local, o.local, sizeof(o.local)) != 0;
    h->hp = *params;
    fill_dev_params(h);
    if (hist_changed) return alloc_hists(h);
    return GPAT_OK;
}

int gpat_finalize(gpat_handle h)
{
    if (!h) return GPAT_OK;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    if (h->comm && nccl().ok) nccl().CommDestroy(h->comm);
    for (int i = 0; i < 2; ++i)
        if (h->registered_host[i]) cudaHostUnregister(const_cast<void*>(h->registered_host[i]));
    void* ptrs[] = {h->ptl_mem, h->esc_mem, h->d_counters, h->d_nptl_split, h->d_leak, h->d_queue,
                    h->w.tile_counts, h->w.tile_offsets, h->idx_a, h->idx_b, h->fld, h->stage, h->stage2,
                    h->d_tags, h->d_tracked, h->d_shock, h->aux, h->ptl_mem2, h->sort_keys, h->sort_tmp,
                    h->d_surf[0], h->d_surf[1], h->d_fesc_loc,
                    h->d_fglobal, h->d_flocal[0], h->d_flocal[1], h->d_flocal[2], h->d_flocal[3],
                    h->d_fesc, h->d_pthr, h->d_sums, h->d_minmax, h->d_quick, h->d_table, h->d_aos};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (auto& ev : h->ev)
        if (ev) cudaEventDestroy(ev);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->st_copy) { cudaStreamSynchronize(h->st_copy); cudaStreamDestroy(h->st_copy); }
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return GPAT_OK;
}

int gpat_upload_fields(gpat_handle h, int slot, const float* f, int nvar, int with_grad)
{
    if (!h || !f) return fail(h, GPAT_ERR_INVALID, "gpat_upload_fields: bad arguments");
    if (nvar != 8 && nvar != 32) return fail(h, GPAT_ERR_INVALID, "gpat_upload_fields: nvar must be 8 or 32");
    if (with_grad && nvar != 32) return fail(h, GPAT_ERR_INVALID, "gpat_upload_fields: with_grad needs nvar = 32");
    if (slot != 0 && slot != 1) return fail(h, GPAT_ERR_INVALID, "gpat_upload_fields: slot must be 0 or 1");
    if (slot == 1 && !h->dp.time_interp)
        return fail(h, GPAT_ERR_INVALID, "gpat_upload_fields: slot 1 (farray2) exists only with time_interp = 1");
    CU(cudaSetDevice(h->device));
    size_t bytes = (size_t)h->dp.nxg * h->dp.nyg_src * h->dp.nzg * nvar * sizeof(float);
    const float* src;
    CU(cudaEventRecord(h->ev[2], h->st));
    if (h->pf_ptr == f && h->pf_nvar == nvar) {
        // the frame was (or is being) copied by gpat_prefetch_fields: only the pack kernel is left
        CU(cudaStreamWaitEvent(h->st, h->ev_copy, 0));
        src = h->stage2;
        h->pf_ptr = nullptr;
    } else {
        // a prefetch is good for the NEXT upload only: an upload of anything else drops it, so that a frame whose
        // upload was skipped (error path, quota break) can never be packed later from a stale device copy
        h->pf_ptr = nullptr;
        if (bytes > h->stage_bytes) {
            if (h->stage) cudaFree(h->stage);
            h->stage = nullptr;
            h->stage_bytes = 0;
            CU(cudaMalloc(&h->stage, bytes));
            h->stage_bytes = bytes;
        }
        CU(cudaMemcpyAsync(h->stage, f, bytes, cudaMemcpyHostToDevice, h->st));
        src = h->stage;
    }
    CU(cudaEventRecord(h->ev[3], h->st));
    int half = (slot == 0) ? h->sel : (h->sel ^ 1);
    launch_pack(src, nvar, with_grad, h->dp, h->layout, h->fld, half, h->sm_count, h->st);
    h->tm.total_launches++;
    CU(cudaEventRecord(h->ev[4], h->st));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->st));
    h->tm.upload_ms = elapsed(h->ev[2], h->ev[4]);
    h->tm.grad_ms = elapsed(h->ev[3], h->ev[4]);
    h->have_field[slot] = true;
    return GPAT_OK;
}

int gpat_upload_turbulence(gpat_handle h, int which, int slot, const float* data)
{
    if (!h || !data || (which != 0 && which != 1) || (slot != 0 && slot != 1))
        return fail(h, GPAT_ERR_INVALID, "gpat_upload_turbulence: bad arguments");
    if (slot == 1 && !h->dp.time_interp)
        return fail(h, GPAT_ERR_INVALID, "gpat_upload_turbulence: slot 1 exists only with time_interp = 1");
    CU(cudaSetDevice(h->device));
    const size_t ncell_dev = (size_t)h->dp.nxg * h->dp.nyg * h->dp.nzg;
    if (!h->aux) {
        CU(cudaMalloc(&h->aux, ncell_dev * 32 * sizeof(float)));
        launch_fill(h->aux, (long long)(ncell_dev * 32), 1.0f, h->sm_count, h->st);  // sigma2 = lc = 1.0, :125, :165
    }
    const size_t bytes = (size_t)h->dp.nxg * h->dp.nyg_src * h->dp.nzg * 2 * sizeof(float);
    if (bytes > h->stage_bytes) {
        if (h->stage) cudaFree(h->stage);
        h->stage = nullptr;
        h->stage_bytes = 0;
        CU(cudaMalloc(&h->stage, bytes));
        h->stage_bytes = bytes;
    }
    CU(cudaMemcpyAsync(h->stage, data, bytes, cudaMemcpyHostToDevice, h->st));
    const int half = (slot == 0) ? h->sel : (h->sel ^ 1);
    launch_pack_aux(h->stage, h->dp, which, h->aux, half, h->sm_count, h->st);
    h->tm.total_launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->st));
    h->have_aux[which] = true;
    return GPAT_OK;
}

int gpat_upload_acc_surface(gpat_handle h, int which, int slot, const double* heights)
{
    if (!h || !heights || (which != 0 && which != 1) || (slot != 0 && slot != 1))
        return fail(h, GPAT_ERR_INVALID, "gpat_upload_acc_surface: bad arguments");
    if (!h->hp.acc_by_surface) return fail(h, GPAT_ERR_STATE, "gpat_upload_acc_surface: acc_by_surface is 0");
    if (which == 1 && !h->hp.surface2_existed)
        return fail(h, GPAT_ERR_INVALID, "gpat_upload_acc_surface: surface2_existed is false");
    if (slot == 1 && !h->dp.time_interp)
        return fail(h, GPAT_ERR_INVALID, "gpat_upload_acc_surface: slot 1 exists only with time_interp = 1");
    CU(cudaSetDevice(h->device));
    const int axis = std::abs(which ? h->hp.surface_norm2 : h->hp.surface_norm1) - 1;
    const int n1 = (axis == 0) ? h->hp.ny + 4 : h->hp.nx + 4;       // acc_region_surface.f90:33-39
    const int n2 = (axis == 2) ? h->hp.ny + 4 : h->hp.nz + 4;
    const size_t n = (size_t)n1 * n2;
    if (!h->d_surf[which]) {
        CU(cudaMalloc(&h->d_surf[which], 2 * n * sizeof(double)));
        CU(cudaMemsetAsync(h->d_surf[which], 0, 2 * n * sizeof(double), h->st));  // acc_surface = 0.0
        h->surf_n1[which] = n1;
        h->surf_n2[which] = n2;
    }
    const int half = h->dp.time_interp ? ((slot == 0) ? h->sel : (h->sel ^ 1)) : 0;
    CU(cudaMemcpyAsync(h->d_surf[which] + (size_t)half * n, heights, n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CU(cudaStreamSynchronize(h->st));
    h->have_surf[which][slot] = true;
    return GPAT_OK;
}

int gpat_prefetch_fields(gpat_handle h, const float* f, int nvar)
{
    if (!h || !f) return fail(h, GPAT_ERR_INVALID, "gpat_prefetch_fields: bad arguments");
    if (nvar != 8 && nvar != 32) return fail(h, GPAT_ERR_INVALID, "gpat_prefetch_fields: nvar must be 8 or 32");
    CU(cudaSetDevice(h->device));
    if (!h->st_copy) {
        CU(cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
    }
    size_t bytes = (size_t)h->dp.nxg * h->dp.nyg_src * h->dp.nzg * nvar * sizeof(float);
    if (bytes > h->stage2_bytes) {
        CU(cudaStreamSynchronize(h->st_copy));
        if (h->stage2) cudaFree(h->stage2);
        h->stage2 = nullptr;
        h->stage2_bytes = 0;
        CU(cudaMalloc(&h->stage2, bytes));
        h->stage2_bytes = bytes;
    }
    // stage2 may still feed the pack kernel of the previous frame on the compute stream
    CU(cudaEventRecord(h->ev[6], h->st));
    CU(cudaStreamWaitEvent(h->st_copy, h->ev[6], 0));
    CU(cudaMemcpyAsync(h->stage2, f, bytes, cudaMemcpyHostToDevice, h->st_copy));
    CU(cudaEventRecord(h->ev_copy, h->st_copy));
    h->pf_ptr = f;
    h->pf_nvar = nvar;
    return GPAT_OK;  // no synchronisation: the copy overlaps whatever the caller launches next
}

int gpat_swap_fields(gpat_handle h)
{
    if (!h) return GPAT_ERR_INVALID;
    if (!h->dp.time_interp) return GPAT_OK;  // copy_fields is only called with time_interp_flag == 1
    if (!h->have_field[1]) return fail(h, GPAT_ERR_STATE, "gpat_swap_fields: farray2 was never uploaded");
    h->sel ^= 1;  // farray1 = farray2 without moving a byte
    h->have_field[0] = true;
    return GPAT_OK;
}

int gpat_inject_uniform(gpat_handle h, int64_t nptl, double dt, int dist_flag, double particle_v0,
                        double t_frame, double dt_mhd, const double part_box[6], double power_index)
{
    if (!h || !part_box || nptl < 0) return fail(h, GPAT_ERR_INVALID, "gpat_inject_uniform: bad arguments");
    if (dist_flag < 0 || dist_flag > 2) return fail(h, GPAT_ERR_INVALID, "gpat_inject_uniform: dist_flag must be 0, 1 or 2");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev[2], h->st));
    launch_inject(h->dp, h->P, nptl, h->nptl_current, h->nptl_max, h->tag_max, dt, dist_flag,
                  particle_v0, t_frame, dt_mhd, part_box, power_index, h->st, 0, 0.0, h->layout, h->fld, h->sel,
                  nullptr, &h->trk);
    if (nptl > 0) h->tm.total_launches++;
    CU(cudaEventRecord(h->ev[3], h->st));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->st));
    h->tm.inject_ms = elapsed(h->ev[2], h->ev[3]);
    h->nptl_current += nptl;
    if (h->nptl_current > h->nptl_max) h->nptl_current = h->nptl_max;  // particle_module.f90:491-492
    h->tag_max += nptl;
    return GPAT_OK;
}

int gpat_inject_targeted(gpat_handle h, int mode, int64_t nptl, double dt, int dist_flag, double particle_v0,
                         double t_frame, double dt_mhd, const double part_box[6], double power_index,
                         int inject_same_nptl, double vmin, int64_t ncells_norm, int64_t* nptl_injected,
                         int64_t* ncells_out)
{
    if (!h || !part_box || nptl < 0) return fail(h, GPAT_ERR_INVALID, "gpat_inject_targeted: bad arguments");
    if (dist_flag < 0 || dist_flag > 2) return fail(h, GPAT_ERR_INVALID, "gpat_inject_targeted: dist_flag must be 0, 1 or 2");
    if (mode == GPAT_INJECT_LARGE_DB2 && (!h->aux || !h->have_aux[0]))
        return fail(h, GPAT_ERR_STATE, "gpat_inject_targeted: inject_large_db2 needs the deltab maps (gpat_upload_turbulence)");
    if (mode != GPAT_INJECT_LARGE_JZ && mode != GPAT_INJECT_LARGE_ABSJ && mode != GPAT_INJECT_LARGE_DIVV &&
        mode != GPAT_INJECT_LARGE_RHO && mode != GPAT_INJECT_LARGE_DB2)
        return fail(h, GPAT_ERR_INVALID, "gpat_inject_targeted: unknown mode");
    if (mode == GPAT_INJECT_LARGE_RHO && !Rec_has_rho(h->layout))
        return fail(h, GPAT_ERR_INVALID, "gpat_inject_targeted: inject_large_rho needs gpat_params.keep_rho = 1");
    if (!h->have_field[0]) return fail(h, GPAT_ERR_STATE, "gpat_inject_targeted: farray1 has not been uploaded");
    if (!inject_same_nptl && ncells_norm <= 0)
        return fail(h, GPAT_ERR_INVALID, "gpat_inject_targeted: ncells_norm must be positive");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev[2], h->st));
    // d_queue doubles as the cell counter and the non-convergence flag (it is idle between pushes)
    launch_ncells(h->dp, h->layout, h->fld, h->sel, mode, vmin, part_box, h->d_queue, h->sm_count, h->st, h->aux);
    h->tm.total_launches++;
    unsigned long long ncells = 0;
    CU(cudaMemcpyAsync(&ncells, h->d_queue, sizeof(ncells), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    const double denom = inject_same_nptl ? (double)ncells : (double)ncells_norm;
    // int(nptl * mpi_sub_size * (dble(ncells) / dble(denominator))), particle_module.f90:843-853
    long long ninj = (denom > 0.0) ? (long long)((double)nptl * ((double)ncells / denom)) : 0;
    if (ncells_out) *ncells_out = (int64_t)ncells;
    if (nptl_injected) *nptl_injected = ninj;
    if (ninj > 0) {
        int* d_fail = reinterpret_cast<int*>(h->d_queue);
        CU(cudaMemsetAsync(d_fail, 0, sizeof(int), h->st));
        launch_inject(h->dp, h->P, ninj, h->nptl_current, h->nptl_max, h->tag_max, dt, dist_flag, particle_v0,
                      t_frame, dt_mhd, part_box, power_index, h->st, mode, vmin, h->layout, h->fld, h->sel, d_fail,
                      &h->trk, nullptr, h->aux);
        h->tm.total_launches++;
        int failed = 0;
        CU(cudaEventRecord(h->ev[3], h->st));
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(&failed, d_fail, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        h->tm.inject_ms = elapsed(h->ev[2], h->ev[3]);
        h->nptl_current += ninj;
        if (h->nptl_current > h->nptl_max) h->nptl_current = h->nptl_max;
        h->tag_max += ninj;
        if (failed)
            return fail(h, GPAT_ERR_STATE, "gpat_inject_targeted: a rejection loop did not find an admissible "
                                           "position in 2^22 draws (the reference would spin forever)");
    }
    return GPAT_OK;
}

int gpat_inject_at_shock(gpat_handle h, int64_t nptl, double dt, int dist_flag, double particle_v0, double t_frame,
                         double power_index)
{
    if (!h || nptl < 0) return fail(h, GPAT_ERR_INVALID, "gpat_inject_at_shock: bad arguments");
    if (dist_flag < 0 || dist_flag > 2) return fail(h, GPAT_ERR_INVALID, "gpat_inject_at_shock: dist_flag must be 0, 1 or 2");
    if (!h->dp.time_interp || !h->have_field[1])
        return fail(h, GPAT_ERR_STATE, "gpat_inject_at_shock: needs time_interp = 1 and farray2 (shock_xpos2 is what "
                                       "interp_shock_location reads at rt = 0)");
    CU(cudaSetDevice(h->device));
    const int nyr = (h->dp.ndim > 1) ? h->dp.ny + 4 : 1, nzr = (h->dp.ndim > 2) ? h->dp.nz + 4 : 1;
    if (!h->d_shock) CU(cudaMalloc(&h->d_shock, (size_t)nyr * nzr * sizeof(int)));
    CU(cudaEventRecord(h->ev[2], h->st));
    launch_shock_xpos(h->dp, h->layout, h->fld, h->sel ^ 1, h->d_shock, h->st);  // farray2
    const double box[6] = {0, 0, 0, 0, 0, 0};
    launch_inject(h->dp, h->P, nptl, h->nptl_current, h->nptl_max, h->tag_max, dt, dist_flag, particle_v0, t_frame,
                  0.0, box, power_index, h->st, GPAT_INJECT_AT_SHOCK, 0.0, h->layout, h->fld, h->sel, nullptr, &h->trk,
                  h->d_shock);
    h->tm.total_launches += 2;
    CU(cudaEventRecord(h->ev[3], h->st));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->st));
    h->tm.inject_ms = elapsed(h->ev[2], h->ev[3]);
    h->nptl_current += nptl;
    if (h->nptl_current > h->nptl_max) h->nptl_current = h->nptl_max;
    h->tag_max += nptl;
    return GPAT_OK;
}

int gpat_init_tracking(gpat_handle h, const int32_t* tags, int ncols, int64_t nptl_tracking, int nsteps_interval)
{
    if (!h || !tags || ncols < 2 || nptl_tracking < 1 || nsteps_interval < 1)
        return fail(h, GPAT_ERR_INVALID, "gpat_init_tracking: bad arguments");
    CU(cudaSetDevice(h->device));
    if (h->d_tags) { cudaFree(h->d_tags); h->d_tags = nullptr; }
    if (h->d_tracked) { cudaFree(h->d_tracked); h->d_tracked = nullptr; }
    h->trk = TrackDev{};
    // nsteps_tracking_max = ceiling((1.0 / dt_min_rel) / nsteps_interval) + 1, particle_module.f90:5876
    const long long nmax = (long long)std::ceil((1.0 / h->hp.dt_min_rel) / nsteps_interval) + 1;
    const size_t rec_bytes = (size_t)nmax * (size_t)nptl_tracking * sizeof(gpat_particle);
    CU(cudaMalloc(&h->d_tags, (size_t)ncols * nptl_tracking * sizeof(int)));
    cudaError_t e = cudaMalloc(&h->d_tracked, rec_bytes);
    if (e != cudaSuccess)
        return fail(h, GPAT_ERR_CUDA, "gpat_init_tracking: particles_tracked(" + std::to_string(nmax) + ", " +
                                          std::to_string((long long)nptl_tracking) + ") does not fit: " +
                                          cudaGetErrorString(e));
    CU(cudaMemcpyAsync(h->d_tags, tags, (size_t)ncols * nptl_tracking * sizeof(int), cudaMemcpyHostToDevice, h->st));
    CU(cudaMemsetAsync(h->d_tracked, 0, rec_bytes, h->st));  // reset_tracked_particles
    CU(cudaStreamSynchronize(h->st));
    h->trk.enabled = 1; h->trk.ncols = ncols; h->trk.ntrack = nptl_tracking; h->trk.nsteps_max = nmax;
    h->trk.tags = h->d_tags; h->trk.rec = h->d_tracked;
    return GPAT_OK;
}

int gpat_tracked_shape(gpat_handle h, int64_t* nsteps_tracking_max, int64_t* nptl_tracking)
{
    if (!h) return GPAT_ERR_INVALID;
    if (nsteps_tracking_max) *nsteps_tracking_max = h->trk.enabled ? h->trk.nsteps_max : 0;
    if (nptl_tracking) *nptl_tracking = h->trk.enabled ? h->trk.ntrack : 0;
    return GPAT_OK;
}

int gpat_download_tracked(gpat_handle h, gpat_particle* out)
{
    if (!h || !out) return fail(h, GPAT_ERR_INVALID, "gpat_download_tracked: bad arguments");
    if (!h->trk.enabled) return fail(h, GPAT_ERR_STATE, "gpat_download_tracked: gpat_init_tracking was not called");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(out, h->d_tracked, (size_t)h->trk.nsteps_max * h->trk.ntrack * sizeof(gpat_particle),
                       cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return GPAT_OK;
}

int gpat_reset_tracked(gpat_handle h)
{
    if (!h) return GPAT_ERR_INVALID;
    if (!h->trk.enabled) return GPAT_OK;
    CU(cudaSetDevice(h->device));
    CU(cudaMemsetAsync(h->d_tracked, 0, (size_t)h->trk.nsteps_max * h->trk.ntrack * sizeof(gpat_particle), h->st));
    CU(cudaStreamSynchronize(h->st));
    return GPAT_OK;
}

int gpat_particle_mover(gpat_handle h, double t0, double dtf, int nsteps_interval, int num_fine_steps,
                        int dump_escaped_dist, uint64_t* steps_done)
{
    if (!h) return GPAT_ERR_INVALID;
    if (!(dtf > 0.0) || num_fine_steps < 1) return fail(h, GPAT_ERR_INVALID, "gpat_particle_mover: need dtf > 0 and num_fine_steps >= 1");
    if (!h->have_field[0] || (h->dp.time_interp && !h->have_field[1]))
        return fail(h, GPAT_ERR_STATE, "gpat_particle_mover: fields have not been uploaded");
    CU(cudaSetDevice(h->device));
    if (dump_escaped_dist) {
        int rc = ensure_escaped(h);
        if (rc) return rc;
    }
    int rc = push_counters(h);
    if (rc) return rc;
    CU(cudaEventRecord(h->ev[4], h->st));
    rc = sort_before_push(h);
    if (rc) return rc;
    rc = run_push(h, t0, dtf, nsteps_interval, num_fine_steps, 0, steps_done);
    if (rc) return rc;
    // remove_particles; (send/recv and add_neighbor_particles are no-ops for one rank per field)
    CU(cudaEventRecord(h->ev[2], h->st));
    launch_remove(h->P, h->E, h->ecap, h->nptl_current, h->d_counters, h->w, h->idx_a, h->idx_b,
                  dump_escaped_dist, h->st);
    h->tm.total_launches += 15;
    rc = sync_counters(h);
    if (rc) return rc;
    // final pass with the un-extended box + second remove (particle_module.f90:1959-1971)
    launch_final_bc(h->dp, h->P, h->nptl_current, h->d_counters, h->d_leak, h->st);
    launch_remove(h->P, h->E, h->ecap, h->nptl_current, h->d_counters, h->w, h->idx_a, h->idx_b,
                  dump_escaped_dist, h->st);
    h->tm.total_launches += 16;
    CU(cudaEventRecord(h->ev[3], h->st));
    CU(cudaEventRecord(h->ev[5], h->st));
    CU(cudaGetLastError());
    rc = sync_counters(h);
    if (rc) return rc;
    h->tm.compact_ms = elapsed(h->ev[2], h->ev[3]);
    h->tm.mover_ms = elapsed(h->ev[4], h->ev[5]);
    return GPAT_OK;
}

int gpat_debug_push_n(gpat_handle h, double t0, double dtf, int nsteps, uint64_t* steps_done)
{
    if (!h || nsteps < 1 || !(dtf > 0.0)) return fail(h, GPAT_ERR_INVALID, "gpat_debug_push_n: bad arguments");
    if (!h->have_field[0] || (h->dp.time_interp && !h->have_field[1]))
        return fail(h, GPAT_ERR_STATE, "gpat_debug_push_n: fields have not been uploaded");
    CU(cudaSetDevice(h->device));
    int rc = push_counters(h);
    if (rc) return rc;
    rc = run_push(h, t0, dtf, 1 << 30, 1, nsteps, steps_done);
    if (rc) return rc;
    long long keep = h->nptl_current;
    rc = sync_counters(h);
    h->nptl_current = keep;
    return rc;
}

int gpat_split(gpat_handle h, double split_ratio, double pmin_split, int nsteps_interval)
{
    (void)nsteps_interval;
    if (!h) return GPAT_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    int rc = push_counters(h);
    if (rc) return rc;
    if (h->nptl_current > 0) {
        CU(cudaEventRecord(h->ev[2], h->st));
        launch_split(h->dp, h->P, h->nptl_current, h->nptl_max, split_ratio, pmin_split, h->d_counters,
                     h->d_nptl_split, h->w, h->idx_a, h->st, &h->trk);
        h->tm.total_launches += 5;
        CU(cudaEventRecord(h->ev[3], h->st));
        CU(cudaGetLastError());
        rc = sync_counters(h);
        if (rc) return rc;
        h->tm.split_ms = elapsed(h->ev[2], h->ev[3]);
    }
    return GPAT_OK;
}

int gpat_download_particles(gpat_handle h, gpat_particle* out, int64_t nmax, int64_t* n)
{
    if (!h) return GPAT_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    if (n) *n = h->nptl_current;
    long long cnt = h->nptl_current < nmax ? h->nptl_current : nmax;
    if (!out || cnt <= 0) return GPAT_OK;
    int rc = ensure_aos(h, cnt);
    if (rc) return rc;
    launch_to_aos(h->P, h->d_aos, cnt, h->st);
    h->tm.total_launches++;
    CU(cudaMemcpyAsync(out, h->d_aos, (size_t)cnt * sizeof(gpat_particle), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return GPAT_OK;
}

int gpat_upload_particles(gpat_handle h, const gpat_particle* in, int64_t n)
{
    if (!h || (!in && n > 0) || n < 0) return fail(h, GPAT_ERR_INVALID, "gpat_upload_particles: bad arguments");
    if (n > h->nptl_max) return fail(h, GPAT_ERR_INVALID, "gpat_upload_particles: n exceeds nptl_max");
    CU(cudaSetDevice(h->device));
    if (n > 0) {
        int rc = ensure_aos(h, n);
        if (rc) return rc;
        CU(cudaMemcpyAsync(h->d_aos, in, (size_t)n * sizeof(gpat_particle), cudaMemcpyHostToDevice, h->st));
        launch_from_aos(h->P, h->d_aos, n, h->st);
        h->tm.total_launches++;
        CU(cudaStreamSynchronize(h->st));
    }
    h->nptl_current = n;
    return GPAT_OK;
}

int gpat_download_escaped(gpat_handle h, gpat_particle* out, int64_t nmax, int64_t* n)
{
    if (!h) return GPAT_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    long long have = h->esc_mem ? (h->nptl_escaped < h->ecap ? h->nptl_escaped : h->ecap) : 0;
    if (n) *n = h->nptl_escaped;
    long long cnt = have < nmax ? have : nmax;
    if (!out || cnt <= 0) return GPAT_OK;
    int rc = ensure_aos(h, cnt);
    if (rc) return rc;
    launch_to_aos(h->E, h->d_aos, cnt, h->st);
    h->tm.total_launches++;
    CU(cudaMemcpyAsync(out, h->d_aos, (size_t)cnt * sizeof(gpat_particle), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return GPAT_OK;
}

int gpat_reset_escaped(gpat_handle h)
{
    if (!h) return GPAT_ERR_INVALID;
    h->nptl_escaped = 0;
    return GPAT_OK;
}

int gpat_get_counters(gpat_handle h, gpat_counters* c)
{
    if (!h || !c) return GPAT_ERR_INVALID;
    c->nptl_current = h->nptl_current; c->nptl_split = h->nptl_split;
    c->nptl_escaped = h->nptl_escaped; c->nptl_max = h->nptl_max; c->tag_max = h->tag_max;
    c->leak = h->leak; c->leak_negp = h->leak_negp;
    return GPAT_OK;
}

int gpat_set_counters(gpat_handle h, const gpat_counters* c)
{
    if (!h || !c) return GPAT_ERR_INVALID;
    if (c->nptl_current < 0 || c->nptl_current > h->nptl_max)
        return fail(h, GPAT_ERR_INVALID, "gpat_set_counters: nptl_current out of range");
    h->nptl_current = c->nptl_current; h->nptl_split = c->nptl_split;
    h->nptl_escaped = c->nptl_escaped; h->tag_max = c->tag_max;
    h->leak = c->leak; h->leak_negp = c->leak_negp;
    return GPAT_OK;
}

int gpat_diagnostics(gpat_handle h, int local_dist, double* fglobal, double* const flocal[4],
                     double quick[8], double* pmax)
{
    if (!h) return GPAT_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    const gpat_params& p = h->hp;
    const size_t nglob = (size_t)p.nmu_global * p.npp_global;
    DiagArgs a{};
    fill_diag_args(h, a, local_dist);
    CU(cudaEventRecord(h->ev[2], h->st));
    CU(cudaMemsetAsync(h->d_fglobal, 0, nglob * sizeof(double), h->st));
    for (int k = 0; k < 4; ++k) {
        bool want = local_dist && flocal && flocal[k] && h->d_flocal[k];
        if (want) CU(cudaMemsetAsync(h->d_flocal[k], 0, h->local_bins[k] * sizeof(double), h->st));
        else { a.loc[k].enabled = 0; a.loc[k].data = nullptr; }
    }
    CU(cudaMemsetAsync(h->d_sums, 0, 2 * sizeof(double), h->st));
    const unsigned long long mm0[3] = {0x7ff0000000000000ull, 0ull, 0ull};
    CU(cudaMemcpyAsync(h->d_minmax, mm0, sizeof(mm0), cudaMemcpyHostToDevice, h->st));
    double lk[2] = {h->leak, h->leak_negp};
    CU(cudaMemcpyAsync(h->d_leak, lk, sizeof(lk), cudaMemcpyHostToDevice, h->st));
    launch_diag(h->P, a, h->sm_count, h->st);
    launch_finalize_quick(h->d_sums, h->d_minmax, h->d_leak, (double)h->nptl_current,
                          (double)h->nptl_split, h->d_quick, h->st);
    h->tm.total_launches += 2;
    if (h->comm) {
        // MPI_REDUCE(... MPI_SUM) of diagnostics.f90:881-905, 143-151, 1707 as one NCCL group
        NcclApi& N = nccl();
        NC(N.GroupStart());
        NC(N.AllReduce(h->d_fglobal, h->d_fglobal, nglob, ncclDouble, ncclSum, h->comm, h->st));
        for (int k = 0; k < 4; ++k)
            if (a.loc[k].enabled)
                NC(N.AllReduce(h->d_flocal[k], h->d_flocal[k], h->local_bins[k], ncclDouble, ncclSum, h->comm, h->st));
        NC(N.AllReduce(h->d_quick, h->d_quick, 6, ncclDouble, ncclSum, h->comm, h->st));
        NC(N.AllReduce(h->d_quick + 6, h->d_quick + 6, 1, ncclDouble, ncclMin, h->comm, h->st));
        NC(N.AllReduce(h->d_quick + 7, h->d_quick + 7, 2, ncclDouble, ncclMax, h->comm, h->st));
        NC(N.GroupEnd());
    }
    if (fglobal) CU(cudaMemcpyAsync(fglobal, h->d_fglobal, nglob * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    for (int k = 0; k < 4; ++k)
        if (a.loc[k].enabled)
            CU(cudaMemcpyAsync(flocal[k], h->d_flocal[k], h->local_bins[k] * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    double q[9];
    CU(cudaMemcpyAsync(q, h->d_quick, sizeof(q), cudaMemcpyDeviceToHost, h->st));
    CU(cudaEventRecord(h->ev[3], h->st));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->st));
    h->tm.diag_ms = elapsed(h->ev[2], h->ev[3]);
    if (quick) memcpy(quick, q, 8 * sizeof(double));
    if (pmax) *pmax = q[8];
    return GPAT_OK;
}

int gpat_escaped_diagnostics(gpat_handle h, double* fescaped)
{
    if (!h || !fescaped) return GPAT_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    const gpat_params& p = h->hp;
    const int nface = 2 * p.ndim;
    const size_t n = (size_t)p.nmu_global * p.npp_global * nface;
    DiagArgs a{};
    fill_diag_args(h, a, 0);
    CU(cudaMemsetAsync(h->d_fesc, 0, n * sizeof(double), h->st));
    long long have = h->esc_mem ? (h->nptl_escaped < h->ecap ? h->nptl_escaped : h->ecap) : 0;
    if (have > 0) {
        launch_escaped_diag(h->E, have, a, nface, h->d_fesc, h->st);
        h->tm.total_launches++;
    }
    if (h->comm) NC(nccl().AllReduce(h->d_fesc, h->d_fesc, n, ncclDouble, ncclSum, h->comm, h->st));
    CU(cudaMemcpyAsync(fescaped, h->d_fesc, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->st));
    return GPAT_OK;
}

int gpat_escaped_local_diagnostics(gpat_handle h, double* const fx[4], double* const fy[4], double* const fz[4])
{
    if (!h) return GPAT_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    const gpat_params& p = h->hp;
    DiagArgs a{};
    fill_diag_args(h, a, 1);
    // one device buffer: per enabled set its x-, y- (ndim > 1) and z-face (ndim > 2) arrays, in that order
    size_t off[4][3], len[4][3], total = 0;
    for (int k = 0; k < 4; ++k) {
        const HistDev& d = a.loc[k];
        const size_t nb = (size_t)d.nmu * d.npbins * 2;
        len[k][0] = d.enabled ? nb * d.nry * d.nrz : 0;
        len[k][1] = (d.enabled && p.ndim > 1) ? nb * d.nrx * d.nrz : 0;
        len[k][2] = (d.enabled && p.ndim > 2) ? nb * d.nrx * d.nry : 0;
        for (int f = 0; f < 3; ++f) { off[k][f] = total; total += len[k][f]; }
    }
    if (total == 0) return GPAT_OK;
    if (h->fesc_loc_n < total) {
        if (h->d_fesc_loc) cudaFree(h->d_fesc_loc);
        h->d_fesc_loc = nullptr;
        h->fesc_loc_n = 0;
        CU(cudaMalloc(&h->d_fesc_loc, total * sizeof(double)));
        h->fesc_loc_n = total;
    }
    CU(cudaMemsetAsync(h->d_fesc_loc, 0, total * sizeof(double), h->st));
    EscLocalDev o{};
    for (int k = 0; k < 4; ++k) {
        o.fx[k] = len[k][0] ? h->d_fesc_loc + off[k][0] : nullptr;
        o.fy[k] = len[k][1] ? h->d_fesc_loc + off[k][1] : nullptr;
        o.fz[k] = len[k][2] ? h->d_fesc_loc + off[k][2] : nullptr;
    }
    long long have = h->esc_mem ? (h->nptl_escaped < h->ecap ? h->nptl_escaped : h->ecap) : 0;
    if (have > 0) {
        launch_escaped_local(h->E, have, a, o, h->st);
        h->tm.total_launches++;
    }
    // MPI_REDUCE of the face arrays, diagnostics.f90:1174-1230
    if (h->comm) NC(nccl().AllReduce(h->d_fesc_loc, h->d_fesc_loc, total, ncclDouble, ncclSum, h->comm, h->st));
    double* const* outs[3] = {fx, fy, fz};
    for (int k = 0; k < 4; ++k)
        for (int f = 0; f < 3; ++f)
            if (len[k][f] && outs[f] && outs[f][k])
                CU(cudaMemcpyAsync(outs[f][k], h->d_fesc_loc + off[k][f], len[k][f] * sizeof(double),
                                   cudaMemcpyDeviceToHost, h->st));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->st));
    return GPAT_OK;
}

int gpat_hist_edges(gpat_handle h, int which, double* pedges, double* muedges)
{
    if (!h || which < 0 || which > 4) return GPAT_ERR_INVALID;
    const gpat_params& p = h->hp;
    double pmin = which ? p.local[which - 1].pmin : p.pmin;
    double pmax = which ? p.local[which - 1].pmax : p.pmax;
    int np = which ? p.local[which - 1].npbins : p.npp_global;
    int nmu = which ? p.local[which - 1].nmu : p.nmu_global;
    // diagnostics.f90:199-209
    double pmin_log = std::log10(pmin), pmax_log = std::log10(pmax);
    double dp_log = (pmax_log - pmin_log) / np;
    if (pedges)
        for (int i = 1; i <= np + 1; ++i) pedges[i - 1] = std::pow(10.0, pmin_log + (i - 1) * dp_log);
    double dmu = (double)(2.0f / (float)nmu);
    if (muedges)
        for (int i = 1; i <= nmu + 1; ++i) muedges[i - 1] = -1.0 + (i - 1) * dmu;
    return GPAT_OK;
}

int gpat_comm_unique_id(char id[128])
{
    if (!id) return GPAT_ERR_INVALID;
    NcclApi& N = nccl();
    if (!N.ok) return fail(nullptr, GPAT_ERR_NCCL, "libnccl.so.2 could not be loaded (set GPAT_NCCL_LIB)");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    if (N.GetUniqueId(&u) != ncclSuccess) return fail(nullptr, GPAT_ERR_NCCL, "ncclGetUniqueId failed");
    memcpy(id, &u, 128);
    return GPAT_OK;
}

int gpat_comm_init(gpat_handle h, const char id[128], int nranks, int rank)
{
    if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, GPAT_ERR_INVALID, "gpat_comm_init: bad arguments");
    NcclApi& N = nccl();
    if (!N.ok) return fail(h, GPAT_ERR_NCCL, "libnccl.so.2 could not be loaded (set GPAT_NCCL_LIB)");
    CU(cudaSetDevice(h->device));
    if (h->comm) { N.CommDestroy(h->comm); h->comm = nullptr; }
    ncclUniqueId u;
    memcpy(&u, id, 128);
    NC(N.CommInitRank(&h->comm, nranks, u, rank));
    h->nranks = nranks;
    h->rank = rank;
    return GPAT_OK;
}

int gpat_comm_destroy(gpat_handle h)
{
    if (!h) return GPAT_ERR_INVALID;
    if (h->comm && nccl().ok) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->st);
        nccl().CommDestroy(h->comm);
    }
    h->comm = nullptr;
    h->nranks = 1;
    h->rank = 0;
    return GPAT_OK;
}

int gpat_get_timings(gpat_handle h, gpat_timings* t)
{
    if (!h || !t) return GPAT_ERR_INVALID;
    *t = h->tm;
    return GPAT_OK;
}

int gpat_set_rng_table(gpat_handle h, const double* u, int64_t nslots, int64_t max_steps)
{
    if (!h) return GPAT_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    if (h->d_table) { cudaFree(h->d_table); h->d_table = nullptr; }
    h->table_slots = h->table_steps = 0;
    if (!u || nslots <= 0 || max_steps <= 0) return GPAT_OK;
    size_t bytes = (size_t)nslots * max_steps * 4 * sizeof(double);
    CU(cudaMalloc(&h->d_table, bytes));
    CU(cudaMemcpy(h->d_table, u, bytes, cudaMemcpyHostToDevice));
    h->table_slots = nslots;
    h->table_steps = max_steps;
    return GPAT_OK;
}

int gpat_debug_gradients(gpat_handle h, const float* f8, float* out32)
{
    if (!h || !f8 || !out32) return GPAT_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    size_t ncell = (size_t)h->dp.nxg * h->dp.nyg_src * h->dp.nzg;
    float *d_in = nullptr, *d_out = nullptr;
    CU(cudaMalloc(&d_in, ncell * 8 * sizeof(float)));
    cudaError_t e = cudaMalloc(&d_out, ncell * 32 * sizeof(float));
    if (e != cudaSuccess) { cudaFree(d_in); return fail(h, GPAT_ERR_CUDA, cudaGetErrorString(e)); }
    cudaMemcpyAsync(d_in, f8, ncell * 8 * sizeof(float), cudaMemcpyHostToDevice, h->st);
    launch_grad32(d_in, h->dp, d_out, h->sm_count, h->st);
    h->tm.total_launches++;
    cudaMemcpyAsync(out32, d_out, ncell * 32 * sizeof(float), cudaMemcpyDeviceToHost, h->st);
    e = cudaStreamSynchronize(h->st);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(h, GPAT_ERR_CUDA, cudaGetErrorString(e));
    return GPAT_OK;
}

int gpat_debug_interp(gpat_handle h, int64_t n, const double* x, const double* y, const double* z,
                      const double* rt, double* fields32)
{
    if (!h || n < 0 || !x || !y || !z || !rt || !fields32) return GPAT_ERR_INVALID;
    if (n == 0) return GPAT_OK;
    CU(cudaSetDevice(h->device));
    double* d = nullptr;
    CU(cudaMalloc(&d, (size_t)n * (4 + 32) * sizeof(double)));
    double *dx = d, *dy = d + n, *dz = d + 2 * n, *drt = d + 3 * n, *dout = d + 4 * n;
    cudaMemcpyAsync(dx, x, n * sizeof(double), cudaMemcpyHostToDevice, h->st);
    cudaMemcpyAsync(dy, y, n * sizeof(double), cudaMemcpyHostToDevice, h->st);
    cudaMemcpyAsync(dz, z, n * sizeof(double), cudaMemcpyHostToDevice, h->st);
    cudaMemcpyAsync(drt, rt, n * sizeof(double), cudaMemcpyHostToDevice, h->st);
    launch_interp_debug(h->layout, h->dp, h->fld, h->sel, n, dx, dy, dz, drt, dout, h->st);
    h->tm.total_launches++;
    cudaMemcpyAsync(fields32, dout, (size_t)n * 32 * sizeof(double), cudaMemcpyDeviceToHost, h->st);
    cudaError_t e = cudaStreamSynchronize(h->st);
    cudaFree(d);
    if (e != cudaSuccess) return fail(h, GPAT_ERR_CUDA, cudaGetErrorString(e));
    return GPAT_OK;
}

}  // extern "C"
