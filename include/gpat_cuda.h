/*
 * gpat_cuda.h -- C ABI of the B200-native pseudo-particle SDE integrator for
 * GPAT (xiaocanli/stochastic-parker).
 *
 * This is the drop-in boundary for ONE hot path of the reference: the Parker
 * transport particle push + RNG + split/compaction + histogram diagnostics
 * (SURVEY.md section 8).  The reference has no FFI of its own; the boundary sits
 * where `program stochastic` (src/programs/stochastic-mhd.f90:12-33) `use`s
 * particle_module / diagnostics / random_number_generator.  Every entry point
 * below names the reference procedure it replaces (file:line under
 * /root/reference/src).  fortran/gpat_cuda_iface.f90 holds the matching
 * ISO_C_BINDING interface block; INTEGRATION.md shows the call-site patch.
 *
 * Conventions
 *  - plain pointers and sizes only; all host arrays are owned by the caller
 *    and are never retained past the call;
 *  - arrays are column-major with the reference's index order (first index
 *    fastest), so a Fortran array can be passed with c_loc();
 *  - every function returns 0 (GPAT_OK) or a GPAT_ERR_* code; the message is
 *    available from gpat_last_error().  The reference's own error style is
 *    "print on rank 0; MPI_FINALIZE; stop" (simulation_setup.f90:78-87,
 *    diagnostics.f90:1989-2014): the Fortran shim does that on non-zero;
 *  - capacity overflow is SILENT, exactly as in the reference
 *    (particle_module.f90:491-492, 5444-5447);
 *  - there is no CPU fallback: without a CUDA device gpat_init fails.
 */
#ifndef GPAT_CUDA_H
#define GPAT_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPAT_OK 0
#define GPAT_ERR_INVALID 1 /* bad argument, or a reference feature outside the GPU path */
#define GPAT_ERR_CUDA 2    /* CUDA runtime error (no device, OOM, launch failure) */
#define GPAT_ERR_STATE 3   /* call sequence error (e.g. mover before fields) */
#define GPAT_ERR_NCCL 4    /* NCCL not loadable or a collective failed */

/* count_flag values, particle_module.f90:55-62 */
#define GPAT_COUNT_FLAG_INBOX 1
#define GPAT_COUNT_FLAG_OTHERS 0
#define GPAT_COUNT_FLAG_ESCAPE_LX (-1)
#define GPAT_COUNT_FLAG_ESCAPE_HX (-2)
#define GPAT_COUNT_FLAG_ESCAPE_LY (-3)
#define GPAT_COUNT_FLAG_ESCAPE_HY (-4)
#define GPAT_COUNT_FLAG_ESCAPE_LZ (-5)
#define GPAT_COUNT_FLAG_ESCAPE_HZ (-6)

/* AoS particle record == `particle_type`, particle_module.f90:38-50 (104 B with
 * natural padding).  Device storage is SoA; this layout is used only for
 * upload/download (restart files, particle dumps: diagnostics.f90:1811-1888).
 * `padding` carries the 64-bit per-particle RNG step counter (bit pattern),
 * which the reference's mt_stream state has no equivalent of. */
typedef struct gpat_particle {
    int8_t split_times;
    int8_t count_flag;
    int8_t pad_[2];
    int32_t origin;
    int32_t nsteps_tracked;
    int32_t nsteps_pushed;
    int32_t tag_injected;
    int32_t tag_splitted;
    double x, y, z, p;
    double v, mu;
    double weight, t, dt;
    double padding;
} gpat_particle;

/* One set of local-distribution parameters, diagnostics.f90:53-65, 2118-2140 */
typedef struct gpat_hist_spec {
    int32_t enabled; /* dump_local_distK */
    int32_t npbins;  /* npbinsK */
    int32_t nmu;     /* nmuK (forced to 1 for Parker, diagnostics.f90:2124-2128) */
    int32_t rx, ry, rz;
    double pmin, pmax; /* pminK, pmaxK */
} gpat_hist_spec;

#define GPAT_RNG_PHILOX 0 /* on-device Philox4x32-10, one stream per particle */
#define GPAT_RNG_TABLE 1  /* replay pre-generated uniforms (gpat_set_rng_table) */

/* Everything the reference keeps in module variables for this path.
 * Sources: read_particle_params (particle_module.f90:2778-2878), set_dpp_params /
 * set_flags_params / set_drift_parameters / set_flag_check_drift_2d
 * (particle_module.f90:260-335), read_diagnostics_params (diagnostics.f90:2049-2200),
 * mhd_config (mhd_config.f90:16-27), fconfig for a 1x1x1 topology
 * (simulation_setup.f90:171-253), particle BCs (simulation_setup.f90:107-123). */
typedef struct gpat_params {
    /* grid */
    int32_t ndim;        /* ndim_field: 1, 2 or 3 (1-D: ny = nz = 1) */
    int32_t nx, ny, nz;  /* mhd_config%nx.. without ghost cells (nz=1 in 2-D) */
    int32_t time_interp; /* time_interp_flag */
    int32_t pbc[3];      /* pbcx,pbcy,pbcz: 0 periodic, 1 open */
    double dx, dy, dz;
    double xmin, ymin, zmin;
    double xmax, ymax, zmax;
    double lx, ly, lz;
    /* particle parameters */
    double b0, p0, pmin, pmax;
    double gamma_turb, pindex; /* pindex = 3 - gamma_turb in the reference */
    double kpara0, kret;
    double dt_min_rel, dt_max_rel;
    int32_t momentum_dependency, mag_dependency;
    int32_t acc_region_flag;
    int32_t pad0_;
    double acc_region[6]; /* xmin,xmax,ymin,ymax,zmin,zmax in [0,1] */
    /* momentum diffusion */
    int32_t dpp_wave, dpp_shear, weak_scattering;
    int32_t keep_rho; /* keep the density slot in the device record (inject_large_rho) */
    double tau0;
    /* drift */
    double drift1, drift2;
    int32_t pcharge;
    int32_t check_drift_2d;
    /* model switches */
    int32_t include_3rd_dim; /* include_3rd_dim_in2d_flag */
    int32_t nlgc;
    double kperp_kpara;
    double duu0; /* duu_init (set_duu_params, particle_module.f90:279-283): focused transport only */
    /* focused_transport = 1: Cartesian push_particle_2d_ft / _2d_include_3rd_ft / _3d_ft (reference-
     * order build); 1-D is rejected (push_particle_1d_ft reads an unassigned dx_dt);
     * spherical_coord and nonuniform_grid must be 0 on the GPU path (error otherwise) */
    int32_t focused_transport, spherical_coord, nonuniform_grid;
    /* deltab_flag / correlation_flag: turbulence maps via gpat_upload_turbulence;
     * acc_by_surface: 3-D only, heights via gpat_upload_acc_surface */
    int32_t deltab_flag, correlation_flag, acc_by_surface;
    /* diagnostics */
    int32_t npp_global, nmu_global;
    gpat_hist_spec local[4];
    /* RNG */
    uint64_t seed;
    int32_t rng_mode; /* GPAT_RNG_* */
    /* rank identity (particle `origin`, particle_module.f90:428) */
    int32_t mpi_rank;
    /* arithmetic: 0 = fast (FMA contraction, fused time-blend), 1 = strict
     * (no contraction, reference operation order; used by the parity tests) */
    int32_t strict_math;
    /* acceleration surfaces (acc_by_surface = 1, 3-D only; acc_region_surface.f90:29-94):
     * surface_normK = sign * (axis + 1) for the reference's "+x" .. "-z" strings, i.e. +1/-1 x,
     * +2/-2 y, +3/-3 z; surface2_existed, is_intersection as in stochastic-mhd.f90:116-117 */
    int32_t surface_norm1, surface_norm2;
    int32_t surface2_existed, is_intersection;
    int32_t pad2_;
} gpat_params;

typedef struct gpat_sim* gpat_handle;

/* ---- life cycle ------------------------------------------------------- */

/* Replaces init_particles (particle_module.f90:171), init_prng
 * (random_number_generator.f90:28), init_field_data (mhd_data_parallel.f90:66),
 * init_particle_distributions (diagnostics.f90:178).  Allocates device SoA
 * particle storage (capacity nptl_max), the field store and the histograms. */
int gpat_init(gpat_handle* h, int device, int64_t nptl_max, const gpat_params* params);

/* Re-reads the physics/diagnostics parameters (not the grid shape). */
int gpat_set_params(gpat_handle h, const gpat_params* params);

/* Replaces free_particles / delete_prng / free_field_data. */
int gpat_finalize(gpat_handle h);

/* Last error text for this handle (h may be NULL for init failures). */
const char* gpat_last_error(gpat_handle h);

/* ---- fields ------------------------------------------------------------ */

/* Consumer side of read_field_data_parallel (mhd_data_parallel.f90:224) +
 * calc_fields_gradients (mhd_data_parallel.f90:504).  `f` points at the first
 * element of an array (nvar, nx+4, ny+4[, nz+4]) column-major, nvar = 8 (the
 * on-disk mhd_data_NNNN record) or 32 (the reference's farray1/farray2).
 * with_grad = 0: only the 8 primaries are read and the device computes the
 * gradients with the reference's FP32/FP64 arithmetic; with_grad = 1 (nvar must
 * be 32): slots 9..32 are taken as computed by the host.
 * slot = 0 -> farray1 (frame at t0), 1 -> farray2 (frame at t0 + dtf). */
int gpat_upload_fields(gpat_handle h, int slot, const float* f, int nvar, int with_grad);

/* Turbulence maps (`-db 1`, `-co 1`).  Replaces read_magnetic_fluctuation + calc_grad_sigma2_slab
 * + calc_grad_sigma2_2d (which = 0; mhd_data_parallel.f90:306-397, 771-1116) and
 * read_correlation_length + calc_grad_lc_slab + calc_grad_lc_2d (which = 1; :406-497, 1259-1604):
 * data holds the slab array followed by the 2-D array, one float per ghosted grid point each --
 * the content of the reference's deltab_NNNN / lc_NNNN file.  slot as in gpat_upload_fields;
 * gpat_swap_fields also stands for copy_magnetic_fluctuation / copy_correlation_length
 * (:1928-1941).  The maps enter kappa (particle_module.f90:2246-2254, 2314-2321, 2505-2517,
 * 2589-2604), D_mumu (:3143-3148) and inject_large_db2; both builds gather them (production build:
 * by the lane group that gathers the particle's field record). */
int gpat_upload_turbulence(gpat_handle h, int which, int slot, const float* data);

/* Acceleration surfaces (`-as 1`, 3-D).  Replaces read_acc_surface (acc_region_surface.f90:121-243):
 * heights = acc_surfaceK1/K2, real(dp), shaped (-1:n1+2, -1:n2+2) column-major with (n1, n2) the
 * grid sizes of the two axes other than the surface normal, in x < y < z order; which = 0/1 for
 * surface 1/2, slot as in gpat_upload_fields (gpat_swap_fields also stands for copy_acc_surface,
 * :390-396).  interp_acc_surface (:255-334) and check_above_acc_surface (:342-388) run inside the
 * 3-D pushers (particle_module.f90:4887-4892, 5297-5303), both builds. */
int gpat_upload_acc_surface(gpat_handle h, int which, int slot, const double* heights);

/* Frame pipeline (stochastic-mhd.f90:401-447 reads frame tf at the top of every iteration,
 * serially).  gpat_prefetch_fields starts the host->device copy of a frame the caller has
 * ALREADY read (e.g. frame tf+1, read while frame tf is being pushed) on a separate copy
 * stream and returns at once, so the copy overlaps the next gpat_particle_mover.  A later
 * gpat_upload_fields with the same host pointer and nvar finds the bytes on the device and
 * only runs the gradient/pack kernel.  The host buffer must stay unchanged (and should be
 * page-locked) until that gpat_upload_fields call returns.  A prefetch is consumed by the NEXT
 * gpat_upload_fields call only: an upload with another pointer or nvar discards it (a caller that
 * reuses one buffer for every frame must not rewrite it between the prefetch and its upload).
 * Optional: without it gpat_upload_fields copies synchronously as before. */
int gpat_prefetch_fields(gpat_handle h, const float* f, int nvar);

/* Replaces copy_fields (mhd_data_parallel.f90:1920): farray1 = farray2.  O(1): the two halves of the
 * device store change roles, nothing is copied.  Slot 1 therefore holds the OLD slot 0 afterwards, not a
 * second copy of the new slot 0 as in the reference; a caller that reads no new frame before the next
 * gpat_particle_mover (tf > tmax_mhd, stochastic-mhd.f90:400) sends the last frame to slot 1 again
 * (both host drivers of this repository do).  The same holds for the turbulence maps and the
 * acceleration surfaces, which swap with the fields. */
int gpat_swap_fields(gpat_handle h);

/* ---- particles --------------------------------------------------------- */

/* Replaces inject_particles_spatial_uniform (particle_module.f90:454-530, whole-field
 * branch) + inject_one_particle (particle_module.f90:385-441).  t_frame =
 * tstamps_mhd(ct_mhd), dt_mhd = tstamps_mhd(ct_mhd+1) - tstamps_mhd(ct_mhd).
 * part_box = xmin,ymin,zmin,xmax,ymax,zmax (stochastic-mhd.f90:375-391). */
int gpat_inject_uniform(gpat_handle h, int64_t nptl, double dt, int dist_flag,
                        double particle_v0, double t_frame, double dt_mhd,
                        const double part_box[6], double power_index);

/* Targeted injection on the device.  Replaces inject_particles_at_large_jz
 * (particle_module.f90:785-905), _at_large_absj (:919-1061), _at_large_divv (:1250-1341) and
 * _at_large_rho (:1356-1468) together with the cell counters get_ncells_large_jz / _absj /
 * _divv / _rho (mhd_data_parallel.f90:2211-2261, 2269-2335, 2385-2455, 2463-2498), for the
 * whole-field-per-rank decomposition (mpi_sub_size = 1):
 *   ncells      = cells of part_box whose farray1 value exceeds vmin (jz_min / absj_min /
 *                 divv_min / rho_min);
 *   nptl_inject = int(nptl * ncells / (inject_same_nptl ? ncells : ncells_norm));
 *   every new particle draws positions uniformly in the WHOLE domain from its own Philox
 *   injection stream until the field interpolated at rt = 0 passes the threshold inside
 *   part_box, then continues like gpat_inject_uniform (mu, momentum, time).
 * nptl_injected / ncells (may be NULL) receive nptl_inject and the cell count.
 * GPAT_INJECT_LARGE_DB2 (inject_particles_at_large_db2, :1075-1236, get_ncells_large_db2
 * mhd_data_parallel.f90:2343-2377) needs the deltab maps (gpat_upload_turbulence);
 * GPAT_INJECT_LARGE_RHO needs gpat_params.keep_rho = 1 unless
 * momentum diffusion already keeps the density slot. */
#define GPAT_INJECT_LARGE_JZ 1
#define GPAT_INJECT_LARGE_ABSJ 2
#define GPAT_INJECT_LARGE_DB2 3
#define GPAT_INJECT_LARGE_DIVV 4
#define GPAT_INJECT_LARGE_RHO 5
#define GPAT_INJECT_AT_SHOCK 6 /* internal mode of gpat_inject_at_shock */
int gpat_inject_targeted(gpat_handle h, int mode, int64_t nptl, double dt, int dist_flag,
                         double particle_v0, double t_frame, double dt_mhd,
                         const double part_box[6], double power_index, int inject_same_nptl,
                         double vmin, int64_t ncells_norm, int64_t* nptl_injected,
                         int64_t* ncells);

/* Replaces locate_shock_xpos (mhd_data_parallel.f90:1988-2006) + inject_particles_at_shock
 * (particle_module.f90:542-633; `-is 1`, config/shock.sh).  Needs time_interp = 1 and both
 * frames uploaded: at rt = 0 the reference's swapped time weights select the LATER frame's
 * shock positions.  The reference accumulates into uninitialised sx1/sx2 in 2-D/3-D
 * (mhd_data_parallel.f90:2037-2049); here they start at zero.  Everything else is kept as
 * written (rz from dpy, weights that do not sum to one, t = frame time, the 0.75 envelope). */
int gpat_inject_at_shock(gpat_handle h, int64_t nptl, double dt, int dist_flag, double particle_v0,
                         double t_frame, double power_index);

/* ---- particle tracking ---------------------------------------------------
 * The second run of the reference's two-run workflow (docs/source/development/
 * particle_module.rst:42-131): particles whose (origin, tag_injected, tag_splitted chain) appears
 * in the tag table are marked by NEGATED tags at injection (particle_module.f90:434-440), sampled
 * every nsteps_interval pushes into particles_tracked (particle_module.f90:1697-1724, 1806-1812)
 * and followed through splits (particle_module.f90:5452-5473).  The per-particle Philox streams
 * are keyed by |tag|, so the tracking run replays the trajectories of the run the tags came from.
 *
 * gpat_init_tracking replaces init_particle_tracking (particle_module.f90:5825-5879) minus the
 * HDF5 read: tags = tags_tracking(ncols = split_times_max + 2, nptl_tracking), column-major, sorted
 * by origin, tag_injected and the tag_splitted chain.  particles_tracked has
 * nsteps_tracking_max = ceiling((1/dt_min_rel)/nsteps_interval) + 1 rows per tracked particle.
 * gpat_download_tracked copies particles_tracked(nsteps_tracking_max, nptl_tracking) (column-major
 * records, what dump_tracked_particles writes, particle_module.f90:6236-6299);
 * gpat_reset_tracked is reset_tracked_particles (particle_module.f90:5884-5902).
 * Tracking runs call gpat_particle_mover with num_fine_steps = 1 (stochastic-mhd.f90:497-499). */
int gpat_init_tracking(gpat_handle h, const int32_t* tags, int ncols, int64_t nptl_tracking,
                       int nsteps_interval);
int gpat_tracked_shape(gpat_handle h, int64_t* nsteps_tracking_max, int64_t* nptl_tracking);
int gpat_download_tracked(gpat_handle h, gpat_particle* out);
int gpat_reset_tracked(gpat_handle h);

/* Particle ORDER.  The reference-order build (strict_math = 1) keeps the reference's order
 * (injection order, swap-with-tail removal, children appended).  The production build sorts the
 * particle arrays by grid cell at the start of every gpat_particle_mover (better locality of the
 * field gathers); particles are identified by (origin, tag_injected, tag_splitted), never by
 * position in the array, and no result of this interface depends on the order except which
 * particle is dropped when nptl_max overflows (particle_module.f90:491-492, 5444-5447). */

/* Replaces particle_mover (particle_module.f90:1846-1974) including both
 * remove_particles passes (particle_module.f90:5365-5403).  t0 = tstamps_mhd(frame),
 * dtf = tstamps_mhd(frame+1) - t0.  Blocking.  steps_done (may be NULL) receives
 * the number of push_particle_* calls executed (the unit of the headline
 * metric). */
int gpat_particle_mover(gpat_handle h, double t0, double dtf, int nsteps_interval,
                        int num_fine_steps, int dump_escaped_dist, uint64_t* steps_done);

/* Replaces split_particle (particle_module.f90:5430-5480). */
int gpat_split(gpat_handle h, double split_ratio, double pmin_split, int nsteps_interval);

/* ptls(1:n) access for dump_particles (diagnostics.f90:1811), read_particles
 * (particle_module.f90:5744) and the module state (particle_module.f90:5532-5664). */
int gpat_download_particles(gpat_handle h, gpat_particle* out, int64_t nmax, int64_t* n);
int gpat_upload_particles(gpat_handle h, const gpat_particle* in, int64_t n);

/* Escaped particles of the current interval (escaped_ptls, particle_module.f90:134-135)
 * and reset_escaped_particles (diagnostics.f90, called at stochastic-mhd.f90:533).  The device
 * array grows like resize_escaped_particles (particle_module.f90:5329-5358) when a caller does not
 * reset it every interval.  ORDER: escapees of one remove pass are stored in ascending particle-
 * array index, the reference stores them in the encounter order of its swap-with-tail loop; the
 * set is the same, records are identified by (origin, tag_injected, tag_splitted). */
int gpat_download_escaped(gpat_handle h, gpat_particle* out, int64_t nmax, int64_t* n);
int gpat_reset_escaped(gpat_handle h);

/* Module counters: nptl_current, nptl_split, nptl_escaped, tag_max, leak,
 * leak_negp (particle_module.f90:73-82). */
typedef struct gpat_counters {
    int64_t nptl_current, nptl_split, nptl_escaped, nptl_max;
    int64_t tag_max;
    double leak, leak_negp;
} gpat_counters;
int gpat_get_counters(gpat_handle h, gpat_counters* c);
int gpat_set_counters(gpat_handle h, const gpat_counters* c);

/* ---- diagnostics -------------------------------------------------------- */

/* Replaces calc_particle_distributions (diagnostics.f90:738-906) + quick_check
 * (diagnostics.f90:116-170) + get_pmax_global (diagnostics.f90:1691-1719) in one
 * pass over the particles.  Outputs are ALREADY reduced over the ranks of the
 * communicator set by gpat_comm_init (the MPI_REDUCE calls at diagnostics.f90:
 * 881-905, 143-151, 1707); with no communicator they are the local values.
 *   fglobal   : (nmu_global, npp_global) doubles
 *   flocal[k] : (nmuK, npbinsK, nrxK, nryK, nrzK) doubles, or NULL to skip
 *   quick[8]  : var_global(1:6) of quick_check = nptl_current, nptl_split,
 *               sum(weight), leak, leak_negp, sum(dt); then pdt_min, pdt_max
 *   pmax      : max particle momentum
 * Any pointer may be NULL. */
int gpat_diagnostics(gpat_handle h, int local_dist, double* fglobal, double* const flocal[4],
                     double quick[8], double* pmax);

/* calc_escaped_distributions (diagnostics.f90:913-1232), global part:
 * fescaped (nmu_global, npp_global, 2*ndim), reduced like fglobal. */
int gpat_escaped_diagnostics(gpat_handle h, double* fescaped);

/* Bin edges, init_particle_distributions (diagnostics.f90:196-209, 270-283).
 * which = 0 global, 1..4 local set.  pedges has npbins+1, muedges nmu+1 values. */
/* calc_escaped_distributions, local part (diagnostics.f90:956-1170; arrays of
 * init_local_escaped_distributions, diagnostics.f90:358-405), already summed over ranks like the
 * MPI_REDUCEs of diagnostics.f90:1174-1230.  For each enabled local set k (0..3):
 *   fx[k] = fescaped{k+1}_x(nmu, npbins, nry, nrz, 2)
 *   fy[k] = fescaped{k+1}_y(nmu, npbins, nrx, nrz, 2)   (ndim > 1)
 *   fz[k] = fescaped{k+1}_z(nmu, npbins, nrx, nry, 2)   (ndim > 2)
 * column-major, last index 1 = low face, 2 = high face.  Null pointers (arrays, or single entries) are
 * skipped.  Needs dump_escaped_dist = 1 in gpat_particle_mover, like gpat_escaped_diagnostics. */
int gpat_escaped_local_diagnostics(gpat_handle h, double* const fx[4], double* const fy[4], double* const fz[4]);
int gpat_hist_edges(gpat_handle h, int which, double* pedges, double* muedges);

/* ---- multi-GPU (one process per GPU) ------------------------------------ */

/* NCCL communicator for the diagnostics all-reduce.  The 128-byte id is created
 * on one rank and distributed by the caller (MPI_Bcast in the Fortran driver,
 * torch.distributed in bench.py). */
int gpat_comm_unique_id(char id[128]);
int gpat_comm_init(gpat_handle h, const char id[128], int nranks, int rank);
int gpat_comm_destroy(gpat_handle h);

/* ---- instrumentation ---------------------------------------------------- */

/* Device times (CUDA events on the library's stream) of the last calls, ms. */
typedef struct gpat_timings {
    float mover_ms;      /* whole gpat_particle_mover */
    float push_ms;       /* push kernel only */
    float compact_ms;    /* remove_particles passes */
    float upload_ms;     /* last gpat_upload_fields: H2D + gradient/pack kernel */
    float grad_ms;       /* gradient/pack kernel only */
    float inject_ms, split_ms, diag_ms;
    uint64_t push_steps; /* steps of the last mover call */
    uint32_t push_launches;
    uint32_t total_launches; /* kernels launched by this handle since init */
} gpat_timings;
int gpat_get_timings(gpat_handle h, gpat_timings* t);

/* ---- test hooks ---------------------------------------------------------- */

/* Uniform table for GPAT_RNG_TABLE: u[(slot*max_steps + step)*4 + j], where
 * slot = tag_injected and step = the particle's RNG step counter. */
int gpat_set_rng_table(gpat_handle h, const double* u, int64_t nslots, int64_t max_steps);

/* All 24 gradients of an 8-variable frame with the device gradient arithmetic,
 * returned in the reference's 32-slot layout (parity check of
 * calc_fields_gradients, mhd_data_parallel.f90:533-566). */
int gpat_debug_gradients(gpat_handle h, const float* f8, float* out32);

/* Push every in-box particle exactly nsteps adaptive steps (inner loop body of
 * particle_mover_one_cycle, particle_module.f90:1600-1704, without the end-of-interval
 * fix-up).  Used for per-step parity and steady-state throughput. */
int gpat_debug_push_n(gpat_handle h, double t0, double dtf, int nsteps, uint64_t* steps_done);

/* Interpolated fields(1:32) at given positions/times for the current field
 * store (interp_fields, mhd_data_parallel.f90:1751-1793); slots the push does not
 * use on this configuration are returned as 0. */
int gpat_debug_interp(gpat_handle h, int64_t n, const double* x, const double* y,
                      const double* z, const double* rt, double* fields32);

#ifdef __cplusplus
}
#endif
#endif /* GPAT_CUDA_H */
