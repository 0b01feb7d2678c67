/* TEST DOUBLE of the C ABI (include/gpat_cuda.h) -- CPU tests of the HOST drivers only.
 *
 * tests/test_cpu_cpp_driver.py builds this into tests/abi_mock/_build/libgpat_cuda.so and puts that
 * directory on LD_LIBRARY_PATH of host/gpat_driver, so that the C++ driver's own logic (switches,
 * conf.dat, frame / map / surface / tag files, call order, quick.dat and spectrum files) can be checked
 * without a GPU against stochastic_parker_b200.run_intervals driving the same oracle.  Every entry point
 * forwards to the CPU oracle (oracle/gpat_oracle.c).  It is never built by __graft_entry__.build(), never
 * shipped next to the product and never loaded by the package: the product library has no CPU path and
 * fails loudly without a CUDA device (tests/test_cpu_host.py::test_no_cuda_device_fails_loudly). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gpat_cuda.h"

typedef struct orc_sim orc_sim;
orc_sim* orc_create(const gpat_params* p, int64_t nptl_max);
void orc_set_params(orc_sim* S, const gpat_params* p);
void orc_destroy(orc_sim* S);
void orc_set_fields(orc_sim* S, int slot, const float* f, int nvar, int with_grad);
void orc_calc_gradients(orc_sim* S, int slot);
void orc_set_acc_surface(orc_sim* S, int which, int slot, const double* heights);
void orc_set_turbulence(orc_sim* S, int which, int slot, const float* data);
void orc_copy_fields(orc_sim* S);
void orc_particle_mover(orc_sim* S, double t0, double dtf, int nsteps_interval, int num_fine_steps,
                        int dump_escaped_dist, uint64_t* steps_done);
void orc_init_tracking(orc_sim* S, const int32_t* tags, int ncols, int64_t nptl_tracking, int nsteps_interval);
void orc_reset_tracked(orc_sim* S);
int64_t orc_get_tracked(const orc_sim* S, gpat_particle* out, int64_t* nsteps_max);
void orc_inject_uniform(orc_sim* S, int64_t nptl, double dt, int dist_flag, double particle_v0, double t_frame,
                        double dt_mhd, const double part_box[6], double power_index);
int64_t orc_inject_at_shock(orc_sim* S, int64_t nptl, double dt, int dist_flag, double particle_v0,
                            double t_frame, double power_index);
int64_t orc_ncells_large(const orc_sim* S, int mode, double vmin, const double part_box[6]);
int64_t orc_inject_targeted(orc_sim* S, int mode, int64_t nptl, double dt, int dist_flag, double particle_v0,
                            double t_frame, double dt_mhd, const double part_box[6], double power_index,
                            int inject_same_nptl, double vmin, int64_t ncells_norm);
void orc_split(orc_sim* S, double split_ratio, double pmin_split, int nsteps_interval);
void orc_hist_edges(const orc_sim* S, int which, double* pedges, double* muedges);
void orc_diagnostics(const orc_sim* S, int local_dist, double* fglobal, double* const flocal[4], double quick[8],
                     double* pmax_out);
void orc_escaped_diagnostics(const orc_sim* S, double* fescaped);
void orc_escaped_local_diagnostics(const orc_sim* S, double* const fx[4], double* const fy[4], double* const fz[4]);
int64_t orc_get_particles(const orc_sim* S, gpat_particle* out, int64_t nmax);
void orc_set_particles(orc_sim* S, const gpat_particle* in, int64_t n);
int64_t orc_get_escaped(const orc_sim* S, gpat_particle* out, int64_t nmax);
void orc_reset_escaped(orc_sim* S);
void orc_get_counters(const orc_sim* S, gpat_counters* c);
void orc_set_counters(orc_sim* S, const gpat_counters* c);

struct gpat_sim {
    orc_sim* S;
};
#define SIM(h) (((struct gpat_sim*)(h))->S)

int gpat_init(gpat_handle* h, int device, int64_t nptl_max, const gpat_params* params)
{
    (void)device;
    struct gpat_sim* g = (struct gpat_sim*)calloc(1, sizeof(*g));
    g->S = orc_create(params, nptl_max);
    *h = (gpat_handle)g;
    return GPAT_OK;
}
int gpat_set_params(gpat_handle h, const gpat_params* p) { orc_set_params(SIM(h), p); return GPAT_OK; }
int gpat_finalize(gpat_handle h)
{
    if (!h) return GPAT_OK;
    orc_destroy(SIM(h));
    free(h);
    return GPAT_OK;
}
const char* gpat_last_error(gpat_handle h) { (void)h; return "(test double)"; }
int gpat_upload_fields(gpat_handle h, int slot, const float* f, int nvar, int with_grad)
{
    orc_set_fields(SIM(h), slot, f, nvar, with_grad);
    if (!with_grad) orc_calc_gradients(SIM(h), slot);
    return GPAT_OK;
}
int gpat_upload_turbulence(gpat_handle h, int which, int slot, const float* data)
{
    orc_set_turbulence(SIM(h), which, slot, data);
    return GPAT_OK;
}
int gpat_upload_acc_surface(gpat_handle h, int which, int slot, const double* heights)
{
    orc_set_acc_surface(SIM(h), which, slot, heights);
    return GPAT_OK;
}
int gpat_prefetch_fields(gpat_handle h, const float* f, int nvar) { (void)h; (void)f; (void)nvar; return GPAT_OK; }
int gpat_swap_fields(gpat_handle h) { orc_copy_fields(SIM(h)); return GPAT_OK; }
int gpat_inject_uniform(gpat_handle h, int64_t nptl, double dt, int dist_flag, double particle_v0, double t_frame,
                        double dt_mhd, const double part_box[6], double power_index)
{
    orc_inject_uniform(SIM(h), nptl, dt, dist_flag, particle_v0, t_frame, dt_mhd, part_box, power_index);
    return GPAT_OK;
}
int gpat_inject_targeted(gpat_handle h, int mode, int64_t nptl, double dt, int dist_flag, double particle_v0,
                         double t_frame, double dt_mhd, const double part_box[6], double power_index,
                         int inject_same_nptl, double vmin, int64_t ncells_norm, int64_t* nptl_injected,
                         int64_t* ncells)
{
    int64_t nc = orc_ncells_large(SIM(h), mode, vmin, part_box);
    int64_t ni = orc_inject_targeted(SIM(h), mode, nptl, dt, dist_flag, particle_v0, t_frame, dt_mhd, part_box,
                                     power_index, inject_same_nptl, vmin, ncells_norm);
    if (ncells) *ncells = nc;
    if (nptl_injected) *nptl_injected = ni;
    return GPAT_OK;
}
int gpat_inject_at_shock(gpat_handle h, int64_t nptl, double dt, int dist_flag, double particle_v0, double t_frame,
                         double power_index)
{
    orc_inject_at_shock(SIM(h), nptl, dt, dist_flag, particle_v0, t_frame, power_index);
    return GPAT_OK;
}
int gpat_init_tracking(gpat_handle h, const int32_t* tags, int ncols, int64_t nptl_tracking, int nsteps_interval)
{
    orc_init_tracking(SIM(h), tags, ncols, nptl_tracking, nsteps_interval);
    return GPAT_OK;
}
int gpat_tracked_shape(gpat_handle h, int64_t* nsteps_tracking_max, int64_t* nptl_tracking)
{
    int64_t nmax = 0;
    int64_t n = orc_get_tracked(SIM(h), NULL, &nmax);
    if (nsteps_tracking_max) *nsteps_tracking_max = nmax;
    if (nptl_tracking) *nptl_tracking = n;
    return GPAT_OK;
}
int gpat_download_tracked(gpat_handle h, gpat_particle* out)
{
    int64_t nmax = 0;
    orc_get_tracked(SIM(h), out, &nmax);
    return GPAT_OK;
}
int gpat_reset_tracked(gpat_handle h) { orc_reset_tracked(SIM(h)); return GPAT_OK; }
int gpat_particle_mover(gpat_handle h, double t0, double dtf, int nsteps_interval, int num_fine_steps,
                        int dump_escaped_dist, uint64_t* steps_done)
{
    orc_particle_mover(SIM(h), t0, dtf, nsteps_interval, num_fine_steps, dump_escaped_dist, steps_done);
    return GPAT_OK;
}
int gpat_split(gpat_handle h, double split_ratio, double pmin_split, int nsteps_interval)
{
    orc_split(SIM(h), split_ratio, pmin_split, nsteps_interval);
    return GPAT_OK;
}
int gpat_download_particles(gpat_handle h, gpat_particle* out, int64_t nmax, int64_t* n)
{
    int64_t k = orc_get_particles(SIM(h), out, nmax);
    if (n) *n = k;
    return GPAT_OK;
}
int gpat_upload_particles(gpat_handle h, const gpat_particle* in, int64_t n) { orc_set_particles(SIM(h), in, n); return GPAT_OK; }
int gpat_download_escaped(gpat_handle h, gpat_particle* out, int64_t nmax, int64_t* n)
{
    int64_t k = orc_get_escaped(SIM(h), out, nmax);
    if (n) *n = k;
    return GPAT_OK;
}
int gpat_reset_escaped(gpat_handle h) { orc_reset_escaped(SIM(h)); return GPAT_OK; }
int gpat_get_counters(gpat_handle h, gpat_counters* c) { orc_get_counters(SIM(h), c); return GPAT_OK; }
int gpat_set_counters(gpat_handle h, const gpat_counters* c) { orc_set_counters(SIM(h), c); return GPAT_OK; }
int gpat_diagnostics(gpat_handle h, int local_dist, double* fglobal, double* const flocal[4], double quick[8],
                     double* pmax)
{
    orc_diagnostics(SIM(h), local_dist, fglobal, flocal, quick, pmax);
    return GPAT_OK;
}
int gpat_escaped_diagnostics(gpat_handle h, double* fescaped) { orc_escaped_diagnostics(SIM(h), fescaped); return GPAT_OK; }
int gpat_escaped_local_diagnostics(gpat_handle h, double* const fx[4], double* const fy[4], double* const fz[4])
{
    orc_escaped_local_diagnostics(SIM(h), fx, fy, fz);
    return GPAT_OK;
}
int gpat_hist_edges(gpat_handle h, int which, double* pedges, double* muedges)
{
    orc_hist_edges(SIM(h), which, pedges, muedges);
    return GPAT_OK;
}
