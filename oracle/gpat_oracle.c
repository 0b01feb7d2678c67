/*
 * gpat_oracle.c -- CPU restatement of GPAT's Parker-transport particle path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call it.  The product (stochastic_parker_b200/) never does.
 *
 * PARITY: PINNED TO THE REFERENCE'S OWN FORTRAN, EXECUTED -- NOT COMPILED.  The reference
 * (xiaocanli/stochastic-parker) ships no tests, golden vectors or known-answer files for this
 * path and cannot be built here or on the B200 box (no Fortran front-end on either:
 * profiles/r02a_fortran_probe.log; it also needs MPI + HDF5 + FoBiS/FLAP + mt_stream_f90-1.11),
 * so there is no oracle/_ref binary.  Instead oracle/f90/f90run.py executes the UNMODIFIED text
 * of the reference's procedures (particle_mover, particle_mover_one_cycle, every push_particle_*,
 * both kappa routines, calc_dpp_*, interp_fields, calc_fields_gradients, the injectors,
 * remove_particles, split_particle, the tracking hooks, calc_particle_distributions,
 * calc_escaped_distributions, quick_check ...) with Fortran's kind / promotion / literal /
 * evaluation-order rules, and tests/golden/ref_f90/*.npz hold what they compute for 18 switch
 * combinations (1-D, 2-D, 3-D, NLGC, D_pp, focused transport, open boundaries with escapes,
 * splitting).  tests/test_cpu_reference_f90.py holds THIS FILE to those vectors BIT FOR BIT
 * (particles, counters, every histogram), live against /root/reference where it exists.
 * What that does not cover: a compiler's freedom to contract a*b+c into an FMA or to vectorise
 * a reduction (the interpreter evaluates strictly, like gfortran -O2 without -ffast-math on a
 * target without FMA contraction), and the 2-D / 3-D shock injector, which reads uninitialised
 * variables in the reference (tests/test_cpu_reference_f90.py proves it by execution).
 * This file restates the cited Fortran line by line: FP64 throughout, the same operation order
 * (Fortran left-to-right association), the same default-real (FP32) literals.  Build the parity
 * copy with `-O2 -ffp-contract=off` so the compiler keeps that order.
 *
 * Third-party arithmetic that is NOT restated: mt_stream_f90-1.11 (multiple-stream
 * MT19937, random_number_generator.f90:9,35-44,100).  north_star defines parity
 * as "fed the same pre-generated random increments", so uniforms are an INPUT
 * here: either a table, or Philox4x32-10 keyed per particle (the stream the GPU
 * library defines; specification in DESIGN.md "RNG"), or -- oracle only, for the
 * statistical comparison -- one sequential MT19937 seeded with the reference's
 * seed array (pinned by mt19937ar's published output).
 *
 * All file:line citations are relative to /root/reference/src/modules/ with
 * PM = particle_module.f90, MD = mhd_data_parallel.f90, DG = diagnostics.f90.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/gpat_cuda.h" /* POD structs only (gpat_params, gpat_particle) */

#define NFIELDS 8
#define NGRADS 24
#define NVAR (NFIELDS + NGRADS)

struct orc_sim;
typedef struct orc_sim {
    gpat_params P;
    int nxg, nyg, nzg;        /* array extents with ghosts: nx+4, ny+4, nz+4 | 1 */
    float* farray1;           /* (32, nxg, nyg, nzg), MD:35,82-102 */
    float* farray2;
    /* sigma2_slab, sigma2_2d, lc_slab, lc_2d (MD:36-41, 107-182): (4 arrays x 4 comps, cells),
     * comp 0 the value, 1..3 its gradients; allocated lazily, 1.0 everywhere like the reference */
    float* aux1;
    float* aux2;
    /* acc_surface11/12, 21/22 (acc_region_surface.f90:12-13): [which][slot], real(dp), (-1:n1+2, -1:n2+2) */
    double* surf[2][2];
    gpat_particle* ptls;      /* PM:65 */
    gpat_particle* escaped;   /* PM:134 */
    int64_t nptl_current, nptl_old, nptl_max, nptl_split, nptl_inject;
    int64_t nptl_escaped, nptl_escaped_max;
    int64_t tag_max;
    double leak, leak_negp;
    double dt_min, dt_max;
    int neighbors[6];         /* SS:299-342 for a 1x1x1 topology */
    const double* rng_table;  /* optional uniform table */
    int64_t rng_slots, rng_max_steps;
    uint64_t steps;           /* push_particle_* calls */
    /* ORC_RNG_MT19937: one sequential MT19937 like the reference's thread-0 stream (RNG:28-101) */
    uint32_t mt[624];
    int mti;
    /* particle tracking, PM:144-150 */
    int track_particle_flag;
    int split_times_max;
    int64_t nptl_tracking, nsteps_tracking_max;
    int32_t* tags_tracking;          /* (split_times_max+2, nptl_tracking), column-major */
    gpat_particle* particles_tracked; /* (nsteps_tracking_max, nptl_tracking), column-major */
} orc_sim;

static inline double sq(double x) { return x * x; }
static int check_above_acc_surface(const struct orc_sim* S, double x, double y, double z, double h1, double h2);
static void track_after_push(orc_sim* S, gpat_particle* ptl); /* particle tracking, below */

/* ------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11), written from the published algorithm. */
/* ------------------------------------------------------------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr, key, out);
}

/* u32 -> [0,1] with 32-bit resolution (the genrand_real1 convention) */
static inline double u01(uint32_t w) { return (double)w / 4294967295.0; }

static inline uint64_t get_rng_step(const gpat_particle* p)
{
    uint64_t s;
    memcpy(&s, &p->padding, 8);
    return s;
}
static inline void set_rng_step(gpat_particle* p, uint64_t s) { memcpy(&p->padding, &s, 8); }

/* ------------------------------------------------------------------------ */
/* MT19937 (Matsumoto & Nishimura 1998, init_by_array + genrand_int32),       */
/* written from the published algorithm.  ORC_RNG_MT19937 is an ORACLE-ONLY   */
/* mode: a single sequential stream seeded like random_number_generator.f90   */
/* seeds mt_stream (iseeda = Z'123',Z'234',Z'345',Z'456', RNG:16,38) and       */
/* mapped to [0,1] with 32-bit resolution (genrand_real1).  It is NOT claimed */
/* to equal mt_stream's jump-ahead sub-streams; it exists so that the GPU's   */
/* Philox statistics can be compared with a run driven by the reference's     */
/* generator family (tests: Poisson agreement).  Serial: one thread.          */
/* ------------------------------------------------------------------------ */
#define ORC_RNG_MT19937 2
static void mt_init_genrand(orc_sim* S, uint32_t s)
{
    S->mt[0] = s;
    for (S->mti = 1; S->mti < 624; S->mti++)
        S->mt[S->mti] = 1812433253u * (S->mt[S->mti - 1] ^ (S->mt[S->mti - 1] >> 30)) + (uint32_t)S->mti;
}
static void mt_init_by_array(orc_sim* S, const uint32_t* key, int klen)
{
    mt_init_genrand(S, 19650218u);
    int i = 1, j = 0;
    for (int k = (624 > klen ? 624 : klen); k; k--) {
        S->mt[i] = (S->mt[i] ^ ((S->mt[i - 1] ^ (S->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { S->mt[0] = S->mt[623]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (int k = 623; k; k--) {
        S->mt[i] = (S->mt[i] ^ ((S->mt[i - 1] ^ (S->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { S->mt[0] = S->mt[623]; i = 1; }
    }
    S->mt[0] = 0x80000000u;
}
static uint32_t mt_genrand_int32(orc_sim* S)
{
    static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
    uint32_t y;
    if (S->mti >= 624) {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (S->mt[kk] & 0x80000000u) | (S->mt[kk + 1] & 0x7fffffffu);
            S->mt[kk] = S->mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < 623; kk++) {
            y = (S->mt[kk] & 0x80000000u) | (S->mt[kk + 1] & 0x7fffffffu);
            S->mt[kk] = S->mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (S->mt[623] & 0x80000000u) | (S->mt[0] & 0x7fffffffu);
        S->mt[623] = S->mt[396] ^ (y >> 1) ^ mag01[y & 1u];
        S->mti = 0;
    }
    y = S->mt[S->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
uint32_t orc_mt19937_next(orc_sim* S) { return mt_genrand_int32(S); }

/* Four uniforms of one push step: ran1, ran2, ran3, ran_p (PM:3548-3550,3589). */
static void step_uniforms(const orc_sim* S, const gpat_particle* ptl, double u[4])
{
    uint64_t step = get_rng_step(ptl);
    if (S->P.rng_mode == ORC_RNG_MT19937) {
        orc_sim* W = (orc_sim*)S; /* the generator state advances */
        for (int j = 0; j < 4; ++j) u[j] = u01(mt_genrand_int32(W));
        return;
    }
    if (S->P.rng_mode == GPAT_RNG_TABLE && S->rng_table) {
        int64_t slot = abs(ptl->tag_injected);
        if (slot >= S->rng_slots || (int64_t)step >= S->rng_max_steps) {
            u[0] = u[1] = u[2] = u[3] = 0.5;
            return;
        }
        const double* t = S->rng_table + ((size_t)slot * S->rng_max_steps + step) * 4;
        u[0] = t[0]; u[1] = t[1]; u[2] = t[2]; u[3] = t[3];
        return;
    }
    /* tracked particles carry NEGATED tags (PM:436-437, 5453-5472); the stream is keyed by the
     * magnitudes so that a tracking run replays the trajectories of the run it was selected from */
    uint32_t ctr[4] = {(uint32_t)step, (uint32_t)(step >> 32), (uint32_t)abs(ptl->tag_injected),
                       (uint32_t)abs(ptl->tag_splitted)};
    uint32_t key[2] = {(uint32_t)S->P.seed, (uint32_t)(S->P.seed >> 32) + (uint32_t)ptl->origin};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    for (int j = 0; j < 4; ++j) u[j] = u01(o[j]);
}

/* The fifth uniform of a step (push_particle_2d_include_3rd_ft / _3d_ft draw ran1..ran3, then one
 * for p and one for mu: PM:4521-4523, 4573-4576): first word of a SECOND Philox block of the same
 * step, counter word 1 with its top bit flipped.  MT19937 mode: the next draw.  Table mode: the
 * tables hold four uniforms per step, so these pushers are rejected there (gpat_init). */
static double step_uniform5(const orc_sim* S, const gpat_particle* ptl)
{
    if (S->P.rng_mode == ORC_RNG_MT19937) return u01(mt_genrand_int32((orc_sim*)S));
    uint64_t step = get_rng_step(ptl);
    uint32_t ctr[4] = {(uint32_t)step, (uint32_t)(step >> 32) ^ 0x80000000u, (uint32_t)abs(ptl->tag_injected),
                       (uint32_t)abs(ptl->tag_splitted)};
    uint32_t key[2] = {(uint32_t)S->P.seed, (uint32_t)(S->P.seed >> 32) + (uint32_t)ptl->origin};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    return u01(o[0]);
}

/* Sequential uniform reader for injection: word k of the stream
 * ctr = (k/4, 0, tag_injected, 0), same key as above. */
typedef struct inj_stream {
    const orc_sim* S;
    uint32_t tag, origin;
    uint32_t k;
    uint32_t buf[4];
} inj_stream;

static double inj_next(inj_stream* s)
{
    if (s->S->P.rng_mode == ORC_RNG_MT19937) return u01(mt_genrand_int32((orc_sim*)s->S));
    if ((s->k & 3u) == 0) {
        uint32_t ctr[4] = {s->k >> 2, 0u, s->tag, 0u};
        uint32_t key[2] = {(uint32_t)s->S->P.seed, (uint32_t)(s->S->P.seed >> 32) + s->origin};
        philox4x32_10(ctr, key, s->buf);
    }
    double u = u01(s->buf[s->k & 3u]);
    s->k++;
    return u;
}

/* ------------------------------------------------------------------------ */
/* life cycle                                                                */
/* ------------------------------------------------------------------------ */
static void set_neighbors(orc_sim* S)
{
    /* SS:267-272 (msize == 1): periodic -> self (rank 0), open -> -1 */
    for (int d = 0; d < 3; ++d) {
        int n = (S->P.pbc[d] == 0) ? 0 : -1;
        S->neighbors[2 * d] = n;
        S->neighbors[2 * d + 1] = n;
    }
}

orc_sim* orc_create(const gpat_params* p, int64_t nptl_max)
{
    orc_sim* S = (orc_sim*)calloc(1, sizeof(orc_sim));
    S->P = *p;
    S->nxg = p->nx + 4;
    S->nyg = (p->ndim > 1) ? p->ny + 4 : p->ny;
    S->nzg = (p->ndim > 2) ? p->nz + 4 : p->nz;
    size_t n = (size_t)NVAR * S->nxg * S->nyg * S->nzg;
    S->farray1 = (float*)calloc(n, sizeof(float));
    S->farray2 = (float*)calloc(n, sizeof(float));
    S->nptl_max = nptl_max;
    S->ptls = (gpat_particle*)calloc((size_t)nptl_max + 1, sizeof(gpat_particle)); /* PM:171-194 */
    S->nptl_escaped_max = nptl_max;
    S->escaped = (gpat_particle*)calloc((size_t)nptl_max + 1, sizeof(gpat_particle));
    set_neighbors(S);
    {
        const uint32_t iseeda[4] = {0x123u, 0x234u, 0x345u, 0x456u}; /* RNG:16 */
        mt_init_by_array(S, iseeda, 4);
    }
    return S;
}

void orc_set_params(orc_sim* S, const gpat_params* p)
{
    S->P = *p;
    set_neighbors(S);
}

void orc_destroy(orc_sim* S)
{
    if (!S) return;
    free(S->farray1); free(S->farray2); free(S->ptls); free(S->escaped);
    free(S->aux1); free(S->aux2);
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) free(S->surf[a][b]);
    free(S->tags_tracking); free(S->particles_tracked);
    free(S);
}

void orc_set_rng_table(orc_sim* S, const double* u, int64_t nslots, int64_t max_steps)
{
    S->rng_table = u; S->rng_slots = nslots; S->rng_max_steps = max_steps;
}

/* ------------------------------------------------------------------------ */
/* fields: load, gradients (MD:504-605), copy (MD:1920)                      */
/* ------------------------------------------------------------------------ */
#define FIDX(S, v, i, j, k) ((size_t)(v) + (size_t)NVAR * ((size_t)(i) + (size_t)(S)->nxg * ((size_t)(j) + (size_t)(S)->nyg * (size_t)(k))))

/* f: (nvar, nxg, nyg, nzg); copies the 8 primaries (and 24 gradients if with_grad) */
void orc_set_fields(orc_sim* S, int slot, const float* f, int nvar, int with_grad)
{
    float* fa = slot ? S->farray2 : S->farray1;
    size_t ncell = (size_t)S->nxg * S->nyg * S->nzg;
    int ncopy = (with_grad && nvar == NVAR) ? NVAR : NFIELDS;
    for (size_t c = 0; c < ncell; ++c)
        for (int v = 0; v < ncopy; ++v) fa[c * NVAR + v] = f[c * nvar + v];
}

void orc_calc_gradients(orc_sim* S, int slot)
{
    float* fa = slot ? S->farray2 : S->farray1;
    const double idxh = 0.5 / S->P.dx; /* MD:512-514 */
    const double idyh = 0.5 / S->P.dy;
    const double idzh = 0.5 / S->P.dz;
    const int nxg = S->nxg, nyg = S->nyg, nzg = S->nzg;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < nzg; ++k)
        for (int j = 0; j < nyg; ++j)
            for (int i = 0; i < nxg; ++i)
                for (int v = 0; v < NFIELDS; ++v) {
                    float g;
                    /* d/dx -> slot nfields+1+3v, MD:535-542 */
                    if (i == 0) {
                        float a = -3.0f * fa[FIDX(S, v, 0, j, k)];
                        float b = 4.0f * fa[FIDX(S, v, 1, j, k)];
                        float s = a + b;
                        g = s - fa[FIDX(S, v, 2, j, k)];
                    } else if (i == nxg - 1) {
                        float a = 3.0f * fa[FIDX(S, v, nxg - 1, j, k)];
                        float b = 4.0f * fa[FIDX(S, v, nxg - 2, j, k)];
                        float s = a - b;
                        g = s + fa[FIDX(S, v, nxg - 3, j, k)];
                    } else {
                        g = fa[FIDX(S, v, i + 1, j, k)] - fa[FIDX(S, v, i - 1, j, k)];
                    }
                    fa[FIDX(S, NFIELDS + 3 * v + 0, i, j, k)] = (float)((double)g * idxh);
                    /* d/dy, MD:545-554 (only if the y extent > 1) */
                    if (nyg > 1) {
                        if (j == 0) {
                            float a = -3.0f * fa[FIDX(S, v, i, 0, k)];
                            float b = 4.0f * fa[FIDX(S, v, i, 1, k)];
                            float s = a + b;
                            g = s - fa[FIDX(S, v, i, 2, k)];
                        } else if (j == nyg - 1) {
                            float a = 3.0f * fa[FIDX(S, v, i, nyg - 1, k)];
                            float b = 4.0f * fa[FIDX(S, v, i, nyg - 2, k)];
                            float s = a - b;
                            g = s + fa[FIDX(S, v, i, nyg - 3, k)];
                        } else {
                            g = fa[FIDX(S, v, i, j + 1, k)] - fa[FIDX(S, v, i, j - 1, k)];
                        }
                        fa[FIDX(S, NFIELDS + 3 * v + 1, i, j, k)] = (float)((double)g * idyh);
                    }
                    /* d/dz, MD:557-566 */
                    if (nzg > 1) {
                        if (k == 0) {
                            float a = -3.0f * fa[FIDX(S, v, i, j, 0)];
                            float b = 4.0f * fa[FIDX(S, v, i, j, 1)];
                            float s = a + b;
                            g = s - fa[FIDX(S, v, i, j, 2)];
                        } else if (k == nzg - 1) {
                            float a = 3.0f * fa[FIDX(S, v, i, j, nzg - 1)];
                            float b = 4.0f * fa[FIDX(S, v, i, j, nzg - 2)];
                            float s = a - b;
                            g = s + fa[FIDX(S, v, i, j, nzg - 3)];
                        } else {
                            g = fa[FIDX(S, v, i, j, k + 1)] - fa[FIDX(S, v, i, j, k - 1)];
                        }
                        fa[FIDX(S, NFIELDS + 3 * v + 2, i, j, k)] = (float)((double)g * idzh);
                    }
                }
}

/* ------------------------------------------------------------------------ */
/* acceleration surfaces: acc_region_surface.f90 (3-D only)                   */
/* ------------------------------------------------------------------------ */
static void surf_dims(const orc_sim* S, int norm, int* n1, int* n2)
{
    int axis = abs(norm) - 1; /* 0 x, 1 y, 2 z */
    if (axis == 0) { *n1 = S->P.ny + 4; *n2 = S->P.nz + 4; }      /* ARS:34-35 */
    else if (axis == 1) { *n1 = S->P.nx + 4; *n2 = S->P.nz + 4; }
    else { *n1 = S->P.nx + 4; *n2 = S->P.ny + 4; }
}

void orc_set_acc_surface(orc_sim* S, int which, int slot, const double* heights)
{
    int n1, n2;
    surf_dims(S, which ? S->P.surface_norm2 : S->P.surface_norm1, &n1, &n2);
    free(S->surf[which][slot]);
    S->surf[which][slot] = (double*)malloc(sizeof(double) * (size_t)n1 * n2);
    memcpy(S->surf[which][slot], heights, sizeof(double) * (size_t)n1 * n2);
}

/* interp_acc_surface, ARS:255-334 */
static void interp_acc_surface(const orc_sim* S, const int pos[3], const double w[8], double rt, double* h1,
                               double* h2)
{
    const gpat_params* P = &S->P;
    int i1 = 0, j1 = 0;
    double w2[4]; /* weights_2d(1,1) (2,1) (1,2) (2,2) */
    int cur_axis = -1;
    double out[2] = {0.0, 0.0};
    for (int k = 0; k < 2; ++k) {
        if (k == 1 && !P->surface2_existed) { out[1] = 0.0; break; }
        int norm = k ? P->surface_norm2 : P->surface_norm1;
        int axis = abs(norm) - 1;
        if (axis != cur_axis) { /* the second surface reuses i1, j1, weights_2d when the axes agree */
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
        int n1, n2;
        surf_dims(S, norm, &n1, &n2);
        int a = i1 + 1, b = j1 + 1; /* Fortran lower bound -1 */
        if (a < 0) a = 0;
        if (a > n1 - 2) a = n1 - 2;
        if (b < 0) b = 0;
        if (b > n2 - 2) b = n2 - 2;
        double hh[2] = {0.0, 0.0};
        for (int slot = 0; slot < (P->time_interp ? 2 : 1); ++slot) {
            const double* sf = S->surf[k][slot];
            double acc = 0.0; /* sum() over the 2x2 section in array-element order */
            acc = acc + sf[a + (size_t)n1 * b] * w2[0];
            acc = acc + sf[(a + 1) + (size_t)n1 * b] * w2[1];
            acc = acc + sf[a + (size_t)n1 * (b + 1)] * w2[2];
            acc = acc + sf[(a + 1) + (size_t)n1 * (b + 1)] * w2[3];
            hh[slot] = acc;
        }
        out[k] = P->time_interp ? hh[0] * (1.0 - rt) + hh[1] * rt : hh[0];
    }
    *h1 = out[0];
    *h2 = out[1];
}

/* check_above_acc_surface, ARS:342-388 */
static int check_above_acc_surface(const orc_sim* S, double x, double y, double z, double h1, double h2)
{
    const gpat_params* P = &S->P;
    const double c[3] = {x, y, z};
    double ph = c[abs(P->surface_norm1) - 1];
    int in = (P->surface_norm1 > 0) ? (ph > h1) : (ph < h1);
    if (P->surface2_existed) {
        ph = c[abs(P->surface_norm2) - 1];
        int in2 = (P->surface_norm2 > 0) ? (ph > h2) : (ph < h2);
        in = P->is_intersection ? (in && in2) : (in || in2);
    }
    return in;
}

/* ------------------------------------------------------------------------ */
/* turbulence maps: read_magnetic_fluctuation / read_correlation_length        */
/* (MD:306-497: the file holds the slab array then the 2-D array),             */
/* calc_grad_sigma2_slab/_2d, calc_grad_lc_slab/_2d (MD:771-1604: the same     */
/* FP32-difference x FP64 0.5/dx arithmetic as the fields)                     */
/* ------------------------------------------------------------------------ */
#define NAUX 16
#define AIDX(S, a, c, i, j, k) ((size_t)((a) * 4 + (c)) + (size_t)NAUX * ((size_t)(i) + (size_t)(S)->nxg * ((size_t)(j) + (size_t)(S)->nyg * (size_t)(k))))
static float* aux_of(orc_sim* S, int slot)
{
    float** pa = slot ? &S->aux2 : &S->aux1;
    if (!*pa) {
        size_t n = (size_t)NAUX * S->nxg * S->nyg * S->nzg;
        *pa = (float*)malloc(n * sizeof(float));
        for (size_t i = 0; i < n; ++i) (*pa)[i] = 1.0f; /* MD:125-126, 165-166 */
    }
    return *pa;
}

void orc_set_turbulence(orc_sim* S, int which, int slot, const float* data)
{
    float* ax = aux_of(S, slot);
    const int nxg = S->nxg, nyg = S->nyg, nzg = S->nzg;
    const size_t ncell = (size_t)nxg * nyg * nzg;
    const double idh[3] = {0.5 / S->P.dx, 0.5 / S->P.dy, 0.5 / S->P.dz};
    for (int t = 0; t < 2; ++t) { /* slab, then 2-D */
        const int a = which * 2 + t;
        const float* src = data + (size_t)t * ncell;
        for (int k = 0; k < nzg; ++k)
            for (int j = 0; j < nyg; ++j)
                for (int i = 0; i < nxg; ++i)
                    ax[AIDX(S, a, 0, i, j, k)] = src[(size_t)i + (size_t)nxg * ((size_t)j + (size_t)nyg * k)];
        for (int k = 0; k < nzg; ++k)
            for (int j = 0; j < nyg; ++j)
                for (int i = 0; i < nxg; ++i)
                    for (int d = 0; d < 3; ++d) {
                        const int n = d == 0 ? nxg : (d == 1 ? nyg : nzg);
                        const int pos = d == 0 ? i : (d == 1 ? j : k);
                        if (n <= 1) continue; /* `if (uny > lny)`: the 1.0 fill stays */
                        int c[3] = {i, j, k};
#define AV(o) (c[d] = pos + (o), ax[AIDX(S, a, 0, c[0], c[1], c[2])])
                        float g;
                        if (pos == 0) {
                            float p0 = -3.0f * AV(0), p1 = 4.0f * AV(1);
                            g = (p0 + p1) - AV(2);
                        } else if (pos == n - 1) {
                            float p0 = 3.0f * AV(0), p1 = 4.0f * AV(-1);
                            g = (p0 - p1) + AV(-2);
                        } else {
                            float hi = AV(1), lo = AV(-1);
                            g = hi - lo;
                        }
#undef AV
                        ax[AIDX(S, a, 1 + d, i, j, k)] = (float)((double)g * idh[d]);
                    }
    }
}

void orc_get_fields(const orc_sim* S, int slot, float* out32)
{
    const float* fa = slot ? S->farray2 : S->farray1;
    memcpy(out32, fa, sizeof(float) * (size_t)NVAR * S->nxg * S->nyg * S->nzg);
}

void orc_copy_fields(orc_sim* S) /* MD:1920-1923; copy_magnetic_fluctuation / _correlation_length MD:1928-1941 */
{
    memcpy(S->farray1, S->farray2, sizeof(float) * (size_t)NVAR * S->nxg * S->nyg * S->nzg);
    if (S->aux2) memcpy(aux_of(S, 0), S->aux2, sizeof(float) * (size_t)NAUX * S->nxg * S->nyg * S->nzg);
    for (int k = 0; k < 2; ++k) /* copy_acc_surface, ARS:390-396 */
        if (S->surf[k][1]) {
            int n1, n2;
            surf_dims(S, k ? S->P.surface_norm2 : S->P.surface_norm1, &n1, &n2);
            if (!S->surf[k][0]) S->surf[k][0] = (double*)malloc(sizeof(double) * (size_t)n1 * n2);
            memcpy(S->surf[k][0], S->surf[k][1], sizeof(double) * (size_t)n1 * n2);
        }
}

/* interp_magnetic_fluctuation + interp_correlation_length (MD:1806-1915): 16 values =
 * db2_slab(1:4), db2_2d(1:4), lc_slab(1:4), lc_2d(1:4) */
static void interp_aux(const orc_sim* S, const int pos[3], const double w[8], double rt, double out[NAUX])
{
    double o2[NAUX];
    const int ze = (S->P.ndim > 2) ? 1 : 0, ye = (S->P.ndim > 1) ? 1 : 0;
    for (int v = 0; v < NAUX; ++v) { out[v] = 0.0; o2[v] = 0.0; }
    for (int k = 0; k <= ze; ++k)
        for (int j = 0; j <= ye; ++j)
            for (int i = 0; i <= 1; ++i) {
                int idx = k * 4 + j * 2 + i;
                int c0 = pos[0] + 1, c1 = pos[1] + 1, c2 = pos[2] + 1;
                if (c0 < 0) c0 = 0;
                if (c0 > S->nxg - 2) c0 = S->nxg - 2;
                if (c1 < 0) c1 = 0;
                if (c1 > S->nyg - 2) c1 = S->nyg - 2;
                if (c2 < 0) c2 = 0;
                if (c2 > S->nzg - 2) c2 = S->nzg - 2;
                int ci = c0 + i, cj = (S->P.ndim > 1) ? c1 + j : 0, ck = (S->P.ndim > 2) ? c2 + k : 0;
                const float* a1 = S->aux1 + AIDX(S, 0, 0, ci, cj, ck);
                for (int v = 0; v < NAUX; ++v) out[v] = out[v] + (double)a1[v] * w[idx];
                if (S->P.time_interp) {
                    const float* a2 = S->aux2 + AIDX(S, 0, 0, ci, cj, ck);
                    for (int v = 0; v < NAUX; ++v) o2[v] = o2[v] + (double)a2[v] * w[idx];
                }
            }
    if (S->P.time_interp) {
        double rt1 = 1.0 - rt;
        for (int v = 0; v < NAUX; ++v) out[v] = out[v] * rt1 + o2[v] * rt;
    }
}

/* ------------------------------------------------------------------------ */
/* interpolation: get_interp_paramters (PM:642-674), interp_fields (MD:1751)  */
/* ------------------------------------------------------------------------ */
static void get_interp_parameters(const orc_sim* S, double px, double py, double pz, int pos[3],
                                  double w[8])
{
    double rx, ry, rz;
    if (S->P.ndim == 1) { /* PM:649-652 */
        pos[0] = (int)floor(px) + 1; pos[1] = 1; pos[2] = 1;
        ry = 0.0;
        rz = 0.0;
    } else if (S->P.ndim == 2) {
        pos[0] = (int)floor(px) + 1; pos[1] = (int)floor(py) + 1; pos[2] = 1;
        ry = py - pos[1] + 1;
        rz = 0.0;
    } else {
        pos[0] = (int)floor(px) + 1; pos[1] = (int)floor(py) + 1; pos[2] = (int)floor(pz) + 1;
        ry = py - pos[1] + 1;
        rz = pz - pos[2] + 1;
    }
    rx = px - pos[0] + 1;
    double rx1 = 1.0 - rx, ry1 = 1.0 - ry, rz1 = 1.0 - rz;
    w[0] = rx1 * ry1 * rz1;
    w[1] = rx * ry1 * rz1;
    w[2] = rx1 * ry * rz1;
    w[3] = rx * ry * rz1;
    w[4] = rx1 * ry1 * rz;
    w[5] = rx * ry1 * rz;
    w[6] = rx1 * ry * rz;
    w[7] = rx * ry * rz;
}

/* Fortran index ix in -1..nx+2 -> C index ix+1; iz: 1-based (extent 1) in 2-D */
static void interp_fields(const orc_sim* S, const int pos[3], const double w[8], double rt,
                          double fields[NVAR])
{
    double fields2[NVAR];
    const int ze = (S->P.ndim > 2) ? 1 : 0;
    const int ye = (S->P.ndim > 1) ? 1 : 0;
    for (int v = 0; v < NVAR; ++v) { fields[v] = 0.0; fields2[v] = 0.0; }
    for (int k = 0; k <= ze; ++k)
        for (int j = 0; j <= ye; ++j)
            for (int i = 0; i <= 1; ++i) {
                int idx = k * 4 + j * 2 + i;
                /* A runaway particle indexes outside farray in the reference (undefined
                 * behaviour); here, as in the GPU library, the cell index is clamped. */
                int c0 = pos[0] + 1, c1 = pos[1] + 1, c2 = pos[2] + 1;
                if (c0 < 0) c0 = 0;
                if (c0 > S->nxg - 2) c0 = S->nxg - 2;
                if (c1 < 0) c1 = 0;
                if (c1 > S->nyg - 2) c1 = S->nyg - 2;
                if (c2 < 0) c2 = 0;
                if (c2 > S->nzg - 2) c2 = S->nzg - 2;
                int ci = c0 + i;
                int cj = (S->P.ndim > 1) ? c1 + j : 0;
                int ck = (S->P.ndim > 2) ? c2 + k : 0;
                const float* f1 = S->farray1 + FIDX(S, 0, ci, cj, ck);
                for (int v = 0; v < NVAR; ++v) fields[v] = fields[v] + (double)f1[v] * w[idx];
                if (S->P.time_interp) {
                    const float* f2 = S->farray2 + FIDX(S, 0, ci, cj, ck);
                    for (int v = 0; v < NVAR; ++v) fields2[v] = fields2[v] + (double)f2[v] * w[idx];
                }
            }
    if (S->P.time_interp) {
        double rt1 = 1.0 - rt;
        for (int v = 0; v < NVAR; ++v) fields[v] = fields[v] * rt1 + fields2[v] * rt;
    }
}

void orc_interp(const orc_sim* S, int64_t n, const double* x, const double* y, const double* z,
                const double* rt, double* fields32)
{
    for (int64_t i = 0; i < n; ++i) {
        double px = (x[i] - S->P.xmin) / S->P.dx;
        double py = (y[i] - S->P.ymin) / S->P.dy;
        double pz = (z[i] - S->P.zmin) / S->P.dz;
        int pos[3];
        double w[8];
        get_interp_parameters(S, px, py, pz, pos, w);
        interp_fields(S, pos, w, rt[i], fields32 + (size_t)i * NVAR);
    }
}

/* ------------------------------------------------------------------------ */
/* kappa: PM:93-104, PM:2208-2450, PM:2464-2771                               */
/* ------------------------------------------------------------------------ */
typedef struct kappa_type {
    double knorm_para, knorm_perp, kpara, kperp, skpara, skperp, skpara_perp;
    double kxx, kyy, kzz, kxy, kxz, kyz;
    double dkxx_dx, dkyy_dy, dkzz_dz, dkxy_dx, dkxy_dy, dkxz_dx, dkxz_dz, dkyz_dy, dkyz_dz;
} kappa_type;

#define F(n) fields[(n) - 1]           /* fields(n), 1-based */
#define FG(n) fields[NFIELDS + (n) - 1] /* fields(nfields+n) */

/* aux: db2_slab(1:4) db2_2d(1:4) lc_slab(1:4) lc_2d(1:4) at the particle (interp_aux) */
#define DB2S(n) aux[(n) - 1]
#define DB22(n) aux[4 + (n) - 1]
#define LCS(n) aux[8 + (n) - 1]
#define LC2(n) aux[12 + (n) - 1]
static void calc_kappa(const orc_sim* S, const gpat_particle* ptl, const double* fields,
                       const double* aux, kappa_type* kp)
{
    const gpat_params* P = &S->P;
    double bx = F(5), by = F(6), bz = F(7);
    double b = sqrt(sq(bx) + sq(by) + sq(bz));
    double ib1 = (b < 2.220446049250313e-16) ? 1.0 : 1.0 / b; /* EPSILON(b), PM:2230-2234 */
    double ib2 = ib1 * ib1;
    double ib3 = ib1 * ib2;
    memset(kp, 0, sizeof(*kp));

    kp->knorm_para = 1.0;
    kp->knorm_perp = 1.0;
    if (P->mag_dependency == 1) kp->knorm_para = kp->knorm_para * pow(b, P->gamma_turb - 2.0);
    if (P->deltab_flag) kp->knorm_para = kp->knorm_para / DB2S(1);                       /* PM:2246-2248 */
    if (P->correlation_flag) kp->knorm_para = kp->knorm_para * pow(LCS(1), P->gamma_turb - 1.0); /* PM:2252-2254 */
    double knorm;
    if (P->momentum_dependency == 1)
        knorm = kp->knorm_para * pow(ptl->p / P->p0, P->pindex);
    else
        knorm = kp->knorm_para;
    kp->knorm_perp = kp->knorm_para;
    kp->kpara = P->kpara0 * knorm;
    kp->kperp = kp->kpara * P->kret;
    kp->skpara = sqrt(2.0 * kp->kpara);
    kp->skperp = sqrt(2.0 * kp->kperp);
    kp->skpara_perp = sqrt(2.0 * (kp->kpara - kp->kperp));

    if (P->ndim == 1) {
        /* PM:2271-2292.  With mag_dependency = 1 the reference multiplies by a db_dx that this
         * branch never assigns (PM:2274): undefined there, rejected by
         * orc_create_checked / gpat_init here, so dkdx is 0 whenever this line is reached. */
        double dkdx = 0.0;
        if (P->deltab_flag) dkdx = dkdx - DB2S(2) / DB2S(1);
        if (P->correlation_flag) dkdx = dkdx + (P->gamma_turb - 1.0) * LCS(2) / LCS(1);
        kp->dkxx_dx = kp->kpara * dkdx; /* not focused transport */
        return;
    }
    int three = (P->ndim == 3) || (P->ndim == 2 && P->include_3rd_dim);
    double dbx_dx = FG(13), dbx_dy = FG(14), dby_dx = FG(16), dby_dy = FG(17);
    double db_dx = FG(22), db_dy = FG(23);
    double dbx_dz = 0.0, dby_dz = 0.0, dbz_dx = 0.0, dbz_dy = 0.0, dbz_dz = 0.0, db_dz = 0.0;
    if (three) { dbz_dx = FG(19); dbz_dy = FG(20); }
    if (P->ndim == 3) { dbx_dz = FG(15); dby_dz = FG(18); dbz_dz = FG(21); db_dz = FG(24); }
    double dkdx = 0.0, dkdy = 0.0, dkdz = 0.0;
    if (P->mag_dependency == 1) {
        if (P->ndim == 3) { /* PM:2405-2409: no ib1 in 3-D */
            dkdx = db_dx * (P->gamma_turb - 2.0);
            dkdy = db_dy * (P->gamma_turb - 2.0);
            dkdz = db_dz * (P->gamma_turb - 2.0);
        } else { /* PM:2310-2313, 2360-2363 */
            dkdx = db_dx * ib1 * (P->gamma_turb - 2.0);
            dkdy = db_dy * ib1 * (P->gamma_turb - 2.0);
        }
    }
    if (P->deltab_flag) { /* PM:2314-2317, 2364-2367, 2410-2414 */
        dkdx = dkdx - DB2S(2) / DB2S(1);
        dkdy = dkdy - DB2S(3) / DB2S(1);
        if (P->ndim == 3) dkdz = dkdz - DB2S(4) / DB2S(1);
    }
    if (P->correlation_flag) { /* PM:2318-2321, 2368-2371, 2415-2419 */
        dkdx = dkdx + (P->gamma_turb - 1.0) * LCS(2) / LCS(1);
        dkdy = dkdy + (P->gamma_turb - 1.0) * LCS(3) / LCS(1);
        if (P->ndim == 3) dkdz = dkdz + (P->gamma_turb - 1.0) * LCS(4) / LCS(1);
    }
    /* PM:2372-2376: the focused-transport equation carries the parallel streaming itself */
    double kpp = P->focused_transport ? -kp->kperp : kp->kpara - kp->kperp;
    kp->dkxx_dx = kp->kperp * dkdx + kpp * dkdx * sq(bx) * ib2 +
                  2.0 * kpp * bx * (dbx_dx * b - bx * db_dx) * ib3;
    kp->dkyy_dy = kp->kperp * dkdy + kpp * dkdy * sq(by) * ib2 +
                  2.0 * kpp * by * (dby_dy * b - by * db_dy) * ib3;
    kp->dkxy_dx = kpp * dkdx * bx * by * ib2 +
                  kpp * ((dbx_dx * by + bx * dby_dx) * ib2 - 2.0 * bx * by * db_dx * ib3);
    kp->dkxy_dy = kpp * dkdy * bx * by * ib2 +
                  kpp * ((dbx_dy * by + bx * dby_dy) * ib2 - 2.0 * bx * by * db_dy * ib3);
    kp->kxx = kp->kperp + kpp * bx * bx * ib2;
    kp->kyy = kp->kperp + kpp * by * by * ib2;
    kp->kxy = kpp * bx * by * ib2;
    if (three) {
        kp->dkzz_dz = kp->kperp * dkdz + kpp * dkdz * sq(bz) * ib2 +
                      2.0 * kpp * bz * (dbz_dz * b - bz * db_dz) * ib3;
        kp->dkxz_dx = kpp * dkdx * bx * bz * ib2 +
                      kpp * ((dbx_dx * bz + bx * dbz_dx) * ib2 - 2.0 * bx * bz * db_dx * ib3);
        kp->dkxz_dz = kpp * dkdz * bx * bz * ib2 +
                      kpp * ((dbx_dz * bz + bx * dbz_dz) * ib2 - 2.0 * bx * bz * db_dz * ib3);
        kp->dkyz_dy = kpp * dkdy * by * bz * ib2 +
                      kpp * ((dby_dy * bz + by * dbz_dy) * ib2 - 2.0 * by * bz * db_dy * ib3);
        kp->dkyz_dz = kpp * dkdz * by * bz * ib2 +
                      kpp * ((dby_dz * bz + by * dbz_dz) * ib2 - 2.0 * by * bz * db_dz * ib3);
        kp->kzz = kp->kperp + kpp * bz * bz * ib2;
        kp->kxz = kpp * bx * bz * ib2;
        kp->kyz = kpp * by * bz * ib2;
    }
}

/* NLGC variant, PM:2464-2771 (deltab/correlation flags off) */
static void calc_kappa_nlgc(const orc_sim* S, const gpat_particle* ptl, const double* fields,
                            const double* aux, kappa_type* kp)
{
    const gpat_params* P = &S->P;
    double bx = F(5), by = F(6), bz = F(7);
    double b = sqrt(sq(bx) + sq(by) + sq(bz));
    double ib1 = (b < 2.220446049250313e-16) ? 1.0 : 1.0 / b;
    double ib2 = ib1 * ib1;
    double ib3 = ib1 * ib2;
    memset(kp, 0, sizeof(*kp));
    kp->knorm_para = 1.0;
    kp->knorm_perp = 1.0;
    if (P->mag_dependency == 1) {
        kp->knorm_para = kp->knorm_para * pow(b, P->gamma_turb - 2.0);
        kp->knorm_perp = kp->knorm_perp * pow(b, (P->gamma_turb - 2.0) / 3.0);
    }
    if (P->deltab_flag) { /* PM:2505-2509 */
        kp->knorm_para = kp->knorm_para / DB2S(1);
        kp->knorm_perp = kp->knorm_perp * pow(DB2S(1), -1.0 / 3.0) * pow(DB22(1), 2.0 / 3.0);
    }
    if (P->correlation_flag) { /* PM:2513-2517 */
        kp->knorm_para = kp->knorm_para * pow(LCS(1), P->gamma_turb - 1.0);
        kp->knorm_perp = kp->knorm_perp * pow(LCS(1), (P->gamma_turb - 1.0) / 3.0) * pow(LC2(1), 2.0 / 3.0);
    }
    double knorm_para, knorm_perp;
    if (P->momentum_dependency == 1) {
        knorm_para = kp->knorm_para * pow(ptl->p / P->p0, P->pindex);
        knorm_perp = kp->knorm_perp * pow(ptl->p / P->p0, (5.0 - P->gamma_turb) / 3.0);
    } else {
        knorm_para = kp->knorm_para;
        knorm_perp = kp->knorm_perp;
    }
    kp->kpara = P->kpara0 * knorm_para;
    kp->kperp = P->kpara0 * P->kperp_kpara * knorm_perp * sq(ptl->mu);
    kp->skpara = sqrt(2.0 * kp->kpara);
    kp->skperp = sqrt(2.0 * kp->kperp);
    kp->skpara_perp = sqrt(2.0 * (kp->kpara - kp->kperp));

    if (P->ndim == 1) { /* PM:2535-2562, same uninitialised db_dx with mag_dependency = 1 */
        double dkpara_dx = 0.0;
        if (P->deltab_flag) dkpara_dx = dkpara_dx - DB2S(2) / DB2S(1);
        if (P->correlation_flag) dkpara_dx = dkpara_dx + (P->gamma_turb - 1.0) * LCS(2) / LCS(1);
        kp->dkxx_dx = kp->kpara * dkpara_dx;
        return;
    }
    int three = (P->ndim == 3) || (P->ndim == 2 && P->include_3rd_dim);
    double dbx_dx = FG(13), dbx_dy = FG(14), dby_dx = FG(16), dby_dy = FG(17);
    double db_dx = FG(22), db_dy = FG(23);
    double dbx_dz = 0.0, dby_dz = 0.0, dbz_dx = 0.0, dbz_dy = 0.0, dbz_dz = 0.0, db_dz = 0.0;
    if (three) { dbz_dx = FG(19); dbz_dy = FG(20); }
    if (P->ndim == 3) { dbx_dz = FG(15); dby_dz = FG(18); dbz_dz = FG(21); db_dz = FG(24); }
    double dkpara_dx = 0.0, dkpara_dy = 0.0, dkpara_dz = 0.0;
    double dkperp_dx = 0.0, dkperp_dy = 0.0, dkperp_dz = 0.0;
    if (P->mag_dependency == 1) {
        dkpara_dx = db_dx * ib1 * (P->gamma_turb - 2.0);
        dkpara_dy = db_dy * ib1 * (P->gamma_turb - 2.0);
        dkperp_dx = db_dx * ib1 * (P->gamma_turb - 2.0) / 3.0;
        dkperp_dy = db_dy * ib1 * (P->gamma_turb - 2.0) / 3.0;
        if (P->ndim == 3) {
            dkpara_dz = db_dz * ib1 * (P->gamma_turb - 2.0);
            dkperp_dz = db_dz * ib1 * (P->gamma_turb - 2.0) / 3.0;
        }
    }
    if (P->deltab_flag) { /* PM:2589-2596 and the 2-D+3rd / 3-D twins */
        dkpara_dx = dkpara_dx - DB2S(2) / DB2S(1);
        dkpara_dy = dkpara_dy - DB2S(3) / DB2S(1);
        dkperp_dx = dkperp_dx - DB2S(2) / DB2S(1) / 3.0 + 2.0 * DB22(2) / DB22(1) / 3.0;
        dkperp_dy = dkperp_dy - DB2S(3) / DB2S(1) / 3.0 + 2.0 * DB22(3) / DB22(1) / 3.0;
        if (P->ndim == 3) {
            dkpara_dz = dkpara_dz - DB2S(4) / DB2S(1);
            dkperp_dz = dkperp_dz - DB2S(4) / DB2S(1) / 3.0 + 2.0 * DB22(4) / DB22(1) / 3.0;
        }
    }
    if (P->correlation_flag) { /* PM:2597-2604 */
        dkpara_dx = dkpara_dx + (P->gamma_turb - 1.0) * LCS(2) / LCS(1);
        dkpara_dy = dkpara_dy + (P->gamma_turb - 1.0) * LCS(3) / LCS(1);
        dkperp_dx = dkperp_dx + (P->gamma_turb - 1.0) * LCS(2) / LCS(1) / 3.0 + 2.0 * LC2(2) / LC2(1) / 3.0;
        dkperp_dy = dkperp_dy + (P->gamma_turb - 1.0) * LCS(3) / LCS(1) / 3.0 + 2.0 * LC2(3) / LC2(1) / 3.0;
        if (P->ndim == 3) {
            dkpara_dz = dkpara_dz + (P->gamma_turb - 1.0) * LCS(4) / LCS(1);
            dkperp_dz = dkperp_dz + (P->gamma_turb - 1.0) * LCS(4) / LCS(1) / 3.0 + 2.0 * LC2(4) / LC2(1) / 3.0;
        }
    }
    double kpp = P->focused_transport ? -kp->kperp : kp->kpara - kp->kperp; /* PM:2605-2609 */
    double kpa = kp->kpara, kpe = kp->kperp;
    kp->dkxx_dx = kpe * dkperp_dx + (kpa * dkpara_dx - kpe * dkperp_dx) * sq(bx) * ib2 +
                  2.0 * kpp * bx * (dbx_dx * b - bx * db_dx) * ib3;
    kp->dkyy_dy = kpe * dkperp_dy + (kpa * dkpara_dy - kpe * dkperp_dy) * sq(by) * ib2 +
                  2.0 * kpp * by * (dby_dy * b - by * db_dy) * ib3;
    kp->dkxy_dx = (kpa * dkpara_dx - kpe * dkperp_dx) * bx * by * ib2 +
                  kpp * ((dbx_dx * by + bx * dby_dx) * ib2 - 2.0 * bx * by * db_dx * ib3);
    kp->dkxy_dy = (kpa * dkpara_dy - kpe * dkperp_dy) * bx * by * ib2 +
                  kpp * ((dbx_dy * by + bx * dby_dy) * ib2 - 2.0 * bx * by * db_dy * ib3);
    kp->kxx = kpe + kpp * bx * bx * ib2;
    kp->kyy = kpe + kpp * by * by * ib2;
    kp->kxy = kpp * bx * by * ib2;
    if (three) {
        kp->dkzz_dz = kpe * dkperp_dz + (kpa * dkpara_dz - kpe * dkperp_dz) * sq(bz) * ib2 +
                      2.0 * kpp * bz * (dbz_dz * b - bz * db_dz) * ib3;
        kp->dkxz_dx = (kpa * dkpara_dx - kpe * dkperp_dx) * bx * bz * ib2 +
                      kpp * ((dbx_dx * bz + bx * dbz_dx) * ib2 - 2.0 * bx * bz * db_dx * ib3);
        kp->dkxz_dz = (kpa * dkpara_dz - kpe * dkperp_dz) * bx * bz * ib2 +
                      kpp * ((dbx_dz * bz + bx * dbz_dz) * ib2 - 2.0 * bx * bz * db_dz * ib3);
        kp->dkyz_dy = (kpa * dkpara_dy - kpe * dkperp_dy) * by * bz * ib2 +
                      kpp * ((dby_dy * bz + by * dbz_dy) * ib2 - 2.0 * by * bz * db_dy * ib3);
        kp->dkyz_dz = (kpa * dkpara_dz - kpe * dkperp_dz) * by * bz * ib2 +
                      kpp * ((dby_dz * bz + by * dbz_dz) * ib2 - 2.0 * by * bz * db_dz * ib3);
        kp->kzz = kpe + kpp * bz * bz * ib2;
        kp->kxz = kpp * bx * bz * ib2;
        kp->kyz = kpp * by * bz * ib2;
    }
}

/* ------------------------------------------------------------------------ */
/* acceleration region (PM:2885-2906), D_pp (PM:2918-2979)                    */
/* ------------------------------------------------------------------------ */
static int particle_in_acceleration_region(const orc_sim* S, const gpat_particle* ptl)
{
    const gpat_params* P = &S->P;
    int inx, iny = 1, inz = 1;
    double xnorm = (ptl->x - P->xmin) / P->lx;
    inx = (xnorm >= P->acc_region[0]) && (xnorm <= P->acc_region[1]);
    if (P->ndim > 1) {
        double ynorm = (ptl->y - P->ymin) / P->ly;
        iny = (ynorm >= P->acc_region[2]) && (ynorm <= P->acc_region[3]);
    }
    if (P->ndim == 3) {
        double znorm = (ptl->z - P->zmin) / P->lz;
        inz = (znorm >= P->acc_region[4]) && (znorm <= P->acc_region[5]);
    }
    return inx && iny && inz;
}

static void calc_dpp_wave_scattering(const orc_sim* S, double rho, double b, double kpara,
                                     const gpat_particle* ptl, double* dp_dt, double* dpp)
{
    double va = b / sqrt(rho);
    if (S->P.momentum_dependency == 1)
        *dp_dt = *dp_dt + (8.0 * ptl->p / (27.0 * kpara)) * sq(va);
    else
        *dp_dt = *dp_dt + (4.0 * ptl->p / (9.0 * kpara)) * sq(va);
    *dpp = *dpp + sq(ptl->p * va) / (9.0 * kpara);
}

static void calc_dpp_flow_shear(const orc_sim* S, double b, double bx, double by, double bz,
                                double knorm_para, double sxx, double syy, double szz, double sxy,
                                double sxz, double syz, const gpat_particle* ptl, double* dp_dt,
                                double* dpp)
{
    const gpat_params* P = &S->P;
    double gshear;
    if (P->weak_scattering) {
        double ib = (b < 2.220446049250313e-16) ? 0.0 : 1.0 / b;
        double bbsigma = sxx * sq(bx) + syy * sq(by) + szz * sq(bz) +
                         2.0 * (sxy * bx * by + sxz * bx * bz + syz * by * bz);
        bbsigma = bbsigma * ib * ib;
        gshear = sq(bbsigma) / 5.0;
    } else {
        gshear = 2.0 * (sq(sxx) + sq(syy) + sq(szz) + 2.0 * (sq(sxy) + sq(sxz) + sq(syz))) / 15.0;
    }
    if (gshear > 0.0) {
        *dp_dt = *dp_dt + (2.0 + P->pindex) * gshear * P->tau0 * knorm_para *
                              pow(ptl->p, P->pindex - 1.0) * pow(P->p0, 2.0 - P->pindex);
        *dpp = *dpp + gshear * P->tau0 * knorm_para * pow(ptl->p, P->pindex) *
                          pow(P->p0, 2.0 - P->pindex);
    }
}

static inline double min2(double a, double b) { return (b < a) ? b : a; }

/* common tail of every pusher: momentum update, PM:3589-3605 */
/* sh: the two interpolated surface heights of acc_by_surface runs (3-D pushers only), or NULL */
static int in_acceleration_region(const orc_sim* S, const gpat_particle* ptl, const double* sh)
{
    int in = particle_in_acceleration_region(S, ptl);
    if (sh) in = in && check_above_acc_surface(S, ptl->x, ptl->y, ptl->z, sh[0], sh[1]); /* PM:4889-4892 */
    return in;
}

static void update_momentum(const orc_sim* S, gpat_particle* ptl, double dp_dt, double dpp,
                            double sdt, double ranp, double* deltap, const double* sh)
{
    const gpat_params* P = &S->P;
    *deltap = dp_dt * ptl->dt + ranp * sqrt(2.0 * dpp) * sdt;
    if (P->acc_region_flag == 1) {
        if (in_acceleration_region(S, ptl, sh))
            ptl->p = ptl->p + *deltap;
        else
            *deltap = 0.0;
    } else {
        ptl->p = ptl->p + *deltap;
    }
    if (ptl->p < 0.25 * P->p0) {
        ptl->p = ptl->p - *deltap;
        *deltap = 0.25 * P->p0 - ptl->p;
        ptl->p = 0.25 * P->p0;
    }
}

static double drift_vdp(const orc_sim* S, const gpat_particle* ptl)
{
    const gpat_params* P = &S->P;
    /* `1.0 / (3 * pcharge)` is a default-real division, PM:3436 */
    float q = 1.0f / (float)(3 * P->pcharge);
    return (double)q / sqrt(sq(P->drift1 * P->p0 / ptl->p) + sq(P->drift2 * sq(P->p0) / sq(ptl->p)));
}

/* ------------------------------------------------------------------------ */
/* push_particle_1d, PM:2993-3111 (Cartesian, uniform grid): two uniforms per */
/* step, ran1 for x then one for p (PM:3085-3088)                             */
/* ------------------------------------------------------------------------ */
static void push_particle_1d(orc_sim* S, gpat_particle* ptl, const double* fields,
                             const kappa_type* kp, int fixed_dt, const double u[4], double* deltax,
                             double* deltap)
{
    const gpat_params* P = &S->P;
    double vx = F(1), bx = F(5), by = F(6), bz = F(7);
    double b = sqrt(sq(bx) + sq(by) + sq(bz));
    double dvx_dx = FG(1);
    double dxm = P->dx;
    double dx_dt = vx + kp->dkxx_dx; /* PM:3036 */
    double divv = dvx_dx;
    double dp_dt = -ptl->p * divv / 3.0;
    double dpp = 0.0;
    if (P->dpp_wave) calc_dpp_wave_scattering(S, F(4), b, kp->kpara, ptl, &dp_dt, &dpp);
    if (P->dpp_shear) { /* PM:3050-3056: `divv / 3` is an integer literal promoted to f64 */
        double sxx = dvx_dx - divv / 3.0, syy = -divv / 3.0, szz = -divv / 3.0;
        calc_dpp_flow_shear(S, b, bx, by, bz, kp->knorm_para, sxx, syy, szz, 0.0, 0.0, 0.0, ptl,
                            &dp_dt, &dpp);
    }
    if (!fixed_dt) {
        if (dx_dt != 0.0 && dp_dt != 0.0) { /* PM:3060-3072 */
            double s = (kp->skperp > 0.0) ? kp->skperp : kp->skpara;
            double d = sq(0.5 * dxm / kp->skpara);
            d = min2(d, sq(s / dx_dt));
            d = min2(d, (double)0.1f * ptl->p / fabs(dp_dt));
            ptl->dt = d;
        } else {
            ptl->dt = S->dt_min;
        }
        if (ptl->dt < S->dt_min) ptl->dt = S->dt_min;
        if (ptl->dt > S->dt_max) ptl->dt = S->dt_max;
    }
    double sdt = sqrt(ptl->dt);
    double sqrt3 = sqrt(3.0);
    double ran1 = (2.0 * u[0] - 1.0) * sqrt3;
    *deltax = dx_dt * ptl->dt + ran1 * kp->skpara * sdt;
    ptl->x = ptl->x + *deltax;
    ptl->t = ptl->t + ptl->dt;
    double ranp = (2.0 * u[1] - 1.0) * sqrt3;
    update_momentum(S, ptl, dp_dt, dpp, sdt, ranp, deltap, NULL);
}

/* ------------------------------------------------------------------------ */
/* push_particle_2d, PM:3358-3606 (Cartesian, uniform grid)                   */
/* ------------------------------------------------------------------------ */
static void push_particle_2d(orc_sim* S, gpat_particle* ptl, const double* fields,
                             const kappa_type* kp, int fixed_dt, const double u[4], double* deltax,
                             double* deltay, double* deltap)
{
    const gpat_params* P = &S->P;
    double vx = F(1), vy = F(2), bx = F(5), by = F(6), bz = F(7);
    double b = sqrt(sq(bx) + sq(by) + sq(bz));
    double dvx_dx = FG(1), dvy_dy = FG(5);
    double dxm = P->dx, dym = P->dy;
    double ib = (b < 2.220446049250313e-16) ? 0.0 : 1.0 / b; /* PM:3414-3418 */
    double dbz_dx = FG(19), dbz_dy = FG(20), db_dx = FG(22), db_dy = FG(23);
    double ib2 = ib * ib;
    double ib3 = ib * ib2;
    double vdp = drift_vdp(S, ptl);
    double vdx = vdp * (dbz_dy * ib2 - 2.0 * bz * db_dy * ib3);
    double vdy = vdp * (-dbz_dx * ib2 + 2.0 * bz * db_dx * ib3);
    double vdz;
    if (P->check_drift_2d) {
        double dbx_dy = FG(14), dby_dx = FG(16);
        vdz = vdp * ((dby_dx - dbx_dy) * ib2 - 2.0 * (by * db_dx - bx * db_dy) * ib3);
    } else {
        vdz = 0.0;
    }
    double dx_dt = vx + vdx + kp->dkxx_dx + kp->dkxy_dy;
    double dy_dt = vy + vdy + kp->dkxy_dx + kp->dkyy_dy;
    double dz_dt = vdz;
    double divv = dvx_dx + dvy_dy;
    double dp_dt = -ptl->p * divv / 3.0;
    double dpp = 0.0;
    if (P->dpp_wave) calc_dpp_wave_scattering(S, F(4), b, kp->kpara, ptl, &dp_dt, &dpp);
    if (P->dpp_shear) {
        double dvx_dy = FG(2), dvy_dx = FG(4);
        double sxx = dvx_dx - divv / 3.0, syy = dvy_dy - divv / 3.0, szz = -divv / 3.0;
        double sxy = (dvx_dy + dvy_dx) / 2.0;
        calc_dpp_flow_shear(S, b, bx, by, bz, kp->knorm_para, sxx, syy, szz, sxy, 0.0, 0.0, ptl,
                            &dp_dt, &dpp);
    }
    if (!fixed_dt) {
        if (dx_dt != 0.0 && dy_dt != 0.0 && dp_dt != 0.0) {
            double s = (kp->skperp > 0.0) ? kp->skperp : kp->skpara; /* PM:3518-3530 */
            double d = sq(0.5 * dxm / kp->skpara);
            d = min2(d, sq(0.5 * dym / kp->skpara));
            d = min2(d, sq(s / dx_dt));
            d = min2(d, sq(s / dy_dt));
            d = min2(d, (double)0.1f * ptl->p / fabs(dp_dt));
            ptl->dt = d;
        } else {
            ptl->dt = S->dt_min;
        }
        if (ptl->dt < S->dt_min) ptl->dt = S->dt_min;
        if (ptl->dt > S->dt_max) ptl->dt = S->dt_max;
    }
    double sdt = sqrt(ptl->dt);
    double sqrt3 = sqrt(3.0);
    double ran1 = (2.0 * u[0] - 1.0) * sqrt3;
    double ran2 = (2.0 * u[1] - 1.0) * sqrt3;
    double ran3 = (2.0 * u[2] - 1.0) * sqrt3;
    *deltax = dx_dt * ptl->dt + ran1 * kp->skperp * sdt + ran3 * kp->skpara_perp * sdt * bx * ib;
    *deltay = dy_dt * ptl->dt + ran2 * kp->skperp * sdt + ran3 * kp->skpara_perp * sdt * by * ib;
    double deltaz = dz_dt * ptl->dt;
    ptl->x = ptl->x + *deltax;
    ptl->y = ptl->y + *deltay;
    ptl->z = ptl->z + deltaz;
    ptl->t = ptl->t + ptl->dt;
    double ranp = (2.0 * u[3] - 1.0) * sqrt3;
    update_momentum(S, ptl, dp_dt, dpp, sdt, ranp, deltap, NULL);
}

/* ------------------------------------------------------------------------ */
/* focused transport, 2-D Cartesian: calc_duu (PM:3116-3155) and              */
/* push_particle_2d_ft (PM:3626-3977).  Four uniforms per step: two for the   */
/* perpendicular displacement, one for p, one for mu (PM:3881-3882, 3918-3921)*/
/* ------------------------------------------------------------------------ */
static void calc_duu(const orc_sim* S, const gpat_particle* ptl, double b, const double* aux, double div_bnorm,
                     double divv, double bb_gradv, double bv_gradv, double mu2, double* dmu_dt,
                     double* duu, double* duu_du)
{
    const gpat_params* P = &S->P;
    *dmu_dt = ptl->v * div_bnorm + ptl->mu * divv - 3 * ptl->mu * bb_gradv - 2 * bv_gradv / ptl->v;
    *dmu_dt = *dmu_dt * (1 - mu2) * 0.5;
    double h0 = (double)0.2f; /* `h0 = 0.2`: a default-real literal assigned to real(dp) */
    double dtmp = pow(fabs(ptl->mu), P->gamma_turb - 1) + h0;
    *duu = P->duu0 * (1 - mu2) * dtmp;
    if (ptl->mu > 0.0)
        *duu_du = P->duu0 * (-2 * ptl->mu * dtmp + (1 - mu2) * pow(fabs(ptl->mu), P->gamma_turb - 2));
    else if (ptl->mu < 0.0)
        *duu_du = P->duu0 * (-2 * ptl->mu * dtmp - (1 - mu2) * pow(fabs(ptl->mu), P->gamma_turb - 2));
    else
        *duu_du = 0.0;
    double duu_norm = 1.0;
    if (P->mag_dependency == 1) duu_norm = duu_norm * pow(b, 2.0 - P->gamma_turb);
    if (P->deltab_flag) duu_norm = duu_norm * DB2S(1);                                        /* PM:3143-3145 */
    if (P->correlation_flag) duu_norm = duu_norm * pow(LCS(1), (double)1.0f - P->gamma_turb); /* PM:3146-3148 */
    if (P->momentum_dependency == 1) duu_norm = duu_norm * pow(ptl->p / P->p0, P->gamma_turb - 1);
    *duu_du = *duu_du * duu_norm;
    *duu = *duu * duu_norm;
    *dmu_dt = *dmu_dt + *duu_du;
}

static void push_particle_2d_ft(orc_sim* S, gpat_particle* ptl, const double* fields, const double* aux,
                                const kappa_type* kp, int fixed_dt, const double u[4], double* deltax,
                                double* deltay, double* deltap, double* deltav, double* deltamu)
{
    const gpat_params* P = &S->P;
    const double mu_max = (double)0.99f;
    double vx = F(1), vy = F(2), vz = F(3), bx = F(5), by = F(6), bz = F(7);
    double b = sqrt(sq(bx) + sq(by) + sq(bz));
    double dxm = P->dx, dym = P->dy;
    double ib = (b < 2.220446049250313e-16) ? 0.0 : 1.0 / b;
    double dbx_dx = FG(13), dbx_dy = FG(14), dby_dx = FG(16), dby_dy = FG(17);
    double dbz_dx = FG(19), dbz_dy = FG(20), db_dx = FG(22), db_dy = FG(23);
    double ib2 = ib * ib, ib3 = ib * ib2;
    /* `1.0 / pcharge` is a default-real quotient, PM:3716 */
    double vdp = (double)(1.0f / (float)P->pcharge) /
                 sqrt(sq(P->drift1 * P->p0 / ptl->p) + sq(P->drift2 * sq(P->p0) / sq(ptl->p)));
    double mu2 = sq(ptl->mu);
    double muf1 = 0.5 * (1.0 - mu2), muf2 = 0.5 * (3.0 * mu2 - 1.0);
    double kx = bx * dbx_dx + by * dbx_dy;
    double ky = bx * dby_dx + by * dby_dy;
    double kz = bx * dbz_dx + by * dbz_dy;
    double bdot_curvb = bx * dbz_dy - by * dbz_dx + bz * (dby_dx - dbx_dy);
    double vdx = vdp * (muf1 * (-bz * db_dy) * ib2 + mu2 * (by * kz - bz * ky) * ib3 +
                        muf1 * bx * bdot_curvb * ib3);
    double vdy = vdp * (muf1 * (bz * db_dx) * ib2 + mu2 * (bz * kx - bx * kz) * ib3 +
                        muf1 * by * bdot_curvb * ib3);
    double vdz = 0.0;
    if (P->check_drift_2d)
        vdz = vdp * (muf1 * (bx * db_dy - by * db_dx) * ib2 + mu2 * (bx * ky - by * kx) * ib3 +
                     muf1 * bz * bdot_curvb * ib3);
    double vbx = ptl->v * ptl->mu * ib;
    double vby = vbx * by; /* particle velocity along the magnetic field */
    vbx = vbx * bx;
    double dvx_dx = FG(1), dvx_dy = FG(2), dvy_dx = FG(4), dvy_dy = FG(5), dvz_dx = FG(7), dvz_dy = FG(8);
    double dx_dt = vx + vdx + vbx + kp->dkxx_dx + kp->dkxy_dy;
    double dy_dt = vy + vdy + vby + kp->dkxy_dx + kp->dkyy_dy;
    double dz_dt = vdz;
    double divv = dvx_dx + dvy_dy;
    double bb_gradv = (bx * (bx * dvx_dx + by * dvx_dy) + by * (bx * dvy_dx + by * dvy_dy) +
                       bz * (bx * dvz_dx + by * dvz_dy)) * ib2;
    double bv_gradv = (bx * (vx * dvx_dx + vy * dvx_dy) + by * (vx * dvy_dx + vy * dvy_dy) +
                       bz * (vx * dvz_dx + vy * dvz_dy)) * ib;
    double acc_rate = -(muf1 * divv + muf2 * bb_gradv + ptl->mu * bv_gradv / ptl->v);
    double dp_dt = ptl->p * acc_rate;
    double dpp = 0.0;
    if (P->dpp_wave) calc_dpp_wave_scattering(S, F(4), b, kp->kpara, ptl, &dp_dt, &dpp);
    if (P->dpp_shear) {
        double sxx = dvx_dx - divv / 3, syy = dvy_dy - divv / 3, szz = -divv / 3;
        double sxy = (dvx_dy + dvy_dx) / 2;
        calc_dpp_flow_shear(S, b, bx, by, bz, kp->knorm_para, sxx, syy, szz, sxy, 0.0, 0.0, ptl,
                            &dp_dt, &dpp);
    }
    double div_bnorm = -(bx * db_dx + by * db_dy) * ib2;
    double dmu_dt, duu, duu_du;
    calc_duu(S, ptl, b, aux, div_bnorm, divv, bb_gradv, bv_gradv, mu2, &dmu_dt, &duu, &duu_du);
    if (!fixed_dt) {
        if (dx_dt != 0.0 && dy_dt != 0.0 && dp_dt != 0.0 && dmu_dt != 0.0) {
            double s = (kp->skperp > 0.0) ? kp->skperp : kp->skpara; /* PM:3847-3864 */
            double d = sq(0.5 * dxm / s);
            d = min2(d, sq(0.5 * dym / s));
            d = min2(d, sq(s / dx_dt));
            d = min2(d, sq(s / dy_dt));
            d = min2(d, (double)0.1f * ptl->p / fabs(dp_dt));
            d = min2(d, (double)0.1f / fabs(dmu_dt));
            d = min2(d, 2.0 * duu / sq(dmu_dt));
            ptl->dt = d;
        } else {
            ptl->dt = S->dt_min;
        }
        if (ptl->dt < S->dt_min) ptl->dt = S->dt_min;
        if (ptl->dt > S->dt_max) ptl->dt = S->dt_max;
    }
    double sdt = sqrt(ptl->dt);
    double sqrt3 = sqrt(3.0);
    double ran1 = (2.0 * u[0] - 1.0) * sqrt3;
    double ran2 = (2.0 * u[1] - 1.0) * sqrt3;
    double bxn = bx * ib, byn = by * ib, bzn = bz * ib;
    double ibxyn = 1.0 / sqrt(sq(bxn) + sq(byn));
    /* the second term uses the UN-normalised by / bx, PM:3912-3913 -- kept */
    *deltax = dx_dt * ptl->dt + kp->skperp * ibxyn * sdt * (-bxn * bzn * ran1 - by * ran2);
    *deltay = dy_dt * ptl->dt + kp->skperp * ibxyn * sdt * (-byn * bzn * ran1 + bx * ran2);
    double deltaz = dz_dt * ptl->dt;
    ran1 = (2.0 * u[2] - 1.0) * sqrt3;
    *deltap = dp_dt * ptl->dt + ran1 * sqrt(2 * dpp) * sdt;
    *deltav = ptl->v * *deltap / ptl->p;
    ran1 = (2.0 * u[3] - 1.0) * sqrt3;
    *deltamu = dmu_dt * ptl->dt + ran1 * sqrt(2 * duu) * sdt;
    ptl->x = ptl->x + *deltax;
    ptl->y = ptl->y + *deltay;
    ptl->z = ptl->z + deltaz;
    ptl->mu = ptl->mu + *deltamu;
    ptl->t = ptl->t + ptl->dt;
    if (ptl->mu > mu_max) {
        *deltamu = mu_max - (ptl->mu - *deltamu);
        ptl->mu = mu_max;
    } else if (ptl->mu < -mu_max) {
        *deltamu = -mu_max - (ptl->mu - *deltamu);
        ptl->mu = -mu_max;
    }
    if (P->acc_region_flag == 1) {
        if (particle_in_acceleration_region(S, ptl)) {
            ptl->p = ptl->p + *deltap;
            ptl->v = ptl->v + *deltav;
        } else {
            *deltap = 0.0;
            *deltav = 0.0;
        }
    } else {
        ptl->p = ptl->p + *deltap;
        ptl->v = ptl->v + *deltav;
    }
    if (ptl->p < 0.25 * P->p0) { /* PM:3967-3974 */
        ptl->v = ptl->v - *deltav;
        *deltav = ptl->v * 0.25 * P->p0 / ptl->p - ptl->v;
        ptl->v = ptl->v + *deltav;
        ptl->p = ptl->p - *deltap;
        *deltap = 0.25 * P->p0 - ptl->p;
        ptl->p = 0.25 * P->p0;
    }
}

/* ------------------------------------------------------------------------ */
/* push_particle_2d_include_3rd_ft (PM:4267-4623) and push_particle_3d_ft     */
/* (PM:4930-5320), Cartesian, no acc_by_surface: one body, the 2-D variant    */
/* has every d/dz equal to zero (x - 0, x + 0 and 0 * x are exact, so the     */
/* shared expressions give the bits of the two-dimensional formulas).         */
/* Five uniforms: ran1..ran3 (ran3 is drawn and unused in Cartesian runs),    */
/* then one for p and one for mu.                                             */
/* ------------------------------------------------------------------------ */
static void push_particle_ft_3d_like(orc_sim* S, gpat_particle* ptl, const double* fields, const double* aux,
                                     const kappa_type* kp, int fixed_dt, const double u[4], double u5,
                                     double* deltax, double* deltay, double* deltaz, double* deltap,
                                     double* deltav, double* deltamu, const double* sh)
{
    const gpat_params* P = &S->P;
    const int full3d = (P->ndim == 3);
    const double mu_max = (double)0.99f;
    double vx = F(1), vy = F(2), vz = F(3), bx = F(5), by = F(6), bz = F(7);
    double b = sqrt(sq(bx) + sq(by) + sq(bz));
    double ib = (b < 2.220446049250313e-16) ? 0.0 : 1.0 / b;
    double bxn = bx * ib, byn = by * ib, bzn = bz * ib;
    double bxyn = sqrt(sq(bxn) + sq(byn));
    double ibxyn = (bxyn < 2.220446049250313e-16) ? 0.0 : 1.0 / bxyn;
    double dxm = P->dx, dym = P->dy, dzm = P->dz;
    double dbx_dx = FG(13), dbx_dy = FG(14), dby_dx = FG(16), dby_dy = FG(17);
    double dbz_dx = FG(19), dbz_dy = FG(20), db_dx = FG(22), db_dy = FG(23);
    double dbx_dz = full3d ? FG(15) : 0.0, dby_dz = full3d ? FG(18) : 0.0;
    double dbz_dz = full3d ? FG(21) : 0.0, db_dz = full3d ? FG(24) : 0.0;
    double ib2 = ib * ib, ib3 = ib * ib2;
    double vdp = (double)(1.0f / (float)P->pcharge) /
                 sqrt(sq(P->drift1 * P->p0 / ptl->p) + sq(P->drift2 * sq(P->p0) / sq(ptl->p)));
    double mu2 = sq(ptl->mu);
    double muf1 = 0.5 * (1.0 - mu2), muf2 = 0.5 * (3.0 * mu2 - 1.0);
    double kx, ky, kz, bdot_curvb, vdx, vdy, vdz;
    if (full3d) { /* PM:5049-5063 */
        kx = bx * dbx_dx + by * dbx_dy + bz * dbx_dz;
        ky = bx * dby_dx + by * dby_dy + bz * dby_dz;
        kz = bx * dbz_dx + by * dbz_dy + bz * dbz_dz;
        bdot_curvb = bx * (dbz_dy - dby_dz) + by * (dbx_dz - dbz_dx) + bz * (dby_dx - dbx_dy);
        vdx = vdp * (muf1 * (by * db_dz - bz * db_dy) * ib2 + mu2 * (by * kz - bz * ky) * ib3 +
                     muf1 * bx * bdot_curvb * ib3);
        vdy = vdp * (muf1 * (bz * db_dx - bx * db_dz) * ib2 + mu2 * (bz * kx - bx * kz) * ib3 +
                     muf1 * by * bdot_curvb * ib3);
    } else { /* PM:4388-4400 */
        kx = bx * dbx_dx + by * dbx_dy;
        ky = bx * dby_dx + by * dby_dy;
        kz = bx * dbz_dx + by * dbz_dy;
        bdot_curvb = bx * (dbz_dy) + by * (-dbz_dx) + bz * (dby_dx - dbx_dy);
        vdx = vdp * (muf1 * (-bz * db_dy) * ib2 + mu2 * (by * kz - bz * ky) * ib3 +
                     muf1 * bx * bdot_curvb * ib3);
        vdy = vdp * (muf1 * (bz * db_dx) * ib2 + mu2 * (bz * kx - bx * kz) * ib3 +
                     muf1 * by * bdot_curvb * ib3);
    }
    vdz = vdp * (muf1 * (bx * db_dy - by * db_dx) * ib2 + mu2 * (bx * ky - by * kx) * ib3 +
                 muf1 * bz * bdot_curvb * ib3);
    double vbx = ptl->v * ptl->mu * ib;
    double vby = vbx * by;
    double vbz = vbx * bz;
    vbx = vbx * bx;
    double dvx_dx = FG(1), dvx_dy = FG(2), dvy_dx = FG(4), dvy_dy = FG(5), dvz_dx = FG(7), dvz_dy = FG(8);
    double dvx_dz = full3d ? FG(3) : 0.0, dvy_dz = full3d ? FG(6) : 0.0, dvz_dz = full3d ? FG(9) : 0.0;
    double dx_dt, dy_dt, dz_dt, divv, bb_gradv, bv_gradv;
    if (full3d) { /* PM:5104-5115 */
        dx_dt = vx + vbx + vdx + kp->dkxx_dx + kp->dkxy_dy + kp->dkxz_dz;
        dy_dt = vy + vby + vdy + kp->dkxy_dx + kp->dkyy_dy + kp->dkyz_dz;
        dz_dt = vz + vbz + vdz + kp->dkxz_dx + kp->dkyz_dy + kp->dkzz_dz;
        divv = dvx_dx + dvy_dy + dvz_dz;
        bb_gradv = (bx * (bx * dvx_dx + by * dvx_dy + bz * dvx_dz) + by * (bx * dvy_dx + by * dvy_dy + bz * dvy_dz) +
                    bz * (bx * dvz_dx + by * dvz_dy + bz * dvz_dz)) * ib2;
        bv_gradv = (bx * (vx * dvx_dx + vy * dvx_dy + vz * dvx_dz) + by * (vx * dvy_dx + vy * dvy_dy + vz * dvy_dz) +
                    bz * (vx * dvz_dx + vy * dvz_dy + vz * dvz_dz)) * ib;
    } else { /* PM:4432-4442 */
        dx_dt = vx + vbx + vdx + kp->dkxx_dx + kp->dkxy_dy;
        dy_dt = vy + vby + vdy + kp->dkxy_dx + kp->dkyy_dy;
        dz_dt = vz + vbz + vdz + kp->dkxz_dx + kp->dkyz_dy;
        divv = dvx_dx + dvy_dy;
        bb_gradv = (bx * (bx * dvx_dx + by * dvx_dy) + by * (bx * dvy_dx + by * dvy_dy) +
                    bz * (bx * dvz_dx + by * dvz_dy)) * ib2;
        bv_gradv = (bx * (vx * dvx_dx + vy * dvx_dy) + by * (vx * dvy_dx + vy * dvy_dy) +
                    bz * (vx * dvz_dx + vy * dvz_dy)) * ib;
    }
    double acc_rate = -(muf1 * divv + muf2 * bb_gradv + ptl->mu * bv_gradv / ptl->v);
    double dp_dt = ptl->p * acc_rate;
    double dpp = 0.0;
    if (P->dpp_wave) calc_dpp_wave_scattering(S, F(4), b, kp->kpara, ptl, &dp_dt, &dpp);
    if (P->dpp_shear) {
        double sxx = dvx_dx - divv / 3, syy = dvy_dy - divv / 3;
        double szz = full3d ? dvz_dz - divv / 3 : -divv / 3;
        double sxy = (dvx_dy + dvy_dx) / 2;
        double sxz = full3d ? (dvx_dz + dvz_dx) / 2 : dvz_dx / 2;
        double syz = full3d ? (dvy_dz + dvz_dy) / 2 : dvz_dy / 2;
        calc_dpp_flow_shear(S, b, bx, by, bz, kp->knorm_para, sxx, syy, szz, sxy, sxz, syz, ptl, &dp_dt, &dpp);
    }
    double div_bnorm = full3d ? -(bx * db_dx + by * db_dy + bz * db_dz) * ib2 : -(bx * db_dx + by * db_dy) * ib2;
    double dmu_dt, duu, duu_du;
    calc_duu(S, ptl, b, aux, div_bnorm, divv, bb_gradv, bv_gradv, mu2, &dmu_dt, &duu, &duu_du);
    if (!fixed_dt) {
        int ok = dx_dt != 0.0 && dy_dt != 0.0 && dp_dt != 0.0 && dmu_dt != 0.0;
        if (full3d) ok = ok && dz_dt != 0.0; /* PM:5156-5160; the 2-D variant does not test dz_dt, PM:4482-4485 */
        if (ok) {
            double s = (kp->skperp > 0.0) ? kp->skperp : kp->skpara;
            double d = sq(0.5 * dxm / s);
            d = min2(d, sq(0.5 * dym / s));
            if (full3d) d = min2(d, sq(0.5 * dzm / s));
            d = min2(d, sq(s / dx_dt));
            d = min2(d, sq(s / dy_dt));
            if (full3d) d = min2(d, sq(s / dz_dt));
            d = min2(d, (double)0.1f * ptl->p / fabs(dp_dt));
            d = min2(d, (double)0.1f / fabs(dmu_dt));
            d = min2(d, 2.0 * duu / sq(dmu_dt));
            ptl->dt = d;
        } else {
            ptl->dt = S->dt_min;
        }
        if (ptl->dt < S->dt_min) ptl->dt = S->dt_min;
        if (ptl->dt > S->dt_max) ptl->dt = S->dt_max;
    }
    double sdt = sqrt(ptl->dt);
    double sqrt3 = sqrt(3.0);
    double ran1 = (2.0 * u[0] - 1.0) * sqrt3;
    double ran2 = (2.0 * u[1] - 1.0) * sqrt3;
    /* ran3 = u[2] is drawn and not used by the Cartesian branch, PM:4523, 5233 */
    *deltax = dx_dt * ptl->dt + (-bxn * bzn * kp->skperp * ibxyn * ran1 - byn * kp->skperp * ibxyn * ran2) * sdt;
    *deltay = dy_dt * ptl->dt + (-byn * bzn * kp->skperp * ibxyn * ran1 + bxn * kp->skperp * ibxyn * ran2) * sdt;
    *deltaz = dz_dt * ptl->dt + bxyn * kp->skperp * ran1 * sdt;
    ran1 = (2.0 * u[3] - 1.0) * sqrt3;
    *deltap = dp_dt * ptl->dt + ran1 * sqrt(2 * dpp) * sdt;
    *deltav = ptl->v * *deltap / ptl->p;
    ran1 = (2.0 * u5 - 1.0) * sqrt3;
    *deltamu = dmu_dt * ptl->dt + ran1 * sqrt(2 * duu) * sdt;
    ptl->x = ptl->x + *deltax;
    ptl->y = ptl->y + *deltay;
    ptl->z = ptl->z + *deltaz;
    ptl->mu = ptl->mu + *deltamu;
    ptl->t = ptl->t + ptl->dt;
    if (ptl->mu > mu_max) {
        *deltamu = mu_max - (ptl->mu - *deltamu);
        ptl->mu = mu_max;
    } else if (ptl->mu < -mu_max) {
        *deltamu = -mu_max - (ptl->mu - *deltamu);
        ptl->mu = -mu_max;
    }
    if (P->acc_region_flag == 1) {
        if (in_acceleration_region(S, ptl, full3d ? sh : NULL)) { /* PM:5297-5303 */
            ptl->p = ptl->p + *deltap;
            ptl->v = ptl->v + *deltav;
        } else {
            *deltap = 0.0;
            *deltav = 0.0;
        }
    } else {
        ptl->p = ptl->p + *deltap;
        ptl->v = ptl->v + *deltav;
    }
    if (ptl->p < 0.25 * P->p0) {
        ptl->v = ptl->v - *deltav;
        *deltav = ptl->v * 0.25 * P->p0 / ptl->p - ptl->v;
        ptl->v = ptl->v + *deltav;
        ptl->p = ptl->p - *deltap;
        *deltap = 0.25 * P->p0 - ptl->p;
        ptl->p = 0.25 * P->p0;
    }
}

/* ------------------------------------------------------------------------ */
/* push_particle_2d_include_3rd (PM:3979-4245) and push_particle_3d           */
/* (PM:4625-4907): identical structure; 2-D sets every d/dz to zero.          */
/* ------------------------------------------------------------------------ */
static void push_particle_3d_like(orc_sim* S, gpat_particle* ptl, const double* fields,
                                  kappa_type* kp, int fixed_dt, const double u[4], double* deltax,
                                  double* deltay, double* deltaz, double* deltap, const double* sh)
{
    const gpat_params* P = &S->P;
    const int full3d = (P->ndim == 3);
    double vx = F(1), vy = F(2), vz = F(3), bx = F(5), by = F(6), bz = F(7);
    double b = sqrt(sq(bx) + sq(by) + sq(bz));
    double ib = (b < 2.220446049250313e-16) ? 0.0 : 1.0 / b;
    double bxn = bx * ib, byn = by * ib, bzn = bz * ib;
    double bxyn = sqrt(sq(bxn) + sq(byn));
    double ibxyn = (bxyn < 2.220446049250313e-16) ? 0.0 : 1.0 / bxyn;
    double dvx_dx = FG(1), dvy_dy = FG(5);
    double dvz_dz = full3d ? FG(9) : 0.0;
    double dxm = P->dx, dym = P->dy, dzm = P->dz;
    double dbx_dy = FG(14), dby_dx = FG(16), dbz_dx = FG(19), dbz_dy = FG(20);
    double db_dx = FG(22), db_dy = FG(23);
    double dbx_dz = full3d ? FG(15) : 0.0;
    double dby_dz = full3d ? FG(18) : 0.0;
    double db_dz = full3d ? FG(24) : 0.0;
    double ib2 = ib * ib;
    double ib3 = ib * ib2;
    double vdp = drift_vdp(S, ptl);
    double vdx = vdp * ((dbz_dy - dby_dz) * ib2 - 2.0 * (bz * db_dy - by * db_dz) * ib3);
    double vdy = vdp * ((dbx_dz - dbz_dx) * ib2 - 2.0 * (bx * db_dz - bz * db_dx) * ib3);
    double vdz = vdp * ((dby_dx - dbx_dy) * ib2 - 2.0 * (by * db_dx - bx * db_dy) * ib3);
    if (!full3d) { /* PM:4094-4096 */
        kp->dkxz_dz = 0.0; kp->dkyz_dz = 0.0; kp->dkzz_dz = 0.0;
    }
    double dx_dt = vx + vdx + kp->dkxx_dx + kp->dkxy_dy + kp->dkxz_dz;
    double dy_dt = vy + vdy + kp->dkxy_dx + kp->dkyy_dy + kp->dkyz_dz;
    double dz_dt = vz + vdz + kp->dkxz_dx + kp->dkyz_dy + kp->dkzz_dz;
    double divv = dvx_dx + dvy_dy + dvz_dz;
    double dp_dt = -ptl->p * divv / 3.0;
    double dpp = 0.0;
    if (P->dpp_wave) calc_dpp_wave_scattering(S, F(4), b, kp->kpara, ptl, &dp_dt, &dpp);
    if (P->dpp_shear) {
        double dvx_dy = FG(2), dvy_dx = FG(4), dvz_dx = FG(7), dvz_dy = FG(8);
        double dvx_dz = full3d ? FG(3) : 0.0;
        double dvy_dz = full3d ? FG(6) : 0.0;
        double sxx = dvx_dx - divv / 3.0, syy = dvy_dy - divv / 3.0, szz = dvz_dz - divv / 3.0;
        double sxy = (dvx_dy + dvy_dx) / 2.0;
        double sxz = (dvx_dz + dvz_dx) / 2.0;
        double syz = (dvy_dz + dvz_dy) / 2.0;
        calc_dpp_flow_shear(S, b, bx, by, bz, kp->knorm_para, sxx, syy, szz, sxy, sxz, syz, ptl,
                            &dp_dt, &dpp);
    }
    if (!fixed_dt) {
        if (dx_dt != 0.0 && dy_dt != 0.0 && dp_dt != 0.0) {
            double s = (kp->skperp > 0.0) ? kp->skperp : kp->skpara;
            double d = sq(0.5 * dxm / kp->skpara);
            d = min2(d, sq(0.5 * dym / kp->skpara));
            if (full3d) d = min2(d, sq(0.5 * dzm / kp->skpara)); /* PM:4815-4821 */
            d = min2(d, sq(s / dx_dt));
            d = min2(d, sq(s / dy_dt));
            if (full3d) d = min2(d, sq(s / dz_dt));
            d = min2(d, (double)0.1f * ptl->p / fabs(dp_dt));
            ptl->dt = d;
        } else {
            ptl->dt = S->dt_min;
        }
        if (ptl->dt < S->dt_min) ptl->dt = S->dt_min;
        if (ptl->dt > S->dt_max) ptl->dt = S->dt_max;
    }
    double sdt = sqrt(ptl->dt);
    double sqrt3 = sqrt(3.0);
    double ran1 = (2.0 * u[0] - 1.0) * sqrt3;
    double ran2 = (2.0 * u[1] - 1.0) * sqrt3;
    double ran3 = (2.0 * u[2] - 1.0) * sqrt3;
    *deltax = dx_dt * ptl->dt + (bxn * kp->skpara * ran1 - bxn * bzn * kp->skperp * ibxyn * ran2 -
                                 byn * kp->skperp * ibxyn * ran3) * sdt;
    *deltay = dy_dt * ptl->dt + (byn * kp->skpara * ran1 - byn * bzn * kp->skperp * ibxyn * ran2 +
                                 bxn * kp->skperp * ibxyn * ran3) * sdt;
    *deltaz = dz_dt * ptl->dt + (bzn * kp->skpara * ran1 + bxyn * kp->skperp * ran2) * sdt;
    ptl->x = ptl->x + *deltax;
    ptl->y = ptl->y + *deltay;
    ptl->z = ptl->z + *deltaz;
    ptl->t = ptl->t + ptl->dt;
    double ranp = (2.0 * u[3] - 1.0) * sqrt3;
    update_momentum(S, ptl, dp_dt, dpp, sdt, ranp, deltap, full3d ? sh : NULL);
}

/* ------------------------------------------------------------------------ */
/* particle_boundary_condition, PM:1984-2129 (single-rank branches)           */
/* ------------------------------------------------------------------------ */
static void particle_boundary_condition(orc_sim* S, gpat_particle* ptl, double xmin, double xmax,
                                        double ymin, double ymax, double zmin, double zmax)
{
    const gpat_params* P = &S->P;
    if (ptl->x < xmin && ptl->count_flag == GPAT_COUNT_FLAG_INBOX) {
        if (S->neighbors[0] < 0) {
#pragma omp atomic update
            S->leak += ptl->weight;
            ptl->count_flag = GPAT_COUNT_FLAG_ESCAPE_LX;
        } else {
            ptl->x = ptl->x - xmin + xmax;
        }
    } else if (ptl->x > xmax && ptl->count_flag == GPAT_COUNT_FLAG_INBOX) {
        if (S->neighbors[1] < 0) {
#pragma omp atomic update
            S->leak += ptl->weight;
            ptl->count_flag = GPAT_COUNT_FLAG_ESCAPE_HX;
        } else {
            ptl->x = ptl->x - xmax + xmin;
        }
    }
    if (P->ndim > 1) {
        if (ptl->y < ymin && ptl->count_flag == GPAT_COUNT_FLAG_INBOX) {
            if (S->neighbors[2] < 0) {
#pragma omp atomic update
                S->leak += ptl->weight;
                ptl->count_flag = GPAT_COUNT_FLAG_ESCAPE_LY;
            } else {
                ptl->y = ptl->y - ymin + ymax;
            }
        } else if (ptl->y > ymax && ptl->count_flag == GPAT_COUNT_FLAG_INBOX) {
            if (S->neighbors[3] < 0) {
#pragma omp atomic update
                S->leak += ptl->weight;
                ptl->count_flag = GPAT_COUNT_FLAG_ESCAPE_HY;
            } else {
                ptl->y = ptl->y - ymax + ymin;
            }
        }
    }
    if (P->ndim == 3 || (P->ndim == 2 && P->include_3rd_dim)) {
        if (ptl->z < zmin && ptl->count_flag == GPAT_COUNT_FLAG_INBOX) {
            if (S->neighbors[4] < 0) {
#pragma omp atomic update
                S->leak += ptl->weight;
                ptl->count_flag = GPAT_COUNT_FLAG_ESCAPE_LZ;
            } else {
                ptl->z = ptl->z - zmin + zmax;
            }
        } else if (ptl->z > zmax && ptl->count_flag == GPAT_COUNT_FLAG_INBOX) {
            if (S->neighbors[5] < 0) {
#pragma omp atomic update
                S->leak += ptl->weight;
                ptl->count_flag = GPAT_COUNT_FLAG_ESCAPE_HZ;
            } else {
                ptl->z = ptl->z - zmax + zmin;
            }
        }
    }
}

static void negp_or_bc(orc_sim* S, gpat_particle* ptl, const double e[6])
{
    if (ptl->p < 0.0) { /* PM:1602-1609 */
        ptl->count_flag = GPAT_COUNT_FLAG_OTHERS;
#pragma omp atomic update
        S->leak_negp += ptl->weight;
    } else {
        particle_boundary_condition(S, ptl, e[0], e[1], e[2], e[3], e[4], e[5]);
    }
}

/* one call of interp + kappa + push_particle_* (PM:1614-1691 / 1726-1802) */
static void one_push(orc_sim* S, gpat_particle* ptl, double t0, double dtf, int fixed_dt,
                     double* deltax, double* deltay, double* deltaz, double* deltap, double* deltav,
                     double* deltamu)
{
    const gpat_params* P = &S->P;
    double px = (ptl->x - P->xmin) / P->dx;
    double py = (ptl->y - P->ymin) / P->dy;
    double pz = (ptl->z - P->zmin) / P->dz;
    double rt = (ptl->t - t0) / dtf;
    int pos[3];
    double w[8], fields[NVAR], u[4];
    kappa_type kp;
    get_interp_parameters(S, px, py, pz, pos, w);
    interp_fields(S, pos, w, rt, fields);
    double aux[NAUX];
    for (int v = 0; v < NAUX; ++v) aux[v] = 1.0;
    if (P->deltab_flag || P->correlation_flag) interp_aux(S, pos, w, rt, aux); /* PM:1634-1639 */
    if (P->nlgc)
        calc_kappa_nlgc(S, ptl, fields, aux, &kp);
    else
        calc_kappa(S, ptl, fields, aux, &kp);
    step_uniforms(S, ptl, u);
    double shv[2] = {0.0, 0.0};
    const double* sh = NULL;
    if (P->acc_by_surface && P->ndim == 3) { /* PM:1662-1665, 1683-1686 */
        interp_acc_surface(S, pos, w, rt, &shv[0], &shv[1]);
        sh = shv;
    }
    if (P->focused_transport && (P->ndim == 3 || (P->ndim == 2 && P->include_3rd_dim))) /* PM:1653-1668 */
        push_particle_ft_3d_like(S, ptl, fields, aux, &kp, fixed_dt, u, step_uniform5(S, ptl), deltax, deltay,
                                 deltaz, deltap, deltav, deltamu, sh);
    else if (P->focused_transport) /* PM:1659-1662; the 1-D FT pusher reads an unassigned dx_dt */
        push_particle_2d_ft(S, ptl, fields, aux, &kp, fixed_dt, u, deltax, deltay, deltap, deltav, deltamu);
    else if (P->ndim == 1)
        push_particle_1d(S, ptl, fields, &kp, fixed_dt, u, deltax, deltap);
    else if (P->ndim == 2 && !P->include_3rd_dim)
        push_particle_2d(S, ptl, fields, &kp, fixed_dt, u, deltax, deltay, deltap);
    else
        push_particle_3d_like(S, ptl, fields, &kp, fixed_dt, u, deltax, deltay, deltaz, deltap, sh);
    set_rng_step(ptl, get_rng_step(ptl) + 1);
}

/* ------------------------------------------------------------------------ */
/* particle_mover_one_cycle, PM:1481-1833                                     */
/* ------------------------------------------------------------------------ */
static void particle_mover_one_cycle(orc_sim* S, double t0, double dtf, int nsteps_interval,
                                     int num_fine_steps)
{
    const gpat_params* P = &S->P;
    const double dt_fine = dtf / num_fine_steps;
    const double e[6] = {P->xmin - P->dx * 0.5, P->xmax + P->dx * 0.5, P->ymin - P->dy * 0.5,
                         P->ymax + P->dy * 0.5, P->zmin - P->dz * 0.5, P->zmax + P->dz * 0.5};
    uint64_t steps = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : steps) if (S->P.rng_mode != ORC_RNG_MT19937)
    for (int64_t i = S->nptl_old; i < S->nptl_current; ++i) {
        gpat_particle ptl = S->ptls[i];
        double deltax = 0.0, deltay = 0.0, deltaz = 0.0, deltap = 0.0, deltav = 0.0, deltamu = 0.0;
        double dt_target;
        int step = (int)ceil((ptl.t - t0) / dt_fine); /* PM:1570 */
        if (step <= 0)
            dt_target = dt_fine;
        else
            dt_target = step * dt_fine;
        if (dt_target > dtf) dt_target = dtf;

        /* safe check, PM:1581-1592 */
        if (ptl.p < 0.0 && ptl.count_flag == GPAT_COUNT_FLAG_INBOX) {
            ptl.count_flag = GPAT_COUNT_FLAG_OTHERS;
#pragma omp atomic update
            S->leak_negp += ptl.weight;
        } else {
            particle_boundary_condition(S, &ptl, e[0], e[1], e[2], e[3], e[4], e[5]);
        }
        if (ptl.count_flag != GPAT_COUNT_FLAG_INBOX) {
            S->ptls[i] = ptl;
            continue;
        }
        while (dt_target < (dtf + dt_fine * (double)0.1f)) { /* PM:1596 */
            if (ptl.count_flag != GPAT_COUNT_FLAG_INBOX) break;
            while ((ptl.t - t0) < dt_target && ptl.count_flag == GPAT_COUNT_FLAG_INBOX) {
                negp_or_bc(S, &ptl, e);
                if (ptl.count_flag != GPAT_COUNT_FLAG_INBOX) break;
                one_push(S, &ptl, t0, dtf, 0, &deltax, &deltay, &deltaz, &deltap, &deltav, &deltamu);
                steps++;
                ptl.nsteps_pushed = (ptl.nsteps_pushed + 1) % nsteps_interval; /* PM:1694 */
                track_after_push(S, &ptl);                                     /* PM:1697-1703 */
            }
            /* make sure ptl%t reaches the target exactly, PM:1707-1826 */
            if ((ptl.t - t0) > dt_target && ptl.count_flag == GPAT_COUNT_FLAG_INBOX) {
                ptl.x = ptl.x - deltax;
                ptl.y = ptl.y - deltay;
                ptl.z = ptl.z - deltaz;
                ptl.p = ptl.p - deltap;
                ptl.v = ptl.v - deltav;   /* PM:1712-1713: zero for Parker transport */
                ptl.mu = ptl.mu - deltamu;
                ptl.t = ptl.t - ptl.dt;
                double dt_old = ptl.dt;
                ptl.dt = t0 + dt_target - ptl.t;
                if (ptl.dt > 0) {
                    if (ptl.tag_splitted < 0 && ptl.nsteps_pushed == 0) { /* PM:1717-1721 */
                        ptl.nsteps_tracked = ptl.nsteps_tracked - 1; /* back one step */
                        ptl.nsteps_pushed = nsteps_interval - 2;
                    } else {
                        ptl.nsteps_pushed = ptl.nsteps_pushed - 1; /* PM:1723 */
                    }
                    one_push(S, &ptl, t0, dtf, 1, &deltax, &deltay, &deltaz, &deltap, &deltav, &deltamu);
                    steps++;
                    /* Fortran mod keeps the sign of the dividend, like C's % */
                    ptl.nsteps_pushed = (ptl.nsteps_pushed + 1) % nsteps_interval;
                    track_after_push(S, &ptl);                      /* PM:1806-1812 */
                }
                ptl.dt = dt_old;
                negp_or_bc(S, &ptl, e);
            }
            dt_target = dt_target + dt_fine;
        }
        S->ptls[i] = ptl;
    }
    S->steps += steps;
}

/* remove_particles, PM:5365-5403 (serial swap-with-tail) */
static void remove_particles(orc_sim* S, int dump_escaped_dist)
{
    if (S->nptl_current > 0) {
        int64_t nremoved = 0;
        int64_t i = 1; /* 1-based like the reference */
        while (i <= S->nptl_current) {
            if ((S->nptl_current - i) == (nremoved - 1)) break;
            gpat_particle* pi = &S->ptls[i - 1];
            if (pi->count_flag == GPAT_COUNT_FLAG_INBOX) {
                i = i + 1;
            } else {
                if (pi->count_flag < 0) {
                    S->nptl_escaped++;
                    if (dump_escaped_dist && S->nptl_escaped <= S->nptl_escaped_max)
                        S->escaped[S->nptl_escaped - 1] = *pi;
                }
                gpat_particle ptl1 = S->ptls[S->nptl_current - nremoved - 1];
                S->ptls[S->nptl_current - nremoved - 1] = *pi;
                *pi = ptl1;
                nremoved++;
            }
        }
        S->nptl_current -= nremoved;
    }
}

static void set_dt_min_max(orc_sim* S, double dtf) /* PM:5519-5524 */
{
    S->dt_min = S->P.dt_min_rel * dtf;
    S->dt_max = S->P.dt_max_rel * dtf;
}

/* particle_mover, PM:1846-1974, for a 1x1x1 topology (one cycle, no exchange) */
void orc_particle_mover(orc_sim* S, double t0, double dtf, int nsteps_interval, int num_fine_steps,
                        int dump_escaped_dist, uint64_t* steps_done)
{
    const gpat_params* P = &S->P;
    uint64_t s0 = S->steps;
    S->nptl_old = 0;
    set_dt_min_max(S, dtf);
    for (int64_t i = 0; i < S->nptl_current; ++i) S->ptls[i].nsteps_tracked = 1; /* PM:1913 */
    if (S->nptl_old < S->nptl_current)
        particle_mover_one_cycle(S, t0, dtf, nsteps_interval, num_fine_steps);
    remove_particles(S, dump_escaped_dist);
    S->nptl_old = S->nptl_current; /* add_neighbor_particles, PM:5411 */
    /* final pass with the un-extended box, PM:1959-1971 */
    for (int64_t i = 0; i < S->nptl_current; ++i) {
        gpat_particle ptl = S->ptls[i];
        if (ptl.p < 0.0 && ptl.count_flag != GPAT_COUNT_FLAG_INBOX) {
            ptl.count_flag = GPAT_COUNT_FLAG_OTHERS;
            S->leak_negp += ptl.weight;
        } else {
            particle_boundary_condition(S, &ptl, P->xmin, P->xmax, P->ymin, P->ymax, P->zmin,
                                        P->zmax);
        }
        S->ptls[i] = ptl;
    }
    remove_particles(S, dump_escaped_dist);
    if (steps_done) *steps_done = S->steps - s0;
}

/* test hook mirrored by gpat_debug_push_n: exactly nsteps adaptive pushes */
void orc_debug_push_n(orc_sim* S, double t0, double dtf, int nsteps, uint64_t* steps_done)
{
    const gpat_params* P = &S->P;
    const double e[6] = {P->xmin - P->dx * 0.5, P->xmax + P->dx * 0.5, P->ymin - P->dy * 0.5,
                         P->ymax + P->dy * 0.5, P->zmin - P->dz * 0.5, P->zmax + P->dz * 0.5};
    set_dt_min_max(S, dtf);
    uint64_t steps = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : steps) if (S->P.rng_mode != ORC_RNG_MT19937)
    for (int64_t i = 0; i < S->nptl_current; ++i) {
        gpat_particle ptl = S->ptls[i];
        double dx_, dy_, dz_ = 0.0, dp_, dv_ = 0.0, dmu_ = 0.0;
        for (int n = 0; n < nsteps && ptl.count_flag == GPAT_COUNT_FLAG_INBOX; ++n) {
            negp_or_bc(S, &ptl, e);
            if (ptl.count_flag != GPAT_COUNT_FLAG_INBOX) break;
            one_push(S, &ptl, t0, dtf, 0, &dx_, &dy_, &dz_, &dp_, &dv_, &dmu_);
            steps++;
            ptl.nsteps_pushed = (ptl.nsteps_pushed + 1) % (1 << 30);
        }
        S->ptls[i] = ptl;
    }
    S->steps += steps;
    if (steps_done) *steps_done = steps;
}

/* ------------------------------------------------------------------------ */
/* injection: PM:454-530 (whole-field branch) + PM:385-441                     */
/* ------------------------------------------------------------------------ */
/* ------------------------------------------------------------------------ */
/* particle tracking: init_particle_tracking (PM:5825-5879), reset (PM:5884),  */
/* is_particle_selected (PM:5920-5959), locate_particle (PM:5967-5990)         */
/* ------------------------------------------------------------------------ */
void orc_init_tracking(orc_sim* S, const int32_t* tags, int ncols, int64_t nptl_tracking,
                       int nsteps_interval)
{
    free(S->tags_tracking); free(S->particles_tracked);
    S->track_particle_flag = 1;
    S->split_times_max = ncols - 2;
    S->nptl_tracking = nptl_tracking;
    S->tags_tracking = (int32_t*)malloc(sizeof(int32_t) * (size_t)ncols * nptl_tracking);
    memcpy(S->tags_tracking, tags, sizeof(int32_t) * (size_t)ncols * nptl_tracking);
    /* `1.0 / dt_min_rel` promotes the default-real literal; ceiling of the f64 quotient */
    S->nsteps_tracking_max = (int64_t)ceil((1.0 / S->P.dt_min_rel) / nsteps_interval) + 1;
    S->particles_tracked = (gpat_particle*)calloc((size_t)S->nsteps_tracking_max * nptl_tracking,
                                                  sizeof(gpat_particle));
}

void orc_reset_tracked(orc_sim* S)
{
    if (S->particles_tracked)
        memset(S->particles_tracked, 0, sizeof(gpat_particle) * (size_t)S->nsteps_tracking_max * S->nptl_tracking);
}

int64_t orc_get_tracked(const orc_sim* S, gpat_particle* out, int64_t* nsteps_max)
{
    if (nsteps_max) *nsteps_max = S->nsteps_tracking_max;
    if (out && S->particles_tracked)
        memcpy(out, S->particles_tracked, sizeof(gpat_particle) * (size_t)S->nsteps_tracking_max * S->nptl_tracking);
    return S->nptl_tracking;
}

#define TAG(r, c) S->tags_tracking[((r) - 1) + (size_t)(S->split_times_max + 2) * ((c) - 1)]
/* findloc(tags(row, lo:hi), v, dim=1 [, back]) relative to lo (1-based), 0 if absent */
static int64_t findloc_row(const orc_sim* S, int row, int64_t lo, int64_t hi, int32_t v, int back)
{
    if (!back) { for (int64_t c = lo; c <= hi; ++c) if (TAG(row, c) == v) return c - lo + 1; }
    else { for (int64_t c = hi; c >= lo; --c) if (TAG(row, c) == v) return c - lo + 1; }
    return 0;
}

static int is_particle_selected(const orc_sim* S, const gpat_particle* ptl, int64_t* iptl_lo,
                                int64_t* iptl_hi)
{
    *iptl_lo = -1; *iptl_hi = -1;
    int nsplit = ptl->split_times;
    if (nsplit > S->split_times_max) return 0;
    int64_t n = S->nptl_tracking;
    int64_t i1 = findloc_row(S, 1, 1, n, ptl->origin, 0);
    if (i1 <= 0) return 0;
    int64_t i2 = findloc_row(S, 1, 1, n, ptl->origin, 1);
    int64_t i3 = findloc_row(S, 2, i1, i2, abs(ptl->tag_injected), 0);
    if (i3 <= 0) return 0;
    int64_t i4 = findloc_row(S, 2, i1, i2, abs(ptl->tag_injected), 1);
    i3 = i3 + i1 - 1;
    i4 = i4 + i1 - 1;
    if (nsplit > 0) {
        int64_t i5 = findloc_row(S, nsplit + 2, i3, i4, abs(ptl->tag_splitted), 0);
        if (i5 <= 0) return 0;
        int64_t i6 = findloc_row(S, nsplit + 2, i3, i4, abs(ptl->tag_splitted), 1);
        *iptl_lo = i5 + i3 - 1;
        *iptl_hi = i6 + i3 - 1;
    } else {
        *iptl_lo = i3;
        *iptl_hi = i4;
    }
    return 1;
}

/* particles_tracked(n, lo:hi) = ptl */
static void record_tracked(orc_sim* S, const gpat_particle* ptl, int64_t lo, int64_t hi)
{
    int64_t n = ptl->nsteps_tracked;
    if (n < 1 || n > S->nsteps_tracking_max) return; /* out of bounds in the reference */
    for (int64_t c = lo; c <= hi; ++c)
        S->particles_tracked[(n - 1) + (size_t)S->nsteps_tracking_max * (c - 1)] = *ptl;
}

/* the tracking block after every push, PM:1697-1703 / 1806-1812 */
static void track_after_push(orc_sim* S, gpat_particle* ptl)
{
    if (S->track_particle_flag && ptl->tag_splitted < 0 && ptl->nsteps_pushed == 0) {
        int64_t lo, hi;
        is_particle_selected(S, ptl, &lo, &hi); /* locate_particle: same search, no checks */
        ptl->nsteps_tracked = ptl->nsteps_tracked + 1;
        if (lo > 0) record_tracked(S, ptl, lo, hi);
    }
}

/* inject_one_particle, PM:385-441; `st` continues the particle's injection stream */
static void inject_one_particle(orc_sim* S, inj_stream* st, double xpos, double ypos, double zpos,
                                int dist_flag, double particle_v0, double mu, double t_frame,
                                double dt_mhd, double dt, double power_index)
{
    const gpat_params* P = &S->P;
    gpat_particle* q = &S->ptls[S->nptl_current - 1];
    memset(q, 0, sizeof(*q));
    q->x = xpos; q->y = ypos; q->z = zpos;
    if (dist_flag == 0) { /* PM:399-407 */
        double ftest = 1.0, fxp = 0.5, ptmp = 0.0;
        while (ftest > fxp) {
            ptmp = (inj_next(st) * (P->pmax - P->pmin) + P->pmin) / P->p0;
            fxp = sq(ptmp) * exp(-sq(ptmp));
            ftest = inj_next(st) * (double)0.37f;
        }
        q->p = ptmp * P->p0;
    } else if (dist_flag == 1) {
        q->p = P->p0;
    } else if (dist_flag == 2) { /* PM:410-418 */
        double r01 = inj_next(st);
        if ((int)power_index == 1) {
            q->p = pow(P->pmax / P->p0, r01) * P->p0;
        } else {
            double norm = pow(P->pmax, -power_index + 1) - pow(P->p0, -power_index + 1);
            q->p = pow(r01 * norm + pow(P->p0, -power_index + 1), 1.0 / (-power_index + 1));
        }
    }
    q->v = particle_v0 * q->p / P->p0;
    q->mu = mu;
    q->weight = 1.0;
    q->t = t_frame + inj_next(st) * dt_mhd;
    q->dt = dt;
    q->split_times = 0;
    q->count_flag = GPAT_COUNT_FLAG_INBOX;
    q->origin = P->mpi_rank;
    q->nsteps_tracked = 0;
    q->nsteps_pushed = 0;
    q->tag_injected = (int32_t)S->tag_max;
    S->tag_max++;
    q->tag_splitted = 1;
    set_rng_step(q, 0);
    if (S->track_particle_flag) { /* PM:434-440 */
        int64_t lo, hi;
        if (is_particle_selected(S, q, &lo, &hi)) {
            q->nsteps_tracked = 1;
            q->tag_injected = -q->tag_injected;
            q->tag_splitted = -1;
            record_tracked(S, q, lo, hi);
        }
    }
}

void orc_inject_uniform(orc_sim* S, int64_t nptl, double dt, int dist_flag, double particle_v0,
                        double t_frame, double dt_mhd, const double part_box[6],
                        double power_index)
{
    const gpat_params* P = &S->P;
    const double mu_max = (double)0.99f; /* PM:121 */
    double xmin_box = part_box[0], ymin_box = part_box[1], zmin_box = part_box[2];
    double xmax_box = part_box[3], ymax_box = part_box[4], zmax_box = part_box[5];
    S->nptl_inject = nptl;
    for (int64_t i = 0; i < nptl; ++i) {
        S->nptl_current++;
        if (S->nptl_current > S->nptl_max) S->nptl_current = S->nptl_max; /* PM:491-492 */
        inj_stream st = {S, (uint32_t)S->tag_max, (uint32_t)P->mpi_rank, 0, {0, 0, 0, 0}};
        double xtmp = inj_next(&st) * (xmax_box - xmin_box) + xmin_box;
        double ytmp = inj_next(&st) * (ymax_box - ymin_box) + ymin_box;
        double ztmp = inj_next(&st) * (zmax_box - zmin_box) + zmin_box;
        double mu_tmp = mu_max * (2.0 * inj_next(&st) - 1.0);
        inject_one_particle(S, &st, xtmp, ytmp, ztmp, dist_flag, particle_v0, mu_tmp, t_frame,
                            dt_mhd, dt, power_index);
    }
}

/* ------------------------------------------------------------------------ */
/* shock injection: locate_shock_xpos (MD:1988-2006), interp_shock_location   */
/* (MD:2022-2045), inject_particles_at_shock (PM:542-633).                    */
/* The reference never zeroes sx1/sx2 in 2-D/3-D (MD:2037-2049: `sx1 = sx1 +`  */
/* on uninitialised locals) -- undefined there; HERE THEY START AT ZERO, the   */
/* value a fresh stack frame usually holds.  Everything else is kept as        */
/* written: the time weights are swapped (`sx2*(1-rt) + sx1*rt`, so rt = 0     */
/* picks the LATER frame), rz is computed from dpy (PM:581), the weights of    */
/* the two rows therefore do not sum to one, and t is the frame time itself.   */
/* ------------------------------------------------------------------------ */
static void locate_shock_xpos(const orc_sim* S, const float* fa, int32_t* out)
{
    for (int k = 0; k < S->nzg; ++k)
        for (int j = 0; j < S->nyg; ++j) {
            float best = -1.0f;
            int at = 1; /* maxloc returns the FIRST maximum, 1-based along the ghosted x extent */
            for (int i = 0; i < S->nxg; ++i) {
                float v = fabsf(fa[FIDX(S, NFIELDS + 0, i, j, k)]);
                if (v > best) { best = v; at = i + 1; }
            }
            out[j + (size_t)S->nyg * k] = at;
        }
}

int64_t orc_inject_at_shock(orc_sim* S, int64_t nptl, double dt, int dist_flag, double particle_v0,
                            double t_frame, double power_index)
{
    const gpat_params* P = &S->P;
    const double mu_max = (double)0.99f;
    const double xmin = P->xmin, xmax = P->xmax, ymin = P->ymin, ymax = P->ymax;
    const double zmin = P->zmin, zmax = P->zmax;
    const int nxg_f = P->nx + 4; /* fconfig%nxg, SS:176 */
    int32_t* sx2map = (int32_t*)malloc(sizeof(int32_t) * (size_t)S->nyg * S->nzg);
    locate_shock_xpos(S, S->farray2, sx2map); /* only shock_xpos2 survives rt = 0 */
    S->nptl_inject = nptl;
    for (int64_t i = 0; i < nptl; ++i) {
        S->nptl_current++;
        if (S->nptl_current > S->nptl_max) S->nptl_current = S->nptl_max;
        gpat_particle* q = &S->ptls[S->nptl_current - 1];
        inj_stream st = {S, (uint32_t)S->tag_max, (uint32_t)P->mpi_rank, 0, {0, 0, 0, 0}};
        memset(q, 0, sizeof(*q));
        q->y = inj_next(&st) * (ymax - ymin) + ymin;
        double dpy = q->y / P->dy;
        int iy = (int)floor(dpy);
        q->z = inj_next(&st) * (zmax - zmin) + zmin;
        double dpz = q->z / P->dz;
        int iz = (int)floor(dpz);
        double ry = dpy - iy;
        double rz = dpy - iy; /* PM:581, sic */
        double w[4] = {(1 - ry) * (1 - rz), ry * (1 - rz), (1 - ry) * rz, ry * rz};
        double sx1 = 0.0, sx2 = 0.0; /* see the header of this section */
        if (P->ndim == 1) {
            sx2 = sx2map[0];
        } else if (P->ndim == 2) {
            for (int j = 0; j <= 1; ++j) {
                int fj = iy + j - 1 + 1; /* Fortran index iy+j-1, lower bound -1 */
                if (fj < 0) fj = 0;
                if (fj > S->nyg - 1) fj = S->nyg - 1;
                sx2 = sx2 + sx2map[fj] * w[j];
            }
        } else {
            for (int k = 0; k <= 1; ++k)
                for (int j = 0; j <= 1; ++j) {
                    int fj = iy + j, fk = iz + k;
                    if (fj < 0) fj = 0;
                    if (fj > S->nyg - 1) fj = S->nyg - 1;
                    if (fk < 0) fk = 0;
                    if (fk > S->nzg - 1) fk = S->nzg - 1;
                    sx2 = sx2 + sx2map[fj + (size_t)S->nyg * fk] * w[k * 2 + j];
                }
        }
        (void)dpz; (void)iz;
        double shock_xpos = (sx2 * (1.0 - 0.0) + sx1 * 0.0) + 2; /* two ghost cells */
        q->x = shock_xpos * (xmax - xmin) / nxg_f;
        if (dist_flag == 0) { /* PM:588-596: a different Maxwellian envelope than inject_one_particle */
            double ftest = 1.0, fxp = 0.5, ptmp = 0.0;
            while (ftest > fxp) {
                ptmp = (inj_next(&st) * (P->pmax - P->pmin) + P->pmin) / P->p0;
                fxp = sq(ptmp) * exp(-0.5 * sq(ptmp));
                ftest = inj_next(&st) * (double)0.75f;
            }
            q->p = ptmp * P->p0;
        } else if (dist_flag == 1) {
            q->p = P->p0;
        } else if (dist_flag == 2) {
            double r01 = inj_next(&st);
            if ((int)power_index == 1) {
                q->p = pow(P->pmax / P->p0, r01) * P->p0;
            } else {
                double norm = pow(P->pmax, -power_index + 1) - pow(P->p0, -power_index + 1);
                q->p = pow(r01 * norm + pow(P->p0, -power_index + 1), 1.0 / (-power_index + 1));
            }
        }
        q->v = particle_v0 * q->p / P->p0;
        q->mu = mu_max * (2.0 * inj_next(&st) - 1.0);
        q->weight = 1.0;
        q->t = t_frame;
        q->dt = dt;
        q->split_times = 0;
        q->count_flag = GPAT_COUNT_FLAG_INBOX;
        q->origin = P->mpi_rank;
        q->nsteps_tracked = 0;
        q->nsteps_pushed = 0;
        q->tag_injected = (int32_t)S->tag_max;
        S->tag_max++;
        q->tag_splitted = 1;
        set_rng_step(q, 0);
        if (S->track_particle_flag) {
            int64_t lo, hi;
            if (is_particle_selected(S, q, &lo, &hi)) {
                q->nsteps_tracked = 1;
                q->tag_injected = -q->tag_injected;
                q->tag_splitted = -1;
                record_tracked(S, q, lo, hi);
            }
        }
    }
    free(sx2map);
    return nptl;
}

/* ------------------------------------------------------------------------ */
/* targeted injection: inject_particles_at_large_jz / _absj / _divv / _rho    */
/* (PM:785-905, 919-1061, 1250-1341, 1356-1468) with the cell counters        */
/* get_ncells_large_* (MD:2211-2261, 2269-2335, 2385-2455, 2463-2498),        */
/* Cartesian uniform grid, one rank per field copy (mpi_sub_size = 1).        */
/* mode: 1 jz, 2 absj, 4 divv, 5 rho (3 = db2 needs the deltab maps).         */
/* ------------------------------------------------------------------------ */
static int cell_in_box(const orc_sim* S, int ix, int iy, int iz, const double* b)
{
    /* xpos_local(ix) = dx*(ix-1) + xmin for the physical cells (MD:2129-2171) */
    const gpat_params* P = &S->P;
    int inx = (P->dx * (ix - 1) + P->xmin) > b[0] && (P->dx * (ix - 1) + P->xmin) < b[3];
    int iny = 1, inz = 1;
    if (P->ndim > 1) iny = (P->dy * (iy - 1) + P->ymin) > b[1] && (P->dy * (iy - 1) + P->ymin) < b[4];
    if (P->ndim > 2) inz = (P->dz * (iz - 1) + P->zmin) > b[2] && (P->dz * (iz - 1) + P->zmin) < b[5];
    return inx && iny && inz;
}

/* farray1(slot, ix, iy, iz) with Fortran indices (lower bound -1 on resolved axes) */
static float fa1(const orc_sim* S, int slot, int ix, int iy, int iz)
{
    int cj = (S->P.ndim > 1) ? iy + 1 : 0, ck = (S->P.ndim > 2) ? iz + 1 : 0;
    return S->farray1[FIDX(S, slot - 1, ix + 1, cj, ck)];
}

int64_t orc_ncells_large(const orc_sim* S, int mode, double vmin, const double part_box[6])
{
    const gpat_params* P = &S->P;
    int64_t n = 0;
    for (int iz = 1; iz <= P->nz; ++iz)
        for (int iy = 1; iy <= P->ny; ++iy)
            for (int ix = 1; ix <= P->nx; ++ix) {
                if (!cell_in_box(S, ix, iy, iz, part_box)) continue;
                double v;
                if (mode == 1) { /* MD:2235-2236: FP32 difference and abs, then promoted */
                    float d = fa1(S, NFIELDS + 16, ix, iy, iz) - fa1(S, NFIELDS + 14, ix, iy, iz);
                    v = (double)fabsf(d);
                } else if (mode == 2) { /* MD:2306-2311: all in FP32 */
                    float a = fa1(S, NFIELDS + 18, ix, iy, iz) - fa1(S, NFIELDS + 20, ix, iy, iz);
                    float b = fa1(S, NFIELDS + 19, ix, iy, iz) - fa1(S, NFIELDS + 15, ix, iy, iz);
                    float c = fa1(S, NFIELDS + 14, ix, iy, iz) - fa1(S, NFIELDS + 16, ix, iy, iz);
                    float s2 = a * a + b * b;
                    s2 = s2 + c * c;
                    v = (double)sqrtf(s2);
                } else if (mode == 4) {
                    /* MD:2417-2424: `divv = farray1(nfields+1, :, :, :)` assigns the WHOLE array
                     * (ghosts included) to an allocatable declared (nx,ny,nz); with Fortran 2003
                     * reallocation-on-assignment (gfortran's default) divv gets lower bounds 1,
                     * so divv(ix,iy,iz) is farray1(.., ix-2, iy-2, iz-2): the counter looks two
                     * cells to the lower-left of the cell whose position it tests.  Kept. */
                    int sx = ix - 2, sy = (P->ndim > 1) ? iy - 2 : iy, sz = (P->ndim > 2) ? iz - 2 : iz;
                    v = (double)fa1(S, NFIELDS + 1, sx, sy, sz);
                    if (P->ndim > 1) {
                        v = v + (double)fa1(S, NFIELDS + 5, sx, sy, sz);
                        if (P->ndim > 2) v = v + (double)fa1(S, NFIELDS + 9, sx, sy, sz);
                    }
                    v = -v;
                } else if (mode == 3) { /* get_ncells_large_db2, MD:2371: sigma2_slab_1(1, ix, iy, iz) */
                    int cj = (P->ndim > 1) ? iy + 1 : 0, ck = (P->ndim > 2) ? iz + 1 : 0;
                    v = (double)S->aux1[AIDX(S, 0, 0, ix + 1, cj, ck)];
                } else { /* MD:2490 */
                    v = (double)fa1(S, 4, ix, iy, iz);
                }
                if (v > vmin) n++;
            }
    return n;
}

int64_t orc_inject_targeted(orc_sim* S, int mode, int64_t nptl, double dt, int dist_flag,
                            double particle_v0, double t_frame, double dt_mhd,
                            const double part_box[6], double power_index, int inject_same_nptl,
                            double vmin, int64_t ncells_norm)
{
    const gpat_params* P = &S->P;
    const double mu_max = (double)0.99f;
    const double xmin = P->xmin, ymin = P->ymin, zmin = P->zmin;
    const double xmax = P->xmax, ymax = P->ymax, zmax = P->zmax;
    int64_t ncells = orc_ncells_large(S, mode, vmin, part_box);
    /* one rank per field copy: mpi_sub_size = 1 and the "global" count is the local one */
    int64_t denom = inject_same_nptl ? ncells : ncells_norm;
    int64_t nptl_inject = (int64_t)((double)(nptl * 1) * ((double)ncells / (double)denom));
    if (denom == 0) nptl_inject = 0; /* 0/0 -> int(NaN) is undefined in the reference */
    S->nptl_inject = nptl_inject;
    for (int64_t i = 0; i < nptl_inject; ++i) {
        S->nptl_current++;
        if (S->nptl_current > S->nptl_max) S->nptl_current = S->nptl_max;
        inj_stream st = {S, (uint32_t)S->tag_max, (uint32_t)P->mpi_rank, 0, {0, 0, 0, 0}};
        double xtmp = part_box[0], ytmp = part_box[1], ztmp = part_box[2];
        double crit = (mode == 4) ? 2.0 : (mode == 5 ? 0.0 : -2.0);
        for (;;) {
            int again = (mode == 4) ? (-crit < vmin) : (crit < vmin);
            if (!again) break;
            xtmp = inj_next(&st) * (xmax - xmin) + xmin;
            ytmp = inj_next(&st) * (ymax - ymin) + ymin;
            ztmp = inj_next(&st) * (zmax - zmin) + zmin;
            if (xtmp >= part_box[0] && xtmp <= part_box[3] && ytmp >= part_box[1] &&
                ytmp <= part_box[4] && ztmp >= part_box[2] && ztmp <= part_box[5]) {
                double px = (xtmp - xmin) / P->dx, py = (ytmp - ymin) / P->dy;
                double pz = (ztmp - zmin) / P->dz;
                int pos[3];
                double w[8], fields[NVAR];
                get_interp_parameters(S, px, py, pz, pos, w);
                interp_fields(S, pos, w, 0.0, fields);
                if (mode == 3) { /* PM:1173-1174 */
                    double ax[NAUX];
                    interp_aux(S, pos, w, 0.0, ax);
                    crit = ax[0];
                } else if (mode == 1) {
                    crit = fabs(FG(16) - FG(14));
                } else if (mode == 2) {
                    crit = sqrt(sq(FG(18) - FG(20)) + sq(FG(19) - FG(15)) + sq(FG(14) - FG(16)));
                } else if (mode == 4) {
                    crit = FG(1);
                    if (P->ndim > 1) {
                        crit = crit + FG(5);
                        if (P->ndim > 2) crit = crit + FG(9);
                    }
                } else {
                    crit = F(4);
                }
            } else {
                crit = (mode == 4) ? 3.0 : (mode == 5 ? 0.0 : -3.0);
            }
        }
        double mu_tmp = mu_max * (2.0 * inj_next(&st) - 1.0);
        inject_one_particle(S, &st, xtmp, ytmp, ztmp, dist_flag, particle_v0, mu_tmp, t_frame,
                            dt_mhd, dt, power_index);
    }
    return nptl_inject;
}

/* ------------------------------------------------------------------------ */
/* split_particle, PM:5430-5480 (untracked particles)                         */
/* ------------------------------------------------------------------------ */
static double powi(double x, int m) /* libgcc __powidf2: what gfortran emits for dp**integer */
{
    unsigned int n = (m < 0) ? -(unsigned int)m : (unsigned int)m;
    double y = (n % 2) ? x : 1.0;
    while (n >>= 1) {
        x = x * x;
        if (n % 2) y *= x;
    }
    return (m < 0) ? 1.0 / y : y;
}

void orc_split(orc_sim* S, double split_ratio, double pmin_split, int nsteps_interval)
{
    (void)nsteps_interval;
    const gpat_params* P = &S->P;
    int64_t nptl = S->nptl_current;
    for (int64_t i = 0; i < nptl; ++i) {
        gpat_particle ptl = S->ptls[i];
        double p_threshold = pmin_split * P->p0 * powi(split_ratio, ptl.split_times);
        if (ptl.p > p_threshold && ptl.p <= P->pmax) {
            S->nptl_current++;
            if (S->nptl_current > S->nptl_max) {
                S->nptl_current = S->nptl_max;
                return;
            }
            S->nptl_split++;
            ptl.weight = (double)powf(0.5f, 1.0f + (float)ptl.split_times); /* PM:5449 */
            ptl.split_times = (int8_t)(ptl.split_times + 1);
            gpat_particle* child = &S->ptls[S->nptl_current - 1];
            *child = ptl;
            if (ptl.tag_splitted < 0) { /* tracked particle, PM:5452-5473 */
                int64_t lo, hi;
                child->tag_splitted = ptl.tag_splitted - (1 << (ptl.split_times - 1));
                if (is_particle_selected(S, child, &lo, &hi)) {
                    if (ptl.nsteps_pushed == 0) {
                        child->nsteps_tracked = child->nsteps_tracked + 1;
                        record_tracked(S, child, lo, hi);
                    }
                } else {
                    child->tag_splitted = -child->tag_splitted;
                }
                if (is_particle_selected(S, &ptl, &lo, &hi)) {
                    if (ptl.nsteps_pushed == 0) {
                        ptl.nsteps_tracked = ptl.nsteps_tracked + 1;
                        record_tracked(S, &ptl, lo, hi);
                    }
                } else {
                    ptl.tag_splitted = -ptl.tag_splitted; /* stop tracking */
                }
            } else {
                child->tag_splitted = ptl.tag_splitted + (1 << (ptl.split_times - 1)); /* PM:5475 */
            }
            S->ptls[i] = ptl;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* diagnostics: DG:196-209 (edges), DG:738-879 (binning), DG:116-170, 1691    */
/* ------------------------------------------------------------------------ */
static int64_t ifloor(double v, int* ok)
{
    if (!(v > -2.0e9 && v < 2.0e9)) { *ok = 0; return 0; } /* NaN/Inf: never a valid bin */
    return (int64_t)floor(v);
}

void orc_hist_edges(const orc_sim* S, int which, double* pedges, double* muedges)
{
    const gpat_params* P = &S->P;
    double pmin = which ? P->local[which - 1].pmin : P->pmin;
    double pmax = which ? P->local[which - 1].pmax : P->pmax;
    int np = which ? P->local[which - 1].npbins : P->npp_global;
    int nmu = which ? P->local[which - 1].nmu : P->nmu_global;
    double pmin_log = log10(pmin), pmax_log = log10(pmax);
    double dp_log = (pmax_log - pmin_log) / np;
    for (int i = 1; i <= np + 1; ++i) pedges[i - 1] = pow(10.0, pmin_log + (i - 1) * dp_log);
    double dmu = (double)(2.0f / (float)nmu); /* DG:206: default-real division */
    for (int i = 1; i <= nmu + 1; ++i) muedges[i - 1] = -1.0 + (i - 1) * dmu;
}

/* fglobal: (nmu_global, npp_global); flocal[k]: (nmu, npbins, nrx, nry, nrz), all
 * column-major and zeroed here.  quick[8] and pmax as in gpat_diagnostics. */
void orc_diagnostics(const orc_sim* S, int local_dist, double* fglobal, double* const flocal[4],
                     double quick[8], double* pmax_out)
{
    const gpat_params* P = &S->P;
    const int nmu_g = P->nmu_global, npp = P->npp_global;
    double pmin_log = log10(P->pmin), pmax_log = log10(P->pmax);
    double dp_log = (pmax_log - pmin_log) / npp;
    double dmu = (double)(2.0f / (float)nmu_g);
    if (fglobal) memset(fglobal, 0, sizeof(double) * (size_t)nmu_g * npp);
    int nrx[4], nry[4], nrz[4];
    double dxd[4], dyd[4], dzd[4], pminl[4], dpl[4], dmul[4];
    for (int k = 0; k < 4; ++k) {
        const gpat_hist_spec* h = &P->local[k];
        if (!h->enabled) continue;
        nrx[k] = (P->nx + h->rx - 1) / h->rx; /* DG:2135-2137 */
        nry[k] = (P->ny + h->ry - 1) / h->ry;
        nrz[k] = (P->nz + h->rz - 1) / h->rz;
        dxd[k] = P->lx / nrx[k]; /* DG:271-276 (nrx_mhd == nrx for 1x1x1) */
        dyd[k] = P->ly / nry[k];
        dzd[k] = P->lz / nrz[k];
        pminl[k] = log10(h->pmin);
        dpl[k] = (log10(h->pmax) - pminl[k]) / h->npbins;
        dmul[k] = (double)(2.0f / (float)h->nmu);
        if (local_dist && flocal && flocal[k])
            memset(flocal[k], 0, sizeof(double) * (size_t)h->nmu * h->npbins * nrx[k] * nry[k] * nrz[k]);
    }
    double q3 = 0.0, q6 = 0.0, pdt_min = 1.0, pdt_max = 0.0, pmx = 0.0;
    for (int64_t n = 0; n < S->nptl_current; ++n) {
        const gpat_particle* ptl = &S->ptls[n];
        double x = ptl->x, y = ptl->y, z = ptl->z, p = ptl->p, mu = ptl->mu, weight = ptl->weight;
        if (fglobal && p > P->pmin && p <= P->pmax && mu >= -1.0 && mu <= 1.0) {
            int ok = 1;
            int64_t ip = ifloor((log10(p) - pmin_log) / dp_log, &ok) + 1;
            int64_t imu = ifloor((mu + 1.0) / dmu, &ok) + 1;
            /* Column-major address like fglobal(imu,ip); an index past the array is
             * undefined behaviour in the reference (DG:775-779) and is dropped here. */
            int64_t lin = (imu - 1) + (ip - 1) * (int64_t)nmu_g;
            if (ok && ip >= 1 && imu >= 1 && lin >= 0 && lin < (int64_t)nmu_g * npp)
                fglobal[lin] += ptl->weight;
        }
        if (local_dist && flocal) {
            for (int k = 0; k < 4; ++k) {
                const gpat_hist_spec* h = &P->local[k];
                if (!h->enabled || !flocal[k]) continue;
                int ok = 1;
                int64_t ix = ifloor((x - P->xmin) / dxd[k], &ok) + 1;
                int64_t iy = ifloor((y - P->ymin) / dyd[k], &ok) + 1;
                int64_t iz = ifloor((z - P->zmin) / dzd[k], &ok) + 1;
                int64_t ip = ifloor((log10(p) - pminl[k]) / dpl[k], &ok) + 1;
                int64_t imu = ifloor((mu + 1.0) / dmul[k], &ok) + 1;
                int condx = ix >= 1 && ix <= nrx[k];
                int condy = iy >= 1 && iy <= nry[k];
                int condz = iz >= 1 && iz <= nrz[k];
                int condp = ip > 0 && ip < h->npbins; /* top bin never filled, DG:799 */
                int condmu = imu >= 1 && imu <= h->nmu;
                if (ok && condx && condy && condz && condp && condmu) {
                    size_t lin = (size_t)(imu - 1) +
                                 (size_t)h->nmu * ((size_t)(ip - 1) +
                                 (size_t)h->npbins * ((size_t)(ix - 1) +
                                 (size_t)nrx[k] * ((size_t)(iy - 1) + (size_t)nry[k] * (size_t)(iz - 1))));
                    flocal[k][lin] += weight;
                }
            }
        }
        q3 += ptl->weight;
        if (ptl->dt < pdt_min) pdt_min = ptl->dt;
        if (ptl->dt > pdt_max) pdt_max = ptl->dt;
        q6 += ptl->dt;
        if (ptl->p > pmx) pmx = ptl->p;
    }
    if (quick) {
        quick[0] = (double)S->nptl_current;
        quick[1] = (double)S->nptl_split;
        quick[2] = q3;
        quick[3] = S->leak;
        quick[4] = S->leak_negp;
        quick[5] = q6;
        quick[6] = pdt_min;
        quick[7] = pdt_max;
    }
    if (pmax_out) *pmax_out = pmx;
}

/* global part of calc_escaped_distributions, DG:913-1000: fescaped(nmu,npp,2*ndim),
 * face index = -count_flag */
void orc_escaped_diagnostics(const orc_sim* S, double* fescaped)
{
    const gpat_params* P = &S->P;
    const int nmu_g = P->nmu_global, npp = P->npp_global, nface = 2 * P->ndim;
    double pmin_log = log10(P->pmin), pmax_log = log10(P->pmax);
    double dp_log = (pmax_log - pmin_log) / npp;
    double dmu = (double)(2.0f / (float)nmu_g);
    memset(fescaped, 0, sizeof(double) * (size_t)nmu_g * npp * nface);
    int64_t n_esc = S->nptl_escaped < S->nptl_escaped_max ? S->nptl_escaped : S->nptl_escaped_max;
    for (int64_t n = 0; n < n_esc; ++n) {
        const gpat_particle* ptl = &S->escaped[n];
        double p = ptl->p, mu = ptl->mu;
        int face = -ptl->count_flag;
        if (face < 1 || face > nface) continue;
        if (p > P->pmin && p <= P->pmax && mu >= -1.0 && mu <= 1.0) {
            int ok = 1;
            int64_t ip = ifloor((log10(p) - pmin_log) / dp_log, &ok) + 1;
            int64_t imu = ifloor((mu + 1.0) / dmu, &ok) + 1;
            int64_t lin = (imu - 1) + (ip - 1) * (int64_t)nmu_g;
            if (ok && ip >= 1 && imu >= 1 && lin >= 0 && lin < (int64_t)nmu_g * npp)
                fescaped[lin + (size_t)(face - 1) * nmu_g * npp] += ptl->weight;
        }
    }
}

/* local part of calc_escaped_distributions, DG:956-1170 (+ init_local_escaped_distributions, DG:358-405):
 * fx[k] = fescaped{k+1}_x(nmu, npbins, nry, nrz, 2), fy[k] = ..._y(nmu, npbins, nrx, nrz, 2) (ndim > 1),
 * fz[k] = ..._z(nmu, npbins, nrx, nry, 2) (ndim > 2); last index 1 = low face, 2 = high face.
 * The local sets add the (spherical-corrected) weight, DG:937-943: the particle's own weight here. */
void orc_escaped_local_diagnostics(const orc_sim* S, double* const fx[4], double* const fy[4], double* const fz[4])
{
    const gpat_params* P = &S->P;
    int64_t n_esc = S->nptl_escaped < S->nptl_escaped_max ? S->nptl_escaped : S->nptl_escaped_max;
    for (int k = 0; k < 4; ++k) {
        const gpat_hist_spec* h = &P->local[k];
        if (!h->enabled) continue;
        const int nrx = (P->nx + h->rx - 1) / h->rx, nry = (P->ny + h->ry - 1) / h->ry, nrz = (P->nz + h->rz - 1) / h->rz;
        const double dxd = P->lx / nrx, dyd = P->ly / nry, dzd = P->lz / nrz;
        const double pminl = log10(h->pmin), dpl = (log10(h->pmax) - pminl) / h->npbins;
        const double dmul = (double)(2.0f / (float)h->nmu);
        const size_t nb = (size_t)h->nmu * h->npbins;
        if (fx && fx[k]) memset(fx[k], 0, sizeof(double) * nb * nry * nrz * 2);
        if (fy && fy[k] && P->ndim > 1) memset(fy[k], 0, sizeof(double) * nb * nrx * nrz * 2);
        if (fz && fz[k] && P->ndim > 2) memset(fz[k], 0, sizeof(double) * nb * nrx * nry * 2);
        for (int64_t n = 0; n < n_esc; ++n) {
            const gpat_particle* ptl = &S->escaped[n];
            int okx = 1, oky = 1, okz = 1, okp = 1, okm = 1;
            int64_t ix = ifloor((ptl->x - P->xmin) / dxd, &okx) + 1;
            int64_t iy = ifloor((ptl->y - P->ymin) / dyd, &oky) + 1;
            int64_t iz = ifloor((ptl->z - P->zmin) / dzd, &okz) + 1;
            int64_t ip = ifloor((log10(ptl->p) - pminl) / dpl, &okp) + 1;
            int64_t imu = ifloor((ptl->mu + 1.0) / dmul, &okm) + 1;
            int condx = okx && ix >= 1 && ix <= nrx;
            int condy = oky && iy >= 1 && iy <= nry;
            int condz = okz && iz >= 1 && iz <= nrz;
            int condp = okp && ip > 0 && ip < h->npbins;
            int condmu = okm && imu >= 1 && imu <= h->nmu;
            if (!(condp && condmu)) continue;
            const size_t b = (size_t)(imu - 1) + (size_t)h->nmu * (size_t)(ip - 1);
            const int face = -ptl->count_flag; /* 1 lx, 2 hx, 3 ly, 4 hy, 5 lz, 6 hz */
            const size_t side = (size_t)((face - 1) & 1);
            if ((face == 1 || face == 2) && condy && condz && fx && fx[k])
                fx[k][b + nb * ((size_t)(iy - 1) + (size_t)nry * ((size_t)(iz - 1) + (size_t)nrz * side))] += ptl->weight;
            else if ((face == 3 || face == 4) && condx && condz && fy && fy[k] && P->ndim > 1)
                fy[k][b + nb * ((size_t)(ix - 1) + (size_t)nrx * ((size_t)(iz - 1) + (size_t)nrz * side))] += ptl->weight;
            else if ((face == 5 || face == 6) && condx && condy && fz && fz[k] && P->ndim > 2)
                fz[k][b + nb * ((size_t)(ix - 1) + (size_t)nrx * ((size_t)(iy - 1) + (size_t)nry * side))] += ptl->weight;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* accessors                                                                  */
/* ------------------------------------------------------------------------ */
int64_t orc_get_particles(const orc_sim* S, gpat_particle* out, int64_t nmax)
{
    int64_t n = S->nptl_current < nmax ? S->nptl_current : nmax;
    memcpy(out, S->ptls, sizeof(gpat_particle) * (size_t)n);
    return S->nptl_current;
}

void orc_set_particles(orc_sim* S, const gpat_particle* in, int64_t n)
{
    if (n > S->nptl_max) n = S->nptl_max;
    memcpy(S->ptls, in, sizeof(gpat_particle) * (size_t)n);
    S->nptl_current = n;
}

int64_t orc_get_escaped(const orc_sim* S, gpat_particle* out, int64_t nmax)
{
    int64_t n = S->nptl_escaped < nmax ? S->nptl_escaped : nmax;
    if (n > S->nptl_escaped_max) n = S->nptl_escaped_max;
    memcpy(out, S->escaped, sizeof(gpat_particle) * (size_t)n);
    return S->nptl_escaped;
}

void orc_reset_escaped(orc_sim* S) { S->nptl_escaped = 0; }

/* test hook: replace the escaped list (lets a test re-bin another implementation's escapees) */
void orc_set_escaped(orc_sim* S, const gpat_particle* in, int64_t n)
{
    if (n > S->nptl_escaped_max) n = S->nptl_escaped_max;
    memcpy(S->escaped, in, sizeof(gpat_particle) * (size_t)n);
    S->nptl_escaped = n;
}

void orc_get_counters(const orc_sim* S, gpat_counters* c)
{
    c->nptl_current = S->nptl_current; c->nptl_split = S->nptl_split;
    c->nptl_escaped = S->nptl_escaped; c->nptl_max = S->nptl_max; c->tag_max = S->tag_max;
    c->leak = S->leak; c->leak_negp = S->leak_negp;
}

void orc_set_counters(orc_sim* S, const gpat_counters* c)
{
    S->nptl_current = c->nptl_current; S->nptl_split = c->nptl_split;
    S->nptl_escaped = c->nptl_escaped; S->tag_max = c->tag_max;
    S->leak = c->leak; S->leak_negp = c->leak_negp;
}

uint64_t orc_total_steps(const orc_sim* S) { return S->steps; }

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's reference arm: torchrun exports OMP_NUM_THREADS=1 to every rank, which would time the
 * CPU path on one core; the arm asks for the host's cores explicitly. */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
