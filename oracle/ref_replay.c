/*
 * ref_replay.c -- replay driver around the UNMODIFIED reference CPU sources.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/smk_oracle.c header).
 *
 * oracle/Makefile (target `ref`) compiles this file together with
 *   /root/reference/src/cpu/kernel.c and /root/reference/src/cpu/init.c
 * (from where they lie; nothing is copied into this repo) into oracle/_ref/*.so.
 * It feeds the reference's own attenuate_segment (kernel.c:75-333, an external
 * symbol, SimpleMOC-kernel_header.h:82-84) with the deterministic stream of
 * oracle/smk_oracle.c, so that the restatement can be checked bit for bit
 * against the reference's object code, and it exposes the reference's own
 * run_kernel (kernel.c:3-73) for timing on the host cores.
 *
 * The reference header is included from /root/reference at build time (-I).
 */
#include "SimpleMOC-kernel_header.h"
#include <stdint.h>

/* stream functions restated in smk_oracle.c (linked into the same .so) */
void smk_oracle_segment_ids(uint64_t seed, int64_t seg_begin, int64_t count,
                            int regions, int fai, int32_t *qsr_out, int32_t *fai_out);
void smk_oracle_track_psi0(uint64_t seed, int64_t track, int groups, float *psi);

static Input *make_input(int regions, int fai, int groups, long segments, int nthreads)
{
    Input *I = set_default_input();              /* init.c:4-24 */
    I->fine_axial_intervals = fai;
    I->source_3D_regions = regions;              /* main.c:18-19 computes this */
    I->egroups = groups;
    I->segments = segments;
    I->nthreads = nthreads;
    return I;
}

/* Which build is this?  bit0: TABLE, bit1: OPENMP */
int ref_build_flags(void)
{
    int f = 0;
#ifdef TABLE
    f |= 1;
#endif
#ifdef OPENMP
    f |= 2;
#endif
    return f;
}

/*
 * Replay tracks [track_begin, track_end) through the reference's
 * attenuate_segment, single-threaded and in track order (deterministic).
 * Slabs use the reference layout (init.c:35-54): the three arrays are
 * contiguous from S[0].fine_source / fine_flux / sigT, so we allocate with the
 * reference's initialize_sources and copy the caller's data over its rand() fill.
 */
int ref_replay_run(int regions, int fai, int groups, int64_t segments,
                   int seg_per_track, uint64_t seed,
                   const float *fine_source, float *fine_flux, const float *sigT,
                   int64_t track_begin, int64_t track_end, float *psi_final)
{
    Input *I = make_input(regions, fai, groups, (long)segments, 1);
    Source *S = initialize_sources(I);           /* init.c:26-78 */
    const size_t n = (size_t)regions * fai * groups;
    memcpy(S[0].fine_source, fine_source, n * sizeof(float));
    memcpy(S[0].fine_flux, fine_flux, n * sizeof(float));
    memcpy(S[0].sigT, sigT, (size_t)regions * groups * sizeof(float));

    Table *table = NULL;
#ifdef TABLE
    table = buildExponentialTable(0.01, 10.0, I); /* main.c:32-34 */
#endif
    SIMD_Vectors simd_vecs = allocate_simd_vectors(I);
    float *state_flux = (float *)malloc((size_t)groups * sizeof(float));

    for (int64_t t = track_begin; t < track_end; t++) {
        smk_oracle_track_psi0(seed, t, groups, state_flux);
        int64_t s0 = t * seg_per_track, s1 = s0 + seg_per_track;
        if (s1 > segments) s1 = segments;
        for (int64_t s = s0; s < s1; s++) {
            int32_t QSR_id, FAI_id;
            smk_oracle_segment_ids(seed, s, 1, regions, fai, &QSR_id, &FAI_id);
            attenuate_segment(I, S, QSR_id, FAI_id, state_flux, &simd_vecs, table);
        }
        if (psi_final)
            memcpy(psi_final + (t - track_begin) * groups, state_flux,
                   (size_t)groups * sizeof(float));
    }
    memcpy(fine_flux, S[0].fine_flux, n * sizeof(float));
    /* the reference never frees its slabs; we do, this is a library */
    free(state_flux);
    free(simd_vecs.q0);
    free(S[0].fine_source); free(S[0].fine_flux); free(S[0].sigT);
#ifdef OPENMP
    free(S[0].locks);
#endif
    free(S); free(I);
    return 0;
}

/* The reference's table, for comparing table constants (init.c:81-117). */
int ref_table(float *values706, float *dx, float *maxVal)
{
    Input *I = set_default_input();
    Table *t = buildExponentialTable(0.01, 10.0, I);
    memcpy(values706, t->values, (size_t)(2 * t->N) * sizeof(float));
    *dx = t->dx; *maxVal = t->maxVal;
    int N = t->N;
    free(t->values); free(t); free(I);
    return N;
}

/*
 * Time the reference's own run_kernel (its own rand_r stream, its own OpenMP
 * schedule and locks) exactly as main.c:44-47 does: wall clock around the call.
 * Returns seconds; data set-up (initialize_sources) is outside, as in main.c.
 */
double ref_time_run_kernel(int regions_2d, int groups, long segments, int nthreads)
{
    Input *I = set_default_input();
    I->source_2D_regions = regions_2d;
    I->egroups = groups;
    I->segments = segments;
    I->source_3D_regions = (int)ceil((double)I->source_2D_regions *
                                     I->coarse_axial_intervals / I->decomp_assemblies_ax);
#ifdef OPENMP
    if (nthreads < 1) nthreads = omp_get_num_procs();   /* io.c:103 */
    I->nthreads = nthreads;
    omp_set_num_threads(I->nthreads);                   /* main.c:24 */
#else
    I->nthreads = 1;
#endif
    srand(12345u);
    Source *S = initialize_sources(I);
    Table *table = NULL;
#ifdef TABLE
    table = buildExponentialTable(0.01, 10.0, I);
#endif
    double start = get_time();
    run_kernel(I, S, table);
    double stop = get_time();
    free(S[0].fine_source); free(S[0].fine_flux); free(S[0].sigT);
#ifdef OPENMP
    free(S[0].locks);
#endif
    free(S); free(I);
    return stop - start;
}

int ref_num_procs(void)
{
#ifdef OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
