/*
 * smk_oracle.c -- CPU oracle for the SimpleMOC-kernel segment-attenuation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or as the timed CPU
 * baseline.  The product (simplemoc-kernel_b200/) never links or calls it.
 *
 * What this is: a plain-C restatement, in the reference's evaluation order, of
 *   - attenuate_segment            /root/reference/src/cpu/kernel.c:75-333
 *   - interpolateTable             /root/reference/src/cpu/kernel.c:337-361
 *   - buildExponentialTable        /root/reference/src/cpu/init.c:81-117
 *   - the run_kernel segment loop  /root/reference/src/cpu/kernel.c:3-73
 * driven by the deterministic counter-based stream that replaces the reference's
 * time-seeded rand_r() draws (kernel.c:15,29-30,47,50) and rand() fills
 * (init.c:64-75).  The stream is specified in DESIGN.md section 3 and restated
 * here independently of the CUDA implementation.
 *
 * Parity pin: the reference has no golden vectors, so this
 * restatement is pinned against the reference's own object code: oracle/Makefile
 * (target `ref`) compiles the unmodified /root/reference/src/cpu/{kernel.c,init.c}
 * into oracle/_ref/ and tests/test_oracle.py checks that this file and the
 * reference's attenuate_segment agree BIT FOR BIT on the same stream (exp and
 * table builds).  Fixtures produced by that reference build are committed under
 * tests/golden/ so the pin also holds where /root/reference is absent.
 * The Philox generator is pinned against the Random123 known-answer vectors.
 *
 * Canonical build: gcc -O2 -ffp-contract=off -fopenmp (no fast-math), so that
 * every expression below is evaluated exactly as written, in IEEE binary32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.x philox.h).              */
/* ------------------------------------------------------------------------- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void smk_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                              uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream domains: the 4th counter word separates the three uses of the stream. */
#define DOMAIN_SEGMENT 0x5345474Du /* 'SEGM' : (QSR_id, FAI_id) of a segment   */
#define DOMAIN_PSI     0x50534930u /* 'PSI0' : incoming angular flux of a track */
#define DOMAIN_FILL    0x46494C4Cu /* 'FILL' : source / flux / sigT slabs       */

/* Mirrors `(float) rand() / RAND_MAX` (init.c:68-69,75; kernel.c:30): rand() is a
 * 31-bit integer and (float)RAND_MAX == 2^31, so the reference's value is
 * (float)r * 2^-31 with r in [0, 2^31).  We take the top 31 bits of a word.  */
static inline float u01(uint32_t w)
{
    return (float)(int32_t)(w >> 1) * 0x1.0p-31f;
}

static inline void stream_words(uint64_t seed, uint64_t index, uint32_t sub,
                                uint32_t domain, uint32_t w[4])
{
    uint32_t ctr[4] = { (uint32_t)index, (uint32_t)(index >> 32), sub, domain };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    smk_oracle_philox4x32_10(ctr, key, w);
}

/* (QSR_id, FAI_id) of global segment `seg`; `%` as in kernel.c:47,50 on a
 * 31-bit draw (rand_r's range).  Words 2,3 carry the per-segment geometry
 * (segment_geometry below; kernel.c:95-104). */
void smk_oracle_segment_ids(uint64_t seed, int64_t seg_begin, int64_t count,
                            int regions, int fai, int32_t *qsr_out,
                            int32_t *fai_out)
{
    for (int64_t i = 0; i < count; i++) {
        uint32_t w[4];
        stream_words(seed, (uint64_t)(seg_begin + i), 0u, DOMAIN_SEGMENT, w);
        qsr_out[i] = (int32_t)((w[0] >> 1) % (uint32_t)regions);
        fai_out[i] = (int32_t)((w[1] >> 1) % (uint32_t)fai);
    }
}

/* psi0 of track `track`, groups [0, groups). */
void smk_oracle_track_psi0(uint64_t seed, int64_t track, int groups, float *psi)
{
    for (int q = 0; q * 4 < groups; q++) {
        uint32_t w[4];
        stream_words(seed, (uint64_t)track, (uint32_t)q, DOMAIN_PSI, w);
        for (int j = 0; j < 4 && q * 4 + j < groups; j++)
            psi[q * 4 + j] = u01(w[j]);
    }
}

/* Deterministic replacement of init.c:64-75.  Arrays are the reference's
 * unpadded layouts: fine_source[R][F][G], fine_flux[R][F][G], sigT[R][G].
 * Element e of array a comes from word (e & 3) of counter (e >> 2, a).
 * sigt_floor = 0 reproduces the reference's U[0,1) cross sections; a positive
 * floor gives the well-conditioned diagnostic data set sigT = floor + u*(1-floor). */
static void fill_array(float *dst, int64_t n, uint32_t array_id, uint64_t seed,
                       float floor_)
{
    const float span = 1.0f - floor_;
    for (int64_t q = 0; q * 4 < n; q++) {
        uint32_t w[4];
        stream_words(seed, (uint64_t)q, array_id, DOMAIN_FILL, w);
        for (int j = 0; j < 4 && q * 4 + j < n; j++) {
            float u = u01(w[j]);
            dst[q * 4 + j] = (floor_ > 0.0f) ? floor_ + u * span : u;
        }
    }
}

void smk_oracle_fill(float *fine_source, float *fine_flux, float *sigT,
                     int regions, int fai, int groups, uint64_t seed,
                     float sigt_floor)
{
    int64_t n = (int64_t)regions * fai * groups;
    if (fine_source) fill_array(fine_source, n, 0u, seed, 0.0f);
    if (fine_flux)   fill_array(fine_flux, n, 1u, seed, 0.0f);
    if (sigT)        fill_array(sigT, (int64_t)regions * groups, 2u, seed, sigt_floor);
}

/* ------------------------------------------------------------------------- */
/* Segment geometry.  kernel.c:95-104: "Some placeholder constants - In the    */
/* full app some of these are calculated based off position in geometry."      */
/* The six placeholders become parameters; with SMK_ORACLE_GEOM they vary per  */
/* segment, drawn from words 2,3 of the segment's stream block (DESIGN.md      */
/* section 3b): four 16-bit fields u -> factor f(u) = 1 + spread*(u*2^-15 - 1) */
/*   ds = ds0 f(w2>>16)   zin = zin0 f(w2&0xFFFF)   mu = mu0 f(w3>>16)         */
/*   mu2 = mu2_0 f(w3>>16)^2   weight = weight0 f(w3&0xFFFF)   dz = dz0        */
/* (dz is a property of the axial mesh, the same for every segment).  Every    */
/* operation is a single IEEE binary32 operation in the order written, so the  */
/* GPU reproduces the values bit for bit; spread = 0 gives f = 1 exactly, i.e.  */
/* the base values, i.e. the reference when the base is kernel.c:99-104.        */
/* ------------------------------------------------------------------------- */
typedef struct {
    float dz, zin, weight, mu, mu2, ds;
} seg_geometry;

static const seg_geometry k_reference_geometry = { /* kernel.c:99-104 */
    0.1f, 0.3f, 0.5f, 0.9f, 0.3f, 0.7f
};

static inline float geom_factor(uint32_t u16, float spread)
{
    const float t = (float)u16 * 0x1.0p-15f; /* exact */
    const float c = t - 1.0f;                /* exact: <= 16 significant bits */
    const float sc = spread * c;
    return 1.0f + sc;
}

/* base[7] = { dz, zin, weight, mu, mu2, ds, spread } */
static inline seg_geometry segment_geometry(const float *base, uint32_t w2, uint32_t w3)
{
    seg_geometry g;
    const float spread = base[6];
    const float f_ds = geom_factor(w2 >> 16, spread);
    const float f_zin = geom_factor(w2 & 0xFFFFu, spread);
    const float f_mu = geom_factor(w3 >> 16, spread);
    const float f_w = geom_factor(w3 & 0xFFFFu, spread);
    const float f_mu_sq = f_mu * f_mu;
    g.dz = base[0];
    g.zin = base[1] * f_zin;
    g.weight = base[2] * f_w;
    g.mu = base[3] * f_mu;
    g.mu2 = base[4] * f_mu_sq;
    g.ds = base[5] * f_ds;
    return g;
}

/* geometry of segments [seg_begin, seg_begin + count): out[i*6 + {0..5}] =
 * dz, zin, weight, mu, mu2, ds (for tests of the GPU's derivation) */
void smk_oracle_segment_geometry(uint64_t seed, int64_t seg_begin, int64_t count,
                                 const float *base7, float *out)
{
    for (int64_t i = 0; i < count; i++) {
        uint32_t w[4];
        stream_words(seed, (uint64_t)(seg_begin + i), 0u, DOMAIN_SEGMENT, w);
        const seg_geometry g = segment_geometry(base7, w[2], w[3]);
        out[i * 6 + 0] = g.dz; out[i * 6 + 1] = g.zin; out[i * 6 + 2] = g.weight;
        out[i * 6 + 3] = g.mu; out[i * 6 + 4] = g.mu2; out[i * 6 + 5] = g.ds;
    }
}

/* ------------------------------------------------------------------------- */
/* Exponential table: init.c:81-117 and kernel.c:337-361.                     */
/* ------------------------------------------------------------------------- */
typedef struct {
    float *values; /* 2*N floats: {slope, intercept} per interval */
    float dx;
    float maxVal;
    int N;
} oracle_table;

/* init.c:81-117 called as buildExponentialTable(0.01, 10.0, I) (main.c:33).
 * Returns N; writes 2*N floats into values (capacity checked by caller: 706). */
int smk_oracle_build_table(float precision, float maxVal, float *values,
                           int capacity, float *dx_out, float *maxval_out)
{
    int N = (int)(maxVal * sqrt(1.0 / (8.0 * precision * 0.01))); /* init.c:88 */
    float dx = maxVal / (float)N;                                 /* init.c:91 */
    if (2 * N > capacity) return -1;
    for (int n = 0; n < N; n++) {
        float exponential = exp(-n * dx);                         /* init.c:105 */
        values[2 * n] = -exponential;                             /* init.c:106 */
        values[2 * n + 1] = 1 + (n * dx - 1) * exponential;       /* init.c:107 */
    }
    *dx_out = dx;
    *maxval_out = maxVal - dx;                                    /* init.c:113 */
    return N;
}

static inline float table_lookup(const oracle_table *t, float x)
{
    if (x > t->maxVal)                                            /* kernel.c:340 */
        return 1.0f;
    int interval = (int)(x / t->dx + 0.5f * t->dx);               /* kernel.c:344 */
    interval = interval * 2;
    float slope = t->values[interval];
    float intercept = t->values[interval + 1];
    return slope * x + intercept;                                 /* kernel.c:358 */
}

float smk_oracle_table_lookup(const float *values, float dx, float maxVal, float x)
{
    oracle_table t = { (float *)values, dx, maxVal, 0 };
    return table_lookup(&t, x);
}

/* libm's expf, exposed so tests can compare the GPU's glibc-faithful expf. */
float smk_oracle_expf(float x) { return expf(x); }

/* out[i] = expf(-tau[i]) with libm, for exhaustive sweeps */
void smk_oracle_expf_neg_array(const float *tau, float *out, int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
        out[i] = expf(-tau[i]);
}

/* ------------------------------------------------------------------------- */
/* attenuate_segment, kernel.c:75-333, one group at a time.                    */
/* The reference stages each sub-expression through a scratch vector and runs  */
/* 15 loops over g; per element the operations and their order are those below */
/* (no cross-group arithmetic exists), so a single loop is bit-identical.       */
/* ------------------------------------------------------------------------- */
static void attenuate_one_segment(int groups, int fai_count, int FAI_id,
                                  const float *src_region, /* [F][G] */
                                  const float *sigT_region, /* [G]    */
                                  float *psi, float *tally,
                                  const oracle_table *table,
                                  const seg_geometry *geom)
{
    const float dz = geom->dz;   /* kernel.c:99-104, as parameters */
    const float zin = geom->zin;
    const float weight = geom->weight;
    const float mu = geom->mu;
    const float mu2 = geom->mu2;
    const float ds = geom->ds;

    const float *f1 = src_region + (int64_t)(FAI_id - 1) * groups;
    const float *f2 = src_region + (int64_t)FAI_id * groups;
    const float *f3 = src_region + (int64_t)(FAI_id + 1) * groups;

    for (int g = 0; g < groups; g++) {
        float q0, q1, q2;
        if (FAI_id == 0) {                           /* kernel.c:111-135 */
            const float y2 = f2[g];
            const float y3 = f3[g];
            const float c0 = y2;
            const float c1 = (y3 - y2) / dz;
            q0 = c0 + c1 * zin;
            q1 = c1;
            q2 = 0;
        } else if (FAI_id == fai_count - 1) {        /* kernel.c:137-161 */
            const float y1 = f1[g];
            const float y2 = f2[g];
            const float c0 = y2;
            const float c1 = (y2 - y1) / dz;
            q0 = c0 + c1 * zin;
            q1 = c1;
            q2 = 0;
        } else {                                     /* kernel.c:163-191 */
            const float y1 = f1[g];
            const float y2 = f2[g];
            const float y3 = f3[g];
            const float c0 = y2;
            const float c1 = (y1 - y3) / (2.f * dz);
            const float c2 = (y1 - 2.f * y2 + y3) / (2.f * dz * dz);
            q0 = c0 + c1 * zin + c2 * zin * zin;
            q1 = c1 + 2.f * c2 * zin;
            q2 = c2;
        }

        const float sigT = sigT_region[g];           /* kernel.c:200-208 */
        const float tau = sigT * ds;
        const float sigT2 = sigT * sigT;

        float expVal;                                /* kernel.c:216-223 */
        if (table)
            expVal = table_lookup(table, tau);
        else
            expVal = 1.f - expf(-tau);

        const float reuse = tau * (tau - 2.f) + 2.f * expVal /* kernel.c:233-237 */
            / (sigT * sigT2);

        const float flux_integral =                  /* kernel.c:245-252 */
            (q0 * tau + (sigT * psi[g] - q0) * expVal) / sigT2
            + q1 * mu * reuse
            + q2 * mu2 * (tau * (tau * (tau - 3.f) + 6.f) - 6.f * expVal)
                  / (3.f * sigT2 * sigT2);

        tally[g] = weight * flux_integral;           /* kernel.c:259-263 */

        const float t1 = q0 * expVal / sigT;                     /* kernel.c:291 */
        const float t2 = q1 * mu * (tau - expVal) / sigT2;       /* kernel.c:301 */
        const float t3 = q2 * mu2 * reuse;                       /* kernel.c:311 */
        const float t4 = psi[g] * (1.f - expVal);                /* kernel.c:321 */
        psi[g] = t1 + t2 + t3 + t4;                              /* kernel.c:331 */
    }
}

/* One segment against caller-supplied rows; used by the known-answer tests. */
void smk_oracle_attenuate_segment(int groups, int fai_count, int FAI_id,
                                  const float *src_region,
                                  const float *sigT_region, float *psi,
                                  float *tally_out, int use_table)
{
    float tv[706];
    oracle_table t = { tv, 0.f, 0.f, 0 };
    if (use_table)
        t.N = smk_oracle_build_table(0.01f, 10.0f, tv, 706, &t.dx, &t.maxVal);
    attenuate_one_segment(groups, fai_count, FAI_id, src_region, sigT_region,
                          psi, tally_out, use_table ? &t : NULL, &k_reference_geometry);
}

/* Same with caller-supplied geometry geom6 = { dz, zin, weight, mu, mu2, ds }. */
void smk_oracle_attenuate_segment_geom(int groups, int fai_count, int FAI_id,
                                       const float *src_region,
                                       const float *sigT_region, float *psi,
                                       float *tally_out, int use_table,
                                       const float *geom6)
{
    float tv[706];
    oracle_table t = { tv, 0.f, 0.f, 0 };
    if (use_table)
        t.N = smk_oracle_build_table(0.01f, 10.0f, tv, 706, &t.dx, &t.maxVal);
    const seg_geometry g = { geom6[0], geom6[1], geom6[2], geom6[3], geom6[4], geom6[5] };
    attenuate_one_segment(groups, fai_count, FAI_id, src_region, sigT_region,
                          psi, tally_out, use_table ? &t : NULL, &g);
}

/* ------------------------------------------------------------------------- */
/* Replay driver: the run_kernel loop (kernel.c:3-73) over tracks.             */
/* A track is seg_per_track consecutive segments that share one carried psi    */
/* (the reference carries a thread's state_flux across its dynamic chunks of   */
/* 100, kernel.c:30,43,331; the CUDA reference batches -p segments per block). */
/* ------------------------------------------------------------------------- */
#define SMK_ORACLE_TABLE   1u  /* use the interpolation table (TABLE build)   */
#define SMK_ORACLE_F64ACC  2u  /* diagnostic: accumulate tallies in double    */
#define SMK_ORACLE_GEOM    4u  /* per-segment geometry from stream words 2,3  */

/* Returns 0 on success.  fine_flux is updated in place (kernel.c:274-277).
 * psi_final, if non-NULL, receives the outgoing psi of each track in
 * [track_begin, track_end): psi_final[(t - track_begin)*groups + g].
 * id_checksum, if non-NULL, receives sum over segments of
 * (QSR_id*fai + FAI_id + 1) * ((seg & 0xFFFF) + 1) mod 2^64 (indexing fingerprint). */
/* geom7 = { dz, zin, weight, mu, mu2, ds, spread } or NULL for kernel.c:99-104.
 * Without SMK_ORACLE_GEOM every segment uses the base values (spread ignored). */
int smk_oracle_run_geom(int regions, int fai, int groups, int64_t segments,
                        int seg_per_track, uint64_t seed,
                        const float *fine_source, float *fine_flux, const float *sigT,
                        int64_t track_begin, int64_t track_end, float *psi_final,
                        uint64_t *id_checksum, int nthreads, unsigned flags,
                        const float *geom7)
{
    if (regions < 1 || fai < 2 || groups < 1 || seg_per_track < 1 || segments < 0)
        return 1;
    const int64_t n_tracks = (segments + seg_per_track - 1) / seg_per_track;
    if (track_begin < 0 || track_end > n_tracks || track_begin > track_end)
        return 2;

    float tv[706];
    oracle_table tab = { tv, 0.f, 0.f, 0 };
    const int use_table = (flags & SMK_ORACLE_TABLE) != 0;
    if (use_table)
        tab.N = smk_oracle_build_table(0.01f, 10.0f, tv, 706, &tab.dx, &tab.maxVal);

    float base7[7] = { k_reference_geometry.dz, k_reference_geometry.zin,
                       k_reference_geometry.weight, k_reference_geometry.mu,
                       k_reference_geometry.mu2, k_reference_geometry.ds, 0.0f };
    if (geom7) memcpy(base7, geom7, sizeof base7);
    const int per_segment = (flags & SMK_ORACLE_GEOM) != 0;
    const seg_geometry fixed = { base7[0], base7[1], base7[2], base7[3], base7[4], base7[5] };

    const int64_t rows = (int64_t)regions * fai;
    double *acc64 = NULL;
    if (flags & SMK_ORACLE_F64ACC) {
        acc64 = (double *)calloc((size_t)(rows * groups), sizeof(double));
        if (!acc64) return 3;
    }
    uint64_t checksum = 0;

#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
    omp_lock_t *locks = NULL;
    if (nthreads > 1) {
        locks = (omp_lock_t *)malloc((size_t)rows * sizeof(omp_lock_t));
        for (int64_t i = 0; i < rows; i++) omp_init_lock(&locks[i]);
    }
#else
    nthreads = 1;
#endif

#pragma omp parallel num_threads(nthreads) reduction(+ : checksum)
    {
        float *psi = (float *)malloc((size_t)groups * sizeof(float));
        float *tally = (float *)malloc((size_t)groups * sizeof(float));

#pragma omp for schedule(dynamic, 16)
        for (int64_t t = track_begin; t < track_end; t++) {
            smk_oracle_track_psi0(seed, t, groups, psi);
            const int64_t s0 = t * seg_per_track;
            int64_t s1 = s0 + seg_per_track;
            if (s1 > segments) s1 = segments;
            for (int64_t s = s0; s < s1; s++) {
                uint32_t w[4];
                stream_words(seed, (uint64_t)s, 0u, DOMAIN_SEGMENT, w);
                const int32_t QSR_id = (int32_t)((w[0] >> 1) % (uint32_t)regions); /* kernel.c:47 */
                const int32_t FAI_id = (int32_t)((w[1] >> 1) % (uint32_t)fai);     /* kernel.c:50 */
                checksum += ((uint64_t)QSR_id * (uint64_t)fai + (uint64_t)FAI_id + 1u)
                            * ((uint64_t)(s & 0xFFFF) + 1u);
                const seg_geometry g = per_segment ? segment_geometry(base7, w[2], w[3]) : fixed;

                attenuate_one_segment(groups, fai, FAI_id,
                                      fine_source + (int64_t)QSR_id * fai * groups,
                                      sigT + (int64_t)QSR_id * groups, psi, tally,
                                      use_table ? &tab : NULL, &g);

                const int64_t row = (int64_t)QSR_id * fai + FAI_id;
                if (acc64) {
                    for (int g = 0; g < groups; g++) {
#pragma omp atomic
                        acc64[row * groups + g] += (double)tally[g];
                    }
                } else {
                    float *FSR_flux = fine_flux + row * groups;
#ifdef _OPENMP
                    if (locks) omp_set_lock(&locks[row]);    /* kernel.c:265 */
#endif
                    for (int g = 0; g < groups; g++)
                        FSR_flux[g] += tally[g];             /* kernel.c:276 */
#ifdef _OPENMP
                    if (locks) omp_unset_lock(&locks[row]);  /* kernel.c:279 */
#endif
                }
            }
            if (psi_final)
                memcpy(psi_final + (t - track_begin) * groups, psi,
                       (size_t)groups * sizeof(float));
        }
        free(psi);
        free(tally);
    }

    if (acc64) {
        for (int64_t i = 0; i < rows * groups; i++)
            fine_flux[i] = (float)((double)fine_flux[i] + acc64[i]);
        free(acc64);
    }
#ifdef _OPENMP
    if (locks) {
        for (int64_t i = 0; i < rows; i++) omp_destroy_lock(&locks[i]);
        free(locks);
    }
#endif
    if (id_checksum) *id_checksum = checksum;
    return 0;
}

int smk_oracle_run(int regions, int fai, int groups, int64_t segments,
                   int seg_per_track, uint64_t seed,
                   const float *fine_source, float *fine_flux, const float *sigT,
                   int64_t track_begin, int64_t track_end, float *psi_final,
                   uint64_t *id_checksum, int nthreads, unsigned flags)
{
    return smk_oracle_run_geom(regions, fai, groups, segments, seg_per_track, seed,
                               fine_source, fine_flux, sigT, track_begin, track_end,
                               psi_final, id_checksum, nthreads, flags, NULL);
}

int smk_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
