/*
 * smk_main.c -- plain-C host driver of the B200-native SimpleMOC-kernel path.
 *
 * Plays the role of /root/reference/src/cpu/main.c + io.c (and of their CUDA twins
 * /root/reference/src/cuda/main.cu + io.cu): same Input fields, same -t/-s/-e/-p/-d
 * options, same banner / INPUT SUMMARY / RESULTS SUMMARY lines, but the sweep is
 * done by libsmk.so through the C ABI of include/smk.h.  Differences, all additive:
 *   - -s is parsed as a long (the reference's atoi overflows at 1e10, io.c:126)
 *   - long options configure the changed subsystems (seed, exp/math mode, 2D regions,
 *     number of GPUs) without colliding with the reference's letters
 *   - a VERIFICATION block prints the indexing fingerprint and flux norms that the
 *     CPU replay oracle reproduces (the reference prints no result at all,
 *     main.c:49-59)
 * Nothing here touches the GPU directly; nothing here calls the oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smk.h"

typedef struct {
    int source_2D_regions;
    int source_3D_regions;
    int coarse_axial_intervals;
    int fine_axial_intervals;
    int decomp_assemblies_ax;
    long segments;
    int egroups;
    int nthreads;        /* kept for CLI compatibility: CPU threads of an external replay */
    int streams;         /* reported: one independent counter stream per track           */
    int seg_per_thread;  /* -p */
    size_t nbytes;
    /* changed subsystems */
    unsigned long long seed;
    int exp_mode, math_mode, device, gpus, allreduce;
    int host_fill;
    int tally_f64;
    int segment_geometry;      /* kernel.c:99-104 vary per segment (SMK_FLAG_SEGMENT_GEOMETRY) */
    int fit_per_sweep;         /* source fit once per (region, interval, group) per sweep (SMK_FLAG_FIT_PER_SWEEP) */
    float geometry_spread;
    float sigt_floor;
    const char *dump_flux;
    const char *verify_flux;   /* raw float32 flux of a CPU replay of the same stream */
    double tolerance;
} Input;

static void rule(void)
{
    for (int i = 0; i < 80; i++) putchar('=');
    putchar('\n');
}

static void centered(const char *s)
{
    int pad = (79 - (int)strlen(s)) / 2;
    printf("%*s%s\n", pad + 1, "", s);
}

/* 1234567 -> "1,234,567"; takes a long so that -s 10000000000 prints correctly */
static void grouped(long v)
{
    char digits[32], out[48];
    int n = snprintf(digits, sizeof digits, "%ld", v), o = 0;
    for (int i = 0; i < n; i++) {
        out[o++] = digits[i];
        int left = n - 1 - i;
        if (left > 0 && left % 3 == 0 && digits[i] != '-') out[o++] = ',';
    }
    out[o] = 0;
    puts(out);
}

static void banner(int version)
{
    static const char *art[] = {
        "   __           __        ___        __   __           ___  __        ___     ",
        "  /__` |  |\\/| |__) |    |__   |\\/| /  \\ /  ` __ |__/ |__  |__) |\\ | |__  |   ",
        "  .__/ |  |  | |    |___ |___  |  | \\__/ \\__,    |  \\ |___ |  \\ | \\| |___ |___",
    };
    char v[64];
    rule();
    for (int i = 0; i < 3; i++) puts(art[i]);
    putchar('\n');
    rule();
    putchar('\n');
    centered("Developed at");
    centered("The Massachusetts Institute of Technology");
    centered("and");
    centered("Argonne National Laboratory");
    putchar('\n');
    snprintf(v, sizeof v, "Version: %d", version);
    centered(v);
    putchar('\n');
    rule();
}

static void usage_and_exit(void)
{
    puts("Usage: ./SimpleMOC <options>");
    puts("Options include:");
    puts("  -t <threads>          Number of OpenMP threads to run");
    puts("  -s <segments>         Number of segments to process");
    puts("  -e <energy groups>    Number of energy groups");
    puts("  -p <segs per thread>  Number of segments per CUDA Block");
    puts("  -d <CUDA device ID>   CUDA GPU device ID number");
    puts("  --regions-2d <n>      2D source regions (default 5000)");
    puts("  --seed <n>            Seed of the counter-based segment stream (default 42)");
    puts("  --exp <mode>          poly | mufu | glibc | table (default poly)");
    puts("  --math <mode>         fast | strict (default fast)");
    puts("  --gpus <n>            Split the segments over n GPUs, all-reduce the tallies");
    puts("  --allreduce <impl>    peer (NVLink peer-memory kernel, default) | nccl");
    puts("  --host-fill           Fill the slabs on the host and upload them");
    puts("  --tally-f64           Diagnostic: accumulate the tallies in double precision");
    puts("  --segment-geometry    dz, zin, weight, mu, mu2, ds vary per segment (stream words 2,3)");
    puts("  --fit-per-sweep       Evaluate the axial source fit once per (region, interval, group) per sweep,");
    puts("                        not once per segment (same results; <= 64 groups; off by default)");
    puts("  --geometry-spread <x> Relative half-width of that variation, in [0, 1) (default 0.25)");
    puts("  --sigt-floor <x>      Well-conditioned diagnostic data: sigT in [x, 1)");
    puts("  --dump-flux <file>    Write the final scalar flux (raw float32)");
    puts("  --verify <file>       Compare the flux with a CPU replay of the same stream (raw float32,");
    puts("                        e.g. from tools/oracle_replay.py) and print PASS/FAIL");
    puts("  --tolerance <x>       L2-relative tolerance of --verify (default 1e-5)");
    puts("See readme for full description of default run values");
    exit(1);
}

static const char *need(int argc, char **argv, int *i)
{
    if (++*i >= argc) usage_and_exit();
    return argv[*i];
}

static int lookup(const char *s, const char *const names[], int n)
{
    for (int i = 0; i < n; i++)
        if (strcmp(s, names[i]) == 0) return i;
    usage_and_exit();
    return -1;
}

static void defaults(Input *I)
{
    memset(I, 0, sizeof *I);
    I->source_2D_regions = 5000;
    I->coarse_axial_intervals = 27;
    I->fine_axial_intervals = 5;
    I->decomp_assemblies_ax = 20;
    I->segments = 50000000;
    I->egroups = 128;
    I->nthreads = 1;
    I->seg_per_thread = 100;
    I->seed = 42ull;
    I->exp_mode = SMK_EXP_POLY;
    I->math_mode = SMK_MATH_FAST;
    I->gpus = 1;
    I->geometry_spread = 0.25f;
    I->tolerance = 1e-5;
}

static void parse(int argc, char **argv, Input *I)
{
    static const char *const exps[] = {"poly", "mufu", "glibc", "table"};
    static const char *const maths[] = {"fast", "strict"};
    static const char *const reduces[] = {"peer", "nccl"};
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (!strcmp(a, "-t")) I->nthreads = atoi(need(argc, argv, &i));
        else if (!strcmp(a, "-s")) I->segments = strtol(need(argc, argv, &i), NULL, 10);
        else if (!strcmp(a, "-e")) I->egroups = atoi(need(argc, argv, &i));
        else if (!strcmp(a, "-p")) I->seg_per_thread = atoi(need(argc, argv, &i));
        else if (!strcmp(a, "-d")) I->device = atoi(need(argc, argv, &i));
        else if (!strcmp(a, "--regions-2d")) I->source_2D_regions = atoi(need(argc, argv, &i));
        else if (!strcmp(a, "--seed")) I->seed = strtoull(need(argc, argv, &i), NULL, 10);
        else if (!strcmp(a, "--exp")) I->exp_mode = lookup(need(argc, argv, &i), exps, 4);
        else if (!strcmp(a, "--math")) I->math_mode = lookup(need(argc, argv, &i), maths, 2);
        else if (!strcmp(a, "--gpus")) I->gpus = atoi(need(argc, argv, &i));
        else if (!strcmp(a, "--allreduce")) I->allreduce = lookup(need(argc, argv, &i), reduces, 2);
        else if (!strcmp(a, "--host-fill")) I->host_fill = 1;
        else if (!strcmp(a, "--tally-f64")) I->tally_f64 = 1;
        else if (!strcmp(a, "--segment-geometry")) I->segment_geometry = 1;
        else if (!strcmp(a, "--fit-per-sweep")) I->fit_per_sweep = 1;
        else if (!strcmp(a, "--geometry-spread")) { I->geometry_spread = (float)atof(need(argc, argv, &i)); I->segment_geometry = 1; }
        else if (!strcmp(a, "--sigt-floor")) I->sigt_floor = (float)atof(need(argc, argv, &i));
        else if (!strcmp(a, "--dump-flux")) I->dump_flux = need(argc, argv, &i);
        else if (!strcmp(a, "--verify")) I->verify_flux = need(argc, argv, &i);
        else if (!strcmp(a, "--tolerance")) I->tolerance = atof(need(argc, argv, &i));
        else usage_and_exit();
    }
    if (I->nthreads < 1 || I->segments < 0 || I->egroups < 1 || I->seg_per_thread < 1 || I->gpus < 1)
        usage_and_exit();
}

static double estimate_mb(const Input *I)
{
    double fine = (double)I->source_3D_regions * I->fine_axial_intervals * I->egroups * sizeof(float);
    double sig = (double)I->source_3D_regions * I->egroups * sizeof(float);
    return (2.0 * fine + sig) / 1024.0 / 1024.0;
}

static void summary(const Input *I, const char *device_name)
{
    centered("INPUT SUMMARY");
    rule();
    printf("%-25s%s\n", "CUDA Device: ", device_name);
    printf("%-25s%d\n", "Energy Groups:", I->egroups);
    printf("%-25s%d\n", "2D Source Regions:", I->source_2D_regions);
    printf("%-25s%d\n", "Coarse Axial Intervals:", I->coarse_axial_intervals);
    printf("%-25s%d\n", "Fine Axial Intervals:", I->fine_axial_intervals);
    printf("%-25s%d\n", "Axial Decomposition:", I->decomp_assemblies_ax);
    printf("%-25s%d\n", "3D Source Regions:", I->source_3D_regions);
    printf("%-25s", "Segments:"); grouped(I->segments);
    printf("%-25s", "Random Number Streams:"); grouped((long)smk_num_tracks(I->segments, I->seg_per_thread));
    printf("%-25s%.2f\n", "Memory Estimate (MB):", estimate_mb(I));
    printf("%-25s%d\n", "Segments per CUDA block:", I->seg_per_thread);
    printf("%-25s%s\n", "Exponential Table:", I->exp_mode == SMK_EXP_TABLE ? "ON" : "OFF");
    printf("%-25s%llu\n", "Stream Seed:", I->seed);
    if (I->segment_geometry)
        printf("%-25sper segment, spread %.3f\n", "Segment Geometry:", I->geometry_spread);
    if (I->gpus > 1)
        printf("%-25s%d (%s all-reduce)\n", "GPUs:", I->gpus, I->allreduce == SMK_ALLREDUCE_NCCL ? "NCCL" : "peer-memory");
    rule();
}

/* host-side deterministic fill is only needed for --host-fill; it is the same stream the
 * device fill uses, written out again here in C so that the driver
 * stays free of any dependency on the oracle. */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = 0xD2511F53ull * c[0], p1 = 0xCD9E8D57ull * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

static void host_fill(float *dst, long n, uint32_t array_id, unsigned long long seed, float floor_)
{
    const float span = 1.0f - floor_;
    for (long q = 0; q * 4 < n; q++) {
        uint32_t c[4] = {(uint32_t)q, (uint32_t)((unsigned long long)q >> 32), array_id, 0x46494C4Cu};
        philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        for (int j = 0; j < 4 && q * 4 + j < n; j++) {
            float u = (float)(int32_t)(c[j] >> 1) * 0x1.0p-31f;
            dst[q * 4 + j] = floor_ > 0.0f ? floor_ + u * span : u;
        }
    }
}

#define CHECK(call)                                                        \
    do {                                                                   \
        if ((call) != SMK_OK) {                                            \
            printf("Error at %s:%d: %s\n", __FILE__, __LINE__, smk_last_error()); \
            return EXIT_FAILURE;                                           \
        }                                                                  \
    } while (0)

int main(int argc, char *argv[])
{
    const int version = 4;
    Input I;
    defaults(&I);
    parse(argc, argv, &I);
    /* main.c:18-19 */
    I.source_3D_regions = (int)ceil((double)I.source_2D_regions * I.coarse_axial_intervals /
                                    I.decomp_assemblies_ax);

    banner(version);

    if (smk_device_count() < 1) {
        printf("Error: no CUDA device visible (this build has no CPU fallback)\n");
        return EXIT_FAILURE;
    }
    char device_name[256];
    CHECK(smk_device_name(I.device, device_name, sizeof device_name));
    summary(&I, device_name);

    centered("INITIALIZATION");
    rule();
    smk_params p;
    memset(&p, 0, sizeof p);
    p.source_3D_regions = I.source_3D_regions;
    p.fine_axial_intervals = I.fine_axial_intervals;
    p.egroups = I.egroups;
    p.seg_per_track = I.seg_per_thread;
    p.segments = I.segments;
    p.seed = I.seed;
    p.exp_mode = I.exp_mode;
    p.math_mode = I.math_mode;
    p.device = I.device;
    p.flags = (I.tally_f64 ? SMK_FLAG_TALLY_F64 : 0) | (I.segment_geometry ? SMK_FLAG_SEGMENT_GEOMETRY : 0) |
              (I.fit_per_sweep ? SMK_FLAG_FIT_PER_SWEEP : 0);

    smk_ctx *ctx = NULL;
    smk_multi *multi = NULL;
    if (I.gpus > 1) CHECK(smk_multi_create(&p, I.gpus, NULL, I.allreduce, &multi));
    else CHECK(smk_create(&p, &ctx));
    if (I.segment_geometry) {
        /* base values = the reference's placeholders (kernel.c:99-104) */
        smk_geometry g = {0.1f, 0.3f, 0.5f, 0.9f, 0.3f, 0.7f, I.geometry_spread};
        if (multi) CHECK(smk_multi_set_geometry(multi, &g));
        else CHECK(smk_set_geometry(ctx, &g));
    }
    const long n_fine = (long)I.source_3D_regions * I.fine_axial_intervals * I.egroups;
    float *flux = (float *)malloc((size_t)n_fine * sizeof(float));
    if (!flux) { printf("Error: out of host memory\n"); return EXIT_FAILURE; }
    printf("Building Source Data Arrays...\n");
    if (I.host_fill) {
        const long n_sig = (long)I.source_3D_regions * I.egroups;
        float *src = (float *)malloc((size_t)n_fine * sizeof(float));
        float *sig = (float *)malloc((size_t)n_sig * sizeof(float));
        if (!src || !sig) { printf("Error: out of host memory\n"); return EXIT_FAILURE; }
        host_fill(src, n_fine, 0u, I.seed, 0.0f);
        host_fill(flux, n_fine, 1u, I.seed, 0.0f);
        host_fill(sig, n_sig, 2u, I.seed, I.sigt_floor);
        if (multi) CHECK(smk_multi_upload(multi, src, flux, sig));
        else CHECK(smk_upload(ctx, src, flux, sig));
        free(src); free(sig);
    } else {
        if (multi) CHECK(smk_multi_fill_device(multi, I.sigt_floor));
        else CHECK(smk_fill_device(ctx, I.sigt_floor));
    }
    printf("Initialization Complete.\n");
    rule();

    centered("SIMULATION");
    rule();
    printf("Attentuating fluxes across segments...\n");
    double seconds = 0.0, kernel_seconds = 0.0;
    if (multi) {
        /* runtime = sweep + the one all-reduce of the tallies */
        CHECK(smk_multi_run(multi, &kernel_seconds, &seconds));
    } else {
        CHECK(smk_run(ctx, 0, smk_num_tracks(I.segments, I.seg_per_thread), &seconds));
        kernel_seconds = seconds;
    }
    printf("Simulation Complete.\n");

    rule();
    centered("RESULTS SUMMARY");
    rule();
    const double intersections = (double)I.segments * (double)I.egroups;
    const double tpi = seconds / intersections * 1.0e9;
    printf("%-25s%.3f seconds\n", "Runtime:", seconds);
    printf("%-25s%.8lf ns\n", "Time per Intersection:", tpi);
    printf("%-25s%.4e\n", "Intersections per second:", seconds > 0 ? intersections / seconds : 0.0);
    if (multi) printf("%-25s%.3f seconds\n", "Slowest GPU kernel:", kernel_seconds);
    rule();

    centered("VERIFICATION");
    rule();
    uint64_t checksum = 0;
    if (multi) {
        CHECK(smk_multi_download_checksum(multi, &checksum));
        CHECK(smk_multi_download_flux(multi, 0, flux));
    } else {
        CHECK(smk_download_checksum(ctx, &checksum));
        CHECK(smk_download_flux(ctx, flux));
    }
    double sum = 0.0, sumsq = 0.0;
    long nonfinite = 0;
    for (long i = 0; i < n_fine; i++) {
        if (!isfinite(flux[i])) { nonfinite++; continue; }
        sum += flux[i];
        sumsq += (double)flux[i] * flux[i];
    }
    printf("%-25s%016llx\n", "Segment Index Checksum:", (unsigned long long)checksum);
    printf("%-25s%.9e\n", "Scalar Flux Sum:", sum);
    printf("%-25s%.9e\n", "Scalar Flux L2 Norm:", sqrt(sumsq));
    printf("%-25s%ld\n", "Non-finite Flux Values:", nonfinite);
    int verdict = 0;
    if (I.verify_flux) {
        /* scalar flux vs a CPU replay of the same stream: same finite pattern and
         * ||gpu - cpu||_2 / ||cpu||_2 <= tolerance */
        float *ref = (float *)malloc((size_t)n_fine * sizeof(float));
        FILE *f = fopen(I.verify_flux, "rb");
        if (!ref || !f || fread(ref, sizeof(float), (size_t)n_fine, f) != (size_t)n_fine) {
            printf("Error: cannot read %ld floats from %s\n", n_fine, I.verify_flux);
            return EXIT_FAILURE;
        }
        fclose(f);
        double num = 0.0, den = 0.0;
        long pattern = 0;
        for (long i = 0; i < n_fine; i++) {
            if (isfinite(flux[i]) != isfinite(ref[i])) { pattern++; continue; }
            if (!isfinite(ref[i])) continue;
            const double d = (double)flux[i] - ref[i];
            num += d * d;
            den += (double)ref[i] * ref[i];
        }
        const double err = den > 0 ? sqrt(num / den) : sqrt(num);
        verdict = (pattern == 0 && err <= I.tolerance) ? 1 : -1;
        printf("%-25s%.3e (tolerance %.1e)\n", "L2-relative Error:", err, I.tolerance);
        printf("%-25s%ld\n", "Finite-pattern Mismatch:", pattern);
        printf("%-25s%s\n", "Verification:", verdict > 0 ? "PASS" : "FAIL");
        free(ref);
    } else {
        printf("%-25s%s\n", "Verification:", "not requested (--verify <cpu replay flux>)");
    }
    rule();

    if (I.dump_flux) {
        FILE *f = fopen(I.dump_flux, "wb");
        if (!f || fwrite(flux, sizeof(float), (size_t)n_fine, f) != (size_t)n_fine) {
            printf("Error: cannot write %s\n", I.dump_flux);
            return EXIT_FAILURE;
        }
        fclose(f);
    }
    free(flux);
    smk_destroy(ctx);
    smk_multi_destroy(multi);
    return verdict < 0 ? 2 : 0;
}
