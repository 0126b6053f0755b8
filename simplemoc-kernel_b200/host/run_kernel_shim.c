/*
 * run_kernel_shim.c -- link-time drop-in for the reference's hot path.
 *
 * Defines the reference's own seam
 *     void run_kernel(Input *I, Source *S, Table *table);      (SimpleMOC-kernel_header.h:81)
 * on top of the C ABI of include/smk.h, so that the UNMODIFIED reference driver
 * (/root/reference/src/cpu/main.c, init.c, io.c) links against libsmk.so instead of its kernel.c:
 *
 *   gcc -std=gnu99 -fopenmp -DOPENMP -I/root/reference/src/cpu -Iinclude \
 *       /root/reference/src/cpu/{main,init,io}.c simplemoc-kernel_b200/host/run_kernel_shim.c \
 *       -Lsimplemoc-kernel_b200/lib -lsmk -lm -o SimpleMOC-kernel
 *
 * (The repository's reference-build recipe builds exactly this while /root/reference is mounted, see
 * INTEGRATION.md section 1; the reference header is taken from there at build time, nothing of it is
 * copied here.)  The reference's main() then
 * prints its own banner, input summary and "Time per Intersection" around a sweep that ran on the GPU;
 * its timer brackets the whole call (main.c:45-47), i.e. context creation, H2D, sweep and D2H.
 *
 * What changes for the user, by design of the north star: segments come from the deterministic counter
 * stream (seed SMK_SEED, default 42) instead of rand_r(), so the run is reproducible.
 * The CPU driver has no -p / -d options (io.c:109-154), so segments per track and the device come from
 * the environment (SMK_SEG_PER_TRACK, default 100 = the OpenMP chunk of kernel.c:43; SMK_DEVICE, default 0);
 * -t only sizes the reference's OpenMP team, which the GPU sweep does not use.  All of this is printed.
 */
#include "SimpleMOC-kernel_header.h"
#include "smk.h"

void run_kernel(Input *I, Source *S, Table *table)
{
    (void)table;
    smk_params p;
    memset(&p, 0, sizeof p);
    p.source_3D_regions = I->source_3D_regions;      /* main.c:18-19 */
    p.fine_axial_intervals = I->fine_axial_intervals;
    p.egroups = I->egroups;
    const char *spt = getenv("SMK_SEG_PER_TRACK");
    p.seg_per_track = spt ? atoi(spt) : 100;          /* cuda init.cu:41 / the OpenMP chunk, kernel.c:43 */
    p.segments = I->segments;
    const char *seed = getenv("SMK_SEED");
    p.seed = seed ? strtoull(seed, NULL, 10) : 42ull;
#ifdef TABLE
    p.exp_mode = SMK_EXP_TABLE;                       /* a TABLE=yes build keeps its table semantics */
#else
    p.exp_mode = SMK_EXP_POLY;
#endif
    p.math_mode = SMK_MATH_FAST;
    const char *dev = getenv("SMK_DEVICE");
    p.device = dev ? atoi(dev) : 0;
    printf("GPU sweep on device %d: %d segments per track, stream seed %llu; -t %d is not used by the GPU path\n",
           p.device, p.seg_per_track, (unsigned long long)p.seed, I->nthreads);

    /* initialize_sources lays the three slabs out contiguously from S[0] (init.c:35-54) */
    double kernel_s = 0.0, total_s = 0.0;
    if (smk_run_host(&p, S[0].fine_source, S[0].fine_flux, S[0].sigT, &kernel_s, &total_s) != SMK_OK) {
        printf("Error: %s\n", smk_last_error());
        exit(EXIT_FAILURE);
    }
    printf("GPU sweep: %.6f s kernel, %.6f s with upload/download (libsmk)\n", kernel_s, total_s);
}
