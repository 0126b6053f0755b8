/*
 * smk.h -- C ABI of the B200-native SimpleMOC-kernel segment-attenuation path.
 *
 * The reference has no plugin/FFI layer: its seam is the plain C call
 *     void run_kernel(Input *I, Source *S, Table *table);
 * (/root/reference/src/cpu/SimpleMOC-kernel_header.h:81, called once from
 * /root/reference/src/cpu/main.c:46, and its CUDA twin launched from
 * /root/reference/src/cuda/main.cu:90).  This header is what a C host driver
 * (ours: simplemoc-kernel_b200/host/smk_main.c; the reference's main.c with the
 * three-line patch shown in INTEGRATION.md) binds instead.  Plain pointers and
 * sizes only; no CUDA or torch types appear in any signature.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SMK_E* code on failure,
 *     and never calls exit(); smk_last_error() returns the message of the last
 *     failure on the calling thread (the reference prints "Error at file:line"
 *     and returns EXIT_FAILURE, /root/reference/src/cuda/SimpleMOC-kernel_header.h:24-26).
 *   - host arrays use the reference's unpadded layouts (init.c:35-54):
 *         fine_source[R][F][G], fine_flux[R][F][G], sigT[R][G]   (float)
 *     with R = source_3D_regions, F = fine_axial_intervals, G = egroups.
 *   - device arrays are padded to G_pad groups per row (smk_padded_groups()); any
 *     G >= 1 is supported (rows wider than 256 groups are swept in blocks of 256).
 *   - a "track" is seg_per_track consecutive segments of the deterministic
 *     stream that share one carried angular flux.
 */
#ifndef SMK_H
#define SMK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMK_ABI_VERSION 2

/* error codes */
#define SMK_OK          0
#define SMK_EINVAL     -1   /* bad argument / unsupported configuration        */
#define SMK_ECUDA      -2   /* a CUDA runtime call failed                      */
#define SMK_ENOMEM     -3   /* host or device allocation failed                */
#define SMK_ESTATE     -4   /* call made in the wrong state (e.g. no upload)   */

/* how 1 - exp(-tau) is evaluated (kernel.c:216-223) */
#define SMK_EXP_POLY    0   /* FMA-pipe polynomial, correctly rounded where the
                               reference formula is ill-conditioned (default).
                               Domain: the polynomial is fitted for tau = sigT*ds
                               <= 0.7 (the reference's own data: sigT < 1, ds =
                               0.7).  The library tracks max(sigT) of the uploaded
                               data; when max(sigT) * max(ds) > 0.7 it switches to
                               a form that evaluates tau > 0.7 with MUFU.EX2 (well
                               conditioned there), so any non-negative sigT is safe */
#define SMK_EXP_MUFU    1   /* MUFU.EX2 (ex2.approx.ftz), 2 ulp                 */
#define SMK_EXP_GLIBC   2   /* double-precision replica of glibc 2.39 expf      */
#define SMK_EXP_TABLE   3   /* the reference's interpolation table
                               (init.c:81-117, kernel.c:337-361; TABLE build)  */

/* arithmetic of the attenuation formulae */
#define SMK_MATH_FAST   0   /* FMA contraction, one MUFU.RCP instead of 5 divides */
#define SMK_MATH_STRICT 1   /* the reference's operation order, IEEE div, no FMA:
                               per-intersection results are bit-identical to a
                               -O2 -ffp-contract=off build of kernel.c          */

/*
 * Problem description.  Fields 1-5 are the reference's Input
 * (/root/reference/src/cpu/SimpleMOC-kernel_header.h:24-33) plus the CUDA
 * variant's seg_per_thread (/root/reference/src/cuda/SimpleMOC-kernel_header.h:39,
 * the -p option, io.cu:141-147); the rest configure the changed subsystems.
 */
typedef struct smk_params {
    int32_t  source_3D_regions;     /* R  (main.c:18-19)                        */
    int32_t  fine_axial_intervals;  /* F  (init.c:10), must be >= 2             */
    int32_t  egroups;               /* G  (-e)                                  */
    int32_t  seg_per_track;         /* -p, default 100 (init.cu:41)             */
    int64_t  segments;              /* N  (-s)                                  */
    uint64_t seed;                  /* key of the counter-based stream          */
    int32_t  exp_mode;              /* SMK_EXP_*                                */
    int32_t  math_mode;             /* SMK_MATH_*                               */
    int32_t  device;                /* CUDA device ordinal (-d, io.cu:148-158)  */
    int32_t  flags;                 /* SMK_FLAG_*                               */
} smk_params;

#define SMK_FLAG_KEEP_PSI  1        /* keep each track's outgoing psi (tests)   */
#define SMK_FLAG_TALLY_F64 2        /* diagnostic: accumulate the tallies in f64 (order-independent
                                       yardstick for fp32 accumulation noise); every kernel */
#define SMK_FLAG_SEGMENT_GEOMETRY 4 /* dz, zin, weight, mu, mu2, ds of kernel.c:99-104 vary per segment
                                       (smk_geometry / smk_set_geometry below)  */
#define SMK_FLAG_FIT_PER_SWEEP 8     /* OFF by default.  The quadratic axial source fit (kernel.c:111-191) depends only
                                       on (region, interval, group): with this flag it is evaluated once per sweep for
                                       every such triple (inside smk_run*, by the pass that lays out the gather records)
                                       instead of once per segment, with the same operations in the same order: results
                                       are bit-identical, 8 of 45 operations per interior intersection leave the segment
                                       loop.  The reference evaluates the fit per segment, so the default does too;
                                       needs SMK_MATH_FAST and the constant geometry (SMK_EINVAL otherwise), and takes
                                       effect up to 128 groups while the derived arrays fit the L2 (f32 tallies) --
                                       smk_kernel_name says which form runs. */

/*
 * Segment geometry.  /root/reference/src/cpu/kernel.c:95-104: "Some placeholder constants - In the
 * full app some of these are calculated based off position in geometry."  With
 * SMK_FLAG_SEGMENT_GEOMETRY the six placeholders are parameters and vary per segment: words 2,3 of
 * the segment's stream block give four 16-bit fields u, each mapped to a factor
 * f(u) = 1 + spread * (u * 2^-15 - 1) in [1 - spread, 1 + spread):
 *     ds = ds0 f(w2 >> 16)    zin = zin0 f(w2 & 0xFFFF)    mu = mu0 f(w3 >> 16)
 *     mu2 = mu2_0 f(w3 >> 16)^2    weight = weight0 f(w3 & 0xFFFF)    dz = dz0 (axial mesh)
 * every step one IEEE binary32 operation, replayed identically by the CPU
 * oracle.  spread = 0 reproduces the base values exactly; base = kernel.c:99-104 is the default.
 */
typedef struct smk_geometry {
    float dz, zin, weight, mu, mu2, ds;   /* base values; defaults 0.1 0.3 0.5 0.9 0.3 0.7 */
    float spread;                         /* in [0, 1); default 0.25                        */
} smk_geometry;

typedef struct smk_ctx smk_ctx;     /* opaque: device buffers, stream, events   */

/* ---- library ---------------------------------------------------------- */
int          smk_abi_version(void);
const char  *smk_last_error(void);
int          smk_device_count(void);
/* name of device `device` copied into buf (the reference prints it, io.cu:87) */
int          smk_device_name(int device, char *buf, size_t buflen);
/* G_pad for G groups: row stride, in floats, of every device array */
int          smk_padded_groups(int egroups);
int64_t      smk_num_tracks(int64_t segments, int seg_per_track);

/* ---- context ---------------------------------------------------------- */
/* allocates device buffers for p on p->device; replaces initialize_device_sources
 * (/root/reference/src/cuda/init.cu:105-127) */
int   smk_create(const smk_params *p, smk_ctx **out);
void  smk_destroy(smk_ctx *ctx);
/* run on a caller-owned cudaStream_t (passed as void*) instead of the context's */
int   smk_set_stream(smk_ctx *ctx, void *cuda_stream);
/* base values and spread of the per-segment geometry; the context must have been created
 * with SMK_FLAG_SEGMENT_GEOMETRY (SMK_ESTATE otherwise).  dz, ds > 0, 0 <= spread < 1. */
int   smk_set_geometry(smk_ctx *ctx, const smk_geometry *g);
int   smk_get_geometry(const smk_ctx *ctx, smk_geometry *g);
/* name of the kernel instantiation the next smk_run will launch, e.g.
 * "attenuate_warp_track<4 groups/lane, poly, f32 tally, const geometry>" */
const char *smk_kernel_name(smk_ctx *ctx);

/* ---- data ------------------------------------------------------------- */
/* host (unpadded, pageable or pinned) -> device (padded); tallies are zeroed.
 * fine_flux may be NULL (initial scalar flux = 0).  Returns when the host arrays may be reused. */
int   smk_upload(smk_ctx *ctx, const float *fine_source, const float *fine_flux,
                 const float *sigT);
/* same, enqueue only: the copies are ordered on the context's stream before any later smk_run*;
 * the host arrays must stay unchanged until smk_synchronize (pinned memory makes the copies truly
 * asynchronous; pageable memory is staged by the driver) */
int   smk_upload_async(smk_ctx *ctx, const float *fine_source, const float *fine_flux,
                       const float *sigT);
/* rows [row_begin, row_begin + rows) of one array (a "row" is G floats: R*F rows for source and
 * flux, R rows for sigT), enqueue only.  For callers that split the upload over ranks and complete
 * the replicas with an all-gather on the device arrays (smk_device_*).  Tallies are NOT reset.
 * After a partial sigT upload the library no longer knows max(sigT): call smk_scan_sigt_max once
 * the device array is complete, or the POLY exponential stays in its (slower) wide-range form. */
#define SMK_ARRAY_SOURCE 0
#define SMK_ARRAY_FLUX   1
#define SMK_ARRAY_SIGT   2
int   smk_upload_rows_async(smk_ctx *ctx, int array, int64_t row_begin, int64_t rows, const float *host);
/* max(sigT) of the device array (one small kernel + 4-byte read back; synchronises the stream) */
int   smk_scan_sigt_max(smk_ctx *ctx, float *max_out);
/* caller-supplied upper bound of sigT over ALL rows of the device array, for data the library did not
 * see on the host (rows gathered from peers): e.g. the all-reduce(max) of the ranks' slice maxima.
 * A bound that is too small makes SMK_EXP_POLY wrong for tau > 0.7; +inf is always safe. */
int   smk_set_sigt_bound(smk_ctx *ctx, float bound);
/* device-side deterministic fill, bit-identical to the host stream fill that
 * replaces init.c:64-75; sigt_floor = 0 for U[0,1) */
int   smk_fill_device(smk_ctx *ctx, float sigt_floor);
/* zero the tally deltas (start of a new sweep) */
int   smk_reset_tallies(smk_ctx *ctx);
/* fine_flux_out[R][F][G] = initial flux + tallies accumulated since the last
 * reset (kernel.c:274-277 applied to every replayed segment) */
int   smk_download_flux(smk_ctx *ctx, float *fine_flux_out);
/* rows [row_begin, row_begin + rows) of the same (out[rows][G]), enqueue only: for callers that
 * reduce-scatter the tallies over ranks and read back one slice per rank */
int   smk_download_flux_rows_async(smk_ctx *ctx, int64_t row_begin, int64_t rows, float *out);
/* Two contexts pipelined on ONE GPU (uploads / downloads of one under the sweep of the other): the sweep is a
 * persistent grid that fills every SM, so a small kernel of the other context that becomes ready at the same moment
 * (the `flux0 + tallies` pass in front of a download) can lose the race and wait a whole sweep for an SM, delaying
 * the download, the caller's next upload and with it the sweep after.  smk_wait_finalized(ctx, other) makes
 * everything enqueued on ctx from now on wait (on the device, cudaStreamWaitEvent) until the `flux0 + tallies` pass
 * of `other`'s most recent smk_download_flux* has run; the copy itself still overlaps.  No-op if `other` has not
 * enqueued a download yet.  Call it before smk_run_async. */
int   smk_wait_finalized(smk_ctx *ctx, smk_ctx *other);
/* outgoing psi of the tracks [track_begin, track_end) swept by the LAST smk_run*:
 * psi_out[(t - track_begin) * G + g]; n_tracks must equal track_end - track_begin of that run
 * (SMK_EINVAL otherwise: it is the capacity of psi_out); needs SMK_FLAG_KEEP_PSI */
int   smk_download_psi(smk_ctx *ctx, float *psi_out, int64_t n_tracks);
/* sum over replayed segments s of (QSR_id*F + FAI_id + 1) * ((s & 0xFFFF) + 1)
 * mod 2^64, accumulated since the last reset: the indexing fingerprint */
int   smk_download_checksum(smk_ctx *ctx, uint64_t *checksum);

/* ---- the hot path ------------------------------------------------------ */
/* attenuate tracks [track_begin, track_end) (run_kernel's segment loop,
 * kernel.c:43-55).  Synchronous; *kernel_seconds (may be NULL) receives the
 * CUDA-event time of the kernel alone, as main.cu:89-96 measures it. */
int   smk_run(smk_ctx *ctx, int64_t track_begin, int64_t track_end,
              double *kernel_seconds);
/* same, enqueue only (no synchronisation, no timing) */
int   smk_run_async(smk_ctx *ctx, int64_t track_begin, int64_t track_end);
int   smk_synchronize(smk_ctx *ctx);
/* number of kernel launches issued through ctx since creation */
int64_t smk_launch_count(const smk_ctx *ctx);

/*
 * Drop-in for run_kernel(I, S, table) with HOST slabs: upload, attenuate all
 * tracks, download.  fine_flux is updated in place like the reference does
 * (kernel.c:276).  kernel_seconds / total_seconds may be NULL.
 */
int   smk_run_host(const smk_params *p, const float *fine_source,
                   float *fine_flux, const float *sigT,
                   double *kernel_seconds, double *total_seconds);

/* ---- plumbing for callers that own device memory (torch, NCCL) --------- */
/* raw device pointers of the context's padded arrays (void* = float*) */
void *smk_device_tally(smk_ctx *ctx);      /* [replicas][R][F][G_pad], the all-reduce operand;
                                              replicas > 1 only when R*F < 4096 (contention relief).
                                              With SMK_MATH_FAST and the constant geometry the sums are
                                              kept WITHOUT the segment weight (kernel.c:262; 0.5): the
                                              download applies it once, bit-identically */
void *smk_device_flux0(smk_ctx *ctx);      /* [R][F][G_pad] initial flux            */
void *smk_device_source(smk_ctx *ctx);     /* [R][F][G_pad]                         */
void *smk_device_sigT(smk_ctx *ctx);       /* [R][G_pad]                            */
int64_t smk_padded_elems(const smk_ctx *ctx); /* replicas*R*F*G_pad: floats behind smk_device_tally */
/* pinned host memory for end-to-end runs */
void *smk_alloc_host(size_t bytes);
void  smk_free_host(void *p);

/* ---- multi-GPU sweep in one process (north star item 4) ----------------- */
/*
 * One context per device; the tracks of the stream are sharded by contiguous range
 * (device k sweeps tracks [k*T/P, (k+1)*T/P)); every device holds a replica of the
 * source data and its own zeroed tally deltas; ONE all-reduce of the tally arrays ends
 * the sweep.  The reference has no multi-device path (only -d <id>, io.cu:148-158).
 * all-reduce implementations:
 *   SMK_ALLREDUCE_PEER  one kernel per device over NVLink peer memory: device k sums slice k
 *                       of every peer's tallies (P2P loads) and writes the sum back to every
 *                       peer (P2P stores); deterministic summation order
 *   SMK_ALLREDUCE_NCCL  ncclAllReduce (libnccl.so.2 is dlopen'ed on first use)
 */
#define SMK_ALLREDUCE_PEER 0
#define SMK_ALLREDUCE_NCCL 1
typedef struct smk_multi smk_multi;
/* devices == NULL means ordinals 0 .. n_devices-1; p->device is ignored */
int   smk_multi_create(const smk_params *p, int n_devices, const int *devices, int allreduce,
                       smk_multi **out);
void  smk_multi_destroy(smk_multi *m);
/* ONE host->device copy (to the first device), then every other device pulls the padded arrays over
 * NVLink peer copies, all devices in parallel */
int   smk_multi_upload(smk_multi *m, const float *fine_source, const float *fine_flux, const float *sigT);
int   smk_multi_set_geometry(smk_multi *m, const smk_geometry *g);
int   smk_multi_fill_device(smk_multi *m, float sigt_floor);
/* sweep all tracks + all-reduce.  kernel_seconds = slowest device's kernel (CUDA events),
 * total_seconds = wall clock from first launch to the end of the all-reduce; either may be NULL */
int   smk_multi_run(smk_multi *m, double *kernel_seconds, double *total_seconds);
/* flux0 + all-reduced tallies, read from device `which` (every device holds the same sum) */
int   smk_multi_download_flux(smk_multi *m, int which, float *fine_flux_out);
int   smk_multi_download_checksum(smk_multi *m, uint64_t *checksum);   /* summed over devices */
int   smk_multi_device_count(const smk_multi *m);

/* ---- diagnostics ------------------------------------------------------- */
/* d_out[i] = exp(-tau[i]) as evaluated by exp_mode (host arrays, n elements);
 * used to sweep the exponential against libm.  exp_mode | SMK_DEBUG_EXP_PACKED evaluates the
 * packed (FP32x2) form of the FAST kernels; SMK_DEBUG_EXP_WIDE selects POLY's wide-range form,
 * SMK_DEBUG_EXP_TRACK its libm-following form (the scalar POLY of SMK_MATH_STRICT always follows libm) */
#define SMK_DEBUG_EXP_PACKED 0x100
#define SMK_DEBUG_EXP_WIDE   0x200
#define SMK_DEBUG_EXP_TRACK  0x400   /* packed POLY as the per-segment-geometry kernels evaluate it: follows
                                        glibc's expf where that is not correctly rounded (tau < 2^-8) */
int   smk_debug_exp(int exp_mode, const float *tau, float *out, int64_t n, int device);
/* (QSR_id, FAI_id) of segments [seg_begin, seg_begin+n) as the kernel draws them */
int   smk_debug_segment_ids(const smk_params *p, int64_t seg_begin, int64_t n,
                            int32_t *qsr_out, int32_t *fai_out);
/* dz, zin, weight, mu, mu2, ds of the same segments as the kernels derive them: geom6_out[n][6] */
int   smk_debug_segment_geometry(const smk_params *p, const smk_geometry *g, int64_t seg_begin, int64_t n,
                                 float *geom6_out);

#ifdef __cplusplus
}
#endif
#endif /* SMK_H */
