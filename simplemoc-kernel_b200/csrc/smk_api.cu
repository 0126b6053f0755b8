// smk_api.cu -- implementation of the C ABI declared in include/smk.h.
//
// Host side of the path: owns the padded device arrays, the stream and the events,
// picks the kernel instantiation for (G, math mode, exp mode) and sizes the grid as
// a multiple of the SM count (persistent CTAs striding over tracks).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/smk.h"
#include "smk_kernels.cuh"

namespace smk {

// 2^(i/32) as IEEE binary64 bit patterns with i << 47 subtracted (glibc __exp2f_data.tab)
static const uint64_t h_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

static thread_local char t_error[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
    return code;
}

#define SMK_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(SMK_ECUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                        cudaGetErrorString(e_));                                            \
    } while (0)

// The reference's table (init.c:81-117, called as buildExponentialTable(0.01, 10.0, I),
// main.c:33): same expressions, same types, evaluated on the host in IEEE arithmetic.
static void build_exp_table(ExpTable &t)
{
    const float precision = 0.01f, maxVal = 10.0f;
    const int N = (int)(maxVal * std::sqrt(1.0 / (8.0 * precision * 0.01)));   // 353
    const float dx = maxVal / (float)N;
    for (int n = 0; n < kTableN; ++n) {
        const float ex = (float)std::exp((double)(-n * dx));
        t.pairs[n].x = -ex;
        t.pairs[n].y = 1 + (n * dx - 1) * ex;
    }
    t.dx = dx;
    t.maxVal = maxVal - dx;
    (void)N;
}

struct Shape {
    int groups_pad;   // floats per padded row
    int lpt;          // lanes per track
    int nchunk;       // float4 per lane per group block
    int group_blocks; // blocks of 4*lpt*nchunk groups per row (> 1 only for more than 256 groups)
};

// G <= 128: one float4 per lane, LPT = next power of two of ceil(G/4) lanes per track.
// G > 128: blocks of 256 groups (two float4 per lane, the fastest measured shape), as many as the row needs.
static bool shape_for(int groups, Shape &s)
{
    if (groups < 1) return false;
    const int f4 = (groups + 3) / 4;
    if (f4 <= 32) {
        int l = 1;
        while (l < f4) l <<= 1;
        s.lpt = l;
        s.nchunk = 1;
        s.group_blocks = 1;
    } else {
        s.lpt = 32;
        s.nchunk = 2;
        s.group_blocks = (f4 + 63) / 64;
    }
    s.groups_pad = 4 * s.lpt * s.nchunk * s.group_blocks;
    return true;
}

typedef void (*AttenuateFn)(const KernelArgs);

// groups per lane of attenuate_record_tracks: four (two 256-bit loads per lane and segment) from 5 groups up, two for
// 1..4 groups, where four would leave one lane per track (measured, profiles/ab_r02.md: 29 groups 5.63e11 vs 5.42e11,
// 13 groups 4.62e11 vs 4.49e11, 7 groups equal, 3 groups 2.25e11 vs 2.80e11)
// With per-segment geometry the 5..8-group shape is faster with two (3.59e11 vs 3.46e11 at 7 groups; 17..32 groups:
// 5.05e11 with four vs 4.88e11).
static int record_groups_per_lane(int groups_pad, bool geom) { return groups_pad > (geom ? 8 : 4) ? 4 : 2; }

struct KernelChoice {
    AttenuateFn fn;
    const char *family;
};

// general kernel: every shape x arithmetic x exponential x geometry
template <int LPT, int NCHUNK, bool GEOM>
static AttenuateFn pick_general_modes(int math, int expm)
{
#define SMK_PICK(M, E) \
    if (math == M && expm == E) return attenuate_tracks<LPT, NCHUNK, M, E, GEOM>;
    SMK_PICK(kMathFast, kExpPoly)
    SMK_PICK(kMathFast, kExpPolyWide)
    SMK_PICK(kMathFast, kExpMufu)
    SMK_PICK(kMathFast, kExpGlibc)
    SMK_PICK(kMathFast, kExpTable)
    SMK_PICK(kMathStrict, kExpPoly)      // the scalar polynomial is range-safe by itself
    SMK_PICK(kMathStrict, kExpMufu)
    SMK_PICK(kMathStrict, kExpGlibc)
    SMK_PICK(kMathStrict, kExpTable)
#undef SMK_PICK
    return nullptr;
}

template <bool GEOM>
static AttenuateFn pick_general(const Shape &s, int math, int expm)
{
    if (math == kMathStrict && expm == kExpPolyWide) expm = kExpPoly;
    if (s.nchunk == 1) {
        switch (s.lpt) {
            case 1: return pick_general_modes<1, 1, GEOM>(math, expm);
            case 2: return pick_general_modes<2, 1, GEOM>(math, expm);
            case 4: return pick_general_modes<4, 1, GEOM>(math, expm);
            case 8: return pick_general_modes<8, 1, GEOM>(math, expm);
            case 16: return pick_general_modes<16, 1, GEOM>(math, expm);
            case 32: return pick_general_modes<32, 1, GEOM>(math, expm);
        }
    } else if (s.lpt == 32 && s.nchunk == 2) {
        return pick_general_modes<32, 2, GEOM>(math, expm);
    }
    return nullptr;
}

// one track per warp, FAST arithmetic: 4 groups per lane (65..128 groups) or 2 (33..64)
template <int GPL, bool F64, bool GEOM, bool A32>
static AttenuateFn pick_warp_track_exp(int expm)
{
    switch (expm) {
        case kExpPoly: return attenuate_warp_track<GPL, kExpPoly, F64, GEOM, A32>;
        case kExpPolyWide: return attenuate_warp_track<GPL, kExpPolyWide, F64, GEOM, A32>;
        case kExpMufu: return attenuate_warp_track<GPL, kExpMufu, F64, GEOM, A32>;
        case kExpGlibc: return attenuate_warp_track<GPL, kExpGlibc, F64, GEOM, A32>;
        case kExpTable: return attenuate_warp_track<GPL, kExpTable, F64, GEOM, A32>;
    }
    return nullptr;
}

// a32: every array < 4 GB -> 32-bit byte offsets (only the f32-tally kernels have that form: the f64 tallies are a
// diagnostic)
template <int GPL>
static AttenuateFn pick_warp_track(int expm, bool f64, bool geom, bool a32)
{
    if (f64) return geom ? pick_warp_track_exp<GPL, true, true, false>(expm) : pick_warp_track_exp<GPL, true, false, false>(expm);
    if (a32) return geom ? pick_warp_track_exp<GPL, false, true, true>(expm) : pick_warp_track_exp<GPL, false, false, true>(expm);
    return geom ? pick_warp_track_exp<GPL, false, true, false>(expm) : pick_warp_track_exp<GPL, false, false, false>(expm);
}

// sub-warp tracks from gather records (<= 32 groups, FAST); lpt = lanes per track = G_pad / gpl
template <int LPT, int GPL, bool F64, bool GEOM>
static AttenuateFn pick_record_exp(int expm)
{
    switch (expm) {
        case kExpPoly: return attenuate_record_tracks<LPT, GPL, kExpPoly, F64, GEOM>;
        case kExpPolyWide: return attenuate_record_tracks<LPT, GPL, kExpPolyWide, F64, GEOM>;
        case kExpMufu: return attenuate_record_tracks<LPT, GPL, kExpMufu, F64, GEOM>;
        case kExpGlibc: return attenuate_record_tracks<LPT, GPL, kExpGlibc, F64, GEOM>;
        case kExpTable: return attenuate_record_tracks<LPT, GPL, kExpTable, F64, GEOM>;
    }
    return nullptr;
}

// SMK_FLAG_FIT_PER_SWEEP: records holding the fitted coefficients (the library's shapes, constant geometry)
template <int LPT, int GPL, bool F64>
static AttenuateFn pick_record_hoist(int expm)
{
    switch (expm) {
        case kExpPoly: return attenuate_record_tracks<LPT, GPL, kExpPoly, F64, false, true>;
        case kExpPolyWide: return attenuate_record_tracks<LPT, GPL, kExpPolyWide, F64, false, true>;
        case kExpMufu: return attenuate_record_tracks<LPT, GPL, kExpMufu, F64, false, true>;
    }
    return nullptr;
}

static AttenuateFn pick_record_hoist(int groups_pad, int gpl, int expm, bool f64)
{
    switch (groups_pad * 10 + gpl) {
        case 42: return f64 ? pick_record_hoist<2, 2, true>(expm) : pick_record_hoist<2, 2, false>(expm);
        case 84: return f64 ? pick_record_hoist<2, 4, true>(expm) : pick_record_hoist<2, 4, false>(expm);
        case 164: return f64 ? pick_record_hoist<4, 4, true>(expm) : pick_record_hoist<4, 4, false>(expm);
        case 324: return f64 ? pick_record_hoist<8, 4, true>(expm) : pick_record_hoist<8, 4, false>(expm);
    }
    return nullptr;
}

static AttenuateFn pick_warp_track_fit(int expm)
{
    switch (expm) {
        case kExpPoly: return attenuate_warp_track_fit<kExpPoly>;
        case kExpPolyWide: return attenuate_warp_track_fit<kExpPolyWide>;
        case kExpMufu: return attenuate_warp_track_fit<kExpMufu>;
    }
    return nullptr;
}

static AttenuateFn pick_warp_track_rec_hoist(int expm)
{
    switch (expm) {
        case kExpPoly: return attenuate_warp_track_rec<kExpPoly, false, true>;
        case kExpPolyWide: return attenuate_warp_track_rec<kExpPolyWide, false, true>;
        case kExpMufu: return attenuate_warp_track_rec<kExpMufu, false, true>;
    }
    return nullptr;
}

// ALL = every exponential (the shapes the library uses, record_groups_per_lane); otherwise POLY only: the other record
// shape of each group count exists for the A/B knob SMK_RECORDS and the cross-check test
template <int LPT, int GPL, bool F64, bool GEOM, bool ALL>
static AttenuateFn pick_record_modes(int expm)
{
    if constexpr (ALL) return pick_record_exp<LPT, GPL, F64, GEOM>(expm);
    if (expm == kExpPoly) return attenuate_record_tracks<LPT, GPL, kExpPoly, F64, GEOM>;
    if (expm == kExpPolyWide) return attenuate_record_tracks<LPT, GPL, kExpPolyWide, F64, GEOM>;
    return nullptr;
}

template <int LPT, int GPL, bool ALL>
static AttenuateFn pick_record_flags(int expm, bool f64, bool geom)
{
    if (f64) return geom ? pick_record_modes<LPT, GPL, true, true, ALL>(expm) : pick_record_modes<LPT, GPL, true, false, ALL>(expm);
    return geom ? pick_record_modes<LPT, GPL, false, true, ALL>(expm) : pick_record_modes<LPT, GPL, false, false, ALL>(expm);
}

static AttenuateFn pick_record(int groups_pad, int gpl, int expm, bool f64, bool geom)
{
    const int key = groups_pad * 10 + gpl;
    switch (key) {
        case 42: return pick_record_flags<2, 2, true>(expm, f64, geom);      // 1..4 groups
        case 44: return pick_record_flags<1, 4, false>(expm, f64, geom);
        case 84: return pick_record_flags<2, 4, true>(expm, f64, geom);      // 5..8 groups (config 3)
        case 82: return pick_record_flags<4, 2, true>(expm, f64, geom);       // the library's shape with per-segment geometry
        case 164: return pick_record_flags<4, 4, true>(expm, f64, geom);     // 9..16 groups
        case 162: return pick_record_flags<8, 2, false>(expm, f64, geom);
        case 324: return pick_record_flags<8, 4, true>(expm, f64, geom);     // 17..32 groups
        case 322: return pick_record_flags<16, 2, false>(expm, f64, geom);
    }
    return nullptr;
}

// 33..64 groups: one track per warp from gather records (f32 tallies)
template <bool GEOM>
static AttenuateFn pick_warp_track_rec(int expm)
{
    switch (expm) {
        case kExpPoly: return attenuate_warp_track_rec<kExpPoly, GEOM>;
        case kExpPolyWide: return attenuate_warp_track_rec<kExpPolyWide, GEOM>;
        case kExpMufu: return attenuate_warp_track_rec<kExpMufu, GEOM>;
        case kExpGlibc: return attenuate_warp_track_rec<kExpGlibc, GEOM>;
        case kExpTable: return attenuate_warp_track_rec<kExpTable, GEOM>;
    }
    return nullptr;
}

// expm is the internal mode (kExpPolyWide resolved by the caller)
static KernelChoice choose_kernel(const Shape &s, int math, int expm, bool f64, bool geom, bool a32)
{
    if (math == kMathFast && s.nchunk == 1 && s.lpt == 32) return {pick_warp_track<4>(expm, f64, geom, a32), "attenuate_warp_track<4 groups/lane"};
    if (math == kMathFast && s.nchunk == 1 && s.lpt == 16) return {pick_warp_track<2>(expm, f64, geom, a32), "attenuate_warp_track<2 groups/lane"};
    return {geom ? pick_general<true>(s, math, expm) : pick_general<false>(s, math, expm), "attenuate_tracks<general"};
}

}  // namespace smk

#ifdef SMK_TUNING
#include "smk_kernels_tuning.cuh"
#endif

using namespace smk;

struct smk_ctx {
    smk_params p;
    Shape shape;
    smk_geometry geom;       // base + spread (reference constants unless SMK_FLAG_SEGMENT_GEOMETRY)
    AttenuateFn kernel;      // instantiation of the last selection
    AttenuateFn tuning_kernel;  // -DSMK_TUNING builds: SMK_KERNEL=<variant> override (nullptr otherwise)
    int exp_internal;        // kExp* of `kernel`
    char kernel_name[160];
    size_t dyn_smem;
    int sm_count;
    int blocks_per_sm;
    int64_t rows;            // R * F
    int replicas;            // copies of the tally array (contention relief for few-row problems)
    int64_t n_tracks;
    float *d_source, *d_sigT, *d_flux0, *d_tally;
    float *d_stage;          // unpadded staging, R*F*G floats
    float *d_psi;            // [tracks of last run][G_pad] (SMK_FLAG_KEEP_PSI)
    int64_t psi_capacity;    // in tracks
    int64_t last_begin, last_end;
    unsigned long long *d_checksum;
    unsigned long long *d_work;  // dynamic track scheduling counter
    unsigned int *d_max_bits;    // smk_scan_sigt_max scratch
    double *d_tally64;           // SMK_FLAG_TALLY_F64: f64 tally accumulators, [R][F][G_pad]
    float *d_records;            // gather records of attenuate_record_tracks, [R*F][G_pad/2][8]; nullptr = not used
    int rec_gpl;                 // groups per lane of the record kernel (2 or 4)
    bool hoist;                  // the selected kernel reads fitted records (SMK_FLAG_FIT_PER_SWEEP)
    float sigt_max;              // max(sigT) of the device data, +inf when unknown (partial uploads)
    cudaStream_t stream;
    bool own_stream;
    cudaEvent_t ev0, ev1;
    cudaEvent_t ev_finalized;    // recorded after the flux0 + tallies pass of the last download (smk_wait_finalized)
    bool finalized_recorded;
    bool have_data;
    int64_t launches;
};

#ifdef SMK_TUNING
// SMK_KERNEL=<variant>: measured alternatives for the 65..128-group FAST shape (smk_kernels_tuning.cuh)
static int smk_tuning_select(smk_ctx *c)
{
    const char *variant = getenv("SMK_KERNEL");
    if (!variant || !*variant || strcmp(variant, "default") == 0) return SMK_OK;
    const bool shape_ok = c->shape.lpt == 32 && c->shape.nchunk == 1 && c->p.math_mode == kMathFast &&
                          !(c->p.flags & (SMK_FLAG_TALLY_F64 | SMK_FLAG_SEGMENT_GEOMETRY)) &&
                          (c->p.exp_mode == kExpPoly || c->p.exp_mode == kExpMufu);
    if (!shape_ok) return SMK_OK;     // variants exist for that shape only; everything else runs the product kernel
    AttenuateFn fn = nullptr;
    if (strcmp(variant, "oldflat") == 0) fn = pick_pf_exp<false, false, false>(c->p.exp_mode);
    else if (strcmp(variant, "prefetch") == 0) fn = pick_pf_exp<true, false, false>(c->p.exp_mode);
    else if (strcmp(variant, "defer") == 0) fn = pick_pf_exp<false, true, false>(c->p.exp_mode);
    else if (strcmp(variant, "l1pf") == 0) fn = pick_pf_exp<false, false, true>(c->p.exp_mode);
    else if (strcmp(variant, "pipe") == 0) fn = pick_pipe_exp(c->p.exp_mode);
    else if (strcmp(variant, "staged2") == 0 || strcmp(variant, "staged3") == 0) {
        const int stages = variant[6] - '0';
        fn = stages == 2 ? pick_staged_exp<1, 2>(c->p.exp_mode) : pick_staged_exp<1, 3>(c->p.exp_mode);
        c->dyn_smem = (size_t)(kThreadsPerBlock / 32) * stages * 4 * c->shape.groups_pad * sizeof(float);
    } else {
        return fail(SMK_EINVAL, "unknown SMK_KERNEL variant '%s'", variant);
    }
    c->tuning_kernel = fn;
    return SMK_OK;
}
#endif

static int validate(const smk_params *p, Shape &shape)
{
    if (!p) return fail(SMK_EINVAL, "params is NULL");
    if (p->source_3D_regions < 1) return fail(SMK_EINVAL, "source_3D_regions must be >= 1");
    if (p->fine_axial_intervals < 2)
        return fail(SMK_EINVAL, "fine_axial_intervals must be >= 2 (kernel.c:111-137 reads row FAI+1 / FAI-1)");
    if (p->egroups < 1) return fail(SMK_EINVAL, "egroups must be >= 1");
    if (p->seg_per_track < 1) return fail(SMK_EINVAL, "seg_per_track must be >= 1");
    if (p->segments < 0) return fail(SMK_EINVAL, "segments must be >= 0");
    if (p->exp_mode < SMK_EXP_POLY || p->exp_mode > SMK_EXP_TABLE)
        return fail(SMK_EINVAL, "unknown exp_mode %d", p->exp_mode);
    if (p->math_mode != SMK_MATH_FAST && p->math_mode != SMK_MATH_STRICT)
        return fail(SMK_EINVAL, "unknown math_mode %d", p->math_mode);
    if (p->flags & ~(SMK_FLAG_KEEP_PSI | SMK_FLAG_TALLY_F64 | SMK_FLAG_SEGMENT_GEOMETRY | SMK_FLAG_FIT_PER_SWEEP))
        return fail(SMK_EINVAL, "unknown flag bits %#x", p->flags);
    if ((p->flags & SMK_FLAG_FIT_PER_SWEEP) && (p->math_mode != SMK_MATH_FAST || (p->flags & SMK_FLAG_SEGMENT_GEOMETRY)))
        return fail(SMK_EINVAL, "SMK_FLAG_FIT_PER_SWEEP needs SMK_MATH_FAST and the constant geometry (the fit is only "
                                "sweep-invariant there; STRICT is the reference's own per-segment arithmetic)");
    if (!shape_for(p->egroups, shape)) return fail(SMK_EINVAL, "egroups = %d unsupported", p->egroups);
    if ((int64_t)p->source_3D_regions * p->fine_axial_intervals * (shape.groups_pad / 4) >= (1ll << 31))
        return fail(SMK_EINVAL, "regions * intervals * padded groups / 4 must be < 2^31 (32-bit row offsets)");
    if (p->source_3D_regions >= (1 << 25))
        return fail(SMK_EINVAL, "regions must be < 2^25 (sigT row index + two flag bits in one word)");
    if ((int64_t)p->source_3D_regions * p->fine_axial_intervals >= (1ll << 30))
        return fail(SMK_EINVAL, "regions * intervals must be < 2^30 (row index + two flag bits in one word)");
    return SMK_OK;
}

static const smk_geometry kReferenceGeometry = {Geometry::dz, Geometry::zin, Geometry::weight, Geometry::mu,
                                                Geometry::mu2, Geometry::ds, 0.0f};

static int check_geometry(const smk_geometry *g)
{
    if (!g) return fail(SMK_EINVAL, "geometry is NULL");
    if (!(g->dz > 0.0f) || !(g->ds > 0.0f) || !(g->spread >= 0.0f && g->spread < 1.0f) || !std::isfinite(g->dz) ||
        !std::isfinite(g->zin) || !std::isfinite(g->weight) || !std::isfinite(g->mu) || !std::isfinite(g->mu2) ||
        !std::isfinite(g->ds))
        return fail(SMK_EINVAL, "geometry needs finite values, dz > 0, ds > 0 and 0 <= spread < 1");
    return SMK_OK;
}

// Resolve the kernel instantiation for the context's current state (exp form depends on the data's
// max(sigT) and on the geometry's largest ds) and its occupancy.
static int select_kernel(smk_ctx *c)
{
    const bool geom = (c->p.flags & SMK_FLAG_SEGMENT_GEOMETRY) != 0;
    const bool f64 = (c->p.flags & SMK_FLAG_TALLY_F64) != 0;
    int expm = c->p.exp_mode;
    const char *exp_name[] = {"poly", "mufu", "glibc", "table", "poly+mufu beyond tau 0.7"};
    if (expm == kExpPoly) {
        const float ds_max = geom ? c->geom.ds * (1.0f + c->geom.spread) : c->geom.ds;
        // tau = sigT * ds <= bound; `!(<=)` also catches an unknown (+inf) or NaN bound
        if (!(c->sigt_max * ds_max <= kPolyMaxTau)) expm = kExpPolyWide;
    }
    AttenuateFn fn = c->tuning_kernel;
    const char *family = "tuning variant";
    // 32-bit byte offsets reach every element when the largest array (R * F padded rows) is below 4 GB
    // (SMK_ADDR64=1 forces the plain form: tests of the > 4 GB path on small data)
    const char *force64 = getenv("SMK_ADDR64");
    const bool a32 = (uint64_t)c->rows * c->shape.groups_pad * sizeof(float) < (1ull << 32) && !(force64 && force64[0] == '1');
    bool warp_track32 = false;
    c->hoist = false;
    const bool want_hoist = (c->p.flags & SMK_FLAG_FIT_PER_SWEEP) != 0;
    if (!fn && c->d_records && c->shape.groups_pad <= 32) {
        if (want_hoist) fn = pick_record_hoist(c->shape.groups_pad, c->rec_gpl, expm, f64);
        c->hoist = fn != nullptr;
        if (!fn) fn = pick_record(c->shape.groups_pad, c->rec_gpl, expm, f64, geom);
        family = c->rec_gpl == 2 ? "attenuate_record_tracks<2 groups/lane" : "attenuate_record_tracks<4 groups/lane";
    } else if (!fn && c->d_records && c->shape.groups_pad == 128) {
        // (allocated only with SMK_FLAG_FIT_PER_SWEEP: the fitted rows of attenuate_warp_track_fit)
        fn = pick_warp_track_fit(expm);
        c->hoist = fn != nullptr;
        family = "attenuate_warp_track_fit<4 groups/lane";
    } else if (!fn && c->d_records) {
        if (want_hoist) fn = pick_warp_track_rec_hoist(expm);
        c->hoist = fn != nullptr;
        if (!fn) fn = geom ? pick_warp_track_rec<true>(expm) : pick_warp_track_rec<false>(expm);
        family = "attenuate_warp_track_rec<2 groups/lane";
    }
    if (!fn) {
        const KernelChoice k = choose_kernel(c->shape, c->p.math_mode, expm, f64, geom, a32);
        warp_track32 = a32 && !f64 && strncmp(k.family, "attenuate_warp_track", 20) == 0;
        fn = k.fn;
        family = k.family;
    }
    if (!fn) return fail(SMK_EINVAL, "no kernel for egroups=%d math=%d exp=%d", c->p.egroups, c->p.math_mode, expm);
    if (fn != c->kernel) {
        c->kernel = fn;
        if (c->dyn_smem > 0)
            SMK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->dyn_smem));
        SMK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->blocks_per_sm, fn, kThreadsPerBlock, c->dyn_smem));
        if (c->blocks_per_sm < 1) c->blocks_per_sm = 1;
    }
    c->exp_internal = expm;
    snprintf(c->kernel_name, sizeof c->kernel_name, "%s, %s, %s, %s tally, %s geometry%s%s>", family,
             c->p.math_mode == kMathStrict ? "strict" : "fast", exp_name[expm], f64 ? "f64" : "f32",
             geom ? "per-segment" : "const", warp_track32 ? ", 32-bit offsets" : "", c->hoist ? ", fit per sweep" : "");
    return SMK_OK;
}

extern "C" {

int smk_abi_version(void) { return SMK_ABI_VERSION; }

const char *smk_last_error(void) { return t_error; }

int smk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int smk_device_name(int device, char *buf, size_t buflen)
{
    if (!buf || buflen == 0) return fail(SMK_EINVAL, "buf is NULL");
    cudaDeviceProp prop;
    SMK_CUDA(cudaGetDeviceProperties(&prop, device));
    snprintf(buf, buflen, "%s", prop.name);
    return SMK_OK;
}

int smk_padded_groups(int egroups)
{
    Shape s;
    return shape_for(egroups, s) ? s.groups_pad : SMK_EINVAL;
}

int64_t smk_num_tracks(int64_t segments, int seg_per_track)
{
    if (segments < 0 || seg_per_track < 1) return SMK_EINVAL;
    return (segments + seg_per_track - 1) / seg_per_track;
}

int smk_create(const smk_params *p, smk_ctx **out)
{
    if (!out) return fail(SMK_EINVAL, "out is NULL");
    *out = nullptr;
    Shape shape;
    int rc = validate(p, shape);
    if (rc != SMK_OK) return rc;
    int ndev = 0;
    SMK_CUDA(cudaGetDeviceCount(&ndev));
    if (p->device < 0 || p->device >= ndev)
        return fail(SMK_EINVAL, "device %d out of range (%d visible)", p->device, ndev);
    SMK_CUDA(cudaSetDevice(p->device));

    smk_ctx *c = new (std::nothrow) smk_ctx();
    if (!c) return fail(SMK_ENOMEM, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->p = *p;
    c->shape = shape;
    c->geom = kReferenceGeometry;
    if (p->flags & SMK_FLAG_SEGMENT_GEOMETRY) c->geom.spread = 0.25f;
    c->sigt_max = INFINITY;           // nothing uploaded yet
    c->rows = (int64_t)p->source_3D_regions * p->fine_axial_intervals;
    // few tally rows => L2 atomics on the same addresses serialise: spread them over replicas so that
    // at least ~4096 rows are in play (1 for the reference's default 33750 rows)
    c->replicas = 1;
    if (c->rows < 4096 && !(p->flags & SMK_FLAG_TALLY_F64)) {
        c->replicas = (int)((4096 + c->rows - 1) / c->rows);
        if (c->replicas > 32) c->replicas = 32;
    }
    c->n_tracks = smk_num_tracks(p->segments, p->seg_per_track);

    // every failure from here on goes through the single cleanup path (smk_destroy)
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) { c->own_stream = true; e = cudaEventCreate(&c->ev0); }
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_finalized, cudaEventDisableTiming);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, p->device);
    if (e == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (e == cudaSuccess) {
        ExpTable tab;
        build_exp_table(tab);
        e = cudaMemcpyToSymbol(c_exp_table, &tab, sizeof(tab));
    }
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_exp2f_tab, h_exp2f_tab, sizeof(h_exp2f_tab));

    const size_t slab = (size_t)c->rows * shape.groups_pad * sizeof(float);
    const size_t sig = (size_t)p->source_3D_regions * shape.groups_pad * sizeof(float);
    const size_t stage = (size_t)c->rows * p->egroups * sizeof(float);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_source, slab);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_flux0, slab);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_tally, slab * c->replicas);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_sigT, sig);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_stage, stage);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_checksum, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_work, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_max_bits, sizeof(unsigned int));
    if (e == cudaSuccess && (p->flags & SMK_FLAG_TALLY_F64)) {
        e = cudaMalloc(&c->d_tally64, 2 * slab);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->d_tally64, 0, 2 * slab, c->stream);
    }
    // <= 32 groups, FAST: the sweep reads gather records (attenuate_record_tracks), 4 x the
    // source array, rebuilt from the canonical rows at every launch.  SMK_RECORDS=0 keeps the general kernel
    // (A/B and the cross-check test), =2 / =4 selects the groups per lane.
    {
        const int gpl_default = record_groups_per_lane(shape.groups_pad, (p->flags & SMK_FLAG_SEGMENT_GEOMETRY) != 0);
        int gpl = gpl_default;
        if (const char *r = getenv("SMK_RECORDS")) gpl = atoi(r);
        // Records trade bytes for instructions (one 32-byte record per lane instead of 2.6 + 1 partial rows: 11 % more
        // bytes per segment): a win while records + tallies live in the L2, a loss once the sweep is HBM-bound.
        // (an explicit SMK_RECORDS=2|4 / SMK_WT_RECORDS=1 overrides this rule: measurements)
        const char *wt = getenv("SMK_WT_RECORDS");
        const bool forced = (getenv("SMK_RECORDS") && (gpl == 2 || gpl == 4)) || (wt && wt[0] == '1');
        const bool l2_resident = forced || (e == cudaSuccess && (double)slab * 5.0 <= 0.75 * (double)prop.l2CacheSize);
        const bool eligible = shape.nchunk == 1 && shape.groups_pad <= 32 && p->math_mode == kMathFast && l2_resident &&
                              slab * 4 < (1ull << 32) &&
                              c->rows * (shape.groups_pad / 2) < (1ll << 30);
        // the non-default record shape is compiled for the POLY exponential only
        if ((gpl == 2 || gpl == 4) && gpl != gpl_default && p->exp_mode != SMK_EXP_POLY) gpl = gpl_default;
        // the record kernels address with 32-bit byte offsets only: SMK_ADDR64=1 (plain 64-bit addressing forced, a test
        // knob) keeps the row-array kernels
        const char *force64 = getenv("SMK_ADDR64");
        const bool plain64 = force64 && force64[0] == '1';
        if (eligible && !plain64 && (gpl == 2 || gpl == 4)) {
            c->rec_gpl = gpl;
            if (e == cudaSuccess) e = cudaMalloc(&c->d_records, slab * 4);
        }
        // 33..64 groups, one track per warp: the same records feed attenuate_warp_track_rec (SMK_WT_RECORDS=0 keeps the
        // row arrays: A/B and the cross-check test)
        const bool wt_eligible = shape.nchunk == 1 && shape.lpt == 16 && p->math_mode == kMathFast && l2_resident &&
                                 !(p->flags & SMK_FLAG_TALLY_F64) && slab * 4 < (1ull << 32) &&
                                 c->rows * 32 < (1ll << 30);
        if (wt_eligible && !plain64 && !(wt && wt[0] == '0')) {
            c->rec_gpl = 2;
            if (e == cudaSuccess) e = cudaMalloc(&c->d_records, slab * 4);
        }
        // 65..128 groups with SMK_FLAG_FIT_PER_SWEEP: three fitted rows per source row (attenuate_warp_track_fit), while
        // they and the tallies fit in 3/4 of the L2
        const bool fit_eligible = (p->flags & SMK_FLAG_FIT_PER_SWEEP) && shape.nchunk == 1 && shape.lpt == 32 &&
                                  !(p->flags & SMK_FLAG_TALLY_F64) && slab * 3 < (1ull << 32) && !plain64 &&
                                  e == cudaSuccess && (double)slab * 4.25 <= 0.75 * (double)prop.l2CacheSize;
        if (fit_eligible) {
            c->rec_gpl = 4;
            e = cudaMalloc(&c->d_records, slab * 3);
        }
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_tally, 0, slab * c->replicas, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_flux0, 0, slab, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_checksum, 0, sizeof(unsigned long long), c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        int code = (e == cudaErrorMemoryAllocation) ? SMK_ENOMEM : SMK_ECUDA;
        fail(code, "smk_create: %s", cudaGetErrorString(e));
        cudaGetLastError();
        smk_destroy(c);
        return code;
    }
#ifdef SMK_TUNING
    rc = smk_tuning_select(c);        // SMK_KERNEL=<variant>; tuning builds only
    if (rc != SMK_OK) { smk_destroy(c); return rc; }
#endif
    if (const char *r = getenv("SMK_TALLY_REPLICAS")) {       // tuning knob
        const int v = atoi(r);
        if (v >= 1 && v <= c->replicas) c->replicas = v;
    }
    rc = select_kernel(c);
    if (rc != SMK_OK) { smk_destroy(c); return rc; }
    *out = c;
    return SMK_OK;
}

void smk_destroy(smk_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->p.device);
    cudaFree(c->d_source);
    cudaFree(c->d_flux0);
    cudaFree(c->d_tally);
    cudaFree(c->d_sigT);
    cudaFree(c->d_stage);
    cudaFree(c->d_psi);
    cudaFree(c->d_checksum);
    cudaFree(c->d_work);
    cudaFree(c->d_max_bits);
    cudaFree(c->d_tally64);
    cudaFree(c->d_records);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_finalized) cudaEventDestroy(c->ev_finalized);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int smk_set_stream(smk_ctx *c, void *cuda_stream)
{
    if (!c) return fail(SMK_EINVAL, "ctx is NULL");
    SMK_CUDA(cudaSetDevice(c->p.device));
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return SMK_OK;
}

int smk_set_geometry(smk_ctx *c, const smk_geometry *g)
{
    if (!c) return fail(SMK_EINVAL, "ctx is NULL");
    if (!(c->p.flags & SMK_FLAG_SEGMENT_GEOMETRY))
        return fail(SMK_ESTATE, "context created without SMK_FLAG_SEGMENT_GEOMETRY: the geometry is kernel.c:99-104");
    int rc = check_geometry(g);
    if (rc != SMK_OK) return rc;
    c->geom = *g;
    return select_kernel(c);
}

int smk_get_geometry(const smk_ctx *c, smk_geometry *g)
{
    if (!c || !g) return fail(SMK_EINVAL, "NULL argument");
    *g = c->geom;
    return SMK_OK;
}

const char *smk_kernel_name(smk_ctx *c)
{
    if (!c || select_kernel(c) != SMK_OK) return "";
    return c->kernel_name;
}

static int layout_grid(int64_t n)
{
    int64_t b = (n + 255) / 256;
    return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

// host unpadded rows -> device padded rows [row_begin, row_begin + rows), through the unpadded
// staging buffer when G != G_pad; enqueue only
static int upload_rows(smk_ctx *c, const float *h, float *d, int64_t row_begin, int64_t rows, float pad)
{
    const int G = c->p.egroups, Gp = c->shape.groups_pad;
    if (rows <= 0) return SMK_OK;
    if (G == Gp) {
        SMK_CUDA(cudaMemcpyAsync(d + row_begin * Gp, h, (size_t)rows * G * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    } else {
        SMK_CUDA(cudaMemcpyAsync(c->d_stage, h, (size_t)rows * G * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        pad_rows<<<layout_grid(rows * Gp), 256, 0, c->stream>>>(c->d_stage, d + row_begin * Gp, rows, G, Gp, pad);
        SMK_CUDA(cudaGetLastError());
        c->launches += 1;
    }
    return SMK_OK;
}

// max of a host array; NaN or negative values make the bound unknown (+inf)
static float host_max(const float *h, int64_t n)
{
    float m = 0.0f;
    bool bad = false;
    for (int64_t i = 0; i < n; ++i) {
        const float v = h[i];
        bad |= !(v >= 0.0f);
        m = v > m ? v : m;
    }
    return bad ? INFINITY : m;
}

int smk_reset_tallies(smk_ctx *c)
{
    if (!c) return fail(SMK_EINVAL, "ctx is NULL");
    SMK_CUDA(cudaSetDevice(c->p.device));
    SMK_CUDA(cudaMemsetAsync(c->d_tally, 0, (size_t)c->rows * c->shape.groups_pad * sizeof(float) * c->replicas,
                             c->stream));
    SMK_CUDA(cudaMemsetAsync(c->d_checksum, 0, sizeof(unsigned long long), c->stream));
    if (c->d_tally64)
        SMK_CUDA(cudaMemsetAsync(c->d_tally64, 0, (size_t)c->rows * c->shape.groups_pad * sizeof(double), c->stream));
    return SMK_OK;
}

int smk_upload_async(smk_ctx *c, const float *fine_source, const float *fine_flux, const float *sigT)
{
    if (!c) return fail(SMK_EINVAL, "ctx is NULL");
    if (!fine_source || !sigT) return fail(SMK_EINVAL, "fine_source and sigT are required");
    SMK_CUDA(cudaSetDevice(c->p.device));
    int rc;
    if ((rc = upload_rows(c, fine_source, c->d_source, 0, c->rows, 0.0f)) != SMK_OK) return rc;
    if ((rc = upload_rows(c, sigT, c->d_sigT, 0, c->p.source_3D_regions, 1.0f)) != SMK_OK) return rc;
    if (fine_flux) {
        if ((rc = upload_rows(c, fine_flux, c->d_flux0, 0, c->rows, 0.0f)) != SMK_OK) return rc;
    } else {
        SMK_CUDA(cudaMemsetAsync(c->d_flux0, 0, (size_t)c->rows * c->shape.groups_pad * sizeof(float), c->stream));
    }
    if ((rc = smk_reset_tallies(c)) != SMK_OK) return rc;
    // while the copies are in flight: bound tau = sigT * ds for the choice of the exponential's form
    c->sigt_max = host_max(sigT, (int64_t)c->p.source_3D_regions * c->p.egroups);
    c->have_data = true;
    return SMK_OK;
}

int smk_upload(smk_ctx *c, const float *fine_source, const float *fine_flux, const float *sigT)
{
    int rc = smk_upload_async(c, fine_source, fine_flux, sigT);
    if (rc != SMK_OK) return rc;
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    return SMK_OK;
}

int smk_upload_rows_async(smk_ctx *c, int array, int64_t row_begin, int64_t rows, const float *host)
{
    if (!c || !host) return fail(SMK_EINVAL, "NULL argument");
    const int64_t total = (array == SMK_ARRAY_SIGT) ? c->p.source_3D_regions : c->rows;
    if (array < SMK_ARRAY_SOURCE || array > SMK_ARRAY_SIGT) return fail(SMK_EINVAL, "unknown array %d", array);
    if (row_begin < 0 || rows < 0 || row_begin + rows > total)
        return fail(SMK_EINVAL, "rows [%lld, %lld) outside [0, %lld)", (long long)row_begin, (long long)(row_begin + rows),
                    (long long)total);
    SMK_CUDA(cudaSetDevice(c->p.device));
    float *d = array == SMK_ARRAY_SOURCE ? c->d_source : array == SMK_ARRAY_FLUX ? c->d_flux0 : c->d_sigT;
    int rc = upload_rows(c, host, d, row_begin, rows, array == SMK_ARRAY_SIGT ? 1.0f : 0.0f);
    if (rc != SMK_OK) return rc;
    if (array == SMK_ARRAY_SIGT) {
        // a full-array upload keeps the bound exact; a partial one leaves the other rows unknown
        c->sigt_max = (rows == total) ? host_max(host, rows * c->p.egroups) : INFINITY;
    }
    c->have_data = true;
    return SMK_OK;
}

int smk_scan_sigt_max(smk_ctx *c, float *max_out)
{
    if (!c) return fail(SMK_EINVAL, "ctx is NULL");
    SMK_CUDA(cudaSetDevice(c->p.device));
    const int64_t n = (int64_t)c->p.source_3D_regions * c->shape.groups_pad;
    SMK_CUDA(cudaMemsetAsync(c->d_max_bits, 0, sizeof(unsigned int), c->stream));
    max_rows<<<layout_grid(n), 256, 0, c->stream>>>(c->d_sigT, n, c->d_max_bits);
    SMK_CUDA(cudaGetLastError());
    c->launches += 1;
    unsigned int bits = 0;
    SMK_CUDA(cudaMemcpyAsync(&bits, c->d_max_bits, sizeof bits, cudaMemcpyDeviceToHost, c->stream));
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(&c->sigt_max, &bits, sizeof bits);        // padding groups hold 1.0: the bound is >= 1 when G != G_pad
    if (max_out) *max_out = c->sigt_max;
    return SMK_OK;
}

int smk_set_sigt_bound(smk_ctx *c, float bound)
{
    if (!c) return fail(SMK_EINVAL, "ctx is NULL");
    if (!(bound >= 0.0f)) return fail(SMK_EINVAL, "bound must be >= 0 (or +inf)");
    c->sigt_max = bound;
    return SMK_OK;
}

int smk_fill_device(smk_ctx *c, float sigt_floor)
{
    if (!c) return fail(SMK_EINVAL, "ctx is NULL");
    if (!(sigt_floor >= 0.0f && sigt_floor < 1.0f)) return fail(SMK_EINVAL, "sigt_floor must be in [0, 1)");
    SMK_CUDA(cudaSetDevice(c->p.device));
    const int G = c->p.egroups, Gp = c->shape.groups_pad;
    const int64_t R = c->p.source_3D_regions;
    fill_rows<<<layout_grid(c->rows * Gp), 256, 0, c->stream>>>(c->d_source, c->rows, G, Gp, 0u, c->p.seed, 0.0f, 0.0f);
    fill_rows<<<layout_grid(c->rows * Gp), 256, 0, c->stream>>>(c->d_flux0, c->rows, G, Gp, 1u, c->p.seed, 0.0f, 0.0f);
    fill_rows<<<layout_grid(R * Gp), 256, 0, c->stream>>>(c->d_sigT, R, G, Gp, 2u, c->p.seed, sigt_floor, 1.0f);
    SMK_CUDA(cudaGetLastError());
    c->launches += 3;
    int rc = smk_reset_tallies(c);
    if (rc != SMK_OK) return rc;
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    c->sigt_max = 1.0f;              // u01() < 1, padding = 1
    c->have_data = true;
    return SMK_OK;
}

static int launch(smk_ctx *c, int64_t track_begin, int64_t track_end)
{
    if (!c) return fail(SMK_EINVAL, "ctx is NULL");
    if (!c->have_data) return fail(SMK_ESTATE, "no source data: call smk_upload or smk_fill_device first");
    if (track_begin < 0 || track_end > c->n_tracks || track_begin > track_end)
        return fail(SMK_EINVAL, "track range [%lld, %lld) outside [0, %lld)", (long long)track_begin,
                    (long long)track_end, (long long)c->n_tracks);
    SMK_CUDA(cudaSetDevice(c->p.device));
    int rc = select_kernel(c);
    if (rc != SMK_OK) return rc;
    c->last_begin = track_begin;
    c->last_end = track_end;
    const int64_t tracks = track_end - track_begin;
    if (tracks == 0) return SMK_OK;

    if (c->p.flags & SMK_FLAG_KEEP_PSI) {
        if (tracks > c->psi_capacity) {
            cudaFree(c->d_psi);
            c->d_psi = nullptr;
            c->psi_capacity = 0;
            cudaError_t e = cudaMalloc(&c->d_psi, (size_t)tracks * c->shape.groups_pad * sizeof(float));
            if (e != cudaSuccess) return fail(SMK_ENOMEM, "psi buffer: %s", cudaGetErrorString(e));
            c->psi_capacity = tracks;
        }
    }

    KernelArgs a;
    a.source = reinterpret_cast<const float4 *>(c->d_source);
    a.sigT = reinterpret_cast<const float4 *>(c->d_sigT);
    a.tally = c->d_tally;
    a.replica_stride = c->rows * c->shape.groups_pad;
    a.replicas = c->replicas;
    a.psi_out = (c->p.flags & SMK_FLAG_KEEP_PSI) ? c->d_psi : nullptr;
    a.checksum = c->d_checksum;
    a.work_counter = c->d_work;
    a.tally64 = c->d_tally64;
    SMK_CUDA(cudaMemsetAsync(c->d_work, 0, sizeof(unsigned long long), c->stream));
    a.segments = c->p.segments;
    a.track_begin = track_begin;
    a.track_end = track_end;
    a.seed = c->p.seed;
    a.keys = make_philox_keys(c->p.seed);
    a.mod_regions = make_fastmod((uint32_t)c->p.source_3D_regions);
    a.mod_fai = make_fastmod((uint32_t)c->p.fine_axial_intervals);
    a.fai_count = c->p.fine_axial_intervals;
    a.row_f4 = c->shape.groups_pad / 4;
    a.seg_per_track = c->p.seg_per_track;
    a.group_blocks = c->shape.group_blocks;
    a.geom = GeometryBase{c->geom.dz, c->geom.zin, c->geom.weight, c->geom.mu, c->geom.mu2, c->geom.ds, c->geom.spread};
    const double dz = (double)c->geom.dz;
    a.mesh = MeshConsts{(float)(1.0 / (2.0 * dz)), (float)(1.0 / (2.0 * dz * dz)), (float)(1.0 / dz)};

    a.records = c->d_records;
    int lanes_per_track = c->shape.lpt;
    if (c->d_records && !c->tuning_kernel) {
        // derived layout, rebuilt from the canonical rows inside every sweep (they may have been written through
        // any of the upload paths or directly on the device since the last one)
        const int Gp = c->shape.groups_pad;
        // form 0: raw values; 1 / 2 (SMK_FLAG_FIT_PER_SWEEP): the fit evaluated here, once per (row, group), in the
        // arithmetic of the kernel that reads it (per-lane-coefficient form for <= 32 groups, typed form for 33..64)
        const int form = !c->hoist ? 0 : (Gp <= 32 ? 1 : 2);
        const int F = c->p.fine_axial_intervals;
#define SMK_BUILD(GPL, FORM) \
    build_records<GPL, FORM><<<layout_grid(c->rows * (Gp / GPL)), 256, 0, c->stream>>>(c->d_source, c->d_sigT, c->d_records, c->rows, F, Gp)
        if (Gp == 128) {      // 65..128 groups (SMK_FLAG_FIT_PER_SWEEP only): three fitted rows per source row
            build_fit_rows<<<layout_grid(c->rows * Gp), 256, 0, c->stream>>>(c->d_source, c->d_records, c->rows, F, Gp);
        } else if (c->rec_gpl == 2) {
            if (form == 0) SMK_BUILD(2, 0); else if (form == 1) SMK_BUILD(2, 1); else SMK_BUILD(2, 2);
        } else {
            if (form == 0) SMK_BUILD(4, 0); else SMK_BUILD(4, 1);
        }
#undef SMK_BUILD
        SMK_CUDA(cudaGetLastError());
        c->launches += 1;
        lanes_per_track = Gp / c->rec_gpl;      // (32 for the one-track-per-warp shapes)
    }

    // persistent grid: a whole number of CTAs per SM; warps claim work items dynamically
    const int slots_per_block = (kThreadsPerBlock / 32) * (32 / lanes_per_track);
    const int64_t work = tracks * c->shape.group_blocks;
    int64_t want = (work + slots_per_block - 1) / slots_per_block;
    int64_t full = (int64_t)c->sm_count * c->blocks_per_sm;
    int grid = (int)(want < full ? want : full);
    c->kernel<<<grid, kThreadsPerBlock, c->dyn_smem, c->stream>>>(a);
    SMK_CUDA(cudaGetLastError());
    c->launches += 1;
    return SMK_OK;
}

int smk_run_async(smk_ctx *c, int64_t track_begin, int64_t track_end)
{
    return launch(c, track_begin, track_end);
}

int smk_synchronize(smk_ctx *c)
{
    if (!c) return fail(SMK_EINVAL, "ctx is NULL");
    SMK_CUDA(cudaSetDevice(c->p.device));
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    return SMK_OK;
}

int smk_run(smk_ctx *c, int64_t track_begin, int64_t track_end, double *kernel_seconds)
{
    if (!c) return fail(SMK_EINVAL, "ctx is NULL");
    SMK_CUDA(cudaSetDevice(c->p.device));
    SMK_CUDA(cudaEventRecord(c->ev0, c->stream));
    int rc = launch(c, track_begin, track_end);
    if (rc != SMK_OK) return rc;
    SMK_CUDA(cudaEventRecord(c->ev1, c->stream));
    SMK_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    SMK_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (kernel_seconds) *kernel_seconds = (double)ms * 1e-3;
    return SMK_OK;
}

int64_t smk_launch_count(const smk_ctx *c) { return c ? c->launches : 0; }

int smk_download_flux_rows_async(smk_ctx *c, int64_t row_begin, int64_t rows, float *out)
{
    if (!c || !out) return fail(SMK_EINVAL, "NULL argument");
    if (row_begin < 0 || rows < 0 || row_begin + rows > c->rows)
        return fail(SMK_EINVAL, "rows [%lld, %lld) outside [0, %lld)", (long long)row_begin, (long long)(row_begin + rows),
                    (long long)c->rows);
    if (rows == 0) return SMK_OK;
    SMK_CUDA(cudaSetDevice(c->p.device));
    const int G = c->p.egroups, Gp = c->shape.groups_pad;
    float *stage = c->d_stage + row_begin * G;
    // FAST kernels of the constant geometry leave the (constant) segment weight to this step
    const bool unweighted = c->p.math_mode == kMathFast && !(c->p.flags & SMK_FLAG_SEGMENT_GEOMETRY);
    const float scale = unweighted ? kTallyScaleConst : 1.0f;
    if (c->d_tally64)
        finalize_flux64<<<layout_grid(rows * G), 256, 0, c->stream>>>(c->d_flux0 + row_begin * Gp, c->d_tally64 + row_begin * Gp,
                                                                     stage, rows, G, Gp, scale);
    else
        finalize_flux<<<layout_grid(rows * G), 256, 0, c->stream>>>(c->d_flux0 + row_begin * Gp, c->d_tally + row_begin * Gp,
                                                                   stage, rows, G, Gp, c->replicas, c->rows * Gp, scale);
    SMK_CUDA(cudaGetLastError());
    c->launches += 1;
    SMK_CUDA(cudaEventRecord(c->ev_finalized, c->stream));
    c->finalized_recorded = true;
    SMK_CUDA(cudaMemcpyAsync(out, stage, (size_t)rows * G * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    return SMK_OK;
}

int smk_wait_finalized(smk_ctx *c, smk_ctx *other)
{
    if (!c || !other) return fail(SMK_EINVAL, "NULL argument");
    if (c == other || !other->finalized_recorded) return SMK_OK;     // stream order / nothing to wait for
    if (c->p.device != other->p.device) return fail(SMK_EINVAL, "contexts on different devices (%d, %d)", c->p.device, other->p.device);
    SMK_CUDA(cudaSetDevice(c->p.device));
    SMK_CUDA(cudaStreamWaitEvent(c->stream, other->ev_finalized, 0));
    return SMK_OK;
}

int smk_download_flux(smk_ctx *c, float *out)
{
    if (!c) return fail(SMK_EINVAL, "NULL argument");
    int rc = smk_download_flux_rows_async(c, 0, c->rows, out);
    if (rc != SMK_OK) return rc;
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    return SMK_OK;
}

int smk_download_psi(smk_ctx *c, float *psi_out, int64_t n_tracks)
{
    if (!c || !psi_out) return fail(SMK_EINVAL, "NULL argument");
    if (!(c->p.flags & SMK_FLAG_KEEP_PSI)) return fail(SMK_ESTATE, "context created without SMK_FLAG_KEEP_PSI");
    const int64_t tracks = c->last_end - c->last_begin;
    if (n_tracks != tracks)
        return fail(SMK_EINVAL, "psi_out holds %lld tracks but the last run swept %lld", (long long)n_tracks, (long long)tracks);
    if (tracks <= 0) return SMK_OK;
    SMK_CUDA(cudaSetDevice(c->p.device));
    SMK_CUDA(cudaMemcpy2DAsync(psi_out, (size_t)c->p.egroups * sizeof(float), c->d_psi,
                               (size_t)c->shape.groups_pad * sizeof(float),
                               (size_t)c->p.egroups * sizeof(float), (size_t)tracks,
                               cudaMemcpyDeviceToHost, c->stream));
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    return SMK_OK;
}

int smk_download_checksum(smk_ctx *c, uint64_t *checksum)
{
    if (!c || !checksum) return fail(SMK_EINVAL, "NULL argument");
    SMK_CUDA(cudaSetDevice(c->p.device));
    unsigned long long v = 0;
    SMK_CUDA(cudaMemcpyAsync(&v, c->d_checksum, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    SMK_CUDA(cudaStreamSynchronize(c->stream));
    *checksum = (uint64_t)v;
    return SMK_OK;
}

int smk_run_host(const smk_params *p, const float *fine_source, float *fine_flux, const float *sigT,
                 double *kernel_seconds, double *total_seconds)
{
    if (!fine_flux) return fail(SMK_EINVAL, "fine_flux is required (updated in place)");
    smk_ctx *c = nullptr;
    int rc = smk_create(p, &c);
    if (rc != SMK_OK) return rc;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (cudaEventCreate(&t0) != cudaSuccess || cudaEventCreate(&t1) != cudaSuccess) {
        if (t0) cudaEventDestroy(t0);
        smk_destroy(c);
        return fail(SMK_ECUDA, "cudaEventCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaEventRecord(t0, c->stream);
    rc = smk_upload_async(c, fine_source, fine_flux, sigT);
    if (rc == SMK_OK) rc = smk_run(c, 0, c->n_tracks, kernel_seconds);
    if (rc == SMK_OK) rc = smk_download_flux(c, fine_flux);
    if (rc == SMK_OK) {
        cudaEventRecord(t1, c->stream);
        cudaEventSynchronize(t1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        if (total_seconds) *total_seconds = (double)ms * 1e-3;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    smk_destroy(c);
    return rc;
}

// ---------------------------------------------------------------------------------------
// multi-GPU, one process
// ---------------------------------------------------------------------------------------
namespace {
// the few NCCL entry points we need, resolved from libnccl.so.2 at first use (no link-time
// dependency: the peer-memory all-reduce needs no library at all)
struct Nccl {
    void *lib = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load()
    {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        CommInitAll = (int (*)(void **, int, const int *))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (int (*)(void *))dlsym(lib, "ncclCommDestroy");
        GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
        GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
        AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(lib, "ncclAllReduce");
        GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && AllReduce;
    }
};
Nccl g_nccl;
constexpr int kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;   // ncclFloat32, ncclFloat64, ncclSum (nccl.h)
}  // namespace

struct smk_multi {
    int n;
    int allreduce;
    smk_ctx *ctx[kMaxDevices];
    cudaEvent_t start[kMaxDevices], swept[kMaxDevices], reduced[kMaxDevices];
    void *comms[kMaxDevices];
    bool have_comms;
};

int smk_multi_device_count(const smk_multi *m) { return m ? m->n : 0; }

void smk_multi_destroy(smk_multi *m)
{
    if (!m) return;
    for (int d = 0; d < m->n; ++d) {
        if (!m->ctx[d]) continue;
        cudaSetDevice(m->ctx[d]->p.device);
        if (m->have_comms && m->comms[d]) g_nccl.CommDestroy(m->comms[d]);
        if (m->start[d]) cudaEventDestroy(m->start[d]);
        if (m->swept[d]) cudaEventDestroy(m->swept[d]);
        if (m->reduced[d]) cudaEventDestroy(m->reduced[d]);
        smk_destroy(m->ctx[d]);
    }
    delete m;
}

int smk_multi_create(const smk_params *p, int n_devices, const int *devices, int allreduce, smk_multi **out)
{
    if (!out) return fail(SMK_EINVAL, "out is NULL");
    *out = nullptr;
    if (!p) return fail(SMK_EINVAL, "params is NULL");
    if (n_devices < 1 || n_devices > kMaxDevices) return fail(SMK_EINVAL, "n_devices must be in [1, %d]", kMaxDevices);
    if (allreduce != SMK_ALLREDUCE_PEER && allreduce != SMK_ALLREDUCE_NCCL)
        return fail(SMK_EINVAL, "unknown all-reduce implementation %d", allreduce);
    int visible = 0;
    SMK_CUDA(cudaGetDeviceCount(&visible));
    smk_multi *m = new (std::nothrow) smk_multi();
    if (!m) return fail(SMK_ENOMEM, "out of host memory");
    memset(m, 0, sizeof(*m));
    m->n = n_devices;
    m->allreduce = allreduce;
    int ids[kMaxDevices];
    for (int d = 0; d < n_devices; ++d) {
        ids[d] = devices ? devices[d] : d;
        if (ids[d] < 0 || ids[d] >= visible) {
            smk_multi_destroy(m);
            return fail(SMK_EINVAL, "device %d out of range (%d visible)", ids[d], visible);
        }
    }
    for (int d = 0; d < n_devices; ++d) {
        smk_params pd = *p;
        pd.device = ids[d];
        int rc = smk_create(&pd, &m->ctx[d]);
        if (rc != SMK_OK) {
            smk_multi_destroy(m);
            return rc;
        }
        cudaEventCreate(&m->start[d]);
        cudaEventCreate(&m->swept[d]);
        cudaEventCreateWithFlags(&m->reduced[d], cudaEventDisableTiming);
    }
    if (n_devices > 1 && allreduce == SMK_ALLREDUCE_PEER) {
        for (int d = 0; d < n_devices; ++d) {
            cudaSetDevice(ids[d]);
            for (int e = 0; e < n_devices; ++e) {
                if (e == d) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, ids[d], ids[e]);
                if (!can) {
                    smk_multi_destroy(m);
                    return fail(SMK_ECUDA, "device %d cannot access peer %d (no NVLink/P2P path)", ids[d], ids[e]);
                }
                cudaError_t err = cudaDeviceEnablePeerAccess(ids[e], 0);
                if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) {
                    smk_multi_destroy(m);
                    return fail(SMK_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", ids[d], ids[e], cudaGetErrorString(err));
                }
                cudaGetLastError();
            }
        }
    }
    if (n_devices > 1 && allreduce == SMK_ALLREDUCE_NCCL) {
        if (!g_nccl.load()) {
            smk_multi_destroy(m);
            return fail(SMK_ESTATE, "libnccl.so.2 could not be loaded: %s", dlerror());
        }
        int rc = g_nccl.CommInitAll(m->comms, n_devices, ids);
        if (rc != 0) {
            smk_multi_destroy(m);
            return fail(SMK_ECUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
        }
        m->have_comms = true;
        // first collective on a communicator sets up its channels (~1 s): do it here, on the
        // zeroed tallies, so that smk_multi_run times the sweep and not NCCL's lazy initialisation
        const size_t n = (size_t)(m->ctx[0]->rows * m->ctx[0]->shape.groups_pad) * m->ctx[0]->replicas;
        rc = g_nccl.GroupStart();
        for (int d = 0; d < n_devices && rc == 0; ++d)
            rc = g_nccl.AllReduce(m->ctx[d]->d_tally, m->ctx[d]->d_tally, n, kNcclFloat32, kNcclSum, m->comms[d],
                                  m->ctx[d]->stream);
        if (rc == 0) rc = g_nccl.GroupEnd();
        for (int d = 0; d < n_devices; ++d) {
            cudaSetDevice(ids[d]);
            cudaStreamSynchronize(m->ctx[d]->stream);
        }
        if (rc != 0) {
            smk_multi_destroy(m);
            return fail(SMK_ECUDA, "NCCL warm-up all-reduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
        }
    }
    *out = m;
    return SMK_OK;
}

int smk_multi_upload(smk_multi *m, const float *fine_source, const float *fine_flux, const float *sigT)
{
    if (!m) return fail(SMK_EINVAL, "multi is NULL");
    // one host -> device copy, to the first device ...
    smk_ctx *c0 = m->ctx[0];
    int rc = smk_upload_async(c0, fine_source, fine_flux, sigT);
    if (rc != SMK_OK) return rc;
    SMK_CUDA(cudaEventRecord(m->reduced[0], c0->stream));
    // ... then every other device pulls the padded arrays over NVLink, all of them in parallel
    const size_t slab = (size_t)c0->rows * c0->shape.groups_pad * sizeof(float);
    const size_t sig = (size_t)c0->p.source_3D_regions * c0->shape.groups_pad * sizeof(float);
    for (int d = 1; d < m->n; ++d) {
        smk_ctx *c = m->ctx[d];
        SMK_CUDA(cudaSetDevice(c->p.device));
        SMK_CUDA(cudaStreamWaitEvent(c->stream, m->reduced[0], 0));
        SMK_CUDA(cudaMemcpyPeerAsync(c->d_source, c->p.device, c0->d_source, c0->p.device, slab, c->stream));
        SMK_CUDA(cudaMemcpyPeerAsync(c->d_sigT, c->p.device, c0->d_sigT, c0->p.device, sig, c->stream));
        SMK_CUDA(cudaMemcpyPeerAsync(c->d_flux0, c->p.device, c0->d_flux0, c0->p.device, slab, c->stream));
        if ((rc = smk_reset_tallies(c)) != SMK_OK) return rc;
        c->sigt_max = c0->sigt_max;
        c->have_data = true;
    }
    for (int d = 0; d < m->n; ++d) {
        SMK_CUDA(cudaSetDevice(m->ctx[d]->p.device));
        SMK_CUDA(cudaStreamSynchronize(m->ctx[d]->stream));
    }
    return SMK_OK;
}

int smk_multi_set_geometry(smk_multi *m, const smk_geometry *g)
{
    if (!m) return fail(SMK_EINVAL, "multi is NULL");
    for (int d = 0; d < m->n; ++d) {
        int rc = smk_set_geometry(m->ctx[d], g);
        if (rc != SMK_OK) return rc;
    }
    return SMK_OK;
}

int smk_multi_fill_device(smk_multi *m, float sigt_floor)
{
    if (!m) return fail(SMK_EINVAL, "multi is NULL");
    for (int d = 0; d < m->n; ++d) {
        int rc = smk_fill_device(m->ctx[d], sigt_floor);
        if (rc != SMK_OK) return rc;
    }
    return SMK_OK;
}

int smk_multi_run(smk_multi *m, double *kernel_seconds, double *total_seconds)
{
    if (!m) return fail(SMK_EINVAL, "multi is NULL");
    const int P = m->n;
    const int64_t T = m->ctx[0]->n_tracks;
    for (int d = 0; d < P; ++d) {
        int rc = smk_reset_tallies(m->ctx[d]);
        if (rc != SMK_OK) return rc;
    }
    for (int d = 0; d < P; ++d) SMK_CUDA(cudaStreamSynchronize(m->ctx[d]->stream));
    const auto t0 = std::chrono::steady_clock::now();

    // 1. the sweep: device d takes tracks [d*T/P, (d+1)*T/P)
    for (int d = 0; d < P; ++d) {
        smk_ctx *c = m->ctx[d];
        SMK_CUDA(cudaSetDevice(c->p.device));
        SMK_CUDA(cudaEventRecord(m->start[d], c->stream));
        int rc = launch(c, d * T / P, (d + 1) * T / P);
        if (rc != SMK_OK) return rc;
        SMK_CUDA(cudaEventRecord(m->swept[d], c->stream));
    }
    // 2. one all-reduce of the tally deltas
    if (P > 1 && m->allreduce == SMK_ALLREDUCE_PEER) {
        const bool f64 = m->ctx[0]->d_tally64 != nullptr;
        PeerArrays arrays;
        for (int d = 0; d < kMaxDevices; ++d)
            arrays.p[d] = d < P ? (f64 ? reinterpret_cast<float4 *>(m->ctx[d]->d_tally64)
                                       : reinterpret_cast<float4 *>(m->ctx[d]->d_tally)) : nullptr;
        // elements per device array: float4 for the fp32 tallies, double for the f64 diagnostic ones
        const int64_t n4 = f64 ? m->ctx[0]->rows * m->ctx[0]->shape.groups_pad
                               : m->ctx[0]->rows * m->ctx[0]->shape.groups_pad / 4 * m->ctx[0]->replicas;
        for (int d = 0; d < P; ++d) {
            smk_ctx *c = m->ctx[d];
            SMK_CUDA(cudaSetDevice(c->p.device));
            for (int e = 0; e < P; ++e)
                if (e != d) SMK_CUDA(cudaStreamWaitEvent(c->stream, m->swept[e], 0));
            const int64_t b = d * n4 / P, e4 = (d + 1) * n4 / P;
            if (f64) allreduce_peer_slices64<<<layout_grid(e4 - b), 256, 0, c->stream>>>(arrays, P, b, e4);
            else allreduce_peer_slices<<<layout_grid(e4 - b), 256, 0, c->stream>>>(arrays, P, b, e4);
            SMK_CUDA(cudaGetLastError());
            c->launches += 1;
            SMK_CUDA(cudaEventRecord(m->reduced[d], c->stream));
        }
        for (int d = 0; d < P; ++d) {       // nobody reads its tallies before every slice is written
            SMK_CUDA(cudaSetDevice(m->ctx[d]->p.device));
            for (int e = 0; e < P; ++e)
                if (e != d) SMK_CUDA(cudaStreamWaitEvent(m->ctx[d]->stream, m->reduced[e], 0));
        }
    } else if (P > 1) {
        const bool f64 = m->ctx[0]->d_tally64 != nullptr;
        const size_t n = f64 ? (size_t)(m->ctx[0]->rows * m->ctx[0]->shape.groups_pad)
                             : (size_t)(m->ctx[0]->rows * m->ctx[0]->shape.groups_pad) * m->ctx[0]->replicas;
        int rc = g_nccl.GroupStart();
        for (int d = 0; d < P && rc == 0; ++d) {
            void *buf = f64 ? (void *)m->ctx[d]->d_tally64 : (void *)m->ctx[d]->d_tally;
            rc = g_nccl.AllReduce(buf, buf, n, f64 ? kNcclFloat64 : kNcclFloat32, kNcclSum, m->comms[d], m->ctx[d]->stream);
        }
        if (rc == 0) rc = g_nccl.GroupEnd();
        if (rc != 0) return fail(SMK_ECUDA, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    }
    for (int d = 0; d < P; ++d) {
        SMK_CUDA(cudaSetDevice(m->ctx[d]->p.device));
        SMK_CUDA(cudaStreamSynchronize(m->ctx[d]->stream));
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (total_seconds) *total_seconds = std::chrono::duration<double>(t1 - t0).count();
    if (kernel_seconds) {
        double worst = 0.0;
        for (int d = 0; d < P; ++d) {
            float ms = 0.f;
            SMK_CUDA(cudaSetDevice(m->ctx[d]->p.device));
            SMK_CUDA(cudaEventElapsedTime(&ms, m->start[d], m->swept[d]));
            if (ms * 1e-3 > worst) worst = ms * 1e-3;
        }
        *kernel_seconds = worst;
    }
    return SMK_OK;
}

int smk_multi_download_flux(smk_multi *m, int which, float *out)
{
    if (!m || which < 0 || which >= m->n) return fail(SMK_EINVAL, "bad device index");
    return smk_download_flux(m->ctx[which], out);
}

int smk_multi_download_checksum(smk_multi *m, uint64_t *checksum)
{
    if (!m || !checksum) return fail(SMK_EINVAL, "NULL argument");
    uint64_t sum = 0;
    for (int d = 0; d < m->n; ++d) {
        uint64_t v = 0;
        int rc = smk_download_checksum(m->ctx[d], &v);
        if (rc != SMK_OK) return rc;
        sum += v;
    }
    *checksum = sum;
    return SMK_OK;
}

void *smk_device_tally(smk_ctx *c) { return c ? c->d_tally : nullptr; }
void *smk_device_flux0(smk_ctx *c) { return c ? c->d_flux0 : nullptr; }
void *smk_device_source(smk_ctx *c) { return c ? c->d_source : nullptr; }
void *smk_device_sigT(smk_ctx *c) { return c ? c->d_sigT : nullptr; }
int64_t smk_padded_elems(const smk_ctx *c) { return c ? c->rows * c->shape.groups_pad * c->replicas : 0; }

void *smk_alloc_host(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        fail(SMK_ENOMEM, "cudaMallocHost(%zu) failed", bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void smk_free_host(void *p)
{
    if (p) cudaFreeHost(p);
}

int smk_debug_exp(int exp_mode, const float *tau, float *out, int64_t n, int device)
{
    if (!tau || !out || n < 0) return fail(SMK_EINVAL, "bad argument");
    if (n == 0) return SMK_OK;
    const bool packed = (exp_mode & SMK_DEBUG_EXP_PACKED) != 0, wide = (exp_mode & SMK_DEBUG_EXP_WIDE) != 0;
    const bool track = (exp_mode & SMK_DEBUG_EXP_TRACK) != 0;
    exp_mode &= 0xFF;
    if (exp_mode < SMK_EXP_POLY || exp_mode > SMK_EXP_TABLE) return fail(SMK_EINVAL, "unknown exp_mode %d", exp_mode);
    SMK_CUDA(cudaSetDevice(device));
    ExpTable tab;
    build_exp_table(tab);
    SMK_CUDA(cudaMemcpyToSymbol(c_exp_table, &tab, sizeof(tab)));
    SMK_CUDA(cudaMemcpyToSymbol(c_exp2f_tab, h_exp2f_tab, sizeof(h_exp2f_tab)));
    float *d_in = nullptr, *d_out = nullptr;
    SMK_CUDA(cudaMalloc(&d_in, (size_t)n * sizeof(float)));
    cudaError_t e = cudaMalloc(&d_out, (size_t)n * sizeof(float));
    if (e != cudaSuccess) {
        cudaFree(d_in);
        return fail(SMK_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
    }
    cudaMemcpy(d_in, tau, (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
    const int grid = layout_grid(n);
    if (packed) {
        switch (exp_mode) {
            case SMK_EXP_POLY:
                if (wide && track) debug_exp2_kernel<kExpPolyWide, true><<<grid, 256>>>(d_in, d_out, n);
                else if (wide) debug_exp2_kernel<kExpPolyWide, false><<<grid, 256>>>(d_in, d_out, n);
                else if (track) debug_exp2_kernel<kExpPoly, true><<<grid, 256>>>(d_in, d_out, n);
                else debug_exp2_kernel<kExpPoly, false><<<grid, 256>>>(d_in, d_out, n);
                break;
            case SMK_EXP_MUFU: debug_exp2_kernel<kExpMufu, false><<<grid, 256>>>(d_in, d_out, n); break;
            case SMK_EXP_GLIBC: debug_exp2_kernel<kExpGlibc, false><<<grid, 256>>>(d_in, d_out, n); break;
            case SMK_EXP_TABLE: debug_exp2_kernel<kExpTable, false><<<grid, 256>>>(d_in, d_out, n); break;
        }
    } else {
        switch (exp_mode) {
            case SMK_EXP_POLY: debug_exp_kernel<kExpPoly><<<grid, 256>>>(d_in, d_out, n); break;
            case SMK_EXP_MUFU: debug_exp_kernel<kExpMufu><<<grid, 256>>>(d_in, d_out, n); break;
            case SMK_EXP_GLIBC: debug_exp_kernel<kExpGlibc><<<grid, 256>>>(d_in, d_out, n); break;
            case SMK_EXP_TABLE: debug_exp_kernel<kExpTable><<<grid, 256>>>(d_in, d_out, n); break;
        }
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(SMK_ECUDA, "debug_exp: %s", cudaGetErrorString(e));
    return SMK_OK;
}

static int debug_ids(const smk_params *p, const smk_geometry *g, int64_t seg_begin, int64_t n, int32_t *qsr_out,
                     int32_t *fai_out, float *geom6_out)
{
    Shape shape;
    int rc = validate(p, shape);
    if (rc != SMK_OK) return rc;
    if (n < 0 || seg_begin < 0) return fail(SMK_EINVAL, "bad argument");
    if (n == 0) return SMK_OK;
    SMK_CUDA(cudaSetDevice(p->device));
    int32_t *d_q = nullptr, *d_f = nullptr;
    float *d_g = nullptr;
    cudaError_t e = cudaMalloc(&d_q, (size_t)n * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_f, (size_t)n * sizeof(int32_t));
    if (e == cudaSuccess && geom6_out) e = cudaMalloc(&d_g, (size_t)n * 6 * sizeof(float));
    if (e == cudaSuccess) {
        const smk_geometry gg = g ? *g : kReferenceGeometry;
        debug_ids_kernel<<<layout_grid(n), 256>>>(p->seed, seg_begin, n, make_fastmod((uint32_t)p->source_3D_regions),
                                                 make_fastmod((uint32_t)p->fine_axial_intervals),
                                                 GeometryBase{gg.dz, gg.zin, gg.weight, gg.mu, gg.mu2, gg.ds, gg.spread},
                                                 d_q, d_f, d_g);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && qsr_out) e = cudaMemcpy(qsr_out, d_q, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && fai_out) e = cudaMemcpy(fai_out, d_f, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && geom6_out) e = cudaMemcpy(geom6_out, d_g, (size_t)n * 6 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_q);
    cudaFree(d_f);
    cudaFree(d_g);
    if (e != cudaSuccess) return fail(SMK_ECUDA, "debug_segment_ids: %s", cudaGetErrorString(e));
    return SMK_OK;
}

int smk_debug_segment_ids(const smk_params *p, int64_t seg_begin, int64_t n, int32_t *qsr_out, int32_t *fai_out)
{
    if (!qsr_out || !fai_out) return fail(SMK_EINVAL, "bad argument");
    return debug_ids(p, nullptr, seg_begin, n, qsr_out, fai_out, nullptr);
}

int smk_debug_segment_geometry(const smk_params *p, const smk_geometry *g, int64_t seg_begin, int64_t n, float *geom6_out)
{
    if (!geom6_out) return fail(SMK_EINVAL, "bad argument");
    int rc = check_geometry(g);
    if (rc != SMK_OK) return rc;
    return debug_ids(p, g, seg_begin, n, nullptr, nullptr, geom6_out);
}

}  // extern "C"
