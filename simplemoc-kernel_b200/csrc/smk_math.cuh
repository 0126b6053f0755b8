// smk_math.cuh -- per-intersection arithmetic of attenuate_segment
// (/root/reference/src/cpu/kernel.c:75-333) for one (segment, energy group).
//
// Two arithmetic modes (include/smk.h):
//   STRICT  every operation of the reference, in its order, with IEEE round-to-nearest
//           intrinsics (never contracted): bit-identical per intersection to a
//           -O2 -ffp-contract=off build of kernel.c.  Verification mode.
//   FAST    the same formulae with the constants folded into the quadratic-fit
//           coefficients, FMA contraction and one MUFU.RCP replacing the five divides
//           (kernel.c:236,249,251,291,301).  Throughput mode; checked against the
//           oracle by the L2-relative gate of DESIGN.md section 6.
//
// Four evaluations of e = exp(-tau) (kernel.c:221), see exp_val<>.
// With GEOM (SMK_FLAG_SEGMENT_GEOMETRY) the six placeholder constants of kernel.c:99-104 are
// per-segment values (smk_stream.cuh: segment_geometry) instead of literals.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#include "smk_stream.cuh"

namespace smk {

// placeholder geometry of the mini-app, kernel.c:99-104
struct Geometry {
    static constexpr float dz = 0.1f;
    static constexpr float zin = 0.3f;
    static constexpr float weight = 0.5f;
    static constexpr float mu = 0.9f;
    static constexpr float mu2 = 0.3f;
    static constexpr float ds = 0.7f;
};

// kExpPolyWide is internal: SMK_EXP_POLY when the data do not guarantee tau <= kPolyMaxTau
enum : int { kExpPoly = 0, kExpMufu = 1, kExpGlibc = 2, kExpTable = 3, kExpPolyWide = 4 };
enum : int { kMathFast = 0, kMathStrict = 1 };

// exp_poly() is fitted on x = -tau in [-0.7, 0] (the reference's own data: sigT < 1, ds = 0.7).
// Beyond that its error grows quickly (2e-6 at tau = 1, 9e-3 at tau = 2, garbage from tau = 3), so
// for tau > kPolyMaxTau the POLY mode switches to MUFU.EX2: 1 - e is well conditioned there
// (e <= 0.4966, so 2 ulp of e is 1.2e-7 of expVal) and the 1e-5 gate does not need the
// correctly-rounded exponential that small tau needs.
constexpr float kPolyMaxTau = 0.7f;

// --------------------------------------------------------------------------
// Exponential table of the reference (init.c:81-117): 353 {slope, intercept}
// pairs, dx = 10/353, maxVal = 10 - dx.  tau = 0.7*sigT < 0.7 only ever reaches
// the first 25 intervals; the kernel keeps kTableReach pairs in shared memory.
// --------------------------------------------------------------------------
constexpr int kTableN = 353;
constexpr int kTableReach = 32;
struct ExpTable {
    float2 pairs[kTableN];   // .x = slope, .y = intercept
    float dx;
    float maxVal;
};
// (defined here: this header is included by exactly one translation unit, smk_api.cu)
__constant__ ExpTable c_exp_table;

// 2^(i/32) bit patterns minus (i << 47): the 32-entry table of glibc's expf
// (sysdeps/ieee754/flt-32/math_config.h, __exp2f_data.tab; glibc 2.39).
__constant__ uint64_t c_exp2f_tab[32];

// e = exp(-tau) the way glibc 2.39's x86-64 FMA variant of expf computes it
// (sysdeps/ieee754/flt-32/e_expf.c built with -mfma -mavx2; verified against the
// disassembly of __expf_fma in this image's libm): double arithmetic, N = 32,
// degree-3 polynomial, fused multiply-adds where gcc contracted them.
__device__ __forceinline__ float expf_glibc(float x)
{
    const double InvLn2N = 0x1.71547652b82fep+5;   // 32 / ln 2
    const double Shift = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-20;       // 0x1.c6af84b912394p-5 / 32^3
    const double C1 = 0x1.ebfce50fac4f3p-13;       // 0x1.ebfce50fac4f3p-3 / 32^2
    const double C2 = 0x1.62e42ff0c52d6p-6;        // 0x1.62e42ff0c52d6p-1 / 32
    const double xd = (double)x;
    double kd = __fma_rn(InvLn2N, xd, Shift);
    const uint64_t ki = (uint64_t)__double_as_longlong(kd);
    kd = __dsub_rn(kd, Shift);
    const double r = __fma_rn(InvLn2N, xd, -kd);
    uint64_t t = c_exp2f_tab[ki & 31u];
    t += ki << 47;
    const double s = __longlong_as_double((long long)t);
    const double z = __fma_rn(C0, r, C1);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(C2, r, 1.0);
    y = __fma_rn(z, r2, y);
    y = __dmul_rn(y, s);
    return __double2float_rn(y);
}

// e = RN(1 + x + x^2 P(x)), x = -tau in [-0.7, 0], with the rounding error of 1 + x
// carried into the last addition (Fast2Sum), so that e equals the correctly rounded
// exp(x) wherever 1 - e is ill-conditioned (|x| small).
// Where the reference's libm is NOT correctly rounded the kernel has to follow libm, not exp: for a
// small cross section one ulp of e moves the cubic term of kernel.c:250-251 by 1.2e-7 / sigT^4 (a third
// of the whole tally at sigT = 1.4e-3).  glibc 2.39's expf evaluates, for |x| < ln2/64 (its k = 0
// interval), the fixed cubic 1 + a1 x + a2 x^2 + a3 x^3 in double with
//     a1 = 1 + 1.8997e-10,  a2 = 1/2 + 4.0489e-6,  a3 = 1/6 - 1.4806e-6     (C2, C1, C0 of e_expf.c
// times powers of 32/ln2), which is up to 9e-4 ulp away from exp(x) -- the probability that it rounds to
// the other neighbour.  Below kPolyGlibcTau the polynomial therefore takes glibc's a2 as its leading
// coefficient and, everywhere, glibc's (a1 - 1) x (2e-3 ulp at tau = 0.7: harmless); a3 and the missing
// x^4 term are worth < 2e-5 ulp there.  Against host libm (tools/expf_sweep.py, profiles/expf_sweep_r02.md):
// no mismatch for tau < 2^-10, 2.7e-5 of the values for tau in [2^-10, 2^-8] (was 3.3e-4), <= 1 ulp everywhere.
constexpr float kPolyGlibcTau = 0x1p-8f;
constexpr float kPolyGlibcA2 = 0x1.000088p-1f;     // RN(C1 * (32/ln2)^2)
constexpr float kPolyGlibcB1 = 0x1.a1bdd2p-33f;    // RN(C2 * (32/ln2) - 1)

__device__ __forceinline__ float exp_poly(float x)
{
    float p = 0x1.415ffep-13f;
    p = __fmaf_rn(p, x, 0x1.6336e4p-10f);
    p = __fmaf_rn(p, x, 0x1.10ac84p-7f);
    p = __fmaf_rn(p, x, 0x1.555146p-5f);
    p = __fmaf_rn(p, x, 0x1.555546p-3f);
    p = __fmaf_rn(p, x, (x > -kPolyGlibcTau) ? kPolyGlibcA2 : 0.5f);
    const float s = __fadd_rn(1.0f, x);
    const float lost = __fsub_rn(x, __fsub_rn(s, 1.0f));   // exact: (1 + x) - s
    return __fadd_rn(s, __fmaf_rn(x, __fmaf_rn(x, p, kPolyGlibcB1), lost));
}

__device__ __forceinline__ float exp_mufu(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}

__device__ __forceinline__ float rcp_mufu(float x)
{
#ifdef SMK_EXPERIMENT_NO_MUFU   // timing experiment only: wrong values, no XU-pipe instruction
    return __fmul_rn(x, 0.9f);
#endif
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// interpolateTable, kernel.c:337-361, against pairs staged in shared memory
__device__ __forceinline__ float table_lookup(const float2 *__restrict__ s_pairs, float dx,
                                              float maxVal, float x)
{
    if (x > maxVal) return 1.0f;
    int interval = (int)__fadd_rn(__fdiv_rn(x, dx), __fmul_rn(0.5f, dx));
    const float2 si = (interval < kTableReach) ? s_pairs[interval] : c_exp_table.pairs[interval];
    return __fadd_rn(__fmul_rn(si.x, x), si.y);
}

// expVal = 1 - exp(-tau) (kernel.c:221) or its table replacement (kernel.c:219).
// *e_out receives exp(-tau) itself where it is available (FAST reuses it for t4).
template <int EXPM>
__device__ __forceinline__ float exp_val(float tau, const float2 *s_pairs, float &e_out)
{
    if constexpr (EXPM == kExpTable) {
        const float ev = table_lookup(s_pairs, c_exp_table.dx, c_exp_table.maxVal, tau);
        e_out = __fsub_rn(1.0f, ev);
        return ev;
    } else {
        float e;
        if constexpr (EXPM == kExpPoly || EXPM == kExpPolyWide) e = (tau > kPolyMaxTau) ? exp_mufu(-tau) : exp_poly(-tau);
        else if constexpr (EXPM == kExpMufu) e = exp_mufu(-tau);
        else e = expf_glibc(-tau);
        e_out = e;
        return __fsub_rn(1.0f, e);
    }
}

// --------------------------------------------------------------------------
// Quadratic / linear axial source fit (kernel.c:111-191) in difference form, constants folded
// (q1, q2 pre-multiplied by mu, mu2):
//   d = y1 - y3,  s = y1 - 2 y2 + y3   with (y1, y2, y3) = fine_source[QSR][FAI-1 .. FAI+1][g]
//   q0 = y2 + q0_d d + q0_s s,   mu q1 = q1_d d + q1_s s,   mu2 q2 = q2_s s
//   interior   : c1 = d / (2 dz), c2 = s / (2 dz^2)   (kernel.c:182-189)
//       q0_d = zin/(2dz)  q0_s = zin^2/(2dz^2)  q1_d = mu/(2dz)  q1_s = mu zin/dz^2  q2_s = mu2/(2dz^2)
//       = (1.5, 4.5, 4.5, 27, 15) for the reference geometry (dz = 0.1, zin = 0.3, mu = 0.9, mu2 = 0.3)
//   FAI == 0   : c1 = (y3 - y2) / dz   (kernel.c:128-134):  q0 = y2 + e0 (y3 - y2), mu q1 = e1 (y3 - y2)
//   FAI == F-1 : c1 = (y2 - y1) / dz   (kernel.c:154-160):  q0 = y2 + e0 (y2 - y1), mu q1 = e1 (y2 - y1)
//       e0 = zin/dz = 3, e1 = mu/dz = 9
// Per-lane coefficients (kFitDynamic) are needed where lanes of one warp serve tracks of different
// types: the edge fits are then written in the interior's form with y1 := 0 or y3 := 0,
//   FAI == 0   : (q0_d, q0_s, q1_d, q1_s, q2_s) = (-e0/2,  e0/2, -e1/2,  e1/2, 0)
//   FAI == F-1 : (q0_d, q0_s, q1_d, q1_s, q2_s) = (-e0/2, -e0/2, -e1/2, -e1/2, 0)
// --------------------------------------------------------------------------
struct FitDiff {
    static constexpr float dz = Geometry::dz, zin = Geometry::zin, mu = Geometry::mu, mu2 = Geometry::mu2;
    static constexpr float k1 = 1.0f / (2.0f * dz), k2 = 1.0f / (2.0f * dz * dz);
    static constexpr float q0_d = k1 * zin, q0_s = k2 * zin * zin;          // 1.5, 4.5
    static constexpr float q1_d = mu * k1, q1_s = mu * 2.0f * k2 * zin;     // 4.5, 27
    static constexpr float q2_s = mu2 * k2;                                 // 15
    static constexpr float e0 = zin / dz, e1 = mu / dz;                     // 3, 9
};

// Per-segment scalars of the FAST arithmetic.  For a statically typed edge body (kFitFirst /
// kFitLast) q0_d holds e0 and q1_d holds e1; the other fit fields are unused there.
struct FitCoeffs {
    float q0_d, q0_s, q1_d, q1_s, q2_s;
    float ds, weight;
};

// 1/(2dz), 1/(2dz^2), 1/dz of the problem's axial mesh (host-computed once)
struct MeshConsts {
    float k1, k2, inv_dz;
};

// reference geometry, per-lane segment type (sub-warp tracks)
__device__ __forceinline__ FitCoeffs fit_coeffs(bool first, bool last)
{
    const bool edge = first || last;
    FitCoeffs f;
    f.q0_d = edge ? -0.5f * FitDiff::e0 : FitDiff::q0_d;
    f.q0_s = first ? 0.5f * FitDiff::e0 : (last ? -0.5f * FitDiff::e0 : FitDiff::q0_s);
    f.q1_d = edge ? -0.5f * FitDiff::e1 : FitDiff::q1_d;
    f.q1_s = first ? 0.5f * FitDiff::e1 : (last ? -0.5f * FitDiff::e1 : FitDiff::q1_s);
    f.q2_s = edge ? 0.0f : FitDiff::q2_s;
    f.ds = Geometry::ds;
    f.weight = Geometry::weight;
    return f;
}

// per-segment geometry, coefficients in the interior's form for any type (kFitDynamic)
__device__ __forceinline__ FitCoeffs fit_coeffs_geom(const SegGeometry &g, const MeshConsts &m, bool first, bool last)
{
    FitCoeffs f;
    if (first || last) {
        const float h0 = 0.5f * g.zin * m.inv_dz, h1 = 0.5f * g.mu * m.inv_dz;
        f.q0_d = -h0;
        f.q0_s = first ? h0 : -h0;
        f.q1_d = -h1;
        f.q1_s = first ? h1 : -h1;
        f.q2_s = 0.0f;
    } else {
        const float kz = m.k2 * g.zin;
        f.q0_d = m.k1 * g.zin;
        f.q0_s = kz * g.zin;
        f.q1_d = g.mu * m.k1;
        f.q1_s = 2.0f * g.mu * kz;
        f.q2_s = g.mu2 * m.k2;
    }
    f.ds = g.ds;
    f.weight = g.weight;
    return f;
}

// per-segment geometry, coefficients for the statically typed bodies (one track per warp: the
// lane that hashed the segment computes them once, the warp shares them)
__device__ __forceinline__ FitCoeffs fit_coeffs_geom_typed(const SegGeometry &g, const MeshConsts &m, bool edge)
{
    FitCoeffs f;
    const float kz = m.k2 * g.zin;
    f.q0_d = edge ? g.zin * m.inv_dz : m.k1 * g.zin;      // e0 | q0_d
    f.q0_s = kz * g.zin;
    f.q1_d = edge ? g.mu * m.inv_dz : g.mu * m.k1;        // e1 | q1_d
    f.q1_s = 2.0f * g.mu * kz;
    f.q2_s = g.mu2 * m.k2;
    f.ds = g.ds;
    f.weight = g.weight;
    return f;
}

// ------------------------------------------------------------------------------
// FAST, packed: two intersections (two adjacent energy groups) per instruction with the
// sm_100 FP32x2 datapath (PTX fma/mul/add.rn.f32x2, SASS FFMA2 / FMUL2 / FADD2).  Every
// packed operation is the IEEE round-to-nearest operation on each half, so this is the
// scalar FAST arithmetic with half the issue slots.  The kernel was issue-bound
// (profiles/ncu_r01a_summary.md: issue active 80.7 %, FMA pipe 60 %).
// ------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
// a - b with one rounding: fma(b, -1, a)
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, f2(-1.0f), a); }

// kFitGiven / kFitGivenEdge: the fitted (q0, mu q1, mu2 q2) arrive in the (y2, y1, y3) arguments (SMK_FLAG_FIT_PER_SWEEP:
// fit_row evaluated them once per (region, interval, group) of the sweep); the edge form skips the quadratic terms
enum : int { kFitInterior = 0, kFitFirst = 1, kFitLast = 2, kFitDynamic = 3, kFitGiven = 4, kFitGivenEdge = 5 };

// what finalize_flux multiplies the summed FAST tallies of the constant geometry by (see attenuate_fast2)
constexpr float kTallyScaleConst = Geometry::weight;

// e = exp(-tau) on both halves; returns expVal = 1 - e.  omt_out = RN(1 - tau), a by-product of the
// polynomial that attenuate_fast2 reuses for tau^2 - 2 tau = (1 - tau)^2 - 1.
// TRACK selects the glibc-following leading coefficient per half (see exp_poly): two FSETP + two FSEL per pair,
// 4 % of the 128-group kernel's time (profiles/ab_r02.md).  The kernels switch it on with per-segment
// geometry, where every intersection draws its own tau; with the constant geometry tau takes one value per
// (region, group), libm's misroundings (2e-5 of the values at tau = 5e-4) hit an element that matters with
// probability ~1e-4 per data set, and the selects are left out.
template <int EXPM, bool TRACK>
__device__ __forceinline__ float2 exp_val2(float2 tau, const float2 *s_pairs, float2 &e_out, float2 &omt_out)
{
    if constexpr (EXPM == kExpPoly || EXPM == kExpPolyWide) {
        // exp_poly() on both halves, written in tau = -x: the odd coefficients change sign, every
        // intermediate is the same number up to sign, so e is bit-identical to exp_poly(-tau)
        float2 p = fma2(f2(-0x1.415ffep-13f), tau, f2(0x1.6336e4p-10f));
        p = fma2(p, tau, f2(-0x1.10ac84p-7f));
        p = fma2(p, tau, f2(0x1.555146p-5f));
        p = fma2(p, tau, f2(-0x1.555546p-3f));
        // leading coefficient: glibc's for small tau (see exp_poly), selected per half
        if constexpr (TRACK)
            p = fma2(p, tau, make_float2(tau.x < kPolyGlibcTau ? kPolyGlibcA2 : 0.5f,
                                         tau.y < kPolyGlibcTau ? kPolyGlibcA2 : 0.5f));
        else
            p = fma2(p, tau, f2(0.5f));
        const float2 s = fma2(tau, f2(-1.0f), f2(1.0f));             // RN(1 - tau)
        omt_out = s;
        const float2 lost = fma2(tau, f2(-1.0f), sub2(f2(1.0f), s)); // exact: (1 - tau) - s
        float2 e = add2(s, fma2(tau, fma2(tau, p, f2(-kPolyGlibcB1)), lost));
        if constexpr (EXPM == kExpPolyWide) {
            // outside the fitted range: MUFU.EX2 (XU pipe, otherwise idle), selected per half
            const float2 t = mul2(tau, f2(-1.4426950408889634f));
            float mx, my;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(mx) : "f"(t.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(my) : "f"(t.y));
            e.x = (tau.x > kPolyMaxTau) ? mx : e.x;
            e.y = (tau.y > kPolyMaxTau) ? my : e.y;
        }
        e_out = e;
        return sub2(f2(1.0f), e);
    } else {
        float ex, ey;
        const float evx = exp_val<EXPM>(tau.x, s_pairs, ex);
        const float evy = exp_val<EXPM>(tau.y, s_pairs, ey);
        omt_out = fma2(tau, f2(-1.0f), f2(1.0f));
        e_out = make_float2(ex, ey);
        return make_float2(evx, evy);
    }
}

// The fit of attenuate_fast2 for ONE group, operation for operation (FFMA2 / FMUL2 / FADD2 are the scalar IEEE operations on
// each half): TYPED = the statically typed bodies (kFitInterior / kFitFirst / kFitLast, constant geometry literals),
// otherwise the per-lane-coefficient form (kFitDynamic with fit_coeffs(first, last), missing neighbour = 0).  Used by
// build_records with SMK_FLAG_FIT_PER_SWEEP; the results are bit-identical to what the sweep kernels would compute.
template <bool TYPED>
__device__ __forceinline__ void fit_row(bool first, bool last, float y1, float y2, float y3, float &q0, float &Q1, float &Q2)
{
    using K = FitDiff;
    if (TYPED && (first || last)) {
        const float d = first ? __fmaf_rn(y2, -1.0f, y3) : __fmaf_rn(y1, -1.0f, y2);      // sub2(y3, y2) | sub2(y2, y1)
        q0 = __fmaf_rn(K::e0, d, y2);
        Q1 = __fmul_rn(K::e1, d);
        Q2 = 0.0f;
        return;
    }
    const FitCoeffs f = TYPED ? fit_coeffs(false, false) : fit_coeffs(first, last);
    if (!TYPED) { y1 = first ? 0.0f : y1; y3 = last ? 0.0f : y3; }
    const float d = __fmaf_rn(y3, -1.0f, y1);
    const float s = __fmaf_rn(y2, -2.0f, __fadd_rn(y1, y3));
    q0 = __fmaf_rn(f.q0_s, s, __fmaf_rn(f.q0_d, d, y2));
    Q1 = __fmaf_rn(f.q1_s, s, __fmul_rn(f.q1_d, d));
    Q2 = __fmul_rn(f.q2_s, s);
}

// Two intersections.  FIT selects the segment type at compile time (the edge types also skip the
// quadratic terms, which are exactly zero there: q2 = 0, kernel.c:134,160) or, for kFitDynamic,
// per-lane coefficients in the interior's form.  The coefficients are literals for the reference
// geometry with a static type and come from `f` otherwise (GEOM: per-segment geometry; kFitDynamic).
template <int EXPM, int FIT, bool GEOM>
__device__ __forceinline__ void attenuate_fast2(const FitCoeffs &f, float2 y1, float2 y2, float2 y3,
                                                float2 sigT, const float2 *s_pairs, float2 &psi,
                                                float2 &tally)
{
    constexpr bool kGiven = (FIT == kFitGiven) || (FIT == kFitGivenEdge);
    constexpr bool kQuadratic = (FIT == kFitInterior) || (FIT == kFitDynamic) || (FIT == kFitGiven);
    constexpr bool kFromF = GEOM || (FIT == kFitDynamic);
    static_assert(!(kGiven && GEOM), "the fit is only sweep-invariant with the constant geometry");
    using K = FitDiff;
    float2 q0, Q1, Q2 = f2(0.0f);
    if constexpr (kGiven) {
        q0 = y2;
        Q1 = y1;
        if constexpr (FIT == kFitGiven) Q2 = y3;
    } else if constexpr (kQuadratic) {
        // d = y1 - y3, s = y1 - 2 y2 + y3:  c1 = d / (2 dz), c2 = s / (2 dz^2)   (kernel.c:182-184)
        const float2 d = sub2(y1, y3);
        const float2 s = fma2(y2, f2(-2.0f), add2(y1, y3));
        q0 = fma2(f2(kFromF ? f.q0_s : K::q0_s), s, fma2(f2(kFromF ? f.q0_d : K::q0_d), d, y2));   // y2 + c1 zin + c2 zin^2
        Q1 = fma2(f2(kFromF ? f.q1_s : K::q1_s), s, mul2(f2(kFromF ? f.q1_d : K::q1_d), d));        // mu (c1 + 2 c2 zin)
        Q2 = mul2(f2(kFromF ? f.q2_s : K::q2_s), s);                                                // mu2 c2
    } else {
        // c1 = (y3 - y2) / dz (kernel.c:128) or (y2 - y1) / dz (kernel.c:154)
        const float2 d = (FIT == kFitFirst) ? sub2(y3, y2) : sub2(y2, y1);
        q0 = fma2(f2(GEOM ? f.q0_d : K::e0), d, y2);
        Q1 = mul2(f2(GEOM ? f.q1_d : K::e1), d);
    }

    const float2 tau = mul2(sigT, f2(GEOM ? f.ds : Geometry::ds));
    float2 e, omt;
    const float2 ev = exp_val2<EXPM, GEOM>(tau, s_pairs, e, omt);
    const float2 tme = sub2(tau, ev);                               // tau - expVal (exact)

    const float2 rs = make_float2(rcp_mufu(sigT.x), rcp_mufu(sigT.y));
    const float2 rs2 = mul2(rs, rs);
    // E = expVal / sigT and Fc = (tau - expVal) / sigT^2 serve both the flux integral
    //   (q0 tau + (sigT psi - q0) expVal) / sigT^2 = q0 Fc + psi E                  (kernel.c:248-249)
    // and the outgoing flux  t1 + t2 = q0 E + mu q1 Fc                              (kernel.c:291,301)
    const float2 E = mul2(ev, rs);
    const float2 Fc = mul2(tme, rs2);
    // reuse = tau (tau - 2) + 2 expVal / sigT^3 = ((1 - tau)^2 - 1) + 2 E / sigT^2   (kernel.c:235-236).
    // (1 - tau) comes rounded from the exponential: 6e-8 absolute on a sum that is >= 0.5 (2 expVal / sigT^3
    // >= 2 ds / sigT^2 - ..., while tau (tau - 2) >= -1), one operation instead of two.
    const float2 reuse = fma2(f2(2.0f), mul2(E, rs2), fma2(omt, omt, f2(-1.0f)));
    // The flux integral (kernel.c:248-251) and the outgoing flux (kernel.c:291-331) are both sums of four products
    // over (psi, mu2 q2, mu q1, q0).  They advance in lock step, largest-index term first, so that the two FMAs of
    // a step share their first operand: ptxas schedules them back to back with the operand held in the reuse
    // cache, and a three-register FFMA2 then costs 2 issue cycles instead of 3 (tools/ubench/fp2_operands.cu).
    float2 fi = mul2(psi, E);                                        // sigT psi expVal / sigT^2
    float2 acc = mul2(psi, e);                                       // t4 = psi (1 - expVal), kernel.c:321
    if constexpr (kQuadratic) {
        // tau (tau (tau - 3) + 6) - 6 expVal, kernel.c:250, in the reference's order of operations: its true value is
        // ~tau^4/4 while its terms are ~6 tau, so for small sigT it is pure rounding noise (cancellation factor
        // 24/tau^3) that, divided by 3 sigT^4, reaches 1e-4 of the dominant term, and parity needs the reference's own
        // roundings of the two large terms: 6 expVal is rounded before the subtraction (sub2 = fma(b, -1, a) adds
        // nothing to that).  Measured (gpurun r01a, reproduced by CPU emulation): with the subtraction contracted into
        // fma(-6, expVal, ...) 1.2e-4 L2-relative, this form 3.6e-8.  (ptxas does fuse tau (tau - 3) + 6 into one
        // FFMA2; that rounding is below the noise floor of the term.)
        const float2 cubic = sub2(mul2(tau, add2(mul2(tau, add2(tau, f2(-3.0f))), f2(6.0f))),
                                  mul2(f2(6.0f), ev));
        const float2 h3 = mul2(mul2(cubic, mul2(rs2, rs2)), f2(1.0f / 3.0f));           // cubic / (3 sigT^4)
        fi = fma2(Q2, h3, fi);                                                          // kernel.c:250-251
        acc = fma2(Q2, reuse, acc);                                                     // t3, kernel.c:311
    }
    fi = fma2(Q1, reuse, fi);                                        // term2, kernel.c:249
    acc = fma2(Q1, Fc, acc);                                         // t2, kernel.c:301
    fi = fma2(q0, Fc, fi);                                           // term1: q0 (tau - expVal) / sigT^2
    // kernel.c:262.  With the constant geometry the weight is the same for every segment and is applied once
    // to the summed tallies by finalize_flux (kTallyScaleConst; 0.5 is a power of two, so the result is
    // bit-identical to weighting every contribution); per-segment weights are applied here.
    if constexpr (GEOM) tally = mul2(f2(f.weight), fi);
    else tally = fi;
    psi = fma2(q0, E, acc);                                          // t1, kernel.c:291; sum kernel.c:331
}

// STRICT: one intersection in the reference's own operation order; g holds kernel.c:99-104
// (the literals for the reference geometry, the segment's values with GEOM).
template <int EXPM>
__device__ __forceinline__ void attenuate_strict(const SegGeometry &g, bool first, bool last, float y1, float y2,
                                                 float y3, float sigT, const float2 *s_pairs,
                                                 float &psi, float &tally)
{
    const float dz = g.dz, zin = g.zin, mu = g.mu, mu2 = g.mu2;
    float q0, q1, q2;
    if (first) {                                                    // kernel.c:111-135
        const float c1 = __fdiv_rn(__fsub_rn(y3, y2), dz);
        q0 = __fadd_rn(y2, __fmul_rn(c1, zin));
        q1 = c1;
        q2 = 0.0f;
    } else if (last) {                                              // kernel.c:137-161
        const float c1 = __fdiv_rn(__fsub_rn(y2, y1), dz);
        q0 = __fadd_rn(y2, __fmul_rn(c1, zin));
        q1 = c1;
        q2 = 0.0f;
    } else {                                                        // kernel.c:163-191
        const float two_dz = __fmul_rn(2.0f, dz);
        const float two_dz2 = __fmul_rn(two_dz, dz);
        const float c1 = __fdiv_rn(__fsub_rn(y1, y3), two_dz);
        const float c2 = __fdiv_rn(__fadd_rn(__fsub_rn(y1, __fmul_rn(2.0f, y2)), y3), two_dz2);
        q0 = __fadd_rn(__fadd_rn(y2, __fmul_rn(c1, zin)), __fmul_rn(__fmul_rn(c2, zin), zin));
        q1 = __fadd_rn(c1, __fmul_rn(__fmul_rn(2.0f, c2), zin));
        q2 = c2;
    }
    const float tau = __fmul_rn(sigT, g.ds);                        // kernel.c:206
    const float sigT2 = __fmul_rn(sigT, sigT);                      // kernel.c:207
    float e;
    const float ev = exp_val<EXPM>(tau, s_pairs, e);                // kernel.c:219/221

    const float reuse = __fadd_rn(__fmul_rn(tau, __fsub_rn(tau, 2.0f)),
                                  __fdiv_rn(__fmul_rn(2.0f, ev), __fmul_rn(sigT, sigT2)));
    const float term1 = __fdiv_rn(
        __fadd_rn(__fmul_rn(q0, tau), __fmul_rn(__fsub_rn(__fmul_rn(sigT, psi), q0), ev)), sigT2);
    const float term2 = __fmul_rn(__fmul_rn(q1, mu), reuse);
    const float cubic = __fsub_rn(
        __fmul_rn(tau, __fadd_rn(__fmul_rn(tau, __fsub_rn(tau, 3.0f)), 6.0f)), __fmul_rn(6.0f, ev));
    const float term3 = __fdiv_rn(__fmul_rn(__fmul_rn(q2, mu2), cubic),
                                  __fmul_rn(__fmul_rn(3.0f, sigT2), sigT2));
    const float flux_integral = __fadd_rn(__fadd_rn(term1, term2), term3);  // kernel.c:248-251
    tally = __fmul_rn(g.weight, flux_integral);                             // kernel.c:262

    const float t1 = __fdiv_rn(__fmul_rn(q0, ev), sigT);                               // :291
    const float t2 = __fdiv_rn(__fmul_rn(__fmul_rn(q1, mu), __fsub_rn(tau, ev)), sigT2);  // :301
    const float t3 = __fmul_rn(__fmul_rn(q2, mu2), reuse);                             // :311
    const float t4 = __fmul_rn(psi, __fsub_rn(1.0f, ev));                              // :321
    psi = __fadd_rn(__fadd_rn(__fadd_rn(t1, t2), t3), t4);                             // :331
}

__device__ __forceinline__ SegGeometry reference_geometry()
{
    return SegGeometry{Geometry::dz, Geometry::zin, Geometry::weight, Geometry::mu, Geometry::mu2, Geometry::ds};
}

}  // namespace smk
