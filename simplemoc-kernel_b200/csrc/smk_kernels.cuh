// smk_kernels.cuh -- sm_100a kernels of the segment-attenuation path.
//
//   attenuate_warp_track<GPL, EXPM, F64, GEOM, A32> the hot kernels: run_kernel's segment loop +
//       attenuate_segment (/root/reference/src/cpu/kernel.c:43-55, 75-333), ONE WARP PER TRACK, FAST
//       arithmetic.  GPL = energy groups per lane: 4 for 65..128 groups (128-bit loads, red.v4), 2 for
//       33..64 groups (64-bit loads, red.v2).
//   attenuate_warp_track_rec<EXPM, GEOM>         33..64 groups: attenuate_warp_track<2> fed by one
//       256-bit load per lane and segment from the same records
//   attenuate_record_tracks<LPT, GPL, EXPM, F64, GEOM> <= 32 groups, FAST: sub-warp tracks fed by one or two
//       256-bit loads per lane and segment from gather records (build_records)
//   attenuate_tracks<LPT, NCHUNK, MATH, EXPM, GEOM> the general kernel: sub-warp tracks for <= 32 groups,
//       blocks of 256 groups for > 128 groups (any group count), and the STRICT verification arithmetic
//   fill_rows                                    device-side deterministic fill
//       (replaces /root/reference/src/cpu/init.c:64-75 + the H2D of init.cu:105-127)
//   pad_rows / finalize_flux[64]                 host layout <-> padded device layout
//   allreduce_peer_slices[64]                    multi-GPU all-reduce of the tallies over NVLink peer memory
//
// Work decomposition of the hot kernels
//   track  = seg_per_track consecutive segments sharing one carried angular flux psi
//   a track is owned by LPT lanes of one warp (LPT = lanes per track, a power of two);
//   each lane keeps the psi of its groups in registers for the whole track.
//   Warps claim tracks dynamically from a global counter (claim_tracks).
//   Segment ids come from the counter stream: every LPT segments each lane of the
//   track hashes ONE upcoming segment and the ids are handed round with shuffles, so
//   the Philox cost per intersection is 1/(groups per lane * LPT) of a hash.
//   The FAST arithmetic is packed FP32x2 (FFMA2/FMUL2/FADD2), see smk_math.cuh.
//
// The variants that were measured and rejected (TMA-staged rows, register / L1 prefetch, deferred RED,
// DESIGN.md section 5.3) live in smk_kernels_tuning.cuh and are only compiled with -DSMK_TUNING.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#include "smk_math.cuh"
#include "smk_stream.cuh"

namespace smk {

struct KernelArgs {
    const float4 *__restrict__ source;   // [R][F][G_pad/4]
    const float4 *__restrict__ sigT;     // [R][G_pad/4]
    float *__restrict__ tally;           // [replicas][R][F][G_pad]
    int64_t replica_stride;              // floats between tally replicas (R*F*G_pad)
    int32_t replicas;                    // >1 for few-row problems: warp w tallies into replica w % replicas
    float *__restrict__ psi_out;         // [tracks in launch][G_pad] or nullptr
    unsigned long long *checksum;        // indexing fingerprint accumulator
    unsigned long long *work_counter;    // next unclaimed work item (relative to track_begin), zeroed per launch
    double *tally64;                     // diagnostic: f64 tally accumulators [R][F][G_pad] (SMK_FLAG_TALLY_F64)
    int64_t segments;                    // N
    int64_t track_begin, track_end;
    uint64_t seed;
    PhiloxKeys keys;                     // expanded Philox key schedule of `seed`
    FastMod mod_regions, mod_fai;
    int32_t fai_count;                   // F
    int32_t row_f4;                      // G_pad / 4: float4 per row
    int32_t seg_per_track;               // p
    int32_t group_blocks;                // general kernel, > 256 groups: a row is this many blocks of 256 groups,
                                         // each (track, block) is swept by its own warp (groups are independent)
    GeometryBase geom;                   // SMK_FLAG_SEGMENT_GEOMETRY: base values + spread (kernel.c:99-104)
    MeshConsts mesh;                     // 1/(2dz), 1/(2dz^2), 1/dz
    const float *__restrict__ records;   // attenuate_record_tracks: gather records [R*F][G_pad/2][8] (build_records)
};

#ifndef SMK_THREADS_PER_BLOCK
#define SMK_THREADS_PER_BLOCK 256
#endif
constexpr int kThreadsPerBlock = SMK_THREADS_PER_BLOCK;
#ifndef SMK_MIN_BLOCKS_FAST
#define SMK_MIN_BLOCKS_FAST (1024 / SMK_THREADS_PER_BLOCK)
#endif
// 4 x 256 threads x 64 registers = the whole register file: 32 warps/SM for the FAST kernels
constexpr int kMinBlocksFast = SMK_MIN_BLOCKS_FAST;
#ifndef SMK_MIN_BLOCKS_GEOM
#define SMK_MIN_BLOCKS_GEOM SMK_MIN_BLOCKS_FAST
#endif
constexpr int kMinBlocksGeom = SMK_MIN_BLOCKS_GEOM;

__device__ __forceinline__ void red_add_v4(float4 *addr, float a, float b, float c, float d)
{
#if defined(SMK_EXPERIMENT_NO_RED)   // timing experiment only: plain store instead of the reduction
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
#else
    // one 16-byte vector reduction at L2 per lane (PTX ISA 8.1, sm_90+): SASS RED.E.ADD.F32x4
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :
                 : "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
#endif
}

__device__ __forceinline__ void red_add_v2(float2 *addr, float a, float b)
{
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// diagnostic f64 tallies: order-independent to ~1e-16, the yardstick for fp32 accumulation noise
__device__ __forceinline__ void red_add_f64(double *addr, float v)
{
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr), "d"((double)v) : "memory");
}

// Tally-contention relief (BASELINE config 4: few source regions): when the tally array has few
// rows, several warps hit the same L2 atomic address at once and serialise there.  The library then
// keeps `replicas` copies of the (small) array; every warp adds into copy (warp % replicas) and
// finalize_flux sums the copies.  replicas == 1 for the normal problem sizes.
__device__ __forceinline__ float *warp_tally(const KernelArgs &a, int64_t warp_global)
{
    return a.tally + (a.replicas > 1 ? (warp_global % a.replicas) * a.replica_stride : 0);
}

// Dynamic track scheduling: warps claim work from a global counter instead of striding statically.
// All CTAs of the persistent grid are resident from the start, so with static striding an SM that runs
// slower than the others (far L2 partition, fewer co-resident CTAs) sets the kernel time while the fast
// ones idle: ncu showed 22.5 of 32 warps active on average.  One 64-bit atomic per track (100 segments).
__device__ __forceinline__ int64_t claim_work(const KernelArgs &a, int lane, int n)
{
    unsigned long long first = 0ull;
    if (lane == 0) first = atomicAdd(a.work_counter, (unsigned long long)n);
    return (int64_t)__shfl_sync(0xFFFFFFFFu, first, 0);
}

__device__ __forceinline__ float4 ldg4(const float4 *p)
{
#ifdef SMK_EXPERIMENT_NO_LDG   // ceiling experiment only: synthesise the row from the address, no memory access
    const uint32_t u = (uint32_t)reinterpret_cast<uintptr_t>(p);
    const float v = __uint_as_float(0x3f000000u | ((u >> 4) & 0x7FFFFFu));
    return make_float4(v, v * 0.5f, v * 0.25f, v * 0.75f);
#else
    return __ldg(p);
#endif
}

// ------------------------------------------------------------------------------
// attenuate_warp_track<GPL, EXPM, F64, GEOM, A32>: one track per warp, GPL groups per lane; A32 = every array is
// smaller than 4 GB, addresses from 32-bit byte offsets (ptr_add_index).
//
// Per 32 segments every lane hashes ONE segment of the batch into two words
//     pk = row * 32                              row = QSR_id * F + FAI_id, in units of the lane vector
//     sg = QSR_id * 32 | first << 31 | last << 30   index of the sigT row + the segment type
// (a padded row is 32 lane vectors, so `| lane` completes either index).  Inside the batch the words are
// broadcast with two shuffles; the segment type (first / interior / last fine axial interval) is warp-uniform,
// so the warp branches into one of three straight-line bodies with literal fit coefficients: each loads only
// the rows its type reads and the edge bodies skip the quadratic terms.  The segment loop is unrolled twice
// (kSegmentUnroll; +0.7 % at 128 groups, +1.5 % at 64, profiles/ab_r02.md).
// With GEOM the hashing lane also derives the segment's geometry and fit coefficients and parks them in
// shared memory; the warp reads them back with broadcast loads.
// ------------------------------------------------------------------------------
// base + idx * sizeof(T) for a 32-bit element index.  Left to itself ptxas emits IMAD.WIDE.U32 for every such
// address (and turns a hand-written shift / add-with-carry sequence back into one), and IMAD runs on the FMA-heavy
// pipe -- the unit that binds the FAST kernels (82 % busy, profiles/ncu_r02_default.md).  With A32 the BYTE offset is
// formed in 32-bit arithmetic and added with an explicit add / add-with-carry pair, which ptxas keeps on the integer
// ALU (SASS: LEA + IADD3.X from a uniform-register base): +3 % at 128 groups, +6 % at 64 (profiles/ab_r02.md).
// Bits of idx that the 32-bit product shifts out (the type flags of `sg`) vanish for free.  The library selects
// A32 when every array is smaller than 4 GB (all BASELINE configurations and the 4.6 GB HBM-resident set, whose
// largest array is 1.5 GB), the plain form otherwise.
template <bool A32, typename T>
__device__ __forceinline__ T *ptr_add_index(T *base, uint32_t idx)
{
    if constexpr (A32) {
        const uint64_t b = reinterpret_cast<uint64_t>(base);
        const uint32_t bytes = idx * (uint32_t)sizeof(T);
        uint32_t lo, hi;
        asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, 0;"
            : "=r"(lo), "=r"(hi)
            : "r"((uint32_t)b), "r"((uint32_t)(b >> 32)), "r"(bytes));
        return reinterpret_cast<T *>(((uint64_t)hi << 32) | lo);
    } else {
        return base + idx;
    }
}

template <int GPL>
struct LaneVec;
template <>
struct LaneVec<4> {
    typedef float4 type;
    static __device__ __forceinline__ float4 load(const float4 *p) { return ldg4(p); }
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <>
struct LaneVec<2> {
    typedef float2 type;
    static __device__ __forceinline__ float2 load(const float2 *p) { return __ldg(p); }
    static __device__ __forceinline__ float2 zero() { return make_float2(0.f, 0.f); }
};

template <int EXPM, int FIT, bool GEOM>
__device__ __forceinline__ void attenuate_lane(const FitCoeffs &fc, float4 y1, float4 y2, float4 y3, float4 st,
                                               const float2 *s_pairs, float4 &psi, float4 &t)
{
    float2 p_lo = make_float2(psi.x, psi.y), p_hi = make_float2(psi.z, psi.w), t_lo, t_hi;
    attenuate_fast2<EXPM, FIT, GEOM>(fc, make_float2(y1.x, y1.y), make_float2(y2.x, y2.y), make_float2(y3.x, y3.y),
                                     make_float2(st.x, st.y), s_pairs, p_lo, t_lo);
    attenuate_fast2<EXPM, FIT, GEOM>(fc, make_float2(y1.z, y1.w), make_float2(y2.z, y2.w), make_float2(y3.z, y3.w),
                                     make_float2(st.z, st.w), s_pairs, p_hi, t_hi);
    psi = make_float4(p_lo.x, p_lo.y, p_hi.x, p_hi.y);                 // kernel.c:331
    t = make_float4(t_lo.x, t_lo.y, t_hi.x, t_hi.y);
}

template <int EXPM, int FIT, bool GEOM>
__device__ __forceinline__ void attenuate_lane(const FitCoeffs &fc, float2 y1, float2 y2, float2 y3, float2 st,
                                               const float2 *s_pairs, float2 &psi, float2 &t)
{
    attenuate_fast2<EXPM, FIT, GEOM>(fc, y1, y2, y3, st, s_pairs, psi, t);
}

// FSR_flux[g] += tally[g] (kernel.c:276): one vector RED per lane, or f64 REDs for the diagnostic tallies
template <bool F64, bool A32>
__device__ __forceinline__ void tally_lane(float *tally, double *tally64, uint32_t idx, const float4 &t)
{
    if constexpr (F64) {
        double *d = tally64 + (size_t)idx * 4;
        red_add_f64(d, t.x); red_add_f64(d + 1, t.y); red_add_f64(d + 2, t.z); red_add_f64(d + 3, t.w);
    } else {
        red_add_v4(ptr_add_index<A32>(reinterpret_cast<float4 *>(tally), idx), t.x, t.y, t.z, t.w);
    }
}
template <bool F64, bool A32>
__device__ __forceinline__ void tally_lane(float *tally, double *tally64, uint32_t idx, const float2 &t)
{
    if constexpr (F64) {
        double *d = tally64 + (size_t)idx * 2;
        red_add_f64(d, t.x); red_add_f64(d + 1, t.y);
    } else {
        red_add_v2(ptr_add_index<A32>(reinterpret_cast<float2 *>(tally), idx), t.x, t.y);
    }
}

#ifndef SMK_UNROLL_K
#define SMK_UNROLL_K 2
#endif
constexpr int kSegmentUnroll = SMK_UNROLL_K;     // segment loop of attenuate_warp_track (tuning knob)
constexpr uint32_t kSgFirst = 0x80000000u, kSgLast = 0x40000000u;   // warp_track: type flags above the sigT index
constexpr uint32_t kRowFirst = 1u << 30, kRowLast = 1u << 31;     // general kernel: flags above the row index

#ifndef SMK_MIN_BLOCKS_HALF
#define SMK_MIN_BLOCKS_HALF SMK_MIN_BLOCKS_FAST
#endif
// two groups per lane need fewer registers (46): more resident warps make up for the single dependency chain per lane
constexpr int kMinBlocksHalf = SMK_MIN_BLOCKS_HALF;

template <int GPL, int EXPM, bool F64, bool GEOM, bool A32>
__global__ void __launch_bounds__(kThreadsPerBlock, GEOM ? kMinBlocksGeom : (GPL == 2 ? kMinBlocksHalf : kMinBlocksFast))
attenuate_warp_track(const KernelArgs a)
{
    typedef typename LaneVec<GPL>::type V;
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int kWarps = kThreadsPerBlock / 32;
    constexpr uint32_t ROWV = 32;                        // lane vectors per padded row (G_pad = 32 * GPL)

    __shared__ float2 s_pairs[kTableReach];
    // GEOM: per warp, per segment of the batch: {ds, weight, q0_d, q0_s} {q1_d, q1_s, q2_s, -}
    __shared__ float4 s_coef[GEOM ? kWarps : 1][GEOM ? 32 : 1][2];
    if constexpr (EXPM == kExpTable) {
        if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
        __syncthreads();
    }
    // read once and opaque to the compiler: otherwise every `| lane` is re-derived from threadIdx.x as a
    // fourth LOP3 input and costs an extra instruction per address
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const int warp = threadIdx.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * kWarps + warp;
    const uint32_t F = (uint32_t)a.fai_count;
    const int p = a.seg_per_track;
    const V *const source = reinterpret_cast<const V *>(a.source);
    const V *const sigT = reinterpret_cast<const V *>(a.sigT);
    float *const tally = warp_tally(a, warp_global);
    const int64_t n_tracks = a.track_end - a.track_begin;
    unsigned long long checksum = 0ull;

    for (int64_t w = claim_work(a, lane, 1); w < n_tracks; w = claim_work(a, lane, 1)) {
        const int64_t track = a.track_begin + w;
        const int64_t s0 = track * p;
        const int64_t left = a.segments - s0;
        const int nseg = left < p ? (int)left : p;

        // incoming angular flux of the track (kernel.c:29-30): one Philox block covers 4 groups
        V psi;
        if constexpr (GPL == 4) {
            const u32x4 r = stream_words(a.keys, (uint64_t)track, (uint32_t)lane, kDomainPsi);
            psi = make_float4(u01(r.x), u01(r.y), u01(r.z), u01(r.w));
        } else {
            const u32x4 r = stream_words(a.keys, (uint64_t)track, (uint32_t)(lane >> 1), kDomainPsi);
            psi = (lane & 1) ? make_float2(u01(r.z), u01(r.w)) : make_float2(u01(r.x), u01(r.y));
        }

        for (int b = 0; b < nseg; b += 32) {
            // this lane's segment of the batch
            uint32_t my_pk = 0u, my_sg = 0u;
            if (b + lane < nseg) {
                const uint64_t seg = (uint64_t)(s0 + b + lane);
                const u32x4 r = stream_words(a.keys, seg, 0u, kDomainSegment);
                const uint32_t qsr = fastmod(r.x >> 1, a.mod_regions);       // kernel.c:47
                const uint32_t fai = fastmod(r.y >> 1, a.mod_fai);           // kernel.c:50
                checksum += checksum_term(qsr, fai, F, seg);
                // type flags in the two top bits of the sigT index (smk_create checks R * 32 < 2^30): `first` is a
                // sign test, and the row index needs no masking
                my_pk = (qsr * F + fai) * ROWV;
                my_sg = (qsr * ROWV) | (fai == 0u ? kSgFirst : 0u) | (fai == F - 1u ? kSgLast : 0u);
                if constexpr (GEOM) {
                    const SegGeometry g = segment_geometry(a.geom, r.z, r.w);
                    const FitCoeffs f = fit_coeffs_geom_typed(g, a.mesh, fai == 0u || fai == F - 1u);
                    s_coef[warp][lane][0] = make_float4(f.ds, f.weight, f.q0_d, f.q0_s);
                    s_coef[warp][lane][1] = make_float4(f.q1_d, f.q1_s, f.q2_s, 0.0f);
                }
            }
            if constexpr (GEOM) __syncwarp();
            const int count = (nseg - b) < 32 ? (nseg - b) : 32;
#pragma unroll kSegmentUnroll
            for (int k = 0; k < count; ++k) {
                const uint32_t pk = __shfl_sync(kFull, my_pk, k);
                const uint32_t idx = pk | (uint32_t)lane;
                const uint32_t sg = __shfl_sync(kFull, my_sg, k);
                const V *src = ptr_add_index<A32>(source, idx);
                const V st = LaneVec<GPL>::load(ptr_add_index<A32>(sigT, (sg & ~(kSgFirst | kSgLast)) | (uint32_t)lane));
                FitCoeffs fc = {};
                if constexpr (GEOM) {
                    const float4 c0 = s_coef[warp][k][0], c1 = s_coef[warp][k][1];
                    fc.ds = c0.x; fc.weight = c0.y; fc.q0_d = c0.z; fc.q0_s = c0.w;
                    fc.q1_d = c1.x; fc.q1_s = c1.y; fc.q2_s = c1.z;
                }
                V t;
                if ((int32_t)sg < 0) {
                    const V y2 = LaneVec<GPL>::load(src);
                    const V y3 = LaneVec<GPL>::load(src + ROWV);
                    attenuate_lane<EXPM, kFitFirst, GEOM>(fc, LaneVec<GPL>::zero(), y2, y3, st, s_pairs, psi, t);
                } else if (sg & kSgLast) {
                    const V y1 = LaneVec<GPL>::load(src - ROWV);
                    const V y2 = LaneVec<GPL>::load(src);
                    attenuate_lane<EXPM, kFitLast, GEOM>(fc, y1, y2, LaneVec<GPL>::zero(), st, s_pairs, psi, t);
                } else {
                    const V y1 = LaneVec<GPL>::load(src - ROWV);
                    const V y2 = LaneVec<GPL>::load(src);
                    const V y3 = LaneVec<GPL>::load(src + ROWV);
                    attenuate_lane<EXPM, kFitInterior, GEOM>(fc, y1, y2, y3, st, s_pairs, psi, t);
                }
                tally_lane<F64, A32>(tally, a.tally64, idx, t);                             // kernel.c:276
            }
            if constexpr (GEOM) __syncwarp();       // everyone is done with s_coef before the next batch overwrites it
        }

        if (a.psi_out != nullptr)
            reinterpret_cast<V *>(a.psi_out)[(track - a.track_begin) * ROWV + lane] = psi;
    }

    // one 64-bit atomic per warp for the indexing fingerprint
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) checksum += __shfl_xor_sync(kFull, checksum, off);
    if (lane == 0 && checksum != 0ull) atomicAdd(a.checksum, checksum);
}

// ------------------------------------------------------------------------------
// attenuate_tracks<LPT, NCHUNK, MATH, EXPM, GEOM>: the general kernel.
//   a track is owned by LPT lanes; each lane owns NCHUNK float4 = 4*NCHUNK groups per block of
//   4*LPT*NCHUNK groups.  G <= 32: LPT = 1..8, 32/LPT tracks per warp, per-lane segment types.
//   G > 128 (FAST) / any G (STRICT): LPT = 32; rows wider than one block are split into
//   `group_blocks` blocks and every (track, block) pair is an independent work item (the energy groups
//   of attenuate_segment never interact, kernel.c:116-332), so any group count is supported.
// ------------------------------------------------------------------------------
template <int LPT, int NCHUNK, int MATH, int EXPM, bool GEOM>
__global__ void __launch_bounds__(kThreadsPerBlock, (NCHUNK == 1 && MATH == kMathFast && !GEOM) ? kMinBlocksFast : 1)
attenuate_tracks(const KernelArgs a)
{
    static_assert(LPT >= 1 && LPT <= 32 && (LPT & (LPT - 1)) == 0, "LPT must be a power of two");
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int kSlotsPerWarp = 32 / LPT;
    constexpr int kBlockF4 = LPT * NCHUNK;           // float4 per group block

    __shared__ float2 s_pairs[kTableReach];
    if constexpr (EXPM == kExpTable) {
        if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
        __syncthreads();
    }

    const int lane = threadIdx.x & 31;
    const int sub = lane & (LPT - 1);               // lane within its track
    const int64_t warp_global = (int64_t)blockIdx.x * (kThreadsPerBlock / 32) + (threadIdx.x >> 5);

    const int F = a.fai_count;
    // sub-warp tracks: a padded row is exactly one group block, so its size is a compile-time constant
    // (row offsets become shifts and the neighbouring rows immediate offsets instead of FMA-pipe IMADs)
    const int row_f4 = (LPT < 32) ? kBlockF4 : a.row_f4;
    const int p = a.seg_per_track;
    const int nblk = (LPT == 32) ? a.group_blocks : 1;
    float *const tally = warp_tally(a, warp_global);
    const int64_t n_work = (a.track_end - a.track_begin) * nblk;
    unsigned long long checksum = 0ull;

    for (int64_t wbase = claim_work(a, lane, kSlotsPerWarp); wbase < n_work; wbase = claim_work(a, lane, kSlotsPerWarp)) {
        const int64_t work = wbase + (lane / LPT);
        const bool tvalid = work < n_work;
        const int64_t track = a.track_begin + work / nblk;
        const int blk = (int)(work % nblk);
        const int boff = blk * kBlockF4;            // float4 offset of this group block inside a row
        const int64_t s0 = track * p;
        int nseg = 0;
        if (tvalid) {
            const int64_t left = a.segments - s0;
            nseg = left < p ? (int)left : p;
        }

        // incoming angular flux of the track (kernel.c:29-30), 4 groups per Philox block
        float4 psi[NCHUNK];
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
            const u32x4 w = stream_words(a.keys, (uint64_t)track, (uint32_t)(boff + c * LPT + sub), kDomainPsi);
            psi[c] = make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w));
        }

        const int nseg_warp = (LPT == 32) ? nseg : __reduce_max_sync(kFull, nseg);

        for (int b = 0; b < nseg_warp; b += LPT) {
            // each lane of the track draws the ids of one of the next LPT segments
            // packed: row = QSR_id * F + FAI_id (< 2^29: smk_create checks R * F * G_pad / 4 < 2^31), first / last flags
            // lanes without a segment (ragged last track, short batch) keep both flags set: no neighbouring row is read
            uint32_t my_qsr = 0u, my_row = kRowFirst | kRowLast, my_w2 = 0u, my_w3 = 0u;
            if (b + sub < nseg) {
                const uint64_t seg = (uint64_t)(s0 + b + sub);
                const u32x4 w = stream_words(a.keys, seg, 0u, kDomainSegment);
                my_qsr = fastmod(w.x >> 1, a.mod_regions);                   // kernel.c:47
                const uint32_t my_fai = fastmod(w.y >> 1, a.mod_fai);        // kernel.c:50
                my_row = (my_qsr * (uint32_t)F + my_fai) | (my_fai == 0u ? kRowFirst : 0u) |
                         (my_fai == (uint32_t)(F - 1) ? kRowLast : 0u);
                my_w2 = w.z;
                my_w3 = w.w;
                if (blk == 0) checksum += checksum_term(my_qsr, my_fai, (uint32_t)F, seg);
            }
            const int count = (nseg_warp - b) < LPT ? (nseg_warp - b) : LPT;
            for (int k = 0; k < count; ++k) {
                const uint32_t qsr = __shfl_sync(kFull, my_qsr, k, LPT);
                const uint32_t prow = __shfl_sync(kFull, my_row, k, LPT);
                // only the stream's ragged last track can be shorter than its warp-mates
                const bool active = (LPT == 32) ? true : (b + k) < nseg;
                const bool first = (prow & kRowFirst) != 0u;
                const bool last = (prow & kRowLast) != 0u;
                const uint32_t row = prow & ~(kRowFirst | kRowLast);
                const uint32_t off = row * (uint32_t)row_f4 + (uint32_t)(boff + sub);
                const float4 *src = a.source + off;
                const float4 *sig = a.sigT + (qsr * (uint32_t)row_f4 + (uint32_t)(boff + sub));

                SegGeometry g = reference_geometry();
                if constexpr (GEOM) {
                    const uint32_t w2 = __shfl_sync(kFull, my_w2, k, LPT);
                    const uint32_t w3 = __shfl_sync(kFull, my_w3, k, LPT);
                    g = segment_geometry(a.geom, w2, w3);
                }
                // LPT == 32 (blocks of 256 groups): one track per warp, so the segment type is warp-uniform and the
                // statically typed bodies apply (edge bodies skip the quadratic terms: 29 instead of 45 operations)
                constexpr bool kTyped = (LPT == 32);
                FitCoeffs fc = {};
                if constexpr (MATH == kMathFast) {
                    if constexpr (kTyped) { if constexpr (GEOM) fc = fit_coeffs_geom_typed(g, a.mesh, first || last); }
                    else fc = GEOM ? fit_coeffs_geom(g, a.mesh, first, last) : fit_coeffs(first, last);
                }

#pragma unroll
                for (int c = 0; c < NCHUNK; ++c) {
                    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 y2 = ldg4(src + c * LPT);
                    const float4 st = ldg4(sig + c * LPT);
                    const float4 y1 = first ? zero : ldg4(src + c * LPT - row_f4);
                    const float4 y3 = last ? zero : ldg4(src + c * LPT + row_f4);
                    float4 t, ps = psi[c];
                    if constexpr (MATH == kMathFast && kTyped) {
                        if (first) attenuate_lane<EXPM, kFitFirst, GEOM>(fc, y1, y2, y3, st, s_pairs, ps, t);
                        else if (last) attenuate_lane<EXPM, kFitLast, GEOM>(fc, y1, y2, y3, st, s_pairs, ps, t);
                        else attenuate_lane<EXPM, kFitInterior, GEOM>(fc, y1, y2, y3, st, s_pairs, ps, t);
                    } else if constexpr (MATH == kMathFast) {
                        // tracks of different types share a warp -> per-lane coefficients
                        attenuate_lane<EXPM, kFitDynamic, GEOM>(fc, y1, y2, y3, st, s_pairs, ps, t);
                    } else {
                        attenuate_strict<EXPM>(g, first, last, y1.x, y2.x, y3.x, st.x, s_pairs, ps.x, t.x);
                        attenuate_strict<EXPM>(g, first, last, y1.y, y2.y, y3.y, st.y, s_pairs, ps.y, t.y);
                        attenuate_strict<EXPM>(g, first, last, y1.z, y2.z, y3.z, st.z, s_pairs, ps.z, t.z);
                        attenuate_strict<EXPM>(g, first, last, y1.w, y2.w, y3.w, st.w, s_pairs, ps.w, t.w);
                    }
                    if (active) {
                        psi[c] = ps;                                                  // kernel.c:331
                        if (a.tally64 == nullptr) {
                            red_add_v4(reinterpret_cast<float4 *>(tally) + off + c * LPT, t.x, t.y, t.z, t.w);   // kernel.c:276
                        } else {
                            double *d = a.tally64 + ((size_t)off + c * LPT) * 4;
                            red_add_f64(d, t.x); red_add_f64(d + 1, t.y); red_add_f64(d + 2, t.z); red_add_f64(d + 3, t.w);
                        }
                    }
                }
            }
        }

        if (a.psi_out != nullptr && tvalid) {
            float4 *out = reinterpret_cast<float4 *>(a.psi_out) + (track - a.track_begin) * row_f4 + boff + sub;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) out[c * LPT] = psi[c];
        }
    }

    // one 64-bit atomic per warp for the indexing fingerprint
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) checksum += __shfl_xor_sync(kFull, checksum, off);
    if (lane == 0 && checksum != 0ull) atomicAdd(a.checksum, checksum);
}

// ------------------------------------------------------------------------------
// attenuate_record_tracks<LPT, GPL, EXPM, F64, GEOM>: sub-warp tracks (G_pad <= 32), FAST arithmetic, fed from
// GATHER RECORDS instead of the row arrays.
//
// With few groups a track is a handful of lanes and every lane group gathers its own rows: each of the 3.6 row
// loads + 1 RED per (track, segment) is its own L1 wavefront (one 32-byte sector per lane pair), and at 7 groups
// the general kernel sits against the L1 wavefront rate (L1/TEX 89 % busy, profiles/ncu_r02_g7_general.md) with the
// FMA pipe half idle.  build_records lays everything one (segment, pair of groups) reads side by side,
//     rec[row = QSR * F + FAI][j] = { sigT[QSR][2j..2j+1], y[FAI-1][2j..2j+1], y[FAI][2j..2j+1], y[FAI+1][2j..2j+1] }
// (32 bytes; the missing neighbour of an edge interval is 0, exactly what the general kernel substitutes), so a lane
// fetches its segment with ONE 256-bit load (SASS LDG.E.256) and the lanes of a track read one contiguous
// 16 * G_pad-byte record row (one 128-byte line at 7 groups): 1 load wavefront per (track, segment) instead of 3.6,
// 1 load instruction per lane instead of 3-4, no sigT index.  GPL = 4 keeps four groups per lane (two 256-bit
// loads: {sigT, y[FAI]} and {y[FAI-1], y[FAI+1]}).  The records are a derived copy rebuilt by the library at every
// launch from the canonical rows (build_records: 4 x the source array, a few microseconds), so uploads, row-range
// transfers and external writes into the device arrays keep their meaning.  The arithmetic is attenuate_fast2 with
// per-lane fit coefficients, i.e. bit-identical per intersection to attenuate_tracks<.., kMathFast, ..>.
// The coefficients come from shared memory: with the constant geometry a 4-entry table indexed by the segment type
// (tracks of different types share a warp); with per-segment geometry (GEOM) the lane that hashes a segment derives
// them once from stream words 2,3 (segment_geometry + fit_coeffs_geom, type folded in) and parks them in its slot,
// and the lanes of the track read the slot of the step's hashing lane (the general kernel makes every lane derive
// them from shuffled words): 7 groups 3.29e11 -> 3.59e11, 29 groups 4.12e11 -> 5.05e11 (profiles/ab_r02.md r02o).
// ------------------------------------------------------------------------------
struct __align__(32) Rec8 {
    float2 a, b, c, d;
};

__device__ __forceinline__ Rec8 ldg256(const void *p)
{
    Rec8 r;
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.c.x), "=f"(r.c.y), "=f"(r.d.x), "=f"(r.d.y)
                 : "l"(p));
    return r;
}

// FORM 0: the raw values; FORM 1 / 2 (SMK_FLAG_FIT_PER_SWEEP): the fitted {q0, mu q1, mu2 q2} in the slots of
// {y[FAI], y[FAI-1], y[FAI+1]}, evaluated by fit_row in the per-lane-coefficient form (1: attenuate_record_tracks) or the
// statically typed form (2: attenuate_warp_track_rec) -- bit for bit what the sweep kernel would compute per segment.
template <int GPL, int FORM>
__global__ void build_records(const float *__restrict__ source, const float *__restrict__ sigT, float *__restrict__ rec,
                              int64_t rows, int fai_count, int groups_pad)
{
    // one thread per (row, GPL groups): GPL = 2: one 32-byte record; GPL = 4: two
    const int per_row = groups_pad / GPL;
    const int64_t n = rows * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / per_row;
        const int j = (int)(i - row * per_row);
        const int64_t qsr = row / fai_count;
        const int fai = (int)(row - qsr * fai_count);
        const float *yc = source + row * groups_pad + j * GPL;
        const float *st = sigT + qsr * groups_pad + j * GPL;
        float *out = rec + i * (4 * GPL);
        float m[GPL], c[GPL], p[GPL], s[GPL];
#pragma unroll
        for (int g = 0; g < GPL; ++g) {
            s[g] = st[g];
            c[g] = yc[g];
            m[g] = fai > 0 ? yc[g - groups_pad] : 0.0f;
            p[g] = fai < fai_count - 1 ? yc[g + groups_pad] : 0.0f;
            if constexpr (FORM != 0) {
                float q0, Q1, Q2;
                fit_row<FORM == 2>(fai == 0, fai == fai_count - 1, m[g], c[g], p[g], q0, Q1, Q2);
                c[g] = q0; m[g] = Q1; p[g] = Q2;
            }
        }
        if constexpr (GPL == 2) {
            out[0] = s[0]; out[1] = s[1]; out[2] = m[0]; out[3] = m[1];
            out[4] = c[0]; out[5] = c[1]; out[6] = p[0]; out[7] = p[1];
        } else {
            // {sigT[4], y[FAI][4]} {y[FAI-1][4], y[FAI+1][4]}
#pragma unroll
            for (int g = 0; g < 4; ++g) { out[g] = s[g]; out[4 + g] = c[g]; out[8 + g] = m[g]; out[12 + g] = p[g]; }
        }
    }
}

#ifndef SMK_MIN_BLOCKS_REC
#define SMK_MIN_BLOCKS_REC SMK_MIN_BLOCKS_FAST
#endif
constexpr int kMinBlocksRec = SMK_MIN_BLOCKS_REC;
#ifndef SMK_REC_UNROLL
#define SMK_REC_UNROLL 2
#endif
constexpr int kRecordUnroll = SMK_REC_UNROLL;     // segment loop of attenuate_record_tracks (tuning knob)

// one segment of one lane of attenuate_record_tracks.  CHECK: the lane may be past its track's end (the stream's ragged
// last track, or a warp slot without a track): it computes along with its warp-mates, tallies nothing, keeps its psi.
// `fit` points at the coefficients of the segment: {q0_d, q0_s, q1_d, q1_s} {q2_s, ds, weight, -}
template <int LPT, int GPL, int EXPM, bool F64, bool GEOM, bool HOIST, bool CHECK>
__device__ __forceinline__ void record_segment(const char *rec, const float4 *fit, const float2 *s_pairs, float *tally,
                                               double *tally64, uint32_t pidx, int sub, bool active,
                                               typename LaneVec<GPL>::type &psi)
{
    typedef typename LaneVec<GPL>::type V;
    const uint32_t idx = (HOIST ? pidx : (pidx & ~(kRowFirst | kRowLast))) | (uint32_t)sub;
    const char *r = ptr_add_index<true>(rec, idx * (16u * GPL));
    constexpr int kFit = HOIST ? kFitGiven : kFitDynamic;      // HOIST: the records hold the fitted coefficients
    FitCoeffs fc = {};
    if constexpr (!HOIST) {
        const float4 f0 = fit[0];
        fc.q0_d = f0.x; fc.q0_s = f0.y; fc.q1_d = f0.z; fc.q1_s = f0.w;
        if constexpr (GEOM) {
            const float4 f1 = fit[1];
            fc.q2_s = f1.x; fc.ds = f1.y; fc.weight = f1.z;
        } else {
            fc.q2_s = *reinterpret_cast<const float *>(fit + 1);
        }
    }
    if constexpr (!GEOM) { fc.ds = Geometry::ds; fc.weight = Geometry::weight; }
    V t;
    if constexpr (GPL == 2) {
        const Rec8 q = ldg256(r);
        float2 ps = psi;
        attenuate_fast2<EXPM, kFit, GEOM>(fc, q.b, q.c, q.d, q.a, s_pairs, ps, t);
        if (!CHECK || active) psi = ps;                                       // kernel.c:331
    } else {
        const Rec8 q0 = ldg256(r), q1 = ldg256(r + 32);
        float2 p_lo = make_float2(psi.x, psi.y), p_hi = make_float2(psi.z, psi.w), t_lo, t_hi;
        attenuate_fast2<EXPM, kFit, GEOM>(fc, q1.a, q0.c, q1.c, q0.a, s_pairs, p_lo, t_lo);
        attenuate_fast2<EXPM, kFit, GEOM>(fc, q1.b, q0.d, q1.d, q0.b, s_pairs, p_hi, t_hi);
        if (!CHECK || active) psi = make_float4(p_lo.x, p_lo.y, p_hi.x, p_hi.y);
        t = make_float4(t_lo.x, t_lo.y, t_hi.x, t_hi.y);
    }
    if (!CHECK || active) tally_lane<F64, true>(tally, tally64, idx, t);      // kernel.c:276
}

template <int LPT, int GPL, int EXPM, bool F64, bool GEOM, bool HOIST = false>
__global__ void __launch_bounds__(kThreadsPerBlock, kMinBlocksRec)
attenuate_record_tracks(const KernelArgs a)
{
    static_assert(!(GEOM && HOIST), "the fit is only sweep-invariant with the constant geometry");
    static_assert(LPT >= 1 && LPT <= 16 && (LPT & (LPT - 1)) == 0, "LPT must be a power of two");
    static_assert(GPL == 2 || GPL == 4, "two or four groups per lane");
    typedef typename LaneVec<GPL>::type V;
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int kSlotsPerWarp = 32 / LPT;
    constexpr int kWarps = kThreadsPerBlock / 32;

    __shared__ float2 s_pairs[kTableReach];
    // constant geometry: {q0_d, q0_s, q1_d, q1_s} {q2_s, -, -, -} per segment type = pidx >> 30 (interior, first, last,
    // no segment): tracks of different types share a warp, and a 3-way broadcast read of this table replaces seven
    // selects and the moves that feed them.
    // GEOM: per warp and hashing lane, the coefficients of that lane's segment of the batch (its type folded in),
    // {q0_d, q0_s, q1_d, q1_s} {q2_s, ds, weight, -}: derived once by the lane that hashed the segment.
    __shared__ float4 s_fit[GEOM ? kWarps * 64 : 8];
    if constexpr (!GEOM && !HOIST) {
        if (threadIdx.x < 4) {
            const FitCoeffs f = fit_coeffs(threadIdx.x == 1, threadIdx.x == 2);
            s_fit[2 * threadIdx.x] = make_float4(f.q0_d, f.q0_s, f.q1_d, f.q1_s);
            s_fit[2 * threadIdx.x + 1] = make_float4(f.q2_s, 0.0f, 0.0f, 0.0f);
        }
    }
    if constexpr (EXPM == kExpTable) {
        if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
    }
    __syncthreads();
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const int sub = lane & (LPT - 1);
    const int64_t warp_global = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    float4 *const my_fit = s_fit + (GEOM ? ((threadIdx.x >> 5) * 64 + 2 * lane) : 0);            // this lane's slot (GEOM)
    const float4 *const track_fit = s_fit + (GEOM ? ((threadIdx.x >> 5) * 64 + 2 * (lane - sub)) : 0);   // slot of the track's lane 0
    const uint32_t F = (uint32_t)a.fai_count;
    const int p = a.seg_per_track;
    const char *const rec = reinterpret_cast<const char *>(a.records);
    float *const tally = warp_tally(a, warp_global);
    const int64_t n_work = a.track_end - a.track_begin;
    unsigned long long checksum = 0ull;

    for (int64_t wbase = claim_work(a, lane, kSlotsPerWarp); wbase < n_work; wbase = claim_work(a, lane, kSlotsPerWarp)) {
        const int64_t work = wbase + (lane / LPT);
        const bool tvalid = work < n_work;
        const int64_t track = a.track_begin + work;
        const int64_t s0 = track * p;
        int nseg = 0;
        if (tvalid) {
            const int64_t left = a.segments - s0;
            nseg = left < p ? (int)left : p;
        }

        // incoming angular flux of the track (kernel.c:29-30): one Philox block covers 4 groups
        V psi;
        if constexpr (GPL == 4) {
            const u32x4 r = stream_words(a.keys, (uint64_t)track, (uint32_t)sub, kDomainPsi);
            psi = make_float4(u01(r.x), u01(r.y), u01(r.z), u01(r.w));
        } else {
            const u32x4 r = stream_words(a.keys, (uint64_t)track, (uint32_t)(sub >> 1), kDomainPsi);
            psi = (sub & 1) ? make_float2(u01(r.z), u01(r.w)) : make_float2(u01(r.x), u01(r.y));
        }

        const int nseg_warp = __reduce_max_sync(kFull, nseg);
        const bool even = __all_sync(kFull, nseg == nseg_warp);      // every slot of the warp has a full-length track
        uint64_t seg = (uint64_t)s0 + (uint64_t)sub;                 // the segment this lane hashes in the next batch
        for (int b = 0; b < nseg_warp; b += LPT, seg += LPT) {
            // each lane of the track draws the ids of one of the next LPT segments: idx = row * LPT (the lane's
            // element of the record row and of the tally row once `| sub` is added) + the two type flags.
            // Lanes without a segment keep row 0 with both flags: a valid address, nothing is tallied.
            // (HOIST: the fitted records carry the segment type implicitly, no flags)
            uint32_t my_idx = HOIST ? 0u : (kRowFirst | kRowLast);
            if (b + sub < nseg) {
                const u32x4 w = stream_words(a.keys, seg, 0u, kDomainSegment);
                const uint32_t qsr = fastmod(w.x >> 1, a.mod_regions);                   // kernel.c:47
                const uint32_t fai = fastmod(w.y >> 1, a.mod_fai);                       // kernel.c:50
                my_idx = (qsr * F + fai) * (uint32_t)LPT;
                if constexpr (!HOIST) my_idx |= (fai == 0u ? kRowFirst : 0u) | (fai == F - 1u ? kRowLast : 0u);
                checksum += checksum_term(qsr, fai, F, seg);
                if constexpr (GEOM) {
                    const FitCoeffs f = fit_coeffs_geom(segment_geometry(a.geom, w.z, w.w), a.mesh, fai == 0u, fai == F - 1u);
                    my_fit[0] = make_float4(f.q0_d, f.q0_s, f.q1_d, f.q1_s);
                    my_fit[1] = make_float4(f.q2_s, f.ds, f.weight, 0.0f);
                }
            } else if constexpr (GEOM) {
                my_fit[0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                my_fit[1] = make_float4(0.0f, 1.0f, 0.0f, 0.0f);
            }
            if constexpr (GEOM) __syncwarp();
            const int count = (nseg_warp - b) < LPT ? (nseg_warp - b) : LPT;
            if (even) {
#pragma unroll kRecordUnroll
                for (int k = 0; k < count; ++k) {
                    const uint32_t pidx = __shfl_sync(kFull, my_idx, k, LPT);
                    record_segment<LPT, GPL, EXPM, F64, GEOM, HOIST, false>(rec, GEOM ? track_fit + 2 * k : s_fit + 2 * (pidx >> 30), s_pairs,
                                                                     tally, a.tally64, pidx, sub, true, psi);
                }
            } else {
                for (int k = 0; k < count; ++k) {
                    const uint32_t pidx = __shfl_sync(kFull, my_idx, k, LPT);
                    record_segment<LPT, GPL, EXPM, F64, GEOM, HOIST, true>(rec, GEOM ? track_fit + 2 * k : s_fit + 2 * (pidx >> 30), s_pairs,
                                                                    tally, a.tally64, pidx, sub, (b + k) < nseg, psi);
                }
            }
            if constexpr (GEOM) __syncwarp();       // everyone is done with the slots before the next batch overwrites them
        }

        if (a.psi_out != nullptr && tvalid)
            reinterpret_cast<V *>(a.psi_out)[(track - a.track_begin) * LPT + sub] = psi;
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) checksum += __shfl_xor_sync(kFull, checksum, off);
    if (lane == 0 && checksum != 0ull) atomicAdd(a.checksum, checksum);
}

// ------------------------------------------------------------------------------
// attenuate_warp_track_rec<EXPM, GEOM>: attenuate_warp_track<2> (33..64 groups: one track per warp, two groups per lane,
// warp-uniform segment types, f32 tallies) fed from the gather records of build_records<2>: ONE
// 256-bit load per lane and segment instead of a sigT load + 2-3 row loads of 64 bits, ONE shuffled word per segment
// (the type flags ride above the row index), no sigT index and no neighbour-row addresses: 68 -> 61 instructions per
// lane and segment.  Same typed bodies, same arithmetic: psi per track is bit-identical to attenuate_warp_track.
// 64 groups 6.84e11 -> 7.24e11, config 4 6.91e11 -> 7.38e11 (profiles/ab_r02.md r02p).  At 65..128 groups (four
// groups per lane) neither form pays: gather records move 11 % more bytes through a 4 x larger footprint (-10 %), and
// interleaving the sigT row with every source row (same loads off one address, one shuffle) is within the noise
// (r02q): that shape is bound by the FP32 pipe, not by its 25 non-FP instructions per segment.
// ------------------------------------------------------------------------------
template <int EXPM, bool GEOM, bool HOIST = false>
__global__ void __launch_bounds__(kThreadsPerBlock, GEOM ? kMinBlocksGeom : kMinBlocksHalf)
attenuate_warp_track_rec(const KernelArgs a)
{
    static_assert(!(GEOM && HOIST), "the fit is only sweep-invariant with the constant geometry");
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int kWarps = kThreadsPerBlock / 32;
    constexpr uint32_t ROWV = 32;

    __shared__ float2 s_pairs[kTableReach];
    // GEOM: per warp, per segment of the batch: {ds, weight, q0_d, q0_s} {q1_d, q1_s, q2_s, -} (as attenuate_warp_track)
    __shared__ float4 s_coef[GEOM ? kWarps : 1][GEOM ? 32 : 1][2];
    if constexpr (EXPM == kExpTable) {
        if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
        __syncthreads();
    }
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const int warp = threadIdx.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * kWarps + warp;
    const uint32_t F = (uint32_t)a.fai_count;
    const int p = a.seg_per_track;
    const char *const rec = reinterpret_cast<const char *>(a.records);
    float *const tally = warp_tally(a, warp_global);
    const int64_t n_tracks = a.track_end - a.track_begin;
    unsigned long long checksum = 0ull;

    for (int64_t w = claim_work(a, lane, 1); w < n_tracks; w = claim_work(a, lane, 1)) {
        const int64_t track = a.track_begin + w;
        const int64_t s0 = track * p;
        const int64_t left = a.segments - s0;
        const int nseg = left < p ? (int)left : p;

        // incoming angular flux of the track (kernel.c:29-30): one Philox block covers 4 groups = two lanes
        const u32x4 r0 = stream_words(a.keys, (uint64_t)track, (uint32_t)(lane >> 1), kDomainPsi);
        float2 psi = (lane & 1) ? make_float2(u01(r0.z), u01(r0.w)) : make_float2(u01(r0.x), u01(r0.y));

        for (int b = 0; b < nseg; b += 32) {
            uint32_t my_pk = 0u;
            if (b + lane < nseg) {
                const uint64_t seg = (uint64_t)(s0 + b + lane);
                const u32x4 r = stream_words(a.keys, seg, 0u, kDomainSegment);
                const uint32_t qsr = fastmod(r.x >> 1, a.mod_regions);       // kernel.c:47
                const uint32_t fai = fastmod(r.y >> 1, a.mod_fai);           // kernel.c:50
                checksum += checksum_term(qsr, fai, F, seg);
                // row * 32 < 2^30 (smk_create checks): the type flags ride in the two top bits
                my_pk = ((qsr * F + fai) * ROWV) | (fai == 0u ? kSgFirst : 0u) | (fai == F - 1u ? kSgLast : 0u);
                if constexpr (GEOM) {
                    const SegGeometry g = segment_geometry(a.geom, r.z, r.w);
                    const FitCoeffs f = fit_coeffs_geom_typed(g, a.mesh, fai == 0u || fai == F - 1u);
                    s_coef[warp][lane][0] = make_float4(f.ds, f.weight, f.q0_d, f.q0_s);
                    s_coef[warp][lane][1] = make_float4(f.q1_d, f.q1_s, f.q2_s, 0.0f);
                }
            }
            if constexpr (GEOM) __syncwarp();
            const int count = (nseg - b) < 32 ? (nseg - b) : 32;
#pragma unroll kSegmentUnroll
            for (int k = 0; k < count; ++k) {
                const uint32_t pk = __shfl_sync(kFull, my_pk, k);
                const uint32_t idx = (pk & ~(kSgFirst | kSgLast)) | (uint32_t)lane;
                const Rec8 q = ldg256(ptr_add_index<true>(rec, idx * 32u));     // {sigT, y[FAI-1], y[FAI], y[FAI+1]}
                FitCoeffs fc = {};
                if constexpr (GEOM) {
                    const float4 c0 = s_coef[warp][k][0], c1 = s_coef[warp][k][1];
                    fc.ds = c0.x; fc.weight = c0.y; fc.q0_d = c0.z; fc.q0_s = c0.w;
                    fc.q1_d = c1.x; fc.q1_s = c1.y; fc.q2_s = c1.z;
                }
                float2 t;
                if constexpr (HOIST) {
                    // the records hold {q0, mu q1, mu2 q2}: one edge body (no quadratic terms) and the interior body
                    if (pk & (kSgFirst | kSgLast)) attenuate_lane<EXPM, kFitGivenEdge, false>(fc, q.b, q.c, q.d, q.a, s_pairs, psi, t);
                    else attenuate_lane<EXPM, kFitGiven, false>(fc, q.b, q.c, q.d, q.a, s_pairs, psi, t);
                } else if ((int32_t)pk < 0) attenuate_lane<EXPM, kFitFirst, GEOM>(fc, q.b, q.c, q.d, q.a, s_pairs, psi, t);
                else if (pk & kSgLast) attenuate_lane<EXPM, kFitLast, GEOM>(fc, q.b, q.c, q.d, q.a, s_pairs, psi, t);
                else attenuate_lane<EXPM, kFitInterior, GEOM>(fc, q.b, q.c, q.d, q.a, s_pairs, psi, t);
                tally_lane<false, true>(tally, nullptr, idx, t);                            // kernel.c:276
            }
            if constexpr (GEOM) __syncwarp();       // everyone is done with s_coef before the next batch overwrites it
        }

        if (a.psi_out != nullptr)
            reinterpret_cast<float2 *>(a.psi_out)[(track - a.track_begin) * ROWV + lane] = psi;
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) checksum += __shfl_xor_sync(kFull, checksum, off);
    if (lane == 0 && checksum != 0ull) atomicAdd(a.checksum, checksum);
}

// ------------------------------------------------------------------------------
// SMK_FLAG_FIT_PER_SWEEP at 65..128 groups: build_fit_rows evaluates the typed fit once per (region, interval, group)
// of the sweep into fit[row][{q0, mu q1, mu2 q2}][G_pad] (three rows of 128 groups side by side, so the loads of a
// segment hang off one address with immediate offsets), and attenuate_warp_track_fit<EXPM> is attenuate_warp_track<4>
// reading them: 37 / 26 operations per interior / edge intersection instead of 45 / 29, the same 3-4 loads of 128 bits
// per lane and segment (an edge interval has no q2), psi bit-identical.  f32 tallies, constant geometry, 32-bit offsets.
// ------------------------------------------------------------------------------
__global__ void build_fit_rows(const float *__restrict__ source, float *__restrict__ fit, int64_t rows, int fai_count,
                               int groups_pad)
{
    const int64_t n = rows * groups_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / groups_pad;
        const int g = (int)(i - row * groups_pad);
        const int fai = (int)(row % fai_count);
        const float y2 = source[i];
        const float y1 = fai > 0 ? source[i - groups_pad] : 0.0f;
        const float y3 = fai < fai_count - 1 ? source[i + groups_pad] : 0.0f;
        float q0, Q1, Q2;
        fit_row<true>(fai == 0, fai == fai_count - 1, y1, y2, y3, q0, Q1, Q2);
        float *out = fit + row * 3 * groups_pad + g;
        out[0] = q0;
        out[groups_pad] = Q1;
        out[2 * groups_pad] = Q2;
    }
}

template <int EXPM>
__global__ void __launch_bounds__(kThreadsPerBlock, kMinBlocksFast)
attenuate_warp_track_fit(const KernelArgs a)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int kWarps = kThreadsPerBlock / 32;
    constexpr uint32_t ROWV = 32;

    __shared__ float2 s_pairs[kTableReach];
    if constexpr (EXPM == kExpTable) {
        if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
        __syncthreads();
    }
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const int warp = threadIdx.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * kWarps + warp;
    const uint32_t F = (uint32_t)a.fai_count;
    const int p = a.seg_per_track;
    const float4 *const fit = reinterpret_cast<const float4 *>(a.records);
    float *const tally = warp_tally(a, warp_global);
    const int64_t n_tracks = a.track_end - a.track_begin;
    unsigned long long checksum = 0ull;
    const FitCoeffs fc = {};

    for (int64_t w = claim_work(a, lane, 1); w < n_tracks; w = claim_work(a, lane, 1)) {
        const int64_t track = a.track_begin + w;
        const int64_t s0 = track * p;
        const int64_t left = a.segments - s0;
        const int nseg = left < p ? (int)left : p;
        const u32x4 r0 = stream_words(a.keys, (uint64_t)track, (uint32_t)lane, kDomainPsi);
        float4 psi = make_float4(u01(r0.x), u01(r0.y), u01(r0.z), u01(r0.w));

        for (int b = 0; b < nseg; b += 32) {
            uint32_t my_pk = 0u, my_sg = 0u;
            if (b + lane < nseg) {
                const uint64_t seg = (uint64_t)(s0 + b + lane);
                const u32x4 r = stream_words(a.keys, seg, 0u, kDomainSegment);
                const uint32_t qsr = fastmod(r.x >> 1, a.mod_regions);       // kernel.c:47
                const uint32_t fai = fastmod(r.y >> 1, a.mod_fai);           // kernel.c:50
                checksum += checksum_term(qsr, fai, F, seg);
                my_pk = (qsr * F + fai) * ROWV;
                my_sg = (qsr * ROWV) | ((fai == 0u || fai == F - 1u) ? kSgFirst : 0u);      // one flag: edge interval
            }
            const int count = (nseg - b) < 32 ? (nseg - b) : 32;
#pragma unroll kSegmentUnroll
            for (int k = 0; k < count; ++k) {
                const uint32_t pk = __shfl_sync(kFull, my_pk, k);
                const uint32_t idx = pk | (uint32_t)lane;
                const uint32_t sg = __shfl_sync(kFull, my_sg, k);
                // row * 96 lane vectors: the fitted rows of `row`
                const float4 *src = ptr_add_index<true>(fit, (pk + (pk << 1)) | (uint32_t)lane);
                const float4 st = ldg4(ptr_add_index<true>(a.sigT, (sg & ~kSgFirst) | (uint32_t)lane));
                const float4 q0 = ldg4(src), Q1 = ldg4(src + ROWV);
                float4 t;
                if ((int32_t)sg < 0) {
                    attenuate_lane<EXPM, kFitGivenEdge, false>(fc, Q1, q0, make_float4(0.f, 0.f, 0.f, 0.f), st, s_pairs, psi, t);
                } else {
                    const float4 Q2 = ldg4(src + 2 * ROWV);
                    attenuate_lane<EXPM, kFitGiven, false>(fc, Q1, q0, Q2, st, s_pairs, psi, t);
                }
                tally_lane<false, true>(tally, nullptr, idx, t);                            // kernel.c:276
            }
        }

        if (a.psi_out != nullptr)
            reinterpret_cast<float4 *>(a.psi_out)[(track - a.track_begin) * ROWV + lane] = psi;
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) checksum += __shfl_xor_sync(kFull, checksum, off);
    if (lane == 0 && checksum != 0ull) atomicAdd(a.checksum, checksum);
}

// ------------------------------------------------------------------------------
// layout kernels
// ------------------------------------------------------------------------------

// dst[row][G_pad] <- src[row][G]; padding groups get `pad` (0 for sources, 1 for sigT so
// that the padded lanes of the hot kernel stay finite)
__global__ void pad_rows(const float *__restrict__ src, float *__restrict__ dst, int64_t rows,
                         int groups, int groups_pad, float pad)
{
    const int64_t n = rows * groups_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups_pad;
        const int g = (int)(i - r * groups_pad);
        dst[i] = (g < groups) ? src[r * groups + g] : pad;
    }
}

// out[row][G] = flux0[row][G_pad] + scale * tally[row][G_pad]   (kernel.c:276 summed over the sweep), rows
// [row_begin, row_begin + rows) of the arrays; `stride` = floats between tally replicas; scale = the
// segment weight where the kernel left it out (constant geometry, FAST), else 1
__global__ void finalize_flux(const float *__restrict__ flux0, const float *__restrict__ tally,
                              float *__restrict__ out, int64_t rows, int groups, int groups_pad, int replicas,
                              int64_t stride, float scale)
{
    const int64_t n = rows * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups;
        const int g = (int)(i - r * groups);
        float t = tally[r * groups_pad + g];
        for (int k = 1; k < replicas; ++k) t += tally[k * stride + r * groups_pad + g];
        out[i] = __fadd_rn(flux0[r * groups_pad + g], __fmul_rn(scale, t));
    }
}

// max over an array (sigT): decides whether the POLY exponential needs its wide-range form
__global__ void max_rows(const float *__restrict__ src, int64_t n, unsigned int *__restrict__ out_bits)
{
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = src[i];
        m = (v > m || v != v) ? (v != v ? __int_as_float(0x7f800000) : v) : m;   // NaN counts as unbounded
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));       // non-negative floats order as integers
}

// Element e of the UNPADDED array `array_id` is word (e & 3) of Philox counter
// (e >> 2, array_id, 'FILL') -- identical to the host fill.
__global__ void fill_rows(float *__restrict__ dst, int64_t rows, int groups, int groups_pad,
                          uint32_t array_id, uint64_t seed, float floor_, float pad)
{
    const float span = __fsub_rn(1.0f, floor_);
    const int64_t n = rows * groups_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups_pad;
        const int g = (int)(i - r * groups_pad);
        float v = pad;
        if (g < groups) {
            const uint64_t e = (uint64_t)(r * groups + g);
            const u32x4 w = stream_words(seed, e >> 2, array_id, kDomainFill);
            const uint32_t word = (e & 3u) == 0u ? w.x : (e & 3u) == 1u ? w.y : (e & 3u) == 2u ? w.z : w.w;
            const float u = u01(word);
            v = (floor_ > 0.0f) ? __fadd_rn(floor_, __fmul_rn(u, span)) : u;
        }
        dst[i] = v;
    }
}

// ------------------------------------------------------------------------------
// all-reduce of the tally deltas over NVLink peer memory (one launch per device):
// this device sums float4 [begin, end) of every peer's array (P2P loads, fixed order so
// that every device computes bit-identical sums) and writes the sum back to every peer.
// ------------------------------------------------------------------------------
constexpr int kMaxDevices = 8;
struct PeerArrays {
    float4 *p[kMaxDevices];
};

__global__ void allreduce_peer_slices(PeerArrays arrays, int n_dev, int64_t begin, int64_t end)
{
    for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 acc = arrays.p[0][i];
        for (int d = 1; d < n_dev; ++d) {
            const float4 v = arrays.p[d][i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        for (int d = 0; d < n_dev; ++d) arrays.p[d][i] = acc;
    }
}

// out[row][G] = (float)(flux0 + tally64): finalize for the diagnostic f64 tallies
__global__ void finalize_flux64(const float *__restrict__ flux0, const double *__restrict__ tally64,
                                float *__restrict__ out, int64_t rows, int groups, int groups_pad, float scale)
{
    const int64_t n = rows * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups;
        const int g = (int)(i - r * groups);
        out[i] = (float)((double)flux0[r * groups_pad + g] + (double)scale * tally64[r * groups_pad + g]);
    }
}

__global__ void allreduce_peer_slices64(PeerArrays arrays, int n_dev, int64_t begin, int64_t end)
{
    for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end;
         i += (int64_t)gridDim.x * blockDim.x) {
        double acc = reinterpret_cast<double *>(arrays.p[0])[i];
        for (int d = 1; d < n_dev; ++d) acc += reinterpret_cast<double *>(arrays.p[d])[i];
        for (int d = 0; d < n_dev; ++d) reinterpret_cast<double *>(arrays.p[d])[i] = acc;
    }
}

// diagnostics
template <int EXPM>
__global__ void debug_exp_kernel(const float *__restrict__ tau, float *__restrict__ out, int64_t n)
{
    __shared__ float2 s_pairs[kTableReach];
    if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float e;
        (void)exp_val<EXPM>(tau[i], s_pairs, e);
        out[i] = e;
    }
}

// the packed FAST exponential as the hot kernels evaluate it (both halves get the same tau); TRACK = the
// form of the per-segment-geometry kernels
template <int EXPM, bool TRACK>
__global__ void debug_exp2_kernel(const float *__restrict__ tau, float *__restrict__ out, int64_t n)
{
    __shared__ float2 s_pairs[kTableReach];
    if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float2 e, t2;
        (void)exp_val2<EXPM, TRACK>(f2(tau[i]), s_pairs, e, t2);
        out[i] = e.x;
    }
}

__global__ void debug_ids_kernel(uint64_t seed, int64_t seg_begin, int64_t n, FastMod mr, FastMod mf,
                                 GeometryBase gb, int32_t *qsr, int32_t *fai, float *geom6)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const u32x4 w = stream_words(seed, (uint64_t)(seg_begin + i), 0u, kDomainSegment);
        qsr[i] = (int32_t)fastmod(w.x >> 1, mr);
        fai[i] = (int32_t)fastmod(w.y >> 1, mf);
        if (geom6 != nullptr) {
            const SegGeometry g = segment_geometry(gb, w.z, w.w);
            float *o = geom6 + i * 6;
            o[0] = g.dz; o[1] = g.zin; o[2] = g.weight; o[3] = g.mu; o[4] = g.mu2; o[5] = g.ds;
        }
    }
}

}  // namespace smk
