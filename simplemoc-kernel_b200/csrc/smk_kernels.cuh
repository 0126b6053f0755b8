// smk_kernels.cuh -- sm_100a kernels of the segment-attenuation path.
//
//   attenuate_tracks_pf<NCHUNK, EXPM, ...>       the hot kernel for 65..128 energy groups (default "flat"
//       variant; "prefetch", "defer", "l1pf" are measured alternatives): run_kernel's segment loop +
//       attenuate_segment (/root/reference/src/cpu/kernel.c:43-55, 75-333), one warp per track
//   attenuate_tracks<LPT, NCHUNK, MATH, EXPM>    the general kernel: every other group count (sub-warp
//       tracks, several float4 per lane) and the STRICT verification arithmetic
//   attenuate_tracks_staged<NCHUNK, EXPM, STAGES> measured alternative: rows staged through shared memory
//       by TMA bulk copies (cp.async.bulk + mbarrier ring)
//   fill_rows                                    device-side deterministic fill
//       (replaces /root/reference/src/cpu/init.c:64-75 + the H2D of init.cu:105-127)
//   pad_rows / finalize_flux[64]                 host layout <-> padded device layout
//   allreduce_peer_slices[64]                    multi-GPU all-reduce of the tallies over NVLink peer memory
//
// Work decomposition of the hot kernels
//   track  = seg_per_track consecutive segments sharing one carried angular flux psi
//   a track is owned by LPT lanes of one warp (LPT = lanes per track, a power of two);
//   each lane owns NCHUNK float4 = 4*NCHUNK energy groups and keeps their psi in
//   registers for the whole track.  G = 128 -> LPT = 32, NCHUNK = 1 (one warp per
//   track, 128-bit loads, one 16-byte vector RED per lane per segment);
//   G = 64 -> LPT = 16 (2 tracks per warp); G = 7 -> G_pad = 8, LPT = 2 (16 tracks
//   per warp, the 8th group is padding).
//   Warps claim tracks dynamically from a global counter (claim_tracks).
//   Segment ids come from the counter stream: every LPT segments each lane of the
//   track hashes ONE upcoming segment and the ids are handed round with shuffles, so
//   the Philox cost per intersection is 1/(4*NCHUNK*LPT) of a hash.
//   The FAST arithmetic is packed FP32x2 (FFMA2/FMUL2/FADD2), see smk_math.cuh.
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
    unsigned long long *work_counter;    // next unclaimed track (relative to track_begin), zeroed per launch
    double *tally64;                     // diagnostic: f64 tally accumulators [R][F][G_pad] (SMK_FLAG_TALLY_F64)
    int64_t segments;                    // N
    int64_t track_begin, track_end;
    uint64_t seed;
    PhiloxKeys keys;                     // expanded Philox key schedule of `seed`
    FastMod mod_regions, mod_fai;
    int32_t fai_count;                   // F
    int32_t row_f4;                      // G_pad / 4: float4 per row
    int32_t seg_per_track;               // p
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
#ifndef SMK_UNROLL_K
#define SMK_UNROLL_K 1
#endif
constexpr int kUnrollSegments = SMK_UNROLL_K;   // unroll factor of the per-segment loop of the flat kernel
#ifndef SMK_MIN_BLOCKS_PREFETCH
#define SMK_MIN_BLOCKS_PREFETCH 3
#endif
constexpr int kMinBlocksPrefetch = SMK_MIN_BLOCKS_PREFETCH;

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

// Tally-contention relief (BASELINE config 4: few source regions): when the tally array has few
// rows, several warps hit the same L2 atomic address at once and serialise there.  The library then
// keeps `replicas` copies of the (small) array; every warp adds into copy (warp % replicas) and
// finalize_flux sums the copies.  replicas == 1 for the normal problem sizes.
__device__ __forceinline__ float *warp_tally(const KernelArgs &a, int64_t warp_global)
{
    return a.tally + (a.replicas > 1 ? (warp_global % a.replicas) * a.replica_stride : 0);
}

// Dynamic track scheduling: warps claim tracks from a global counter instead of striding statically.
// All CTAs of the persistent grid are resident from the start, so with static striding an SM that runs
// slower than the others (far L2 partition, fewer co-resident CTAs) sets the kernel time while the fast
// ones idle: ncu showed 22.5 of 32 warps active on average.  One 64-bit atomic per track (100 segments).
__device__ __forceinline__ int64_t claim_tracks(const KernelArgs &a, int lane, int n)
{
    unsigned long long first = 0ull;
    if (lane == 0) first = atomicAdd(a.work_counter, (unsigned long long)n);
    return a.track_begin + (int64_t)__shfl_sync(0xFFFFFFFFu, first, 0);
}

// diagnostic f64 tallies: order-independent to ~1e-16, the yardstick for fp32 accumulation noise
__device__ __forceinline__ void red_add_f64x4(double *addr, float a, float b, float c, float d)
{
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr), "d"((double)a) : "memory");
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr + 1), "d"((double)b) : "memory");
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr + 2), "d"((double)c) : "memory");
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr + 3), "d"((double)d) : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
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

// One segment of one track, FAST arithmetic, for the NCHUNK float4 this lane owns: loads,
// two packed (FP32x2) attenuations per float4, psi carry and the vector RED.
template <int LPT, int NCHUNK, int EXPM, int FIT>
__device__ __forceinline__ void segment_fast(const float4 *__restrict__ src, const float4 *__restrict__ sig,
                                             float4 *tal, int row_f4, const FitCoeffs fc,
                                             const float2 *s_pairs, float4 (&psi)[NCHUNK], bool active,
                                             bool first = false, bool last = false)
{
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 y2 = ldg4(src + c * LPT);
        const float4 st = ldg4(sig + c * LPT);
        float4 y1 = zero, y3 = zero;
        if constexpr (FIT == kFitDynamic) {
            if (!first) y1 = ldg4(src + c * LPT - row_f4);
            if (!last) y3 = ldg4(src + c * LPT + row_f4);
        } else {
            if constexpr (FIT != kFitFirst) y1 = ldg4(src + c * LPT - row_f4);
            if constexpr (FIT != kFitLast) y3 = ldg4(src + c * LPT + row_f4);
        }
        float2 p_lo = make_float2(psi[c].x, psi[c].y), p_hi = make_float2(psi[c].z, psi[c].w);
        float2 t_lo, t_hi;
        attenuate_fast2<EXPM, FIT>(fc, make_float2(y1.x, y1.y), make_float2(y2.x, y2.y), make_float2(y3.x, y3.y),
                                   make_float2(st.x, st.y), s_pairs, p_lo, t_lo);
        attenuate_fast2<EXPM, FIT>(fc, make_float2(y1.z, y1.w), make_float2(y2.z, y2.w), make_float2(y3.z, y3.w),
                                   make_float2(st.z, st.w), s_pairs, p_hi, t_hi);
        if (active) {
            psi[c] = make_float4(p_lo.x, p_lo.y, p_hi.x, p_hi.y);                 // kernel.c:331
            red_add_v4(tal + c * LPT, t_lo.x, t_lo.y, t_hi.x, t_hi.y);            // kernel.c:276
        }
    }
}

template <int LPT, int NCHUNK, int MATH, int EXPM>
__global__ void __launch_bounds__(kThreadsPerBlock, (NCHUNK == 1 && MATH == kMathFast) ? kMinBlocksFast : 1)
attenuate_tracks(const KernelArgs a)
{
    static_assert(LPT >= 1 && LPT <= 32 && (LPT & (LPT - 1)) == 0, "LPT must be a power of two");
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int kSlotsPerWarp = 32 / LPT;

    __shared__ float2 s_pairs[kTableReach];
    if constexpr (EXPM == kExpTable) {
        if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
        __syncthreads();
    }

    const int lane = threadIdx.x & 31;
    const int sub = lane & (LPT - 1);               // lane within its track
    const int64_t warp_global = (int64_t)blockIdx.x * (kThreadsPerBlock / 32) + (threadIdx.x >> 5);

    const int F = a.fai_count;
    const int row_f4 = a.row_f4;
    const int p = a.seg_per_track;
    float *const tally = warp_tally(a, warp_global);
    unsigned long long checksum = 0ull;

    for (int64_t tbase = claim_tracks(a, lane, kSlotsPerWarp); tbase < a.track_end;
         tbase = claim_tracks(a, lane, kSlotsPerWarp)) {
        const int64_t track = tbase + (lane / LPT);
        const bool tvalid = track < a.track_end;
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
            const u32x4 w = stream_words(a.keys, (uint64_t)track, (uint32_t)(c * LPT + sub), kDomainPsi);
            psi[c] = make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w));
        }

        const int nseg_warp = (LPT == 32) ? nseg : __reduce_max_sync(kFull, nseg);

        for (int b = 0; b < nseg_warp; b += LPT) {
            // each lane of the track draws the ids of one of the next LPT segments
            uint32_t my_qsr = 0u, my_fai = 0u;
            if (b + sub < nseg) {
                const uint64_t seg = (uint64_t)(s0 + b + sub);
                const SegmentIds id = segment_ids(a.keys, seg, a.mod_regions, a.mod_fai);
                my_qsr = id.qsr;
                my_fai = id.fai;
                checksum += checksum_term(id.qsr, id.fai, (uint32_t)F, seg);
            }
            const int count = (nseg_warp - b) < LPT ? (nseg_warp - b) : LPT;
            for (int k = 0; k < count; ++k) {
                const uint32_t qsr = __shfl_sync(kFull, my_qsr, k, LPT);
                const uint32_t fai = __shfl_sync(kFull, my_fai, k, LPT);
                // only the stream's ragged last track can be shorter than its warp-mates
                const bool active = (LPT == 32) ? true : (b + k) < nseg;
                const bool first = (fai == 0u);
                const bool last = (fai == (uint32_t)(F - 1));
                // 32-bit row offsets (smk_create checks R * F * G_pad / 4 < 2^31)
                const uint32_t row = qsr * (uint32_t)F + fai;
                const uint32_t off = row * (uint32_t)row_f4 + (uint32_t)sub;
                const float4 *src = a.source + off;
                const float4 *sig = a.sigT + (qsr * (uint32_t)row_f4 + (uint32_t)sub);
                float4 *tal = reinterpret_cast<float4 *>(tally) + off;

                if constexpr (MATH == kMathFast && LPT == 32) {
                    // one track per warp: the segment type is warp-uniform, so branch on it and
                    // run code specialised for the type (literal coefficients; the edge types
                    // load 2 rows and skip the quadratic terms)
                    if (first)
                        segment_fast<LPT, NCHUNK, EXPM, kFitFirst>(src, sig, tal, row_f4, FitCoeffs{}, s_pairs, psi, true);
                    else if (last)
                        segment_fast<LPT, NCHUNK, EXPM, kFitLast>(src, sig, tal, row_f4, FitCoeffs{}, s_pairs, psi, true);
                    else
                        segment_fast<LPT, NCHUNK, EXPM, kFitInterior>(src, sig, tal, row_f4, FitCoeffs{}, s_pairs, psi, true);
                } else if constexpr (MATH == kMathFast) {
                    // several tracks per warp: types differ between lanes -> per-lane coefficients
                    segment_fast<LPT, NCHUNK, EXPM, kFitDynamic>(src, sig, tal, row_f4, fit_coeffs(first, last),
                                                                 s_pairs, psi, active, first, last);
                } else {
#pragma unroll
                    for (int c = 0; c < NCHUNK; ++c) {
                        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 y2 = ldg4(src + c * LPT);
                        const float4 st = ldg4(sig + c * LPT);
                        const float4 y1 = first ? zero : ldg4(src + c * LPT - row_f4);
                        const float4 y3 = last ? zero : ldg4(src + c * LPT + row_f4);
                        float4 t, ps = psi[c];
                        attenuate_strict<EXPM>(first, last, y1.x, y2.x, y3.x, st.x, s_pairs, ps.x, t.x);
                        attenuate_strict<EXPM>(first, last, y1.y, y2.y, y3.y, st.y, s_pairs, ps.y, t.y);
                        attenuate_strict<EXPM>(first, last, y1.z, y2.z, y3.z, st.z, s_pairs, ps.z, t.z);
                        attenuate_strict<EXPM>(first, last, y1.w, y2.w, y3.w, st.w, s_pairs, ps.w, t.w);
                        if (active) {
                            psi[c] = ps;                                          // kernel.c:331
                            red_add_v4(tal + c * LPT, t.x, t.y, t.z, t.w);        // kernel.c:276
                        }
                    }
                }
            }
        }

        if (a.psi_out != nullptr && tvalid) {
            float4 *out = reinterpret_cast<float4 *>(a.psi_out) + (track - a.track_begin) * row_f4 + sub;
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
// attenuate_tracks_staged<NCHUNK, EXPM, STAGES>: the one-track-per-warp FAST kernel with the
// source-region rows staged through shared memory by the TMA engine (north star item 3).
//
// Every warp owns a ring of STAGES buffers of 4 rows (y1, y2, y3 = fine_source[QSR][FAI-1..FAI+1]
// and sigT[QSR]) and one mbarrier per buffer.  The rows of a segment are contiguous in HBM
// (init.c:39-40), so a segment is TWO 1-D bulk copies (cp.async.bulk, SASS UBLKCP): 2 or 3
// source rows, and the sigT row.  The lane that hashed segment s+STAGES-1 issues its copies
// while the warp computes segment s, so the L2 latency that showed up as 24 % long-scoreboard
// stall samples at the first use of the loaded rows (profiles/ncu_r01d_summary.md) is taken off
// the critical path without holding the rows in registers.
// ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SMK_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SMK_DONE_%=;\n\t"
        "bra SMK_WAIT_%=;\n\t"
        "SMK_DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr uint32_t kFlagFirst = 0x80000000u, kFlagLast = 0x40000000u, kRowMask = 0x3FFFFFFFu;

template <int NCHUNK, int EXPM, int FIT>
__device__ __forceinline__ void segment_staged(const float4 *stage, float4 *tal, const float2 *s_pairs,
                                               float4 (&psi)[NCHUNK])
{
    constexpr int ROWF4 = 32 * NCHUNK;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 y2 = stage[ROWF4 + c * 32 + lane];
        const float4 st = stage[3 * ROWF4 + c * 32 + lane];
        float4 y1 = zero, y3 = zero;
        if constexpr (FIT != kFitFirst) y1 = stage[c * 32 + lane];
        if constexpr (FIT != kFitLast) y3 = stage[2 * ROWF4 + c * 32 + lane];
        float2 p_lo = make_float2(psi[c].x, psi[c].y), p_hi = make_float2(psi[c].z, psi[c].w);
        float2 t_lo, t_hi;
        attenuate_fast2<EXPM, FIT>(FitCoeffs{}, make_float2(y1.x, y1.y), make_float2(y2.x, y2.y),
                                   make_float2(y3.x, y3.y), make_float2(st.x, st.y), s_pairs, p_lo, t_lo);
        attenuate_fast2<EXPM, FIT>(FitCoeffs{}, make_float2(y1.z, y1.w), make_float2(y2.z, y2.w),
                                   make_float2(y3.z, y3.w), make_float2(st.z, st.w), s_pairs, p_hi, t_hi);
        psi[c] = make_float4(p_lo.x, p_lo.y, p_hi.x, p_hi.y);                     // kernel.c:331
        red_add_v4(tal + c * 32, t_lo.x, t_lo.y, t_hi.x, t_hi.y);                 // kernel.c:276
    }
}

template <int NCHUNK, int EXPM, int STAGES>
__global__ void __launch_bounds__(kThreadsPerBlock, (NCHUNK == 1) ? kMinBlocksFast : 1)
attenuate_tracks_staged(const KernelArgs a)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int kWarps = kThreadsPerBlock / 32;
    constexpr int ROWF4 = 32 * NCHUNK;                 // float4 per padded row
    constexpr uint32_t ROWB = ROWF4 * 16;              // bytes per padded row
    constexpr uint32_t STAGEB = 4 * ROWB;              // y1, y2, y3, sigT
    constexpr int AHEAD = STAGES - 1;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ float2 s_pairs[kTableReach];
    __shared__ __align__(8) unsigned long long s_bars[kWarps][STAGES];
    if constexpr (EXPM == kExpTable) {
        if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
    }

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *ring = smem_raw + (size_t)warp * STAGES * STAGEB;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t bars_u32 = smem_u32(&s_bars[warp][0]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) mbar_init(bars_u32 + 8u * i, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t warp_global = (int64_t)blockIdx.x * kWarps + warp;
    const uint32_t F = (uint32_t)a.fai_count;
    const int p = a.seg_per_track;
    const char *src_bytes = reinterpret_cast<const char *>(a.source);
    const char *sig_bytes = reinterpret_cast<const char *>(a.sigT);
    float *const tally = warp_tally(a, warp_global);
    unsigned long long checksum = 0ull;
    uint32_t prod_stage = 0, cons_stage = 0, cons_parity = 0;   // ring positions persist across tracks

    // ids of one hashed segment, packed: row = QSR*F + FAI | type flags; qsr kept for the sigT row
    auto draw = [&](int64_t s0, int idx, int nseg, uint32_t &packed, uint32_t &qsr) {
        packed = 0u;
        qsr = 0u;
        if (idx < nseg) {
            const uint64_t seg = (uint64_t)(s0 + idx);
            const SegmentIds id = segment_ids(a.keys, seg, a.mod_regions, a.mod_fai);
            checksum += checksum_term(id.qsr, id.fai, F, seg);
            qsr = id.qsr;
            packed = (id.qsr * F + id.fai) | (id.fai == 0u ? kFlagFirst : 0u) | (id.fai == F - 1u ? kFlagLast : 0u);
        }
    };
    // executed by the ONE lane that drew the segment: two bulk copies into ring slot `stage`
    auto issue = [&](uint32_t packed, uint32_t qsr, uint32_t stage) {
        const uint32_t row = packed & kRowMask;
        const bool first = (packed & kFlagFirst) != 0u, last = (packed & kFlagLast) != 0u;
        const uint32_t nrows = (first || last) ? 2u : 3u;
        const uint32_t bar = bars_u32 + 8u * stage;
        const uint32_t dst = ring_u32 + stage * STAGEB;
        mbar_expect_tx(bar, (nrows + 1u) * ROWB);
        bulk_g2s(dst + (first ? ROWB : 0u), src_bytes + (size_t)(row - (first ? 0u : 1u)) * ROWB, nrows * ROWB, bar);
        bulk_g2s(dst + 3u * ROWB, sig_bytes + (size_t)qsr * ROWB, ROWB, bar);
    };

    for (int64_t track = claim_tracks(a, lane, 1); track < a.track_end; track = claim_tracks(a, lane, 1)) {
        const int64_t s0 = track * p;
        const int64_t left = a.segments - s0;
        const int nseg = left < p ? (int)left : p;

        float4 psi[NCHUNK];
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
            const u32x4 w = stream_words(a.keys, (uint64_t)track, (uint32_t)(c * 32 + lane), kDomainPsi);
            psi[c] = make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w));
        }

        uint32_t cur_packed, cur_qsr, nxt_packed, nxt_qsr;
        draw(s0, lane, nseg, cur_packed, cur_qsr);
        draw(s0, 32 + lane, nseg, nxt_packed, nxt_qsr);

        // prologue: segments 0 .. AHEAD-1 (AHEAD < 32, so they are all in the current batch)
#pragma unroll
        for (int j = 0; j < AHEAD; ++j) {
            if (j < nseg) {
                if (lane == j) issue(cur_packed, cur_qsr, prod_stage);
                prod_stage = (prod_stage + 1 == STAGES) ? 0u : prod_stage + 1;
            }
        }

        for (int s = 0; s < nseg; ++s) {
            const int sp = s + AHEAD;                      // segment to prefetch
            if (sp < nseg) {
                if (lane == (sp & 31)) {
                    const bool same_batch = (sp >> 5) == (s >> 5);
                    issue(same_batch ? cur_packed : nxt_packed, same_batch ? cur_qsr : nxt_qsr, prod_stage);
                }
                prod_stage = (prod_stage + 1 == STAGES) ? 0u : prod_stage + 1;
            }

            const uint32_t packed = __shfl_sync(kFull, cur_packed, s & 31);
            const float4 *stage = reinterpret_cast<const float4 *>(ring + cons_stage * STAGEB);
            float4 *tal = reinterpret_cast<float4 *>(tally) + ((packed & kRowMask) * (uint32_t)ROWF4 + (uint32_t)lane);
            mbar_wait(bars_u32 + 8u * cons_stage, cons_parity);
            if (packed & kFlagFirst)
                segment_staged<NCHUNK, EXPM, kFitFirst>(stage, tal, s_pairs, psi);
            else if (packed & kFlagLast)
                segment_staged<NCHUNK, EXPM, kFitLast>(stage, tal, s_pairs, psi);
            else
                segment_staged<NCHUNK, EXPM, kFitInterior>(stage, tal, s_pairs, psi);
            __syncwarp();                                  // all lanes done reading before the slot is refilled
            cons_stage = (cons_stage + 1 == STAGES) ? 0u : cons_stage + 1;
            cons_parity ^= (cons_stage == 0u) ? 1u : 0u;

            if ((s & 31) == 31) {                          // next batch of ids
                cur_packed = nxt_packed;
                cur_qsr = nxt_qsr;
                draw(s0, s + 33 + lane, nseg, nxt_packed, nxt_qsr);
            }
        }

        if (a.psi_out != nullptr) {
            float4 *out = reinterpret_cast<float4 *>(a.psi_out) + (track - a.track_begin) * ROWF4 + lane;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) out[c * 32] = psi[c];
        }
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) checksum += __shfl_xor_sync(kFull, checksum, off);
    if (lane == 0 && checksum != 0ull) atomicAdd(a.checksum, checksum);
}

// ------------------------------------------------------------------------------
// attenuate_tracks_pf<NCHUNK, EXPM>: one track per warp, FAST math, with the rows of segment
// s+1 requested (128-bit read-only loads into a second register set) before segment s is
// computed.  Same loads and arithmetic as attenuate_tracks; the software pipeline removes the
// long-scoreboard stall at the first use of the loaded rows (24 % of stall samples in
// profiles/ncu_r01d_summary.md) at the price of 16 registers.
// ------------------------------------------------------------------------------
template <int NCHUNK>
struct SegRows {
    float4 y1[NCHUNK], y2[NCHUNK], y3[NCHUNK], st[NCHUNK];
};

template <int NCHUNK>
__device__ __forceinline__ void load_rows(SegRows<NCHUNK> &r, const float4 *__restrict__ source,
                                          const float4 *__restrict__ sigT, uint32_t packed, uint32_t qsr,
                                          int lane)
{
    constexpr uint32_t ROWF4 = 32 * NCHUNK;
    const float4 *src = source + ((packed & kRowMask) * ROWF4 + (uint32_t)lane);
    const float4 *sig = sigT + (qsr * ROWF4 + (uint32_t)lane);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        r.y2[c] = ldg4(src + c * 32);
        r.st[c] = ldg4(sig + c * 32);
        r.y1[c] = (packed & kFlagFirst) ? zero : ldg4(src + c * 32 - ROWF4);
        r.y3[c] = (packed & kFlagLast) ? zero : ldg4(src + c * 32 + ROWF4);
    }
}

template <int NCHUNK, int EXPM, int FIT>
__device__ __forceinline__ void compute_rows(const SegRows<NCHUNK> &r, const float2 *s_pairs, float4 (&psi)[NCHUNK],
                                             float4 (&tally)[NCHUNK])
{
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        float2 p_lo = make_float2(psi[c].x, psi[c].y), p_hi = make_float2(psi[c].z, psi[c].w);
        float2 t_lo, t_hi;
        attenuate_fast2<EXPM, FIT>(FitCoeffs{}, make_float2(r.y1[c].x, r.y1[c].y), make_float2(r.y2[c].x, r.y2[c].y),
                                   make_float2(r.y3[c].x, r.y3[c].y), make_float2(r.st[c].x, r.st[c].y), s_pairs,
                                   p_lo, t_lo);
        attenuate_fast2<EXPM, FIT>(FitCoeffs{}, make_float2(r.y1[c].z, r.y1[c].w), make_float2(r.y2[c].z, r.y2[c].w),
                                   make_float2(r.y3[c].z, r.y3[c].w), make_float2(r.st[c].z, r.st[c].w), s_pairs,
                                   p_hi, t_hi);
        psi[c] = make_float4(p_lo.x, p_lo.y, p_hi.x, p_hi.y);                     // kernel.c:331
        tally[c] = make_float4(t_lo.x, t_lo.y, t_hi.x, t_hi.y);
    }
}

// attenuation of one segment by its (warp-uniform) type; the tally comes back in registers
template <int NCHUNK, int EXPM>
__device__ __forceinline__ void compute_by_type(const SegRows<NCHUNK> &r, uint32_t packed, const float2 *s_pairs,
                                                float4 (&psi)[NCHUNK], float4 (&tally)[NCHUNK])
{
#ifdef SMK_EXPERIMENT_ONE_TYPE   // timing experiment only: every segment runs the interior body
    compute_rows<NCHUNK, EXPM, kFitInterior>(r, s_pairs, psi, tally);
    return;
#endif
    if (packed & kFlagFirst)
        compute_rows<NCHUNK, EXPM, kFitFirst>(r, s_pairs, psi, tally);
    else if (packed & kFlagLast)
        compute_rows<NCHUNK, EXPM, kFitLast>(r, s_pairs, psi, tally);
    else
        compute_rows<NCHUNK, EXPM, kFitInterior>(r, s_pairs, psi, tally);
}

// FSR_flux[g] += tally[g] (kernel.c:276) for the row of `packed`: one vector RED per lane
template <int NCHUNK>
__device__ __forceinline__ void red_row(float *tally_base, uint32_t packed, int lane, const float4 (&t)[NCHUNK])
{
    float4 *tal = reinterpret_cast<float4 *>(tally_base) + ((packed & kRowMask) * (uint32_t)(32 * NCHUNK) + (uint32_t)lane);
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) red_add_v4(tal + c * 32, t[c].x, t[c].y, t[c].z, t[c].w);
}

template <int NCHUNK, int EXPM, bool PREFETCH, bool DEFER, bool L1PF = false>
__global__ void __launch_bounds__(kThreadsPerBlock, (NCHUNK == 1) ? (PREFETCH ? kMinBlocksPrefetch : kMinBlocksFast) : 1)
attenuate_tracks_pf(const KernelArgs a)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int kWarps = kThreadsPerBlock / 32;
    constexpr int ROWF4 = 32 * NCHUNK;

    __shared__ float2 s_pairs[kTableReach];
    if constexpr (EXPM == kExpTable) {
        if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const uint32_t F = (uint32_t)a.fai_count;
    const int p = a.seg_per_track;
    float *const tally = warp_tally(a, warp_global);
    unsigned long long checksum = 0ull;

    auto draw = [&](int64_t s0, int idx, int nseg, uint32_t &packed, uint32_t &qsr) {
        packed = 0u;
        qsr = 0u;
        if (idx < nseg) {
            const uint64_t seg = (uint64_t)(s0 + idx);
            const SegmentIds id = segment_ids(a.keys, seg, a.mod_regions, a.mod_fai);
            checksum += checksum_term(id.qsr, id.fai, F, seg);
            qsr = id.qsr;
            packed = (id.qsr * F + id.fai) | (id.fai == 0u ? kFlagFirst : 0u) | (id.fai == F - 1u ? kFlagLast : 0u);
        }
    };

    for (int64_t track = claim_tracks(a, lane, 1); track < a.track_end; track = claim_tracks(a, lane, 1)) {
        const int64_t s0 = track * p;
        const int64_t left = a.segments - s0;
        const int nseg = left < p ? (int)left : p;

        float4 psi[NCHUNK];
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
            const u32x4 w = stream_words(a.keys, (uint64_t)track, (uint32_t)(c * 32 + lane), kDomainPsi);
            psi[c] = make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w));
        }

        // ids of segments [32b, 32b+32) live in cur_*, of the following 32 in nxt_*
        uint32_t cur_packed, cur_qsr, nxt_packed, nxt_qsr;
        draw(s0, lane, nseg, cur_packed, cur_qsr);
        draw(s0, 32 + lane, nseg, nxt_packed, nxt_qsr);

        // ids of segment i, valid while i is in the current or the next batch of segment `at`
        auto ids_of = [&](int i, int at, uint32_t &packed, uint32_t &qsr) {
            const bool same = (i >> 5) == (at >> 5);
            packed = __shfl_sync(kFull, same ? cur_packed : nxt_packed, i & 31);
            qsr = __shfl_sync(kFull, same ? cur_qsr : nxt_qsr, i & 31);
        };
        auto rotate = [&](int s) {
            if ((s & 31) == 31) {
                cur_packed = nxt_packed;
                cur_qsr = nxt_qsr;
                draw(s0, s + 33 + lane, nseg, nxt_packed, nxt_qsr);
            }
        };

        float4 t[NCHUNK];
        if constexpr (PREFETCH) {
            SegRows<NCHUNK> ra, rb;
            uint32_t pa, qa, pb = 0u, qb = 0u;
            ids_of(0, 0, pa, qa);
            load_rows<NCHUNK>(ra, a.source, a.sigT, pa, qa, lane);
            for (int s = 0; s < nseg; s += 2) {
                if (s + 1 < nseg) {                                   // request s+1, compute s
                    ids_of(s + 1, s, pb, qb);
                    load_rows<NCHUNK>(rb, a.source, a.sigT, pb, qb, lane);
                }
                compute_by_type<NCHUNK, EXPM>(ra, pa, s_pairs, psi, t);
                red_row<NCHUNK>(tally, pa, lane, t);
                rotate(s);
                if (s + 1 >= nseg) break;
                if (s + 2 < nseg) {                                   // request s+2, compute s+1
                    ids_of(s + 2, s + 1, pa, qa);
                    load_rows<NCHUNK>(ra, a.source, a.sigT, pa, qa, lane);
                }
                compute_by_type<NCHUNK, EXPM>(rb, pb, s_pairs, psi, t);
                red_row<NCHUNK>(tally, pb, lane, t);
                rotate(s + 1);
            }
        } else if constexpr (DEFER) {
            // the RED of segment s-1 is issued right after the loads of segment s, i.e. while the
            // warp would be waiting for those loads anyway
            uint32_t pend = 0u;
            for (int s = 0; s < nseg; ++s) {
                SegRows<NCHUNK> r;
                const uint32_t pk = __shfl_sync(kFull, cur_packed, s & 31);
                const uint32_t qs = __shfl_sync(kFull, cur_qsr, s & 31);
                load_rows<NCHUNK>(r, a.source, a.sigT, pk, qs, lane);
                if (s > 0) red_row<NCHUNK>(tally, pend, lane, t);
                compute_by_type<NCHUNK, EXPM>(r, pk, s_pairs, psi, t);
                pend = pk;
                rotate(s);
            }
            if (nseg > 0) red_row<NCHUNK>(tally, pend, lane, t);
        } else {
            // batches of 32 segments (one id per lane); inside a batch: branch on the warp-uniform
            // segment type first, then load only the rows that type reads, compute, RED
            for (int b = 0; b < nseg; b += 32) {
                const int count = (nseg - b) < 32 ? (nseg - b) : 32;
#pragma unroll(kUnrollSegments)
                for (int k = 0; k < count; ++k) {
                    const uint32_t pk = __shfl_sync(kFull, cur_packed, k);
                    const uint32_t qs = __shfl_sync(kFull, cur_qsr, k);
                    const uint32_t off = (pk & kRowMask) * (uint32_t)ROWF4 + (uint32_t)lane;
                    const float4 *src = a.source + off;
                    const float4 *sig = a.sigT + (qs * (uint32_t)ROWF4 + (uint32_t)lane);
                    if constexpr (L1PF) {
                        // pull the rows of the NEXT segment of this batch into L1 while this one is computed
                        // (no registers held; the loads of the next iteration then hit L1)
                        const int kn = (k + 1 < count) ? k + 1 : k;
                        const uint32_t pkn = __shfl_sync(kFull, cur_packed, kn);
                        const uint32_t qsn = __shfl_sync(kFull, cur_qsr, kn);
                        const float4 *srcn = a.source + ((pkn & kRowMask) * (uint32_t)ROWF4 + (uint32_t)lane);
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c) {
                            prefetch_l1(srcn + c * 32);
                            if (!(pkn & kFlagFirst)) prefetch_l1(srcn + c * 32 - ROWF4);
                            if (!(pkn & kFlagLast)) prefetch_l1(srcn + c * 32 + ROWF4);
                            prefetch_l1(a.sigT + (qsn * (uint32_t)ROWF4 + (uint32_t)lane) + c * 32);
                        }
                    }
                    SegRows<NCHUNK> r;
                    if (pk & kFlagFirst) {
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c) {
                            r.y2[c] = ldg4(src + c * 32);
                            r.y3[c] = ldg4(src + c * 32 + ROWF4);
                            r.st[c] = ldg4(sig + c * 32);
                        }
                        compute_rows<NCHUNK, EXPM, kFitFirst>(r, s_pairs, psi, t);
                    } else if (pk & kFlagLast) {
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c) {
                            r.y1[c] = ldg4(src + c * 32 - ROWF4);
                            r.y2[c] = ldg4(src + c * 32);
                            r.st[c] = ldg4(sig + c * 32);
                        }
                        compute_rows<NCHUNK, EXPM, kFitLast>(r, s_pairs, psi, t);
                    } else {
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c) {
                            r.y1[c] = ldg4(src + c * 32 - ROWF4);
                            r.y2[c] = ldg4(src + c * 32);
                            r.y3[c] = ldg4(src + c * 32 + ROWF4);
                            r.st[c] = ldg4(sig + c * 32);
                        }
                        compute_rows<NCHUNK, EXPM, kFitInterior>(r, s_pairs, psi, t);
                    }
                    if (a.tally64 == nullptr) {
                        float4 *tal = reinterpret_cast<float4 *>(tally) + off;
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c) red_add_v4(tal + c * 32, t[c].x, t[c].y, t[c].z, t[c].w);
                    } else {
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c)
                            red_add_f64x4(a.tally64 + ((size_t)off + c * 32) * 4, t[c].x, t[c].y, t[c].z, t[c].w);
                    }
                }
                cur_packed = nxt_packed;
                cur_qsr = nxt_qsr;
                draw(s0, b + 64 + lane, nseg, nxt_packed, nxt_qsr);
            }
        }

        if (a.psi_out != nullptr) {
            float4 *out = reinterpret_cast<float4 *>(a.psi_out) + (track - a.track_begin) * ROWF4 + lane;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) out[c * 32] = psi[c];
        }
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) checksum += __shfl_xor_sync(kFull, checksum, off);
    if (lane == 0 && checksum != 0ull) atomicAdd(a.checksum, checksum);
}

// ------------------------------------------------------------------------------
// attenuate_tracks_half<EXPM>: the flat one-track-per-warp kernel for 33..64 energy groups
// (G_pad = 64): each lane owns TWO groups (one packed FP32x2 pair), 64-bit loads, one 8-byte vector
// RED per lane per segment.  Compared with the general kernel (two tracks per warp, per-lane fit
// coefficients) the segment type is warp-uniform again, so the edge bodies skip the quadratic terms.
// ------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v2(float2 *addr, float a, float b)
{
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

template <int EXPM>
__global__ void __launch_bounds__(kThreadsPerBlock, kMinBlocksFast)
attenuate_tracks_half(const KernelArgs a)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int kWarps = kThreadsPerBlock / 32;
    constexpr uint32_t ROWF2 = 32;                         // float2 per padded row (G_pad = 64)

    __shared__ float2 s_pairs[kTableReach];
    if constexpr (EXPM == kExpTable) {
        if (threadIdx.x < kTableReach) s_pairs[threadIdx.x] = c_exp_table.pairs[threadIdx.x];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const uint32_t F = (uint32_t)a.fai_count;
    const int p = a.seg_per_track;
    const float2 *const source = reinterpret_cast<const float2 *>(a.source);
    const float2 *const sigT = reinterpret_cast<const float2 *>(a.sigT);
    float2 *const tally = reinterpret_cast<float2 *>(warp_tally(a, warp_global));
    unsigned long long checksum = 0ull;

    auto draw = [&](int64_t s0, int idx, int nseg, uint32_t &packed, uint32_t &qsr) {
        packed = 0u;
        qsr = 0u;
        if (idx < nseg) {
            const uint64_t seg = (uint64_t)(s0 + idx);
            const SegmentIds id = segment_ids(a.keys, seg, a.mod_regions, a.mod_fai);
            checksum += checksum_term(id.qsr, id.fai, F, seg);
            qsr = id.qsr;
            packed = (id.qsr * F + id.fai) | (id.fai == 0u ? kFlagFirst : 0u) | (id.fai == F - 1u ? kFlagLast : 0u);
        }
    };

    for (int64_t track = claim_tracks(a, lane, 1); track < a.track_end; track = claim_tracks(a, lane, 1)) {
        const int64_t s0 = track * p;
        const int64_t left = a.segments - s0;
        const int nseg = left < p ? (int)left : p;

        // psi0: one Philox block covers 4 groups = the two groups of lanes 2j and 2j+1
        const u32x4 w = stream_words(a.keys, (uint64_t)track, (uint32_t)(lane >> 1), kDomainPsi);
        float2 psi = (lane & 1) ? make_float2(u01(w.z), u01(w.w)) : make_float2(u01(w.x), u01(w.y));

        uint32_t cur_packed, cur_qsr, nxt_packed, nxt_qsr;
        draw(s0, lane, nseg, cur_packed, cur_qsr);
        draw(s0, 32 + lane, nseg, nxt_packed, nxt_qsr);

        for (int b = 0; b < nseg; b += 32) {
            const int count = (nseg - b) < 32 ? (nseg - b) : 32;
            for (int k = 0; k < count; ++k) {
                const uint32_t pk = __shfl_sync(kFull, cur_packed, k);
                const uint32_t qs = __shfl_sync(kFull, cur_qsr, k);
                const uint32_t off = (pk & kRowMask) * ROWF2 + (uint32_t)lane;
                const float2 *src = source + off;
                const float2 st = __ldg(sigT + (qs * ROWF2 + (uint32_t)lane));
                const float2 y2 = __ldg(src);
                const float2 zero = make_float2(0.f, 0.f);
                float2 t;
                if (pk & kFlagFirst) {
                    const float2 y3 = __ldg(src + ROWF2);
                    attenuate_fast2<EXPM, kFitFirst>(FitCoeffs{}, zero, y2, y3, st, s_pairs, psi, t);
                } else if (pk & kFlagLast) {
                    const float2 y1 = __ldg(src - ROWF2);
                    attenuate_fast2<EXPM, kFitLast>(FitCoeffs{}, y1, y2, zero, st, s_pairs, psi, t);
                } else {
                    const float2 y1 = __ldg(src - ROWF2);
                    const float2 y3 = __ldg(src + ROWF2);
                    attenuate_fast2<EXPM, kFitInterior>(FitCoeffs{}, y1, y2, y3, st, s_pairs, psi, t);
                }
                red_add_v2(tally + off, t.x, t.y);                                  // kernel.c:276
            }
            cur_packed = nxt_packed;
            cur_qsr = nxt_qsr;
            draw(s0, b + 64 + lane, nseg, nxt_packed, nxt_qsr);
        }

        if (a.psi_out != nullptr)
            reinterpret_cast<float2 *>(a.psi_out)[(track - a.track_begin) * ROWF2 + lane] = psi;
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

// out[row][G] = flux0[row][G_pad] + tally[row][G_pad]   (kernel.c:276 summed over the sweep)
__global__ void finalize_flux(const float *__restrict__ flux0, const float *__restrict__ tally,
                              float *__restrict__ out, int64_t rows, int groups, int groups_pad, int replicas)
{
    const int64_t n = rows * groups;
    const int64_t stride = rows * groups_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups;
        const int g = (int)(i - r * groups);
        float t = tally[r * groups_pad + g];
        for (int k = 1; k < replicas; ++k) t += tally[k * stride + r * groups_pad + g];
        out[i] = flux0[r * groups_pad + g] + t;
    }
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
                                float *__restrict__ out, int64_t rows, int groups, int groups_pad)
{
    const int64_t n = rows * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups;
        const int g = (int)(i - r * groups);
        out[i] = (float)((double)flux0[r * groups_pad + g] + tally64[r * groups_pad + g]);
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

__global__ void debug_ids_kernel(uint64_t seed, int64_t seg_begin, int64_t n, FastMod mr, FastMod mf,
                                 int32_t *qsr, int32_t *fai)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const SegmentIds id = segment_ids(seed, (uint64_t)(seg_begin + i), mr, mf);
        qsr[i] = (int32_t)id.qsr;
        fai[i] = (int32_t)id.fai;
    }
}

}  // namespace smk
