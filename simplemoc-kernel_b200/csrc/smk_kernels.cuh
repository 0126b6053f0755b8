// smk_kernels.cuh -- sm_100a kernels of the segment-attenuation path.
//
//   attenuate_tracks<LPT, NCHUNK, MATH, EXPM>   the hot kernel: run_kernel's segment
//       loop + attenuate_segment (/root/reference/src/cpu/kernel.c:43-55, 75-333)
//   fill_rows                                    device-side deterministic fill
//       (replaces /root/reference/src/cpu/init.c:64-75 + the H2D of init.cu:105-127)
//   pad_rows / finalize_flux                     host layout <-> padded device layout
//
// Work decomposition of the hot kernel
//   track  = seg_per_track consecutive segments sharing one carried angular flux psi
//   a track is owned by LPT lanes of one warp (LPT = lanes per track, a power of two);
//   each lane owns NCHUNK float4 = 4*NCHUNK energy groups and keeps their psi in
//   registers for the whole track.  G = 128 -> LPT = 32, NCHUNK = 1 (one warp per
//   track, 128-bit loads, one 16-byte vector RED per lane per segment);
//   G = 64 -> LPT = 16 (2 tracks per warp); G = 7 -> G_pad = 8, LPT = 2 (16 tracks
//   per warp, the 8th group is padding).
//   Segment ids come from the counter stream: every LPT segments each lane of the
//   track hashes ONE upcoming segment and the ids are handed round with shuffles, so
//   the Philox cost per intersection is 1/(4*NCHUNK*LPT) of a hash.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#include "smk_math.cuh"
#include "smk_stream.cuh"

namespace smk {

struct KernelArgs {
    const float4 *__restrict__ source;   // [R][F][G_pad/4]
    const float4 *__restrict__ sigT;     // [R][G_pad/4]
    float *__restrict__ tally;           // [R][F][G_pad]
    float *__restrict__ psi_out;         // [tracks in launch][G_pad] or nullptr
    unsigned long long *checksum;        // indexing fingerprint accumulator
    int64_t segments;                    // N
    int64_t track_begin, track_end;
    uint64_t seed;
    FastMod mod_regions, mod_fai;
    int32_t fai_count;                   // F
    int32_t row_f4;                      // G_pad / 4: float4 per row
    int32_t seg_per_track;               // p
};

constexpr int kThreadsPerBlock = 256;
#ifndef SMK_MIN_BLOCKS_FAST
#define SMK_MIN_BLOCKS_FAST 4
#endif
// 4 x 256 threads x 64 registers = the whole register file: 32 warps/SM for the issue-bound FAST kernels
constexpr int kMinBlocksFast = SMK_MIN_BLOCKS_FAST;

__device__ __forceinline__ void red_add_v4(float4 *addr, float a, float b, float c, float d)
{
    // one 16-byte vector reduction at L2 per lane (PTX ISA 8.1, sm_90+): SASS RED.E.ADD.F32x4
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :
                 : "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

__device__ __forceinline__ float4 ldg4(const float4 *p)
{
    return __ldg(p);
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
    const int64_t total_slots = (int64_t)gridDim.x * (kThreadsPerBlock / 32) * kSlotsPerWarp;
    const int64_t slot = warp_global * kSlotsPerWarp + (lane / LPT);

    const int F = a.fai_count;
    const int row_f4 = a.row_f4;
    const int p = a.seg_per_track;
    unsigned long long checksum = 0ull;

    for (int64_t tbase = a.track_begin; tbase < a.track_end; tbase += total_slots) {
        const int64_t track = tbase + slot;
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
            const u32x4 w = stream_words(a.seed, (uint64_t)track, (uint32_t)(c * LPT + sub), kDomainPsi);
            psi[c] = make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w));
        }

        const int nseg_warp = (LPT == 32) ? nseg : __reduce_max_sync(kFull, nseg);

        for (int b = 0; b < nseg_warp; b += LPT) {
            // each lane of the track draws the ids of one of the next LPT segments
            uint32_t my_qsr = 0u, my_fai = 0u;
            if (b + sub < nseg) {
                const uint64_t seg = (uint64_t)(s0 + b + sub);
                const SegmentIds id = segment_ids(a.seed, seg, a.mod_regions, a.mod_fai);
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
                float4 *tal = reinterpret_cast<float4 *>(a.tally) + off;

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
                              float *__restrict__ out, int64_t rows, int groups, int groups_pad)
{
    const int64_t n = rows * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups;
        const int g = (int)(i - r * groups);
        out[i] = flux0[r * groups_pad + g] + tally[r * groups_pad + g];
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
