// smk_kernels_tuning.cuh -- kernel variants that were MEASURED AND REJECTED.
//
// Only compiled with -DSMK_TUNING (`make -C simplemoc-kernel_b200 tuning` -> lib/libsmk_tuning.so); the
// shipped libsmk.so does not contain them.  In a tuning build SMK_KERNEL=<variant> selects one for the
// 65..128-group FAST shape (reference geometry, fp32 tallies):
//   staged2 | staged3   rows staged through shared memory by TMA bulk copies (cp.async.bulk + mbarrier ring)
//   prefetch            rows of segment s+1 loaded into a second register set before segment s is computed
//   defer               RED of segment s-1 issued under the loads of segment s
//   l1pf                prefetch.global.L1 of the next segment's rows
//   oldflat             round 1's flat loop (double-buffered ids, 64-bit index arithmetic on the FMA pipe)
// tests/test_gpu_parity.py::test_tuning_variants_parity runs every variant against the oracle when the
// tuning library is present.
#pragma once

namespace smk {

#ifndef SMK_MIN_BLOCKS_PREFETCH
#define SMK_MIN_BLOCKS_PREFETCH 3
#endif
constexpr int kMinBlocksPrefetch = SMK_MIN_BLOCKS_PREFETCH;

__device__ __forceinline__ int64_t claim_tracks(const KernelArgs &a, int lane, int n)
{
    const int64_t w = claim_work(a, lane, n);
    return a.track_begin + w;
}

__device__ __forceinline__ void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
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
        attenuate_fast2<EXPM, FIT, false>(FitCoeffs{}, make_float2(y1.x, y1.y), make_float2(y2.x, y2.y),
                                   make_float2(y3.x, y3.y), make_float2(st.x, st.y), s_pairs, p_lo, t_lo);
        attenuate_fast2<EXPM, FIT, false>(FitCoeffs{}, make_float2(y1.z, y1.w), make_float2(y2.z, y2.w),
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
        attenuate_fast2<EXPM, FIT, false>(FitCoeffs{}, make_float2(r.y1[c].x, r.y1[c].y), make_float2(r.y2[c].x, r.y2[c].y),
                                   make_float2(r.y3[c].x, r.y3[c].y), make_float2(r.st[c].x, r.st[c].y), s_pairs,
                                   p_lo, t_lo);
        attenuate_fast2<EXPM, FIT, false>(FitCoeffs{}, make_float2(r.y1[c].z, r.y1[c].w), make_float2(r.y2[c].z, r.y2[c].w),
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
                    {
                        float4 *tal = reinterpret_cast<float4 *>(tally) + off;
#pragma unroll
                        for (int c = 0; c < NCHUNK; ++c) red_add_v4(tal + c * 32, t[c].x, t[c].y, t[c].z, t[c].w);
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


template <int NCHUNK, int STAGES>
static AttenuateFn pick_staged_exp(int expm)
{
    switch (expm) {
        case kExpPoly: return attenuate_tracks_staged<NCHUNK, kExpPoly, STAGES>;
        case kExpMufu: return attenuate_tracks_staged<NCHUNK, kExpMufu, STAGES>;
    }
    return nullptr;
}

template <bool PREFETCH, bool DEFER, bool L1PF>
static AttenuateFn pick_pf_exp(int expm)
{
    switch (expm) {
        case kExpPoly: return attenuate_tracks_pf<1, kExpPoly, PREFETCH, DEFER, L1PF>;
        case kExpMufu: return attenuate_tracks_pf<1, kExpMufu, PREFETCH, DEFER, L1PF>;
    }
    return nullptr;
}

// ------------------------------------------------------------------------------
// attenuate_warp_track_pipe<EXPM>: attenuate_warp_track<4> (65..128 groups, constant geometry, f32 tallies, 32-bit
// offsets) with the rows of segment k+1 copied global -> shared by per-lane cp.async (SASS LDGSTS, no registers held,
// no barrier: every lane copies and later reads only its own 16 bytes of each row) while segment k is computed.
// Measured (profiles/ab_r02.md r02x): psi bit-identical, 7.01e11 vs 7.45e11 (-6 %): the one-track-per-warp kernel is not
// bound by its load latency (32 resident warps cover it); the extra LDGSTS / LDS / DEPBAR instructions cost more.
// ------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int EXPM>
__global__ void __launch_bounds__(kThreadsPerBlock, kMinBlocksFast)
attenuate_warp_track_pipe(const KernelArgs a)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int kWarps = kThreadsPerBlock / 32;
    constexpr uint32_t ROWV = 32;

    __shared__ float2 s_pairs[kTableReach];
    __shared__ float4 s_stage[kWarps][2][4][32];          // [warp][slot][sigT, y1, y2, y3][lane]
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
    float *const tally = warp_tally(a, warp_global);
    const int64_t n_tracks = a.track_end - a.track_begin;
    unsigned long long checksum = 0ull;
    const FitCoeffs fc = {};
    float4 *const stage = &s_stage[warp][0][0][lane];     // slot s, row r at stage[(s * 4 + r) * 32]

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
                my_sg = (qsr * ROWV) | (fai == 0u ? kSgFirst : 0u) | (fai == F - 1u ? kSgLast : 0u);
            }
            const int count = (nseg - b) < 32 ? (nseg - b) : 32;
            // copies of segment kk of the batch into slot kk & 1; returns its (idx, sg)
            auto issue = [&](int kk, uint32_t &idx_out, uint32_t &sg_out) {
                const uint32_t idx = __shfl_sync(kFull, my_pk, kk) | (uint32_t)lane;
                const uint32_t sg = __shfl_sync(kFull, my_sg, kk);
                const float4 *src = ptr_add_index<true>(a.source, idx);
                float4 *dst = stage + (kk & 1) * 4 * 32;
                cp_async16(dst, ptr_add_index<true>(a.sigT, (sg & ~(kSgFirst | kSgLast)) | (uint32_t)lane));
                cp_async16(dst + 2 * 32, src);
                if ((int32_t)sg >= 0) cp_async16(dst + 1 * 32, src - ROWV);          // not the first interval
                if (!(sg & kSgLast)) cp_async16(dst + 3 * 32, src + ROWV);           // not the last interval
                cp_async_commit();
                idx_out = idx;
                sg_out = sg;
            };
            uint32_t idx, sg, idx_next = 0u, sg_next = 0u;
            issue(0, idx, sg);
            for (int k = 0; k < count; ++k) {
                if (k + 1 < count) {
                    issue(k + 1, idx_next, sg_next);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                const float4 *slot = stage + (k & 1) * 4 * 32;
                const float4 st = slot[0], y2 = slot[2 * 32];
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 t;
                if ((int32_t)sg < 0) {
                    attenuate_lane<EXPM, kFitFirst, false>(fc, zero, y2, slot[3 * 32], st, s_pairs, psi, t);
                } else if (sg & kSgLast) {
                    attenuate_lane<EXPM, kFitLast, false>(fc, slot[1 * 32], y2, zero, st, s_pairs, psi, t);
                } else {
                    attenuate_lane<EXPM, kFitInterior, false>(fc, slot[1 * 32], y2, slot[3 * 32], st, s_pairs, psi, t);
                }
                tally_lane<false, true>(tally, nullptr, idx, t);                            // kernel.c:276
                idx = idx_next;
                sg = sg_next;
            }
        }

        if (a.psi_out != nullptr)
            reinterpret_cast<float4 *>(a.psi_out)[(track - a.track_begin) * ROWV + lane] = psi;
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) checksum += __shfl_xor_sync(kFull, checksum, off);
    if (lane == 0 && checksum != 0ull) atomicAdd(a.checksum, checksum);
}

static AttenuateFn pick_pipe_exp(int expm)
{
    switch (expm) {
        case kExpPoly: return attenuate_warp_track_pipe<kExpPoly>;
        case kExpMufu: return attenuate_warp_track_pipe<kExpMufu>;
    }
    return nullptr;
}

}  // namespace smk

struct smk_ctx;
static int smk_tuning_select(smk_ctx *c);
