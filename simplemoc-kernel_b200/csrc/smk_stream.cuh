// smk_stream.cuh -- the deterministic counter-based stream.
//
// Replaces the reference's time-seeded draws: rand_r() for (QSR_id, FAI_id) and the
// initial state_flux (/root/reference/src/cpu/kernel.c:15,29-30,47,50; the CUDA
// reference's per-block XORWOW states, /root/reference/src/cuda/kernel.cu:22,52-60)
// and rand() for the source slabs (/root/reference/src/cpu/init.c:64-75).
// Pure 32-bit integer arithmetic, identical on host and device; the CPU oracle
// restates it independently in its own C file and both are pinned to the
// Random123 Philox4x32-10 known-answer vectors.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SMK_HD __host__ __device__ __forceinline__
#else
#define SMK_HD inline
#endif

namespace smk {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

constexpr uint32_t kDomainSegment = 0x5345474Du;  // 'SEGM'
constexpr uint32_t kDomainPsi     = 0x50534930u;  // 'PSI0'
constexpr uint32_t kDomainFill    = 0x46494C4Cu;  // 'FILL'

struct u32x4 { uint32_t x, y, z, w; };

SMK_HD void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

// Philox4x32-10.  The key schedule only depends on the seed, so the compiler
// hoists the ten bumped keys out of loops.
SMK_HD u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                           uint32_t k0, uint32_t k1)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(kPhiloxM0, c0, hi0, lo0);
        mulhilo(kPhiloxM1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return u32x4{c0, c1, c2, c3};
}

// The key schedule (k0 + r W0, k1 + r W1), r = 0..9, only depends on the seed: the host expands it once
// and the kernels read the 20 words straight from the constant bank (kernel parameters), which removes
// two integer adds per round from every hash.
struct PhiloxKeys {
    uint32_t k[20];
};

inline PhiloxKeys make_philox_keys(uint64_t seed)
{
    PhiloxKeys ks;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        ks.k[2 * r] = k0;
        ks.k[2 * r + 1] = k1;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return ks;
}

SMK_HD u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &ks)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(kPhiloxM0, c0, hi0, lo0);
        mulhilo(kPhiloxM1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ ks.k[2 * r];
        uint32_t n2 = hi0 ^ c3 ^ ks.k[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return u32x4{c0, c1, c2, c3};
}

SMK_HD u32x4 stream_words(const PhiloxKeys &ks, uint64_t index, uint32_t sub, uint32_t domain)
{
    return philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), sub, domain, ks);
}

SMK_HD u32x4 stream_words(uint64_t seed, uint64_t index, uint32_t sub, uint32_t domain)
{
    return philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), sub, domain,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// (float) rand() / RAND_MAX of the reference == (float)r * 2^-31, r a 31-bit draw.
SMK_HD float u01(uint32_t w)
{
    return (float)(int32_t)(w >> 1) * 4.656612873077392578125e-10f;  // 2^-31
}

// n % d for 31-bit n by multiply-high: q = floor(n * M / 2^(31+s)), exact for all
// n < 2^31 with s = ceil(log2 d), M = floor(2^(31+s) / d) + 1 < 2^32 (Granlund &
// Montgomery).  `%` itself mirrors kernel.c:47,50.
struct FastMod {
    uint32_t d, M, s;
};

inline FastMod make_fastmod(uint32_t d)
{
    FastMod f;
    f.d = d;
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;
    f.s = s;
    f.M = (d == 1) ? 0u : (uint32_t)(((1ull << (31 + s)) / d) + 1ull);
    return f;
}

SMK_HD uint32_t fastmod(uint32_t n, const FastMod &f)  // n < 2^31
{
    if (f.d == 1) return 0u;
#if defined(__CUDA_ARCH__)
    uint32_t q = __umulhi(n, f.M) >> (f.s - 1);
#else
    uint32_t q = (uint32_t)(((uint64_t)n * f.M) >> 32) >> (f.s - 1);
#endif
    return n - q * f.d;
}

// --------------------------------------------------------------------------
// Per-segment geometry (SMK_FLAG_SEGMENT_GEOMETRY).  kernel.c:95-104: "Some placeholder
// constants - In the full app some of these are calculated based off position in geometry":
// dz, zin, weight, mu, mu2, ds become parameters (smk_geometry) and, with the flag, vary per
// segment.  Words 2,3 of the segment's stream block give four 16-bit fields u, each mapped to
// a factor f(u) = 1 + spread * (u * 2^-15 - 1) in [1 - spread, 1 + spread):
//   ds = ds0 f(w2 >> 16)       zin = zin0 f(w2 & 0xFFFF)
//   mu = mu0 f(w3 >> 16)       mu2 = mu2_0 f(w3 >> 16)^2      weight = weight0 f(w3 & 0xFFFF)
// dz (the axial mesh spacing) is the same for every segment.  Each step is ONE IEEE binary32
// operation, never contracted, so host, oracle and every kernel derive bit-identical values;
// spread = 0 gives f = 1 exactly, i.e. the base values.
// --------------------------------------------------------------------------
struct SegGeometry {
    float dz, zin, weight, mu, mu2, ds;
};

struct GeometryBase {       // smk_geometry as the kernels take it
    float dz, zin, weight, mu, mu2, ds, spread;
};

SMK_HD float geom_factor(uint32_t u16, float spread)
{
#if defined(__CUDA_ARCH__)
    const float c = __fadd_rn(__fmul_rn((float)u16, 3.0517578125e-05f), -1.0f);   // u * 2^-15 - 1, exact
    return __fadd_rn(1.0f, __fmul_rn(spread, c));
#else
    const float c = (float)u16 * 3.0517578125e-05f - 1.0f;
    const float sc = spread * c;
    return 1.0f + sc;
#endif
}

SMK_HD float mul_rn(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}

SMK_HD SegGeometry segment_geometry(const GeometryBase &b, uint32_t w2, uint32_t w3)
{
    const float f_ds = geom_factor(w2 >> 16, b.spread);
    const float f_zin = geom_factor(w2 & 0xFFFFu, b.spread);
    const float f_mu = geom_factor(w3 >> 16, b.spread);
    const float f_w = geom_factor(w3 & 0xFFFFu, b.spread);
    SegGeometry g;
    g.dz = b.dz;
    g.zin = mul_rn(b.zin, f_zin);
    g.weight = mul_rn(b.weight, f_w);
    g.mu = mul_rn(b.mu, f_mu);
    g.mu2 = mul_rn(b.mu2, mul_rn(f_mu, f_mu));
    g.ds = mul_rn(b.ds, f_ds);
    return g;
}

struct SegmentIds { uint32_t qsr, fai; };

SMK_HD SegmentIds segment_ids(const PhiloxKeys &ks, uint64_t seg, const FastMod &mod_regions,
                              const FastMod &mod_fai)
{
    u32x4 w = stream_words(ks, seg, 0u, kDomainSegment);
    return SegmentIds{fastmod(w.x >> 1, mod_regions), fastmod(w.y >> 1, mod_fai)};
}

SMK_HD SegmentIds segment_ids(uint64_t seed, uint64_t seg, const FastMod &mod_regions,
                              const FastMod &mod_fai)
{
    u32x4 w = stream_words(seed, seg, 0u, kDomainSegment);
    // words z, w carry the per-segment geometry (segment_geometry above; kernel.c:95-104)
    return SegmentIds{fastmod(w.x >> 1, mod_regions), fastmod(w.y >> 1, mod_fai)};
}

// weight of segment s in the indexing fingerprint (include/smk.h: smk_download_checksum)
// (row + 1 < 2^31: smk_create checks regions * intervals < 2^30, so the first factor fits 32 bits and the term is ONE
// 32 x 32 -> 64-bit multiply-add on the device instead of a 64 x 32-bit product)
SMK_HD uint64_t checksum_term(uint32_t qsr, uint32_t fai, uint32_t fai_count, uint64_t seg)
{
    return (uint64_t)(qsr * fai_count + fai + 1u) * (uint64_t)(((uint32_t)seg & 0xFFFFu) + 1u);
}

}  // namespace smk
