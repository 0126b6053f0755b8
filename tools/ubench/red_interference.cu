// Microbenchmark: how much does one tally per ~100 FFMA2 slow the FMA pipe down, for the LSU
// path (red.global.add.v4.f32) and the TMA path (st.shared + cp.reduce.async.bulk)?
// Mirrors attenuate_tracks at 128 groups: 32 warps/SM, one 512-byte row tally per warp per
// ~100 packed FP32 instructions, rows pseudo-random among 33750.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int MODE>   // 0 math only, 1 + red.v4, 2 + TMA bulk reduce, 3 + 4 scalar reds
__global__ void __launch_bounds__(256, 4) k(float *tally, uint32_t rows, int iters, float a, float b)
{
    __shared__ __align__(128) float4 stage[8][2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t h = (blockIdx.x * 8 + warp) * 2654435761u + 12345u;
    float2 x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = make_float2(lane + i, lane - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 25; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
        h = hash32(h + it);
        const uint32_t row = h % rows;
        float *dst = tally + (size_t)row * 128 + lane * 4;
        if (MODE == 1) {
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x[0].x), "f"(x[1].x), "f"(x[2].x), "f"(x[3].x) : "memory");
        } else if (MODE == 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(tally + (size_t)row * 128 + j * 32 + lane), "f"(x[j].x) : "memory");
        } else if (MODE == 2) {
            float4 *buf = &stage[warp][it & 1][0];
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            buf[lane] = make_float4(x[0].x, x[1].x, x[2].x, x[3].x);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t s = (uint32_t)__cvta_generic_to_shared(buf);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;" ::"l"(tally + (size_t)row * 128), "r"(s) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (MODE == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (x[0].x + x[1].y + x[2].x + x[3].y == 12345.678f) tally[0] = 1.f;
}

template <int MODE>
void run(const char *name, float *tally, uint32_t rows)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int iters = 20000, blocks = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(tally, rows, 200, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(tally, rows, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc_per_iter_sm = ms * 1e-3 * khz * 1e3 / ((double)iters * 32);   // per warp-iteration per SM
    printf("%-40s %8.3f ms  %6.1f SM-cycles per warp-iteration (100 FFMA2 = 50.0 at peak)\n", name, ms, cyc_per_iter_sm);
}

int main()
{
    const uint32_t rows = 33750;
    float *tally;
    cudaMalloc(&tally, (size_t)rows * 128 * sizeof(float));
    cudaMemset(tally, 0, (size_t)rows * 128 * sizeof(float));
    run<0>("math only", tally, rows);
    run<1>("math + red.v4.f32 (LSU)", tally, rows);
    run<3>("math + 4 x red.f32 (LSU)", tally, rows);
    run<2>("math + st.shared + bulk reduce (TMA)", tally, rows);
    return 0;
}
