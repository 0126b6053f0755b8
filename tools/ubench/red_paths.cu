// Microbenchmark: throughput of the tally path on sm_100a.
//   A  red.global.add.f32        (scalar, 4 per lane)           LSU path
//   B  red.global.add.v2.f32     (2 per lane)
//   C  red.global.add.v4.f32     (1 per lane, what the kernel uses)
//   D  cp.reduce.async.bulk .add.f32 of a 512-byte row staged in shared memory (TMA path)
// Each warp adds 512-byte rows at pseudo-random row indices of a [rows][128] float array, the
// access pattern of attenuate_tracks at 128 energy groups.  Reports warp-rows/s and bytes/clk/SM
// for a full grid and for a single SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float *tally, uint32_t rows, int iters)
{
    __shared__ __align__(128) float4 stage[8][2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t h = (blockIdx.x * 8 + warp) * 2654435761u + 12345u;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int it = 0; it < iters; ++it) {
        h = hash32(h + it);
        const uint32_t row = h % rows;
        float *dst = tally + (size_t)row * 128 + lane * 4;
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(tally + (size_t)row * 128 + j * 32 + lane), "f"(v.x) : "memory");
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(tally + (size_t)row * 128 + j * 64 + lane * 2), "f"(v.x), "f"(v.y) : "memory");
        } else if (MODE == 2) {
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        } else {
            float4 *buf = &stage[warp][it & 1][0];
            // the bulk reduce issued 2 iterations ago must have finished READING this buffer
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            buf[lane] = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t s = (uint32_t)__cvta_generic_to_shared(buf);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;" ::"l"(tally + (size_t)row * 128), "r"(s) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (MODE == 3 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>
void run(const char *name, int blocks, float *tally, uint32_t rows)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(tally, rows, 200);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(tally, rows, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    const double warp_rows = (double)blocks * 8 * iters;
    const int used_sms = blocks < sms * 4 ? (blocks + 3) / 4 : sms;
    printf("%-34s blocks=%4d  %8.3f ms  %.3e rows/s  %6.2f B/clk/SM  %7.1f GB/s  %s\n", name, blocks, ms, warp_rows / (ms * 1e-3),
           warp_rows * 512 / (ms * 1e-3) / (khz * 1e3) / used_sms, warp_rows * 512 / (ms * 1e-3) / 1e9,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const uint32_t rows = 33750;
    float *tally;
    cudaMalloc(&tally, (size_t)rows * 128 * sizeof(float));
    cudaMemset(tally, 0, (size_t)rows * 128 * sizeof(float));
    for (int blocks : {sms * 4, 4}) {
        run<0>("A red.f32 x4 per lane", blocks, tally, rows);
        run<1>("B red.v2.f32 x2 per lane", blocks, tally, rows);
        run<2>("C red.v4.f32 x1 per lane", blocks, tally, rows);
        run<3>("D cp.reduce.async.bulk 512B (TMA)", blocks, tally, rows);
    }
    return 0;
}
