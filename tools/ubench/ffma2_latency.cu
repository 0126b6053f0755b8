// Microbenchmark: FFMA2 / FFMA dependent-chain throughput at the attenuation kernel's occupancy
// (32 warps/SM = 8 per SM sub-partition) as a function of per-thread ILP.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool PACKED>
__global__ void __launch_bounds__(256) k(float *out, float a, float b, int iters)
{
    float2 x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (PACKED) x[i] = __ffma2_rn(x[i], a2, b2);
                else { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP, bool PACKED>
void run(int blocks_per_sm)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int blocks = sms * blocks_per_sm, threads = 256, iters = 4000;
    float *out;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP, PACKED><<<blocks, threads>>>(out, 1.0001f, 0.5f, 10);
    cudaEventRecord(e0);
    k<ILP, PACKED><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 2.0 * ILP * 8 * iters * (double)blocks * threads;
    printf("%s ILP=%d warps/SM=%2d : %6.1f FMA/clk/SM (peak 128)\n", PACKED ? "FFMA2" : "FFMA ", ILP, blocks_per_sm * 8,
           fma / (ms * 1e-3) / (khz * 1e3) / sms);
    cudaFree(out);
}

int main()
{
    for (int b = 1; b <= 4; b *= 2) {
        run<1, true>(b); run<2, true>(b); run<4, true>(b);
        run<1, false>(b); run<2, false>(b); run<4, false>(b);
    }
    return 0;
}
