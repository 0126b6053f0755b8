// Microbenchmark: what FP32 rate can sm_100a sustain with scalar FFMA, packed FFMA2, and a mix?
// Gives the FP32-pipe ceiling that the attenuation kernel (FMA-pipe bound) is compared against.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_peak ffma2_peak.cu ; run on the GPU box
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0 scalar FFMA, 1 FFMA2, 2 mix (1 FFMA2 : 1 FFMA), 3 mix (1 FFMA2 : 2 FFMA)
__global__ void __launch_bounds__(256) k(float *out, float a, float b, int iters)
{
    float2 x[8];
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x + i, threadIdx.x - i); y[i] = threadIdx.x * 0.5f + i; }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
            if (MODE >= 1) x[i] = __ffma2_rn(x[i], a2, b2);
            if (MODE == 2) y[i] = fmaf(y[i], a, b);
            if (MODE == 3) { y[i] = fmaf(y[i], a, b); y[i] = fmaf(y[i], b, a); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double fma_per_thread_iter)
{
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    float *out;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, 100);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = fma_per_thread_iter * iters * (double)blocks * threads;
    const double per_clk_sm = fma / (ms * 1e-3) / (khz * 1e3) / sms;
    printf("%-28s %8.3f ms  %7.2f TFMA/s  %6.1f FMA/clk/SM at max clock %d MHz (peak 128)\n", name, ms, fma / (ms * 1e-3) / 1e12,
           per_clk_sm, khz / 1000);
    cudaFree(out);
}

int main()
{
    run<0>("scalar FFMA", 16);
    run<1>("packed FFMA2", 16);
    run<2>("FFMA2 + FFMA (1:1)", 24);
    run<3>("FFMA2 + 2 FFMA (1:2)", 32);
    return 0;
}
