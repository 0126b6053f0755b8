// Microbenchmark: cost of a packed FP32x2 instruction as a function of how many distinct 64-bit
// REGISTER operands it reads (immediates and loop-invariant operands served by the reuse cache are free).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 4) k(float *out, float a, float b, int iters)
{
    float2 x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = make_float2(1.0f + threadIdx.x * 1e-3f + i, 1.0f - i * 1e-3f);
        y[i] = make_float2(a + i * 1e-6f, a - i * 1e-6f);
        z[i] = make_float2(b + i * 1e-6f, b - i * 1e-6f);
    }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = __ffma2_rn(x[i], a2, b2);                                   // 1 reg (+2 invariant)
                if (MODE == 1) x[i] = __fmul2_rn(x[i], y[i]);                                     // 2 regs
                if (MODE == 2) x[i] = __fadd2_rn(x[i], z[i]);                                     // 2 regs
                if (MODE == 3) x[i] = __ffma2_rn(x[i], y[i], make_float2(0.5f, 0.5f));            // 2 regs + imm
                if (MODE == 4) x[i] = __ffma2_rn(x[i], make_float2(1.0001f, 1.0001f), z[i]);      // 2 regs + imm
                if (MODE == 5) x[i] = __ffma2_rn(x[i], y[i], z[i]);                               // 3 regs
                if (MODE == 6) x[i] = __fmul2_rn(x[i], make_float2(1.0001f, 1.0001f));            // 1 reg + imm
                if (MODE == 7) x[i] = __ffma2_rn(x[i], y[i], x[(i + 1) & 7]);                     // 3 regs, all fresh
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int blocks = sms * 4, iters = 4000;
    float *out;
    cudaMalloc(&out, (size_t)blocks * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 1e-4f, 10);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 1e-4f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double rate = 48.0 * iters * 8 / (ms * 1e-3 * khz * 1e3);
    printf("%-40s %7.3f ms  %.3f packed/cycle/SMSP = %.2f cycles per instruction\n", name, ms, rate, 1.0 / rate);
    cudaFree(out);
}

int main()
{
    run<0>("FFMA2 x, inv, inv      (1 reg)");
    run<6>("FMUL2 x, imm           (1 reg)");
    run<1>("FMUL2 x, y             (2 regs)");
    run<2>("FADD2 x, z             (2 regs)");
    run<3>("FFMA2 x, y, imm        (2 regs)");
    run<4>("FFMA2 x, imm, z        (2 regs)");
    run<5>("FFMA2 x, y, z          (3 regs)");
    run<7>("FFMA2 x, y, x'         (3 regs)");
    return 0;
}
