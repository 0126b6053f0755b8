// Microbenchmark: sustained rate of the packed FP32x2 instructions (FFMA2 / FMUL2 / FADD2), alone,
// mixed as in the attenuation kernel, and with ~35 % integer ALU instructions interleaved.
// 32 warps/SM, ILP 8 per thread.  Reports packed instructions per cycle per SM sub-partition
// (peak 0.5: one 2-cycle packed instruction every other cycle).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 4) k(float *out, float a, float b, int iters, uint32_t salt)
{
    float2 x[8];
    uint32_t n[4] = {threadIdx.x, threadIdx.x * 3u, salt, salt ^ threadIdx.x};
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(1.0f + threadIdx.x * 1e-3f + i, 1.0f - i * 1e-3f);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = __ffma2_rn(x[i], a2, b2);
                if (MODE == 1) x[i] = __fmul2_rn(x[i], a2);
                if (MODE == 2) x[i] = __fadd2_rn(x[i], b2);
                if (MODE == 3 || MODE == 4) {            // kernel-like mix 46 : 44 : 12
                    const int sel = (r * 8 + i) % 8;
                    if (sel < 4) x[i] = __ffma2_rn(x[i], a2, b2);
                    else if (sel < 7) x[i] = __fmul2_rn(x[i], a2);
                    else x[i] = __fadd2_rn(x[i], b2);
                }
                if (MODE == 5) x[i] = __ffma2_rn(x[i], x[(i + 1) & 7], x[(i + 3) & 7]);   // 3 register operands
                if (MODE == 4 && (i & 1)) {              // + integer ALU work, ~1 per 2 packed ops
                    n[i >> 1] = (n[i >> 1] ^ (n[i >> 1] >> 3)) + salt;
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(n[0] + n[1] + n[2] + n[3]);
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
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 1e-4f, 10, 7u);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 1e-4f, iters, 7u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double packed = 48.0 * iters * 8 /*warps per SMSP*/;
    printf("%-44s %7.3f ms  %.3f packed instr / cycle / SMSP (peak 0.5)\n", name, ms, packed / (ms * 1e-3 * khz * 1e3));
    cudaFree(out);
}

int main()
{
    run<0>("FFMA2 (imm-free, 1 reg + 2 invariant)");
    run<1>("FMUL2");
    run<2>("FADD2");
    run<3>("mix FFMA2:FMUL2:FADD2 = 4:3:1");
    run<4>("same mix + integer ALU (1 per 2 packed)");
    run<5>("FFMA2 with 3 distinct register operands");
    return 0;
}
