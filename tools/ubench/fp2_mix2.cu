// Microbenchmark: packed FP32x2 throughput at the attenuation kernel's shape: 8 warps per SM
// sub-partition, ILP 2 per thread (two independent packed chains), with 0 / 1-per-4 / 1-per-2 integer
// ALU instructions interleaved, and with 3-register-operand FFMA2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ILP, int ALU_EVERY, bool THREE_REG>
__global__ void __launch_bounds__(256, 4) k(float *out, float a, float b, int iters, uint32_t salt)
{
    float2 x[ILP], y[ILP];
    uint32_t n = threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = make_float2(1.0f + threadIdx.x * 1e-3f + i, 1.0f - i * 1e-3f); y[i] = make_float2(a + i, b - i); }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 48 / ILP; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (THREE_REG) x[i] = __ffma2_rn(x[i], y[i], y[(i + 1) % ILP]);
                else x[i] = __ffma2_rn(x[i], a2, b2);
                if (ALU_EVERY > 0 && ((r * ILP + i) % ALU_EVERY) == 0) n = n * 1u + (n >> 7) + salt;   // SHF/LEA-type ALU work
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)n;
}

template <int ILP, int ALU_EVERY, bool THREE_REG>
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
    k<ILP, ALU_EVERY, THREE_REG><<<blocks, 256>>>(out, 1.0001f, 1e-4f, 10, 7u);
    cudaEventRecord(e0);
    k<ILP, ALU_EVERY, THREE_REG><<<blocks, 256>>>(out, 1.0001f, 1e-4f, iters, 7u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-52s %7.3f ms  %.3f packed instr / cycle / SMSP (peak 0.5)\n", name, ms, 48.0 * iters * 8 / (ms * 1e-3 * khz * 1e3));
    cudaFree(out);
}

int main()
{
    run<2, 0, false>("ILP 2, FFMA2 only");
    run<2, 4, false>("ILP 2, + 1 ALU per 4 packed");
    run<2, 2, false>("ILP 2, + 1 ALU per 2 packed");
    run<2, 1, false>("ILP 2, + 1 ALU per packed");
    run<2, 0, true>("ILP 2, 3-register FFMA2");
    run<4, 0, true>("ILP 4, 3-register FFMA2");
    run<1, 0, false>("ILP 1, FFMA2 only");
    run<1, 2, false>("ILP 1, + 1 ALU per 2 packed");
    return 0;
}
