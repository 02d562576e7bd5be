// Microbenchmark: can scalar FFMA (fmalite?) run alongside packed FFMA2 (fmaheavy?) above the FFMA2-only rate?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mix ubench_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*(unsigned long long*)&a), "l"(*(unsigned long long*)&b), "l"(*(unsigned long long*)&c));
    return *(float2*)&d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
template <int NP, int NS>  // NP packed + NS scalar FMAs per group, 8 independent chains each
__global__ void __launch_bounds__(256) k(float* out, float a, int iters) {
    float2 p[8];
    float s[8];
    for (int i = 0; i < 8; ++i) { p[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); s[i] = i * 0.25f + threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (NP > 0) p[i] = ffma2(p[i], make_float2(0.999f, 0.998f), make_float2(1e-4f, 2e-4f));   // immediates
                if (NP > 1) p[i] = ffma2(p[i], make_float2(1.001f, 1.002f), make_float2(-1e-4f, -2e-4f));
                if (NS > 0) s[i] = ffma1(s[i], 0.999f, 1e-4f);
                if (NS > 1) s[i] = ffma1(s[i], 1.001f, -1e-4f);
            }
        }
    }
    float sum = 0;
    for (int i = 0; i < 8; ++i) sum += p[i].x + p[i].y + s[i];
    if (sum == 123.456f) out[0] = sum;
}
template <int NP, int NS>
void run(const char* name, int sms) {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048, blocks = sms * 8;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<NP, NS><<<blocks, 256>>>(d, 0.9f, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fma = (double)blocks * 256 * iters * 8 * 8 * (NP * 2 + NS);
        const double tf = fma * 2 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    printf("%-28s %6.1f TFLOP/s\n", name, best);
    cudaFree(d);
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<1, 0>("FFMA2 only (imm)", sms);
    run<2, 0>("FFMA2 x2", sms);
    run<0, 1>("FFMA only (imm)", sms);
    run<0, 2>("FFMA x2", sms);
    run<1, 1>("FFMA2 + FFMA (1:1)", sms);
    run<2, 1>("FFMA2 + FFMA (2:1)", sms);
    run<1, 2>("FFMA2 + FFMA (1:2)", sms);
    return 0;
}
