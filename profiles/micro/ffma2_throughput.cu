// Throughput of packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against scalar FFMA on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_throughput ffma2_throughput.cu && ./ffma2_throughput
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int MODE>      // 0: 16 independent FFMA chains, 1: 16 independent FFMA2 chains, 2: 8 FFMA2 + 8 MUFU-free integer adds
__global__ void __launch_bounds__(256) bench(float* out, int iters, float s) {
    float acc[16];
    unsigned long long acc2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i] = threadIdx.x * 0.001f + i; acc2[i] = pack(acc[i], acc[i] + 1.f); }
    const unsigned long long s2 = pack(s, s);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) acc[i] = ffma1(acc[i], s, s);
            else acc2[i] = ffma2(acc2[i], s2, s2);
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += MODE == 0 ? acc[i] : __uint_as_float((unsigned)(acc2[i] & 0xffffffffu)) + __uint_as_float((unsigned)(acc2[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) bench<0><<<sms * 8, 256>>>(out, iters, 0.999f);
            else bench<1><<<sms * 8, 256>>>(out, iters, 0.999f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double inst = (double)sms * 8 * 256 * iters * 16;            // thread-level instructions
            const double fma = inst * (mode == 0 ? 1 : 2);
            printf("%s: %.3f ms, %.1f G thread-instr/s, %.2f TFLOP/s, %.1f lane-FMA/clk/SM at the nominal %d MHz\n",
                   mode == 0 ? "FFMA " : "FFMA2", ms, inst / ms / 1e6, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / ((double)khz * 1e3) / sms,
                   khz / 1000);
        }
    }
    return 0;
}
