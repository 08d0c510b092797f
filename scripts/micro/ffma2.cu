// Micro-benchmark: issue cost of packed fp32 (FADD2 / FFMA2, sm_100a) against scalar FADD / FFMA.
// Each thread runs 8 independent chains; 20 warps per SM like the STFT kernel (4 CTAs x 160 threads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
template <int MODE>
__global__ void __launch_bounds__(160, 4) bench(float* out, int iters, float seed) {
    float s[16];
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pk(s[2 * i], s[2 * i + 1]);
    const unsigned long long c = pk(1.0001f, 0.9999f), d = pk(0.5f, -0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) {  // 16 scalar FADD = 16 flop-lanes
                s[2 * u] += 1.0001f;
                s[2 * u + 1] += 0.9999f;
            } else if (MODE == 1) {  // 8 FADD2 = the same 16 adds
                asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[u]) : "l"(c));
            } else if (MODE == 2) {  // 16 scalar FFMA
                s[2 * u] = fmaf(s[2 * u], 1.0001f, 0.5f);
                s[2 * u + 1] = fmaf(s[2 * u + 1], 0.9999f, -0.5f);
            } else {  // 8 FFMA2
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[u]) : "l"(c), "l"(d));
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += s[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float a, b;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i]));
        acc += a + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
float run(float* out, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    bench<MODE><<<148 * 4, 160>>>(out, 16, 1.f);
    cudaEventRecord(a);
    bench<MODE><<<148 * 4, 160>>>(out, iters, 1.f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 4 * 160 * sizeof(float));
    const int iters = 200000;
    const char* names[4] = {"16 x FADD ", " 8 x FADD2", "16 x FFMA ", " 8 x FFMA2"};
    float ms[4] = {run<0>(out, iters), run<1>(out, iters), run<2>(out, iters), run<3>(out, iters)};
    for (int m = 0; m < 4; ++m) {
        // lane-operations per second: 148 SMs * 4 CTAs * 160 threads * iters * 16 per iteration
        const double ops = 148.0 * 4 * 160 * (double)iters * 16;
        printf("%s per iteration: %.3f ms -> %.2f T lane-ops/s (%.1f per clock per SM at 1.9 GHz)\n", names[m], ms[m],
               ops / ms[m] / 1e9, ops / ms[m] / 1e9 * 1e3 / 148 / 1.9e3 * 1e-0);
    }
    return 0;
}
