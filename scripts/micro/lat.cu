// Micro-benchmarks of the serial building blocks of the linkage merge loop (one warp, dependent chains).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Top { double v; int i; int c; };
__device__ __forceinline__ Top warp_top(double v, int i, int c) {
    const unsigned full = 0xffffffffu;
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(full, hi);
    const bool cand = (hi == mh) && c > 0;
    const unsigned b = __ballot_sync(full, cand);
    Top t;
    if (b == 0u) { t.v = INFINITY; t.i = -1; t.c = 0; }
    else if ((b & (b - 1u)) == 0u) {
        const int src = __ffs(b) - 1;
        t.v = __shfl_sync(full, v, src); t.i = __shfl_sync(full, i, src); t.c = __shfl_sync(full, c, src);
    } else {
        const unsigned ml = __reduce_min_sync(full, cand ? lo : 0xffffffffu);
        const bool is_min = cand && (lo == ml);
        t.c = (int)__reduce_add_sync(full, is_min ? (unsigned)c : 0u);
        t.i = (int)__reduce_min_sync(full, is_min ? (unsigned)i : 0xffffffffu);
        t.v = __hiloint2double((int)mh, (int)ml);
    }
    return t;
}

__global__ void k(double* g, long long* out, int n) {
    __shared__ double sm[1024];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 37) % 1024);
    __syncthreads();
    long long t0, t1;
    int r = 0;
    const int R = 64;
    // 1. warp_top chain
    double v = 1.0 + lane * 0.5 + g[lane];
    t0 = clock64();
    for (int q = 0; q < R; ++q) { Top t = warp_top(v, lane, 1); v = t.v + lane * 0.5 + 1.0; }
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 2. dependent LDS chain
    int idx = lane;
    t0 = clock64();
    for (int q = 0; q < R; ++q) idx = (int)sm[idx & 1023];
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 3. dependent shfl (64-bit) chain
    t0 = clock64();
    for (int q = 0; q < R; ++q) v = __shfl_sync(0xffffffffu, v, (lane + 1) & 31) + 1.0;
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 4. dependent DFMA chain
    double a = v;
    t0 = clock64();
    for (int q = 0; q < R; ++q) a = __fma_rn(a, 1.0000001, 0.5);
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 5. dsqrt chain
    t0 = clock64();
    for (int q = 0; q < R; ++q) a = __dsqrt_rn(a + 3.0);
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 6. ddiv chain
    t0 = clock64();
    for (int q = 0; q < R; ++q) a = __ddiv_rn(a + 3.0, 1.7);
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 7. dependent ldcg chain (L2 hits after first pass)
    long off = lane;
    for (int q = 0; q < R; ++q) off = (long)__ldcg(g + 64 + (off & 1023));
    t0 = clock64();
    for (int q = 0; q < R; ++q) off = (long)__ldcg(g + 64 + (off & 1023));
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 8. scattered stores (32 lanes x stride) followed by __threadfence
    t0 = clock64();
    for (int q = 0; q < 8; ++q) {
        for (int u = 0; u < 2; ++u) __stcg(g + 4096 + (size_t)(lane + 32 * u) * 2048 + q, a + q);
        __threadfence();
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / 8; r++;
    // 9. __threadfence with nothing outstanding
    t0 = clock64();
    for (int q = 0; q < 8; ++q) __threadfence();
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / 8; r++;
    // 10. __syncthreads
    t0 = clock64();
    for (int q = 0; q < R; ++q) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 11. clock64 overhead
    t0 = clock64();
    long long acc = 0;
    for (int q = 0; q < R; ++q) acc += clock64();
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 12. redux chain
    unsigned uu = lane;
    t0 = clock64();
    for (int q = 0; q < R; ++q) uu = __reduce_min_sync(0xffffffffu, uu + lane) + 1;
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 13. shfl 32-bit chain
    t0 = clock64();
    for (int q = 0; q < R; ++q) uu = __shfl_sync(0xffffffffu, uu, (lane + 1) & 31) + 1;
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 14. ballot chain
    t0 = clock64();
    for (int q = 0; q < R; ++q) uu = __ballot_sync(0xffffffffu, (uu + lane) & 1) + 1;
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / R; r++;
    // 15. DADD / DMUL chain
    t0 = clock64();
    for (int q = 0; q < R; ++q) a = __dmul_rn(__dadd_rn(a, 0.25), 0.999);
    t1 = clock64();
    if (threadIdx.x == 0) out[r] = (t1 - t0) / (2 * R); r++;
    g[threadIdx.x] = a + v + idx + off + acc + uu;
}

int main() {
    double* g; long long* out;
    cudaMalloc(&g, sizeof(double) * (4096 + 64 * 2048 + 64));
    cudaMemset(g, 0, sizeof(double) * (4096 + 64 * 2048 + 64));
    cudaMalloc(&out, 64 * 8);
    const char* names[] = {"warp_top", "LDS dep", "SHFL64+DADD dep", "DFMA dep", "dsqrt_rn(+add)", "ddiv_rn(+add)", "ldcg L2 dep",
                           "2x32 scattered stcg + threadfence", "threadfence idle", "syncthreads", "clock64", "redux.min(+add)",
                           "SHFL32(+add)", "ballot(+2)", "DADD/DMUL each"};
    for (int threads : {32, 256}) {
        k<<<1, threads>>>(g, out, 1);
        cudaDeviceSynchronize();
        k<<<1, threads>>>(g, out, 1);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[16];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads=%d (%s)\n", threads, cudaGetErrorString(e));
        for (int i = 0; i < 15; ++i) printf("  %-36s %lld cycles\n", names[i], h[i]);
    }
    return 0;
}
