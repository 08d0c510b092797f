// Ingest: PCM16 -> float scaling (frontend/wav.h:98-104 + speakerDiarizer.cpp:2948-2951), the chunk
// geometry of SegmentModel::slide (1407-1470) and SegmentModel::crop (1641-1662) as a batched zero-padding gather,
// so a caller can keep the whole file on the device and hand chunk batches to ONNX Runtime by device pointer.
#include "common.cuh"

#include <cmath>
#include <vector>

namespace sdb {

int upload_small(sd_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);

// int16 -> float, * 1.0f / 32768.0 (a power of two: exact in fp32, so the double division of the reference
// collapses to one multiply)
__global__ void __launch_bounds__(256) ingest_pcm16_kernel(const short* __restrict__ pcm, long n, float* __restrict__ out) {
    const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i + 8 <= n && ((reinterpret_cast<uintptr_t>(pcm + i) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out + i) & 15) == 0)) {
        const int4 v = *reinterpret_cast<const int4*>(pcm + i);
        const int w[4] = {v.x, v.y, v.z, v.w};
        float4 a, b;
        float* o = reinterpret_cast<float*>(&a);
        float* o2 = reinterpret_cast<float*>(&b);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float lo = (float)(short)(w[q] & 0xffff) * (1.0f / 32768.0f);
            const float hi = (float)(short)(w[q] >> 16) * (1.0f / 32768.0f);
            if (q < 2) {
                o[2 * q] = lo;
                o[2 * q + 1] = hi;
            } else {
                o2[2 * (q - 2)] = lo;
                o2[2 * (q - 2) + 1] = hi;
            }
        }
        *reinterpret_cast<float4*>(out + i) = a;
        *reinterpret_cast<float4*>(out + i + 4) = b;
    } else {
        for (long j = i; j < n && j < i + 8; ++j) out[j] = (float)pcm[j] * (1.0f / 32768.0f);
    }
}

// out[c][j] = wave[first[c] + j] inside [0, n), else 0
__global__ void __launch_bounds__(256)
    crop_chunks_kernel(const float* __restrict__ wave, long n, const long* __restrict__ first, int L,
                       float* __restrict__ out) {
    const int c = blockIdx.y;
    const long base = first[c];
    float* o = out + (size_t)c * L;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) {
        const long s = base + j;
        o[j] = (s >= 0 && s < n) ? wave[s] : 0.f;
    }
}

int ingest_pcm16_launch(sd_ctx* ctx, const short* d_pcm, long n, float* d_out) {
    const long threads = (n + 7) / 8;
    ingest_pcm16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(d_pcm, n, d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

int crop_chunks_launch(sd_ctx* ctx, const float* d_wave, long n, const double* starts_s, int n_chunks, double duration,
                       int sample_rate, float* d_out) {
    const int L = (int)std::floor(duration * sample_rate);  // SD:1646
    std::vector<long> first((size_t)n_chunks);
    for (int c = 0; c < n_chunks; ++c) first[c] = (long)(int)std::floor(starts_s[c] * sample_rate);  // SD:1643
    long* d_first = (long*)ctx->scratch(BUF_CL_MISC, sizeof(long) * (size_t)n_chunks);
    if (!d_first) return SD_ERR_NOMEM;
    int rc = upload_small(ctx, d_first, first.data(), sizeof(long) * (size_t)n_chunks);
    if (rc) return rc;
    const dim3 grid((unsigned)std::min(64, (L + 255) / 256), (unsigned)n_chunks);
    crop_chunks_kernel<<<grid, 256, 0, ctx->stream>>>(d_wave, n, d_first, L, d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

}  // namespace sdb
