// Pre-embedding masking: Helper::interpolate (speakerDiarizer.cpp:746-767), Helper::padSequence
// (770-797) and the wav_lens / too-short logic of getEmbedding (2466-2510), plus the per-(chunk, speaker) choice
// between the clean and the raw mask made in speakerDiarization() (3047-3082).
//
// The nearest-neighbour upsampled mask is constant over the samples of one segmentation frame
// (src = j*F/L), so stream compaction reduces to copying whole runs: a prefix sum over the <= F active frames
// gives every run its destination, then all threads copy with coalesced reads.  HBM-bound: 4*L bytes read and
// 4*L written per item (the zero tail included).
#include "common.cuh"

namespace sdb {

// first sample j with j*F/L == f  (integer division as in the reference)
__host__ __device__ inline int frame_first_sample(int f, int F, int L) { return (int)(((long)f * L + F - 1) / F); }

// One CTA per item.  wav_base[b] = offset of the item's first sample inside `wav` (samples beyond wav_limit are
// the zero padding of SegmentModel::crop, speakerDiarizer.cpp:1641-1662).
// Shared memory: first[f] = first sample of frame f (f = 0..F, first[F] = L) and off[f] = destination of the run of
// frame f, so the per-sample work has no division: a thread takes 8 consecutive samples (two 16-byte loads when the
// source is aligned), looks its frame up once and only compares against the next boundary.
__global__ void __launch_bounds__(512)
    mask_compact_kernel(const float* __restrict__ wav, const long* __restrict__ wav_base, long item_stride,
                        long wav_limit, const float* __restrict__ masks, int L, int F, float* __restrict__ signals,
                        float* __restrict__ counts) {
    extern __shared__ int sm[];
    int* first = sm;          // [F + 1]
    int* off = sm + (F + 1);  // [F]: destination offset of frame f's run, -1 when the frame is masked out
    __shared__ int warp_tot[16];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* m = masks + (size_t)b * F;
    for (int f = tid; f <= F; f += 512) first[f] = f < F ? frame_first_sample(f, F, L) : L;
    // exclusive prefix of the run lengths of the active frames
    int carry = 0;
    for (int f0 = 0; f0 < F; f0 += 512) {
        const int f = f0 + tid;
        const bool on = f < F && m[f] > 0.5f;
        const int n = on ? frame_first_sample(f + 1, F, L) - frame_first_sample(f, F, L) : 0;
        int inc = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        int before = carry;
        for (int w2 = 0; w2 < warp; ++w2) before += warp_tot[w2];
        if (f < F) off[f] = on ? before + inc - n : -1;
        int tot = 0;
        for (int w2 = 0; w2 < 16; ++w2) tot += warp_tot[w2];
        carry += tot;
        __syncthreads();
    }
    const int count = carry;
    const long base = wav_base ? wav_base[b] : (long)b * item_stride;
    const float* src = wav + base;
    float* out = signals + (size_t)b * L;
    // A warp takes 256 consecutive samples per step: lane l handles samples l, l + 32, ... so that both the loads and
    // the (shifted) stores of one instruction cover whole 128-byte lines.  The frame of a sample is estimated in
    // fp32 and corrected against the exact boundaries in shared memory.
    const float scale = (float)F / (float)L;
    for (int j0 = warp * 256; j0 < L; j0 += 16 * 256) {
        // frame of this lane's first sample (fp32 estimate corrected against the exact boundaries), then only
        // boundary comparisons (samples ascend with u); masked-out samples are never loaded
        const int jfirst = j0 + lane;
        int f = min(F - 1, (int)((float)jfirst * scale));
        if (jfirst < L) {
            while (jfirst < first[f]) --f;
            while (jfirst >= first[f + 1]) ++f;
        }
        int fb = first[f], nextb = first[f + 1], o = off[f];
        int dst[8];
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + 32 * u + lane;
            dst[u] = -1;
            v[u] = 0.f;
            if (j < L) {
                if (j >= nextb) {
                    do ++f; while (j >= first[f + 1]);
                    fb = first[f];
                    nextb = first[f + 1];
                    o = off[f];
                }
                if (o >= 0) {
                    dst[u] = o + (j - fb);
                    if (base + j < wav_limit) v[u] = src[j];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (dst[u] >= 0) out[dst[u]] = v[u];
    }
    for (int j = count + tid; j < L; j += 512) out[j] = 0.f;  // padSequence's zero tail
    if (tid == 0) counts[b] = (float)count;
}

// wav_lens normalisation per batch of `batch` items (getEmbedding is called per batch of 32, speakerDiarizer.cpp:
// 3083-3105): lens /= max_len of the batch, too-short items get 1.0; a batch whose longest item is too short is
// flagged (the reference then returns NaN embeddings without running the model).
__global__ void wav_lens_kernel(const float* __restrict__ counts, int R, int batch, int min_num_samples,
                                float* __restrict__ wav_lens, unsigned char* __restrict__ too_short,
                                unsigned char* __restrict__ batch_invalid) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int ngroups = (R + batch - 1) / batch;
    if (g >= ngroups) return;
    const int lo = g * batch, hi = min(R, lo + batch);
    float max_len = 0.f;
    for (int i = lo; i < hi; ++i) max_len = fmaxf(max_len, counts[i]);
    const bool invalid = max_len < (float)min_num_samples;
    if (batch_invalid) batch_invalid[g] = invalid ? 1 : 0;
    for (int i = lo; i < hi; ++i) {
        const float c = counts[i];
        if (invalid) {  // reference returns before normalising: lens stay raw counts
            wav_lens[i] = c;
            too_short[i] = 0;
        } else if (c < (float)min_num_samples) {
            wav_lens[i] = 1.0f;
            too_short[i] = 1;
        } else {
            wav_lens[i] = c / max_len;
            too_short[i] = 0;
        }
    }
}

// used mask per (chunk, speaker): the clean mask (frames where at most one speaker is active) when it keeps more
// than min_num_frames frames, else the raw one (speakerDiarizer.cpp:3056-3078).  One warp per (chunk, speaker);
// out[(c*K + k)][F] fp32.
__global__ void __launch_bounds__(256)
    select_masks_kernel(const double* __restrict__ binarized, int C, int F, int K, double min_num_frames,
                        float* __restrict__ out) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= C * K) return;
    const int c = row / K, k = row - c * K;
    const double* b = binarized + (size_t)c * F * K;
    float sum = 0.f;  // the reference sums the clean mask in float
    for (int f = lane; f < F; f += 32) {
        double s = 0.0;
        for (int q = 0; q < K; ++q) s += b[(size_t)f * K + q];
        if (s < 2.0) sum += (float)b[(size_t)f * K + k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const bool use_clean = (double)sum > min_num_frames;
    for (int f = lane; f < F; f += 32) {
        double v = b[(size_t)f * K + k];
        if (use_clean) {
            double s = 0.0;
            for (int q = 0; q < K; ++q) s += b[(size_t)f * K + q];
            if (!(s < 2.0)) v = 0.0;
        }
        out[(size_t)row * F + f] = (float)v;
    }
}

int mask_compact_launch(sd_ctx* ctx, const float* d_wav, const long* d_wav_base, long item_stride, long wav_limit,
                        const float* d_masks, int R, int L, int F, int batch, int min_num_samples, float* d_signals,
                        float* d_wav_lens, unsigned char* d_too_short, unsigned char* d_batch_invalid) {
    float* d_counts = (float*)ctx->scratch(BUF_GENERIC_A, sizeof(float) * (size_t)R);
    if (!d_counts) return SD_ERR_NOMEM;
    mask_compact_kernel<<<R, 512, sizeof(int) * (size_t)(2 * F + 2), ctx->stream>>>(d_wav, d_wav_base, item_stride, wav_limit,
                                                                                d_masks, L, F, d_signals, d_counts);
    SD_LAUNCH_CHECK(ctx);
    const int ngroups = (R + batch - 1) / batch;
    wav_lens_kernel<<<(ngroups + 127) / 128, 128, 0, ctx->stream>>>(d_counts, R, batch, min_num_samples, d_wav_lens,
                                                                    d_too_short, d_batch_invalid);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

int select_masks_launch(sd_ctx* ctx, const double* d_binarized, int C, int F, int K, double min_num_frames,
                        float* d_out) {
    const long rows = (long)C * K;
    select_masks_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, ctx->stream>>>(d_binarized, C, F, K, min_num_frames,
                                                                                      d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

}  // namespace sdb
