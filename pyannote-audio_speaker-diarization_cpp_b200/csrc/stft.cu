// STFT front-end of the embedding stage.
// Replaces EmbeddingModel1::infer (speakerDiarizer.cpp:1977-2036): fp32 batch -> centre-padded frames ->
// periodic Hamming window -> 400-point real DFT -> [B][T][201][2] fp32, plus the batch padding of
// EmbeddingModel1::_infer (speakerDiarizer.cpp:1889-1917).
//
// HBM-bound by design: 4*L bytes read + 4*T*402 bytes written per item (1 125 608 B at L=80 000).
// Persistent CTAs loop over (item, 16*GROUPS/8-frame tile) work units; each tile's samples are brought
// into shared memory with 16-byte cp.async (double-buffered against the FFT of the previous tile); a group
// of 20 threads transforms two real frames at once as one 400-point complex FFT (20 x 20 Cooley-Tukey,
// each 20-point DFT a twiddle-free 4 x 5 prime-factor transform held in registers), exchanging through
// padded, conflict-free shared-memory transposes; the two one-sided spectra are separated and stored with
// 8-byte stores that cover whole 1608-byte output rows.
#include "common.cuh"
#include "fft400.cuh"

#include <cmath>

namespace sdb {

template <int GROUPS>
struct StftCfg {
    static constexpr int kThreads = GROUPS * kRadix;
    static constexpr int kTileFrames = GROUPS * 2;
    static constexpr int kSigFloats = (kTileFrames - 1) * kHop + kNfft;  // samples covered by one tile
    static constexpr int kSigChunks = kSigFloats / 4;                    // 16-byte chunks (kSigFloats % 4 == 0)
    static constexpr size_t kSmemBytes =
        2 * kSigFloats * sizeof(float) + 2 * (size_t)GROUPS * kGroupStride * sizeof(float2);
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Stage the samples of tile `tile` (item b, frames [t0, t0 + kTileFrames)) into sig[0 .. kSigFloats).
// Sample index of sig[j] inside the item: t0*hop - n_fft/2 + j; out-of-range samples are the zeros of
// torch::stft's centre padding (pad_mode "constant").
template <int GROUPS>
__device__ __forceinline__ void stage_tile(float* sig, const float* __restrict__ wav, int L, long tile,
                                           int tiles_per_item, bool aligned16) {
    using Cfg = StftCfg<GROUPS>;
    const int b = (int)(tile / tiles_per_item);
    const int ti = (int)(tile - (long)b * tiles_per_item);
    const long s0 = (long)ti * Cfg::kTileFrames * kHop - kNfft / 2;
    const float* src = wav + (size_t)b * L;
    if (aligned16) {
        for (int c = threadIdx.x; c < Cfg::kSigChunks; c += Cfg::kThreads) {
            const long s = s0 + 4L * c;
            if (s >= 0 && s + 3 < L)
                cp_async16(sig + 4 * c, src + s);
            else {
                float4 v;
                v.x = (s >= 0 && s < L) ? src[s] : 0.f;
                v.y = (s + 1 >= 0 && s + 1 < L) ? src[s + 1] : 0.f;
                v.z = (s + 2 >= 0 && s + 2 < L) ? src[s + 2] : 0.f;
                v.w = (s + 3 >= 0 && s + 3 < L) ? src[s + 3] : 0.f;
                *reinterpret_cast<float4*>(sig + 4 * c) = v;
            }
        }
    } else {
        for (int j = threadIdx.x; j < Cfg::kSigFloats; j += Cfg::kThreads) {
            const long s = s0 + j;
            if (s >= 0 && s < L)
                cp_async4(sig + j, src + s);
            else
                sig[j] = 0.f;
        }
    }
}

template <int GROUPS>
__global__ void __launch_bounds__(GROUPS* kRadix)
    stft400_kernel(const float* __restrict__ wav, float* __restrict__ out, int L, int T, int tiles_per_item,
                   long total_tiles, const float* __restrict__ window, const float2* __restrict__ twiddle,
                   int aligned16) {
    using Cfg = StftCfg<GROUPS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sig0 = reinterpret_cast<float*>(smem_raw);
    float* sig1 = sig0 + Cfg::kSigFloats;
    float2* xchg = reinterpret_cast<float2*>(sig1 + Cfg::kSigFloats);
    float2* zbuf = xchg + GROUPS * kGroupStride;

    const int g = threadIdx.x / kRadix;
    const int r = threadIdx.x - g * kRadix;

    // per-thread constants: this thread always plays role r
    float win[20];
    float2 tw[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) {
        win[i] = window[20 * i + r];
        tw[i] = twiddle[r * 20 + i];
    }

    long tile = blockIdx.x;
    if (tile >= total_tiles) return;
    stage_tile<GROUPS>(sig0, wav, L, tile, tiles_per_item, aligned16 != 0);
    cp_async_commit();
    int cur = 0;
    for (; tile < total_tiles; tile += gridDim.x) {
        float* sig = cur ? sig1 : sig0;
        const long next = tile + gridDim.x;
        if (next < total_tiles) stage_tile<GROUPS>(cur ? sig0 : sig1, wav, L, next, tiles_per_item, aligned16 != 0);
        cp_async_commit();
        cp_async_wait<1>();  // everything but the group just committed has landed
        __syncthreads();

        stft_phase1(sig, (2 * g) * kHop, (2 * g + 1) * kHop, win, tw, g, r, xchg);
        __syncthreads();
        stft_phase2(xchg, g, r, zbuf);
        __syncthreads();

        const int b = (int)(tile / tiles_per_item);
        const int ti = (int)(tile - (long)b * tiles_per_item);
        const int tA = ti * Cfg::kTileFrames + 2 * g;
        float* rowA = out + ((size_t)b * T + tA) * (kBins * 2);
        stft_phase3(zbuf, g, r, tA < T ? rowA : nullptr, tA + 1 < T ? rowA + kBins * 2 : nullptr);
        cur ^= 1;
        // No barrier here: the next iteration's first __syncthreads already orders this iteration's reads of
        // zbuf (phase 3), xchg (phase 2) and sig (phase 1) against their next writes.
    }
    cp_async_wait<0>();
}

static void make_window(int kind, const float* custom, std::vector<float>& w) {
    w.resize(kNfft);
    if (kind == SD_WINDOW_CUSTOM) {
        for (int i = 0; i < kNfft; ++i) w[i] = custom[i];
    } else if (kind == SD_WINDOW_POVEY) {
        // Kaldi "povey": pow(0.5 - 0.5 cos(2 pi n / (N-1)), 0.85)
        for (int i = 0; i < kNfft; ++i)
            w[i] = (float)std::pow(0.5 - 0.5 * std::cos(2.0 * M_PI * i / (double)(kNfft - 1)), 0.85);
    } else {
        // at::hamming_window(400) (periodic, fp32): every step rounded to float (speakerDiarizer.cpp:2007)
        const float scale = (float)(M_PI * 2.0 / (double)kNfft);
        for (int i = 0; i < kNfft; ++i) {
            float a = (float)i * scale;
            float c = std::cos(a);
            float m = c * (float)(-0.46);
            w[i] = m + (float)0.54;
        }
    }
}

static int ensure_tables(sd_ctx* ctx, const sd_stft_params* p) {
    if (!ctx->d_twiddle) {
        std::vector<float> tw(20 * 20 * 2);
        for (int r = 0; r < 20; ++r)
            for (int k = 0; k < 20; ++k) {
                double a = -2.0 * M_PI * (double)(r * k) / (double)kNfft;
                tw[(r * 20 + k) * 2] = (float)std::cos(a);
                tw[(r * 20 + k) * 2 + 1] = (float)std::sin(a);
            }
        SD_CUDA(ctx, cudaMalloc(&ctx->d_twiddle, tw.size() * sizeof(float)));
        SD_CUDA(ctx, cudaMemcpy(ctx->d_twiddle, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    std::vector<float> w;
    make_window(p->window_kind, p->window, w);
    if (!ctx->d_window || ctx->window_kind != p->window_kind || p->window_kind == SD_WINDOW_CUSTOM ||
        ctx->h_window != w) {
        if (!ctx->d_window) SD_CUDA(ctx, cudaMalloc(&ctx->d_window, kNfft * sizeof(float)));
        SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SD_CUDA(ctx, cudaMemcpy(ctx->d_window, w.data(), kNfft * sizeof(float), cudaMemcpyHostToDevice));
        ctx->h_window = w;
        ctx->window_kind = p->window_kind;
    }
    return SD_OK;
}

template <int GROUPS>
static int launch_cfg(sd_ctx* ctx, const float* d_wav, int B, int L, int T, float* d_out) {
    using Cfg = StftCfg<GROUPS>;
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        SD_CUDA(ctx, cudaFuncSetAttribute(stft400_kernel<GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)Cfg::kSmemBytes));
        SD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, stft400_kernel<GROUPS>,
                                                                   Cfg::kThreads, Cfg::kSmemBytes));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const int tiles_per_item = (T + Cfg::kTileFrames - 1) / Cfg::kTileFrames;
    const long total = (long)B * tiles_per_item;
    long grid = (long)ctx->num_sms * blocks_per_sm;
    if (grid > total) grid = total;
    const int aligned = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_wav) & 15) == 0);
    stft400_kernel<GROUPS><<<(unsigned)grid, Cfg::kThreads, Cfg::kSmemBytes, ctx->stream>>>(
        d_wav, d_out, L, T, tiles_per_item, total, ctx->d_window, reinterpret_cast<const float2*>(ctx->d_twiddle),
        aligned);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

int stft_launch(sd_ctx* ctx, const float* d_wav, int B, int L, const sd_stft_params* p, float* d_out) {
    if (p->n_fft != kNfft || p->hop != kHop)
        return ctx->fail(SD_ERR_UNSUPPORTED, "sd_stft: only n_fft=400 / hop=160 has a kernel (got %d / %d)", p->n_fft,
                         p->hop);
    if (p->preemph != 0.f) return ctx->fail(SD_ERR_UNSUPPORTED, "sd_stft: pre-emphasis is not implemented yet");
    if (p->window_kind == SD_WINDOW_CUSTOM && !p->window)
        return ctx->fail(SD_ERR_INVALID, "sd_stft: SD_WINDOW_CUSTOM needs a window pointer");
    int rc = ensure_tables(ctx, p);
    if (rc) return rc;
    const int T = 1 + L / kHop;
    rc = launch_cfg<8>(ctx, d_wav, B, L, T, d_out);
    if (rc) return rc;
    if (p->pad_batch_to > B) {  // _infer: rows beyond the real batch are zeros (speakerDiarizer.cpp:1904)
        size_t row = (size_t)T * kBins * 2 * sizeof(float);
        SD_CUDA(ctx, cudaMemsetAsync(reinterpret_cast<char*>(d_out) + (size_t)B * row, 0,
                                     (size_t)(p->pad_batch_to - B) * row, ctx->stream));
    }
    return SD_OK;
}

}  // namespace sdb
