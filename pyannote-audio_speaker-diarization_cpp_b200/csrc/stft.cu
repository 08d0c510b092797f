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
    static constexpr int kSigHops = (kSigFloats + kHop - 1) / kHop;      // hop-sized segments, each padded
    static constexpr int kSigPadded = kSigFloats + kPad * kSigHops + 4;  // staged floats
    static constexpr size_t kSigBytes = (size_t)((kSigPadded + 3) / 4 * 4) * sizeof(float);
    static constexpr size_t kZBytes = (size_t)GROUPS * kZStride * sizeof(float2);
    // the upper-spectrum buffer shares storage with the staged samples (dead after phase 1)
    static constexpr size_t kSmemBytes = (kSigBytes > kZBytes ? kSigBytes : kZBytes) +
                                         (size_t)GROUPS * kGroupStride * sizeof(float2) + kNfft * sizeof(float) +
                                         kTwTableUnits * sizeof(float2);
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

// Stage the samples of an edge tile (item b, frames [t0, t0 + kTileFrames)) into the padded buffer: sample j of
// the tile (item index t0*hop - n_fft/2 + j) goes to sig[j + kPad * (j / kHop)].  Out-of-range samples are the
// zeros of torch::stft's centre padding (pad_mode "constant").  16-byte pieces never straddle a hop boundary.
template <int GROUPS>
__device__ __forceinline__ void stage_tile(float* sig, const float* __restrict__ wav, int L, long tile,
                                           int tiles_per_item, bool aligned16) {
    using Cfg = StftCfg<GROUPS>;
    const int b = (int)(tile / tiles_per_item);
    const int ti = (int)(tile - (long)b * tiles_per_item);
    const long s0 = (long)ti * Cfg::kTileFrames * kHop - kNfft / 2;
    const float* src = wav + (size_t)b * L;
    if (aligned16) {
        for (int c = threadIdx.x; c < Cfg::kSigFloats / 4; c += Cfg::kThreads) {
            const int j = 4 * c;
            const long s = s0 + j;
            float* dst = sig + j + kPad * (j / kHop);
            if (s >= 0 && s + 3 < L)
                cp_async16(dst, src + s);
            else {
                float4 v;
                v.x = (s >= 0 && s < L) ? src[s] : 0.f;
                v.y = (s + 1 >= 0 && s + 1 < L) ? src[s + 1] : 0.f;
                v.z = (s + 2 >= 0 && s + 2 < L) ? src[s + 2] : 0.f;
                v.w = (s + 3 >= 0 && s + 3 < L) ? src[s + 3] : 0.f;
                *reinterpret_cast<float4*>(dst) = v;
            }
        }
    } else {
        for (int j = threadIdx.x; j < Cfg::kSigFloats; j += Cfg::kThreads) {
            const long s = s0 + j;
            float* dst = sig + j + kPad * (j / kHop);
            if (s >= 0 && s < L)
                cp_async4(dst, src + s);
            else
                dst[0] = 0.f;
        }
    }
}

// ---- mbarrier / bulk-copy (TMA) helpers ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(a),
        "r"(parity));
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <int GROUPS, int MINB>
__global__ void __launch_bounds__(GROUPS* kRadix, MINB)
    stft400_kernel(const float* __restrict__ wav, float* __restrict__ out, int L, int T, int tiles_per_item,
                   long total_tiles, const float* __restrict__ window, const float2* __restrict__ twiddle,
                   int aligned16) {
    using Cfg = StftCfg<GROUPS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* xbuf = reinterpret_cast<float2*>(smem_raw);                 // phase 1 -> phase 2 transpose
    float2* twT = xbuf + GROUPS * kGroupStride;                         // twiddle table (see tw_thread_offset)
    float* wtab = reinterpret_cast<float*>(twT + kTwTableUnits);        // window, pre-scaled by 1/2
    float* sig = wtab + kNfft;                                          // padded samples of the tile
    float2* zup = reinterpret_cast<float2*>(sig);                       // upper half of the spectrum (aliases sig)
    __shared__ __align__(8) uint64_t bar;

    const int g = threadIdx.x / kRadix;
    const int r = threadIdx.x - g * kRadix;
    for (int i = threadIdx.x; i < kNfft; i += Cfg::kThreads)
        wtab[i] = 0.5f * window[i];  // exact scaling; lets phase 3 drop its multiplications
    for (int e = threadIdx.x; e < kTwTableUnits; e += Cfg::kThreads) twT[e] = twiddle[tw_table_source(e)];
    const float2* twp = twT + tw_thread_offset(threadIdx.x);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    __syncthreads();

    // A tile is "interior" when all its samples exist: it is then fetched by one thread with bulk (TMA) copies,
    // one per hop-sized segment (the padded layout), completing on an mbarrier.
    auto interior = [&](long tile) -> bool {
        const int ti = (int)(tile % tiles_per_item);
        const long s0 = (long)ti * Cfg::kTileFrames * kHop - kNfft / 2;
        return aligned16 && s0 >= 0 && s0 + Cfg::kSigFloats <= L;
    };
    auto issue_bulk = [&](long tile) {  // thread 0 only
        const int b = (int)(tile / tiles_per_item);
        const int ti = (int)(tile - (long)b * tiles_per_item);
        const long s0 = (long)ti * Cfg::kTileFrames * kHop - kNfft / 2;
        const float* src = wav + (size_t)b * L + s0;
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // earlier generic reads of sig vs async writes
        mbar_expect_tx(&bar, Cfg::kSigFloats * sizeof(float));
#pragma unroll 1
        for (int j = 0; j < Cfg::kSigFloats; j += kHop) {
            const int n = (Cfg::kSigFloats - j) < kHop ? (Cfg::kSigFloats - j) : kHop;
            bulk_g2s(sig + j + kPad * (j / kHop), src + j, n * sizeof(float), &bar);
        }
    };

    unsigned parity = 0;
    long tile = blockIdx.x;
    bool fetched = false;
    if (tile < total_tiles && interior(tile)) {
        if (threadIdx.x == 0) issue_bulk(tile);
        fetched = true;
    }
    for (; tile < total_tiles; tile += gridDim.x) {
        if (fetched) {
            mbar_wait(&bar, parity);
            parity ^= 1;
        } else {
            stage_tile<GROUPS>(sig, wav, L, tile, tiles_per_item, aligned16 != 0);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
        }
        stft_phase1_tab(sig, (2 * g) * kHopP, (2 * g + 1) * kHopP, wtab, twp, g, r, xbuf);
        __syncthreads();  // sig is free again, the transpose is complete
        float2 v[20];
        stft_phase2_load(xbuf, g, r, v);
        stft_publish_upper(v, g, r, zup);
        __syncthreads();  // upper halves are visible; xbuf may be overwritten by the next tile's phase 1

        const int b = (int)(tile / tiles_per_item);
        const int ti = (int)(tile - (long)b * tiles_per_item);
        const int tA = ti * Cfg::kTileFrames + 2 * g;
        float* rowA = out + ((size_t)b * T + tA) * (kBins * 2);
        if (tA + 1 < T)
            stft_split_store<true, true>(v, zup, g, r, rowA, rowA + kBins * 2);
        else if (tA < T)
            stft_split_store<true, false>(v, zup, g, r, rowA, nullptr);
        __syncthreads();  // zup (which aliases sig) is consumed: the next tile's samples may land
        const long next = tile + gridDim.x;
        fetched = next < total_tiles && interior(next);
        if (fetched && threadIdx.x == 0) issue_bulk(next);
    }
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

template <int GROUPS, int MINB>
static int launch_cfg(sd_ctx* ctx, const float* d_wav, int B, int L, int T, float* d_out) {
    using Cfg = StftCfg<GROUPS>;
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        SD_CUDA(ctx, cudaFuncSetAttribute(stft400_kernel<GROUPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)Cfg::kSmemBytes));
        SD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, stft400_kernel<GROUPS, MINB>,
                                                                   Cfg::kThreads, Cfg::kSmemBytes));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const int tiles_per_item = (T + Cfg::kTileFrames - 1) / Cfg::kTileFrames;
    const long total = (long)B * tiles_per_item;
    long grid = (long)ctx->num_sms * blocks_per_sm;
    if (grid > total) grid = total;
    const int aligned = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_wav) & 15) == 0);
    stft400_kernel<GROUPS, MINB><<<(unsigned)grid, Cfg::kThreads, Cfg::kSmemBytes, ctx->stream>>>(
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
    rc = ctx->stft_variant == 1 ? launch_cfg<8, 3>(ctx, d_wav, B, L, T, d_out) : launch_cfg<8, 4>(ctx, d_wav, B, L, T, d_out);
    if (rc) return rc;
    if (p->pad_batch_to > B) {  // _infer: rows beyond the real batch are zeros (speakerDiarizer.cpp:1904)
        size_t row = (size_t)T * kBins * 2 * sizeof(float);
        SD_CUDA(ctx, cudaMemsetAsync(reinterpret_cast<char*>(d_out) + (size_t)B * row, 0,
                                     (size_t)(p->pad_batch_to - B) * row, ctx->stream));
    }
    return SD_OK;
}

}  // namespace sdb
