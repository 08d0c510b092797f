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
#include "mel_table.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace sdb {

template <int GROUPS>
struct StftCfg {
    static constexpr int kThreads = GROUPS * kRadix;
    static constexpr int kTileFrames = GROUPS * 2;
    static constexpr int kSigFloats = (kTileFrames - 1) * kHop + kNfft;  // samples covered by one tile
    static constexpr int kSigPadded = sig_padded_size(kSigFloats) + 4;   // staged floats (hop segments padded 8 / 12)
    static constexpr size_t kSigBytes = (size_t)((kSigPadded + 3) / 4 * 4) * sizeof(float);
    static constexpr size_t kZBytes = (size_t)GROUPS * kZStride * sizeof(float2);
    static constexpr size_t kTablesBytes = (size_t)GROUPS * kGroupStride * sizeof(float2) + kNfft * sizeof(wtab_t) +
                                           kTwTableUnits * sizeof(float2);
    // fbank kernel: the upper-spectrum buffer shares storage with the staged samples (dead after phase 1)
    static constexpr size_t kSmemBytes = (kSigBytes > kZBytes ? kSigBytes : kZBytes) + kTablesBytes;
    // STFT kernel: nothing aliases the sample buffer (the next tile's samples land while phase 2 runs); the variant
    // that computes the Hamming window in registers has no window table: 41.1 KB, five CTAs per SM fit
    static constexpr size_t kSmemBytesStft = kSigBytes + kTablesBytes;
    static constexpr size_t kSmemBytesStftNoWindow = kSmemBytesStft - kNfft * sizeof(wtab_t);
    static constexpr size_t kSmemBytesStftNoWindow2 = kSmemBytesStftNoWindow + kSigBytes;  // two sample buffers
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

// Where frame t starts (sample t*hop - lead) and what lies beyond the ends of the item:
//   lead 200, zeros      torch::stft(center=true, pad_mode="constant") -- the reference (speakerDiarizer.cpp:2008)
//   lead 120, reflection Kaldi snip_edges=false: frame midpoints at t*hop + hop/2, the signal mirrored about its ends
//                        (x[-1-s] / x[2L-1-s]; torchaudio.compliance.kaldi._get_strided)
//   lead 0               Kaldi snip_edges=true: only whole frames (nothing lies outside)
struct FrameGeom {
    int lead;
    int reflect;
};
__device__ __forceinline__ float edge_sample(const float* __restrict__ src, long s, int L, int reflect) {
    if (reflect) s = s < 0 ? -s - 1 : (s >= L ? 2L * L - 1 - s : s);
    return (s >= 0 && s < L) ? src[s] : 0.f;
}

// Stage the samples of an edge tile (item b, frames [t0, t0 + kTileFrames)) into the padded buffer: sample j of
// the tile (item index t0*hop - lead + j) goes to sig[sig_pos(j)].  Out-of-range samples are zeros or the
// mirrored signal (FrameGeom).  16-byte pieces never straddle a hop boundary.
template <int GROUPS>
__device__ __forceinline__ void stage_tile(float* sig, const float* __restrict__ wav, int L, int b, int ti,
                                           bool aligned16, FrameGeom fg) {
    using Cfg = StftCfg<GROUPS>;
    const long s0 = (long)ti * Cfg::kTileFrames * kHop - fg.lead;
    const float* src = wav + (size_t)b * L;
    if (aligned16) {
        for (int c = threadIdx.x; c < Cfg::kSigFloats / 4; c += Cfg::kThreads) {
            const int j = 4 * c;
            const long s = s0 + j;
            float* dst = sig + sig_pos(j);
            if (s >= 0 && s + 3 < L)
                cp_async16(dst, src + s);
            else {
                float4 v;
                v.x = edge_sample(src, s, L, fg.reflect);
                v.y = edge_sample(src, s + 1, L, fg.reflect);
                v.z = edge_sample(src, s + 2, L, fg.reflect);
                v.w = edge_sample(src, s + 3, L, fg.reflect);
                *reinterpret_cast<float4*>(dst) = v;
            }
        }
    } else {
        for (int j = threadIdx.x; j < Cfg::kSigFloats; j += Cfg::kThreads) {
            const long s = s0 + j;
            float* dst = sig + sig_pos(j);
            if (s >= 0 && s < L)
                cp_async4(dst, src + s);
            else
                dst[0] = edge_sample(src, s, L, fg.reflect);
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
        "WAIT_LOOP_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_LOOP_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// Kaldi-mode arguments: pre-emphasis coefficient and whether the frame mean is removed first
struct KaldiArgs {
    float preemph;
    int remove_dc;
};

// (1 - c) * mean of the two frames of pair g, for every thread of the group (block-wide: contains a barrier)
template <int GROUPS>
__device__ __forceinline__ float2 frame_dc_offsets(const float* sig, int g, int r, float2* part, KaldiArgs ka) {
    if (!ka.remove_dc) return make_float2(0.f, 0.f);
    part[threadIdx.x] = stft_frame_partial_sums(sig, sig_frame_off(2 * g), sig_frame_off(2 * g + 1), r);
    __syncthreads();
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int k = 0; k < kRadix; ++k) {
        const float2 q = part[g * kRadix + k];
        sa += q.x;
        sb += q.y;
    }
    const float f = (1.f - ka.preemph) / (float)kNfft;
    return make_float2(sa * f, sb * f);
}

// KALDI: per-frame pre-emphasis / DC removal (table window).  HAMMING: the reference's periodic Hamming window computed
// in registers instead of read from the shared table (only without KALDI).
// registers per thread that let MINB CTAs of GROUPS * 20 threads share an SM (64 K registers, allocated per warp in
// units of 8 per thread); given to ptxas as __maxnreg__ because with a minimum-blocks launch bound it stops at 72 for
// five CTAs although 80 fit
constexpr int stft_max_regs(int groups, int minb) {
    const int warps = (groups * kRadix + 31) / 32;
    const int r = 65536 / (minb * warps * 32) / 8 * 8;
    return r > 255 ? 255 : r;
}

// DEPTH: sample buffers.  1: the next tile's bulk load is issued once phase 1 has released the buffer and has phase 2 +
// stores to land; 2: two buffers, the load of tile i + 1 is issued before phase 1 of tile i and has the whole tile.
template <int GROUPS, int MINB, bool KALDI, bool HAMMING, int DEPTH>
__global__ void __launch_bounds__(GROUPS* kRadix) __maxnreg__(stft_max_regs(GROUPS, MINB))
    stft400_kernel(const float* __restrict__ wav, float* __restrict__ out, int L, int T, int tiles_per_item,
                   long total_tiles, const float* __restrict__ window, const float2* __restrict__ twiddle,
                   int aligned16, FrameGeom fg, KaldiArgs ka, unsigned long long* __restrict__ tile_ctr,
                   float* __restrict__ dump) {
    using Cfg = StftCfg<GROUPS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* xbuf = reinterpret_cast<float2*>(smem_raw);                 // phase 1 -> phase 2 transpose
    float2* twT = xbuf + GROUPS * kGroupStride;                         // twiddle table (see tw_thread_offset)
    wtab_t* wtab = reinterpret_cast<wtab_t*>(twT + kTwTableUnits);      // window, pre-scaled by 1/2 (not with HAMMING)
    float* sig0 = reinterpret_cast<float*>(wtab + (HAMMING ? 0 : kNfft));  // padded samples of the tile (x DEPTH)
    constexpr int kSigStride = (int)(Cfg::kSigBytes / sizeof(float));
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int s_plan_b, s_plan_ti, s_plan_fetched;  // warp 0's copy of the next tile (DEPTH 2: known a phase earlier)
    // the next tile, worked out by thread 0 only (64-bit division and the interior test cost ~100 instructions, which
    // every thread used to spend per tile): item, tile of the item, >= 0 when there is one; bulk-fetched or not
    __shared__ int s_next_b, s_next_ti, s_next_fetched;
    __shared__ float2 dc_part[KALDI ? GROUPS * kRadix : 1];

    // Tiles are handed out dynamically: the first one is the CTA's own index, every further one comes from a global
    // counter.  A static stride would assume that every CTA of the grid is resident at once; when other kernels hold
    // SM resources (the merge loops of other files of a batch: one register-heavy CTA on most SMs), some CTAs only
    // start when others exit and the kernel takes two waves -- measured 1.55 ms instead of 0.93 ms per launch under
    // eight concurrent merge loops.  With the counter a late CTA just finds less work.  Thread 0 fetches the index
    // one tile ahead of its use, so the round trip of the atomic is off the critical path.
    unsigned long long nxt_raw = 0;  // counter value; the tile index is gridDim.x + this (added at the point of use, so
                                     // that nothing waits for the atomic's round trip before the next iteration)
    if (threadIdx.x == 0) nxt_raw = atomicAdd(tile_ctr, 1ull);
    const int g = threadIdx.x / kRadix;
    const int r = threadIdx.x - g * kRadix;
    const PairShuffle pair_lane(r);
    float ham_c, ham_s;  // (cos, sin)(2 pi r / 400) for the table-free Hamming window
    sincospif((float)r * (1.0f / 200.0f), &ham_s, &ham_c);
    if (!HAMMING)
        for (int i = threadIdx.x; i < kNfft; i += Cfg::kThreads)
            wtab[i] = wtab_make(0.5f * window[i]);  // exact scaling; lets phase 3 drop its multiplications
    for (int e = threadIdx.x; e < kTwTableUnits; e += Cfg::kThreads) twT[e] = twiddle[tw_table_source(e)];
    const float2* twp = twT + tw_thread_offset(threadIdx.x);
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    __syncthreads();

    // A tile is "interior" when all its samples exist: it is then fetched by one thread with bulk (TMA) copies,
    // one per hop-sized segment (the padded layout), completing on an mbarrier.
    auto interior = [&](int ti) -> bool {
        const long s0 = (long)ti * Cfg::kTileFrames * kHop - fg.lead;
        return aligned16 && s0 >= 0 && s0 + Cfg::kSigFloats <= L;
    };
    // Bulk (TMA) copies of a tile: lane 0 of warp 0 announces the bytes on the mbarrier, then lane j issues the copy
    // of segment pair j (nine copies; the hardware takes them one lane at a time, see fft400.cuh) instead of a loop in
    // thread 0, whose warp every other warp of the CTA would wait for at the end-of-tile barrier.  Call with all
    // lanes of warp 0.
    auto issue_bulk = [&](int b, int ti, int k) {  // into sample buffer k, completing on bar[k]
        const int lane = (int)threadIdx.x;
        const long s0 = (long)ti * Cfg::kTileFrames * kHop - fg.lead;
        const float* src = wav + (size_t)b * L + s0;
        float* sig = sig0 + k * kSigStride;
        if (lane == 0) mbar_expect_tx(&bar[k], Cfg::kSigFloats * sizeof(float));
        __syncwarp();
        const int j = lane * 2 * kHop;  // a pair of hop segments is contiguous in the staged layout (fft400.cuh)
        if (j < Cfg::kSigFloats) {
            const int n = (Cfg::kSigFloats - j) < 2 * kHop ? (Cfg::kSigFloats - j) : 2 * kHop;
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // earlier generic reads of sig vs async writes
            bulk_g2s(sig + sig_pos(j), src + j, n * sizeof(float), &bar[k]);
        }
    };
    static_assert(kPadEven == 0, "an even hop segment and the next one must be contiguous in shared memory");
    static_assert((Cfg::kSigFloats + 2 * kHop - 1) / (2 * kHop) <= 32, "one segment pair per lane of warp 0");
    // warp 0: work out the next tile (thread 0: counter value fetched one tile ago, 32-bit division, interior test),
    // start its bulk load into buffer k, and ask the counter for the tile after it
    auto plan_and_issue = [&](int k) -> bool {
        if (threadIdx.x == 0) {
            int nb = -1, nti = 0;
            bool w = false;
            const unsigned long long nxt = (unsigned long long)gridDim.x + nxt_raw;
            if (nxt < (unsigned long long)total_tiles) {  // total_tiles < 2^31 (checked at launch)
                nb = (int)((unsigned)nxt / (unsigned)tiles_per_item);
                nti = (int)((unsigned)nxt - (unsigned)nb * (unsigned)tiles_per_item);
                w = interior(nti);
            }
            s_plan_b = nb;
            s_plan_ti = nti;
            s_plan_fetched = w;
            nxt_raw = atomicAdd(tile_ctr, 1ull);  // used one iteration from now
        }
        __syncwarp();  // lane 0's shared-memory writes are visible to its warp
        const bool w = s_plan_fetched != 0;
        if (w) issue_bulk(s_plan_b, s_plan_ti, k);
        return w;
    };

    // Only thread 0 polls the mbarrier (a polling warp burns issue slots: with every warp polling, 16 % of the
    // kernel's issued instructions were try_wait / branch / yield); the others learn about the arrival through the
    // block barrier that follows, which is the barrier the tile loop needs anyway.
    unsigned parity = 0;  // bit k: phase of bar[k]
    int cur = 0;          // sample buffer of the current tile
    int b = -1, ti = 0;   // current tile: item and tile of the item (b < 0: none left)
    bool fetched = false;
    if ((long)blockIdx.x < total_tiles) {
        b = (int)(blockIdx.x / (unsigned)tiles_per_item);
        ti = (int)(blockIdx.x - (unsigned)b * (unsigned)tiles_per_item);
        if (interior(ti)) {
            if (threadIdx.x < 32) {
                issue_bulk(b, ti, 0);
                if (threadIdx.x == 0) mbar_wait(&bar[0], 0);
            }
            parity ^= 1;
            fetched = true;
            __syncthreads();
        }
    }
    while (b >= 0) {
        float* sig = sig0 + cur * kSigStride;
        bool wait_next = false;  // warp 0: a bulk load is in flight
        // DEPTH 2: the other buffer was released by the previous tile's phase 1 (two barriers ago)
        if (DEPTH == 2 && threadIdx.x < 32) wait_next = plan_and_issue(cur ^ 1);
        if (!fetched) {
            stage_tile<GROUPS>(sig, wav, L, b, ti, aligned16 != 0, fg);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
        }
        if (KALDI) {
            const float2 dc = frame_dc_offsets<GROUPS>(sig, g, r, dc_part, ka);
            stft_phase1_kaldi(sig, sig_frame_off(2 * g), sig_frame_off(2 * g + 1), wtab, twp, g, r, xbuf, ka.preemph, dc);
        } else if (HAMMING) {
            stft_phase1_hamming(sig, sig_frame_off(2 * g), sig_frame_off(2 * g + 1), ham_c, ham_s, twp, g, r, xbuf);
        } else {
            stft_phase1_tab(sig, sig_frame_off(2 * g), sig_frame_off(2 * g + 1), wtab, twp, g, r, xbuf);
        }
        __syncthreads();  // the transpose is complete and sig is free (DEPTH 1: the next tile's samples may land
                          // already, under the shadow of phase 2 and the stores)
        if (DEPTH == 1 && threadIdx.x < 32) wait_next = plan_and_issue(0);
        if (threadIdx.x == 0) {  // for everybody, read after the barrier at the end of the tile
            s_next_b = s_plan_b;
            s_next_ti = s_plan_ti;
            s_next_fetched = s_plan_fetched;
        }
        float2 v[20];
        stft_phase2_load(xbuf, g, r, v);  // slot r plays role pair_role(r) from here on

        // The real-pair split: Z[400 - k] lives in the adjacent lane (fft400.cuh), so the second exchange is a
        // shuffle -- no shared-memory round trip and no block barrier around it (two barriers per tile instead of four).
        const int tA = ti * Cfg::kTileFrames + 2 * g;
        float* rowA = tA < T ? out + ((size_t)b * T + tA) * (kBins * 2) : dump;  // frames past the item's end: scratch row
        float* rowB = tA + 1 < T ? out + ((size_t)b * T + tA + 1) * (kBins * 2) : dump;
        stft_split_store_pair(v, r, rowA, rowB, pair_lane);
        const int nk = DEPTH == 2 ? cur ^ 1 : 0;  // buffer (and mbarrier) of the next tile
        if (wait_next && threadIdx.x == 0) mbar_wait(&bar[nk], (parity >> nk) & 1u);  // its samples have landed
        __syncthreads();  // every phase-2 load of the transpose buffer is done: the next tile's phase 1 may overwrite
                          // it; the next tile's coordinates and its landed samples are visible to everybody
        b = s_next_b;  // (rewritten only after the next tile's first barrier)
        ti = s_next_ti;
        fetched = s_next_fetched != 0;
        if (fetched) parity ^= 1u << nk;
        cur = nk;
    }
    // the last CTA to leave zeroes the counters for the next launch on this context (launches of a context are
    // stream-ordered; contexts do not share counters)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(tile_ctr + 1, 1ull) == (unsigned long long)gridDim.x - 1) {
            tile_ctr[0] = 0;
            tile_ctr[1] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Fused log-mel front-end: same tile pipeline as stft400_kernel, but the one-sided spectra never
// leave the SM.  Epilogue per tile: |X|^2 -> shared memory, triangular mel projection (each bin feeds at most
// two filters, so the matrix is kept as per-filter bin ranges), 10 log10(max(., amin)), coalesced [T][n_mels]
// stores, and the utterance maximum (needed by the top_db clamp) through one atomicMax per warp.
// Spec: embeddings/threeModel.py:212-221 (spectral_magnitude -> Filterbank(n_mels=80) -> MyNormalization);
// parity unpinned (speechbrain 0.5.14 is not vendored) -- checked against oracle/sd_oracle.c only.

__device__ __forceinline__ int float_order_key(float v) {  // monotone float -> int map for atomicMax
    const int i = __float_as_int(v);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float float_from_order_key(int k) {
    const int i = k >= 0 ? k : k ^ 0x7fffffff;
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f;
    memcpy(&f, &i, 4);
    return f;
#endif
}

template <bool PARTS>
struct MelSmem {
    typedef MelTable type;
};
template <>
struct MelSmem<true> {
    typedef MelParts type;
};

// PARTS: the mel projection by (frame, 20-bin part) threads (mel_table.h) instead of one filter per thread
template <int GROUPS, int MINB, bool KALDI, bool HAMMING, bool PARTS>
__global__ void __launch_bounds__(GROUPS* kRadix, MINB)
    fbank400_kernel(const float* __restrict__ wav, float* __restrict__ out, int L, int T, int tiles_per_item,
                    long total_tiles, const float* __restrict__ window, const float2* __restrict__ twiddle,
                    int aligned16, const void* __restrict__ mel, int n_mels, float amin, float log_scale,
                    int* __restrict__ item_max, FrameGeom fg, KaldiArgs ka) {
    using Cfg = StftCfg<GROUPS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* xbuf = reinterpret_cast<float2*>(smem_raw);
    float2* twT = xbuf + GROUPS * kGroupStride;
    wtab_t* wtab = reinterpret_cast<wtab_t*>(twT + kTwTableUnits);
    float* sig = reinterpret_cast<float*>(wtab + kNfft);
    float* pw = reinterpret_cast<float*>(xbuf);  // power spectra [frame][201], aliases the transpose buffer
    __shared__ __align__(8) uint64_t bar;
    typedef typename MelSmem<PARTS>::type MelT;
    __shared__ MelT smel;
    __shared__ float2 dc_part[KALDI ? GROUPS * kRadix : 1];

    const int g = threadIdx.x / kRadix;
    const int r = threadIdx.x - g * kRadix;
    float ham_c, ham_s;  // (cos, sin)(2 pi r / 400) for the table-free Hamming window (stft400_kernel)
    sincospif((float)r * (1.0f / 200.0f), &ham_s, &ham_c);
    for (int i = threadIdx.x; i < kNfft; i += Cfg::kThreads) wtab[i] = wtab_make(0.5f * window[i]);
    for (int e = threadIdx.x; e < kTwTableUnits; e += Cfg::kThreads) twT[e] = twiddle[tw_table_source(e)];
    for (int i = threadIdx.x; i < (int)(sizeof(MelT) / 4); i += Cfg::kThreads)
        reinterpret_cast<int*>(&smel)[i] = reinterpret_cast<const int*>(mel)[i];
    const float2* twp = twT + tw_thread_offset(threadIdx.x);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    __syncthreads();

    auto interior = [&](long tile) -> bool {
        const int ti = (int)(tile % tiles_per_item);
        const long s0 = (long)ti * Cfg::kTileFrames * kHop - fg.lead;
        return aligned16 && s0 >= 0 && s0 + Cfg::kSigFloats <= L;
    };
    auto issue_bulk = [&](long tile) {  // all lanes of warp 0: lane j copies hop segment j (see stft400_kernel)
        const int lane = (int)threadIdx.x;
        const int b = (int)(tile / tiles_per_item);
        const int ti = (int)(tile - (long)b * tiles_per_item);
        const long s0 = (long)ti * Cfg::kTileFrames * kHop - fg.lead;
        const float* src = wav + (size_t)b * L + s0;
        if (lane == 0) mbar_expect_tx(&bar, Cfg::kSigFloats * sizeof(float));
        __syncwarp();
        const int j = lane * 2 * kHop;  // one copy per pair of hop segments (contiguous in the staged layout)
        if (j < Cfg::kSigFloats) {
            const int n = (Cfg::kSigFloats - j) < 2 * kHop ? (Cfg::kSigFloats - j) : 2 * kHop;
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            bulk_g2s(sig + sig_pos(j), src + j, n * sizeof(float), &bar);
        }
    };

    // One thread polls the mbarrier, the others learn about the arrival through the block barrier that the tile loop
    // needs anyway (as in stft400_kernel: with every warp polling, a sixth of the issued instructions were
    // try_wait / branch).
    unsigned parity = 0;
    long tile = blockIdx.x;
    bool fetched = false;
    if (tile < total_tiles && interior(tile)) {
        if (threadIdx.x < 32) {
            issue_bulk(tile);
            if (threadIdx.x == 0) mbar_wait(&bar, parity);
        }
        parity ^= 1;
        fetched = true;
        __syncthreads();
    }
    for (; tile < total_tiles; tile += gridDim.x) {
        if (!fetched) {
            stage_tile<GROUPS>(sig, wav, L, (int)(tile / tiles_per_item), (int)(tile % tiles_per_item), aligned16 != 0, fg);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
        }
        if (KALDI) {
            const float2 dc = frame_dc_offsets<GROUPS>(sig, g, r, dc_part, ka);
            stft_phase1_kaldi(sig, sig_frame_off(2 * g), sig_frame_off(2 * g + 1), wtab, twp, g, r, xbuf, ka.preemph, dc);
        } else if (HAMMING) {
            stft_phase1_hamming(sig, sig_frame_off(2 * g), sig_frame_off(2 * g + 1), ham_c, ham_s, twp, g, r, xbuf);
        } else {
            stft_phase1_tab(sig, sig_frame_off(2 * g), sig_frame_off(2 * g + 1), wtab, twp, g, r, xbuf);
        }
        __syncthreads();  // the transpose is complete and sig is free: the next tile's samples may land from here on
        const long next = tile + gridDim.x;
        fetched = next < total_tiles && interior(next);
        if (fetched && threadIdx.x < 32) issue_bulk(next);
        float2 v[20];
        stft_phase2_load(xbuf, g, r, v);  // slot r plays role pair_role(r) from here on
        __syncthreads();  // xbuf is dead from here: it receives the power spectra

        // |X|^2 of both frames of the pair into pw[frame][bin]; Z[400 - k] comes from the adjacent lane (fft400.cuh)
        stft_split_power_pair(v, r, pw + (2 * g) * kBins, pw + (2 * g + 1) * kBins, PairShuffle(r));
        __syncthreads();  // power spectra complete

        // mel projection + dB.  Thread -> one mel filter and one half of the tile's frames: the filter's weights are
        // read once and reused for every frame.  A warp takes as long as its widest filter (1 bin at the bottom of the
        // mel scale, 13 at the top), so with 160 threads and 80 filters warp w gets filters 16 (4 - w) .. + 15 for both
        // halves: similar widths share a warp (issue slots 240 instead of 320 per tile, the widest warp starts first);
        // 16 consecutive filters are still one 64-byte run of a row.
        const int b = (int)(tile / tiles_per_item);
        const int ti = (int)(tile - (long)b * tiles_per_item);
        const int t0 = ti * Cfg::kTileFrames;
        float vmax = -INFINITY;
        constexpr int kFramesPerPass = Cfg::kThreads / 80 > 0 ? Cfg::kThreads / 80 : 1;  // frame slots per pass
        if constexpr (PARTS) {
            // Thread = (frame f, 20-bin part q): the part's bins are read once (lanes of a half-warp: 16 frames, stride
            // 201 = 9 mod 32 banks; the other half-warp works 80 bins = 16 banks further on), every table entry is a
            // broadcast, and the two accumulators are flushed to the filters' columns of mo[f][.] as the filters change
            // (mel_table.h).  40 multiply-adds per thread whatever the part -- the per-filter form has 8 to 104.
            static_assert(!PARTS || Cfg::kThreads == 16 * kMelParts, "one thread per (frame, part)");
            float* mo = pw + Cfg::kTileFrames * kBins;  // [frame][n_mels + extra columns], odd row stride
            const int RS = n_mels + kMelExtraCols + 1;
            {
                const int f = (int)(threadIdx.x & 15), h = (int)(threadIdx.x >> 4);
                const int q = (h >> 1) < 4 ? (h >> 1) + 4 * (h & 1) : 8 + (h & 1);
                const MelEntry* tab = smel.e[q];
                const float* prow = pw + f * kBins + q * kMelPartBins;
                float* orow = mo + f * RS;
                float accE = 0.f, accO = 0.f;
#pragma unroll
                for (int i = 0; i < kMelEntries - 1; ++i) {
                    const int4 en = *reinterpret_cast<const int4*>(tab + i);  // (wE, wO, fE, fO): one 16-byte broadcast
                    if (en.z >= 0) {
                        orow[en.z] = accE;
                        accE = 0.f;
                    }
                    if (en.w >= 0) {
                        orow[en.w] = accO;
                        accO = 0.f;
                    }
                    const float p = prow[i];  // the 21st entry of parts 0..8 has zero weights; the bin exists
                    accE = fmaf(__int_as_float(en.x), p, accE);
                    accO = fmaf(__int_as_float(en.y), p, accO);
                }
                const int4 en = *reinterpret_cast<const int4*>(tab + kMelEntries - 1);
                if (en.z >= 0) orow[en.z] = accE;
                if (en.w >= 0) orow[en.w] = accO;
            }
            __syncthreads();  // every partial sum is in mo
            for (int m0 = 0; m0 < n_mels; m0 += 80) {
                const int m = m0 + (int)(threadIdx.x % 80), slot = threadIdx.x / 80;
                if (m < n_mels && slot < kFramesPerPass) {
                    const int has = smel.has[m], ex = smel.extra[m];
#pragma unroll
                    for (int j = 0; j < Cfg::kTileFrames / kFramesPerPass; ++j) {
                        const int f = slot + j * kFramesPerPass;
                        float a = has ? mo[f * RS + m] : 0.f;
                        if (ex >= 0) a += mo[f * RS + n_mels + ex];
                        if (t0 + f < T) {
                            const float db = log_scale * __log2f(fmaxf(a, amin));
                            out[((size_t)b * T + t0 + f) * n_mels + m] = db;
                            vmax = fmaxf(vmax, db);
                        }
                    }
                }
            }
        } else {
        for (int m0 = 0; m0 < n_mels; m0 += 80) {
            int m = m0 + (int)(threadIdx.x % 80), slot = threadIdx.x / 80;
            if (Cfg::kThreads == 160) {
                m = m0 + 16 * (4 - (int)(threadIdx.x >> 5)) + (int)(threadIdx.x & 15);
                slot = (threadIdx.x >> 4) & 1;
            }
            if (m < n_mels && slot < kFramesPerPass) {
                const int lo = smel.lo[m], cnt = smel.cnt[m];
                const float* wq = smel.w + smel.off[m];
                float acc[Cfg::kTileFrames / kFramesPerPass];
#pragma unroll
                for (int j = 0; j < Cfg::kTileFrames / kFramesPerPass; ++j) acc[j] = 0.f;
                for (int i = 0; i < cnt; ++i) {
                    const float wv = wq[i];
                    const float* p = pw + lo + i;
#pragma unroll
                    for (int j = 0; j < Cfg::kTileFrames / kFramesPerPass; ++j)
                        acc[j] = fmaf(p[(slot + j * kFramesPerPass) * kBins], wv, acc[j]);
                }
#pragma unroll
                for (int j = 0; j < Cfg::kTileFrames / kFramesPerPass; ++j) {
                    const int f = slot + j * kFramesPerPass;
                    if (t0 + f < T) {
                        // 10 log10(x) = (10 / log2(10)) log2(x) (log_scale = 3.0103; ln 2 for Kaldi's natural log);
                        // MUFU.LG2 is accurate to ~1e-6 dB here
                        const float db = log_scale * __log2f(fmaxf(acc[j], amin));
                        out[((size_t)b * T + t0 + f) * n_mels + m] = db;
                        vmax = fmaxf(vmax, db);
                    }
                }
            }
        }
        }  // per-filter projection
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        if ((threadIdx.x & 31) == 0 && vmax > -INFINITY) atomicMax(item_max + b, float_order_key(vmax));
        if (fetched) {
            if (threadIdx.x == 0) mbar_wait(&bar, parity);  // the next tile's samples have landed
            parity ^= 1;
        }
        __syncthreads();  // pw consumed before the next tile's phase 1 writes xbuf
    }
}

// top_db clamp and sentence-level mean normalisation (MyNormalization, threeModel.py:333-369): one CTA per item,
// thread -> (mel, frame slot) so that every access is coalesced and column sums need no atomics.
// Sweep 1 accumulates the clamped column sums of the first round(len * T) frames; sweep 2 re-reads the block (an L2
// hit: it is <= 320 KB and was just touched), clamps, subtracts the mean and writes in place -- one DRAM read and one
// DRAM write per element.
constexpr int FN_THREADS = 960;
__global__ void __launch_bounds__(FN_THREADS)
    fbank_norm_kernel(float* __restrict__ feats, int T, int n_mels, const int* __restrict__ item_max,
                      const float* __restrict__ wav_lens, float top_db, int mean_norm) {
    extern __shared__ double colsum[];  // [slots][n_mels] partial sums, then [n_mels] means
    const int b = blockIdx.x;
    float* x = feats + (size_t)b * T * n_mels;
    const float floor_db = float_from_order_key(item_max[b]) - top_db;
    long n = lrintf(wav_lens[b] * (float)T);  // torch.round: half to even
    if (n > T) n = T;
    const int slots = FN_THREADS / n_mels;  // frames handled per pass
    const int m = threadIdx.x % n_mels, slot = threadIdx.x / n_mels;
    const bool active = slot < slots;
    if (mean_norm) {
        double acc = 0.0;
        if (active) {
            int f = slot;
            for (; f + 7 * slots < n; f += 8 * slots) {  // eight independent loads in flight
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = x[(size_t)(f + u * slots) * n_mels + m];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += (double)fmaxf(v[u], floor_db);
            }
            for (; f < n; f += slots) acc += (double)fmaxf(x[(size_t)f * n_mels + m], floor_db);
            colsum[slot * n_mels + m] = acc;
        }
        __syncthreads();
        if (threadIdx.x < n_mels) {
            double s2 = 0.0;
            for (int q = 0; q < slots; ++q) s2 += colsum[q * n_mels + threadIdx.x];
            colsum[threadIdx.x] = n > 0 ? s2 / (double)n : (double)NAN;
        }
        __syncthreads();
    }
    if (active) {
        const float mean = mean_norm ? (float)colsum[m] : 0.f;
        int f = slot;
        for (; f + 7 * slots < T; f += 8 * slots) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = x[(size_t)(f + u * slots) * n_mels + m];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[(size_t)(f + u * slots) * n_mels + m] = fmaxf(v[u], floor_db) - mean;
        }
        for (; f < T; f += slots) {
            const size_t e = (size_t)f * n_mels + m;
            x[e] = fmaxf(x[e], floor_db) - mean;
        }
    }
}

__global__ void fill_int_kernel2(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
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

// frame geometry of a parameter set; returns the number of frames of an L-sample item
static int frame_geometry(const sd_stft_params* p, int L, FrameGeom* fg) {
    switch (p->frame_mode) {
        case SD_FRAMES_KALDI_REFLECT:
            *fg = FrameGeom{kNfft / 2 - kHop / 2, 1};
            return (L + kHop / 2) / kHop;
        case SD_FRAMES_KALDI_SNIP:
            *fg = FrameGeom{0, 0};
            return L < kNfft ? 0 : 1 + (L - kNfft) / kHop;
        default:
            *fg = FrameGeom{kNfft / 2, 0};
            return 1 + L / kHop;
    }
}
static bool kaldi_conditioning(const sd_stft_params* p) { return p->preemph != 0.f || p->remove_dc_offset != 0; }

template <int GROUPS, int MINB, bool KALDI, bool HAMMING, int DEPTH = 1>
static int launch_cfg(sd_ctx* ctx, const float* d_wav, int B, int L, int T, float* d_out, FrameGeom fg, KaldiArgs ka) {
    using Cfg = StftCfg<GROUPS>;
    static_assert(DEPTH == 1 || HAMMING, "two sample buffers only in the register-window build");
    const size_t smem = DEPTH == 2 ? Cfg::kSmemBytesStftNoWindow2 : HAMMING ? Cfg::kSmemBytesStftNoWindow : Cfg::kSmemBytesStft;
    const int blocks_per_sm = kernel_setup(ctx, stft400_kernel<GROUPS, MINB, KALDI, HAMMING, DEPTH>, (int)smem, Cfg::kThreads, smem);
    if (blocks_per_sm < 0) return SD_ERR_CUDA;
    const int tiles_per_item = (T + Cfg::kTileFrames - 1) / Cfg::kTileFrames;
    const long total = (long)B * tiles_per_item;
    if (total >= (1l << 31)) return ctx->fail(SD_ERR_INVALID, "sd_stft: %ld tiles in one call (limit 2^31)", total);
    // stft_waves > 1: that many CTAs per resident slot, each looping over proportionally fewer tiles, so that slots
    // are handed back every 1 / waves of the kernel instead of at its end (other streams' small kernels get in)
    long grid = (long)ctx->num_sms * blocks_per_sm * (ctx->stft_waves > 1 ? ctx->stft_waves : 1);
    if (grid > total) grid = total;
    const int aligned = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_wav) & 15) == 0);
    stft400_kernel<GROUPS, MINB, KALDI, HAMMING, DEPTH><<<(unsigned)grid, Cfg::kThreads, smem, ctx->stream>>>(
        d_wav, d_out, L, T, tiles_per_item, total, ctx->d_window, reinterpret_cast<const float2*>(ctx->d_twiddle),
        aligned, fg, ka, ctx->d_stats + 8, reinterpret_cast<float*>(ctx->d_stats + 16));
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

int stft_launch(sd_ctx* ctx, const float* d_wav, int B, int L, const sd_stft_params* p, float* d_out) {
    if (p->n_fft != kNfft || p->hop != kHop)
        return ctx->fail(SD_ERR_UNSUPPORTED, "sd_stft: only n_fft=400 / hop=160 has a kernel (got %d / %d)", p->n_fft,
                         p->hop);
    if (p->window_kind == SD_WINDOW_CUSTOM && !p->window)
        return ctx->fail(SD_ERR_INVALID, "sd_stft: SD_WINDOW_CUSTOM needs a window pointer");
    if (p->frame_mode < 0 || p->frame_mode > SD_FRAMES_KALDI_SNIP)
        return ctx->fail(SD_ERR_INVALID, "sd_stft: unknown frame_mode %d", p->frame_mode);
    int rc = ensure_tables(ctx, p);
    if (rc) return rc;
    FrameGeom fg;
    const int T = frame_geometry(p, L, &fg);
    if (T < 1) return ctx->fail(SD_ERR_INVALID, "sd_stft: %d samples give no frame in this frame_mode", L);
    if (fg.reflect && L < kNfft) return ctx->fail(SD_ERR_INVALID, "sd_stft: reflection needs at least n_fft samples");
    const KaldiArgs ka{p->preemph, p->remove_dc_offset};
    // stft_variant (tuning hook; 1 773 items x 160 000 samples, profiles/r02_stft_v12.log): 0 = the reference's Hamming
    // window computed in registers, four CTAs per SM at 96 registers, one sample buffer (0.747 ms); 3 = five CTAs per SM
    // (80 registers, 8 bytes of spills: 0.767 ms); 6 = two sample buffers, the next tile's load issued a whole tile
    // ahead (0.805 ms: warp 0 issues it before its own phase 1 and everybody waits for that warp); 5 = 8-frame tiles in
    // 80-thread CTAs (0.786 ms); 1 = window table, 3 CTAs/SM; 2 = window table, 4 CTAs/SM
    const bool hamming = p->window_kind == SD_WINDOW_HAMMING_PERIODIC && ctx->stft_variant != 2 && ctx->stft_variant != 1;
    if (kaldi_conditioning(p))
        rc = launch_cfg<8, 3, true, false>(ctx, d_wav, B, L, T, d_out, fg, ka);
    else if (ctx->stft_variant == 1)
        rc = launch_cfg<8, 3, false, false>(ctx, d_wav, B, L, T, d_out, fg, ka);
    else if (hamming && ctx->stft_variant == 6)
        rc = launch_cfg<8, 4, false, true, 2>(ctx, d_wav, B, L, T, d_out, fg, ka);
    else if (hamming && ctx->stft_variant == 3)
        rc = launch_cfg<8, 5, false, true>(ctx, d_wav, B, L, T, d_out, fg, ka);
    else if (hamming && ctx->stft_variant == 5)
        rc = launch_cfg<4, 9, false, true>(ctx, d_wav, B, L, T, d_out, fg, ka);
    else if (hamming)
        rc = launch_cfg<8, 4, false, true>(ctx, d_wav, B, L, T, d_out, fg, ka);
    else
        rc = launch_cfg<8, 4, false, false>(ctx, d_wav, B, L, T, d_out, fg, ka);
    if (rc) return rc;
    if (p->pad_batch_to > B) {  // _infer: rows beyond the real batch are zeros (speakerDiarizer.cpp:1904)
        size_t row = (size_t)T * kBins * 2 * sizeof(float);
        SD_CUDA(ctx, cudaMemsetAsync(reinterpret_cast<char*>(d_out) + (size_t)B * row, 0,
                                     (size_t)(p->pad_batch_to - B) * row, ctx->stream));
    }
    return SD_OK;
}

int fbank_launch(sd_ctx* ctx, const float* d_wav, int B, int L, const float* d_lens, const sd_fbank_params* p,
                 float* d_out) {
    const sd_stft_params* sp = &p->stft;
    if (sp->n_fft != kNfft || sp->hop != kHop)
        return ctx->fail(SD_ERR_UNSUPPORTED, "sd_fbank: only n_fft=400 / hop=160 has a kernel");
    if (sp->frame_mode < 0 || sp->frame_mode > SD_FRAMES_KALDI_SNIP)
        return ctx->fail(SD_ERR_INVALID, "sd_fbank: unknown frame_mode %d", sp->frame_mode);
    int rc = ensure_tables(ctx, sp);
    if (rc) return rc;
    // SD_OPT_STFT_VARIANT 8 (tuning): the mel projection by (frame, 20-bin part) threads (mel_table.h).  Conflict-free
    // (7 M instead of 18 M bank-conflict wavefronts per 600 items) but not faster: its 16-byte table broadcasts cost as
    // many wavefronts as the conflicts they remove and it needs one more block barrier -- 1.477 ms against 1.319 ms
    // for the per-filter form, which therefore stays the default.
    const bool want_parts = ctx->stft_variant == 8;
    const int key = p->n_mels * 1000003 + (int)p->f_min * 7919 + (int)p->f_max + p->sample_rate * 31 + p->mel_kind * 104729 +
                    (want_parts ? 15485863 : 0);
    if (!ctx->d_mel || ctx->mel_key != key) {
        MelTable t;
        std::memset(&t, 0, sizeof(t));
        rc = p->mel_kind == 1 ? build_mel_table_kaldi(p, t) : build_mel_table(p, t);
        if (rc) return ctx->fail(rc, "sd_fbank: unsupported mel configuration (n_mels=%d)", p->n_mels);
        static_assert(sizeof(MelParts) <= 2 * sizeof(MelTable), "d_mel holds either table");
        if (!ctx->d_mel) SD_CUDA(ctx, cudaMalloc(&ctx->d_mel, 2 * sizeof(MelTable)));
        SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        // the (frame, part) form of the projection only on request and when the filterbank has the structure it needs
        // (mel_table.h: true for the 80-filter speechbrain and Kaldi banks), else one filter per thread
        std::vector<MelParts> mp(1);
        ctx->mel_parts = want_parts && build_mel_parts(t, p->n_mels, mp[0]) == SD_OK && 16 * (p->n_mels + kMelExtraCols + 1) * sizeof(float) +
                                 16 * kBins * sizeof(float) <= 8 * kGroupStride * sizeof(float2);
        if (ctx->mel_parts)
            SD_CUDA(ctx, cudaMemcpy(ctx->d_mel, &mp[0], sizeof(MelParts), cudaMemcpyHostToDevice));
        else
            SD_CUDA(ctx, cudaMemcpy(ctx->d_mel, &t, sizeof(MelTable), cudaMemcpyHostToDevice));
        ctx->mel_key = key;
    }
    FrameGeom fg;
    const int T = frame_geometry(sp, L, &fg);
    if (T < 1) return ctx->fail(SD_ERR_INVALID, "sd_fbank: %d samples give no frame in this frame_mode", L);
    if (fg.reflect && L < kNfft) return ctx->fail(SD_ERR_INVALID, "sd_fbank: reflection needs at least n_fft samples");
    const KaldiArgs ka{sp->preemph, sp->remove_dc_offset};
    const bool kaldi = kaldi_conditioning(sp);
    const float log_scale = p->log_kind == 1 ? 0.69314718055994531f : 3.01029995663981195f;  // ln 2 | 10 / log2(10)
    const float top_db = p->log_kind == 1 ? INFINITY : p->top_db;
    int* d_max = (int*)ctx->scratch(BUF_FB_TMP, sizeof(int) * (size_t)B);
    if (!d_max) return SD_ERR_NOMEM;
    fill_int_kernel2<<<(B + 255) / 256, 256, 0, ctx->stream>>>(d_max, B, (int)0x80000000);  // below every key
    SD_LAUNCH_CHECK(ctx);
    using Cfg = StftCfg<8>;
    // the reference's window is computed in registers (two FFMA per sample instead of a table read) -- stft400_kernel
    const bool hamming = !kaldi && sp->window_kind == SD_WINDOW_HAMMING_PERIODIC && ctx->stft_variant != 2;
    auto launch = [&](auto kernel) -> int {
        const int blocks_per_sm = kernel_setup(ctx, kernel, (int)Cfg::kSmemBytes, Cfg::kThreads, Cfg::kSmemBytes);
        if (blocks_per_sm < 0) return SD_ERR_CUDA;
        const int tiles_per_item = (T + Cfg::kTileFrames - 1) / Cfg::kTileFrames;
        const long total = (long)B * tiles_per_item;
        long grid = (long)ctx->num_sms * blocks_per_sm;
        if (grid > total) grid = total;
        const int aligned = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_wav) & 15) == 0);
        kernel<<<(unsigned)grid, Cfg::kThreads, Cfg::kSmemBytes, ctx->stream>>>(
            d_wav, d_out, L, T, tiles_per_item, total, ctx->d_window, reinterpret_cast<const float2*>(ctx->d_twiddle),
            aligned, ctx->d_mel, p->n_mels, p->amin, log_scale, d_max, fg, ka);
        return SD_OK;
    };
    // 96 registers (4 CTAs per SM) without spills for the plain front-end; the Kaldi conditioning needs 128 (3 per SM)
    const bool parts = ctx->mel_parts != 0;  // decided when the table was uploaded (same key)
    if (parts)
        rc = kaldi ? launch(fbank400_kernel<8, 3, true, false, true>)
                   : hamming ? launch(fbank400_kernel<8, 4, false, true, true>)
                             : launch(fbank400_kernel<8, 4, false, false, true>);
    else
        rc = kaldi ? launch(fbank400_kernel<8, 3, true, false, false>)
                   : hamming ? launch(fbank400_kernel<8, 4, false, true, false>)
                             : launch(fbank400_kernel<8, 4, false, false, false>);
    if (rc) return rc;
    SD_LAUNCH_CHECK(ctx);
    const int total_e = T * p->n_mels;
    const size_t nsm = sizeof(double) * (size_t)FN_THREADS;
    (void)total_e;
    if (p->log_kind == 1 && !p->mean_norm) return SD_OK;  // Kaldi default: no clamp, no normalisation -> nothing to do
    fbank_norm_kernel<<<B, FN_THREADS, nsm, ctx->stream>>>(d_out, T, p->n_mels, d_max, d_lens, top_db, p->mean_norm);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

}  // namespace sdb
