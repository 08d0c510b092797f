// C-ABI of libsdb200.so (see include/sdb200.h).  Host-pointer entry points = H2D + kernels + D2H + sync;
// *_dev entry points enqueue on the context stream only.
#include "common.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace sdb {
int normalize_launch(sd_ctx* ctx, const double* d_x, int N, int D, double* d_xn);
int pdist_condensed_launch(sd_ctx* ctx, const double* d_x, int N, int D, int mode, double* d_cond);
int linkage_launch(sd_ctx* ctx, const double* d_x, int N, int D, double* d_Z, int mode);
int fcluster_launch(sd_ctx* ctx, const double* d_Z, int N, double cutoff, int* d_T, int* d_num);
int cosine_cdist_launch(sd_ctx* ctx, const double* d_a, int na, const double* d_b, int nb, int D, double* d_out);
int cluster_labels_launch(sd_ctx* ctx, const double* d_x, int N, int D, const sd_cluster_params* p, int* d_labels,
                          int* d_num, int k_cap);
int clustering_launch(sd_ctx* ctx, const double* d_emb, int C, int S, int D, const int* h_keep, int n_keep,
                      const sd_cluster_params* p, const double* d_binarized, int F, int* d_hard, double* d_soft,
                      int soft_k_cap, int* num_clusters_out, int* d_num_out, double* d_dist = nullptr);
int row_valid_launch(sd_ctx* ctx, const double* d_emb, int R, int D, unsigned char* d_valid);
int mask_compact_launch(sd_ctx* ctx, const float* d_wav, const long* d_wav_base, long item_stride, long wav_limit,
                        const float* d_masks, int R, int L, int F, int batch, int min_num_samples, float* d_signals,
                        float* d_wav_lens, unsigned char* d_too_short, unsigned char* d_batch_invalid);
int reconstruct_rows(int C, const sd_window* chunks, int64_t n_count, const sd_window* cf, int64_t* rows,
                     sd_window* frames_out);
int reconstruct_launch(sd_ctx* ctx, const float* d_seg, int C, int F, int K, const sd_window* chunks, const int* d_hard,
                       int Kc, const int* d_count, int64_t n_count, const sd_window* cf, double* d_out,
                       int64_t cap_elems, int64_t* rows_out, sd_window* frames_out);
int to_annotation_launch(sd_ctx* ctx, const double* d_scores, int64_t rows, int cols, const sd_window* frames,
                         double onset, double offset, double min_on, double min_off, double* d_seg, int* d_label,
                         int64_t cap, long* d_n);
int ingest_pcm16_launch(sd_ctx* ctx, const short* d_pcm, long n, float* d_out);
int crop_chunks_launch(sd_ctx* ctx, const float* d_wave, long n, const double* starts_s, int n_chunks, double duration,
                       int sample_rate, float* d_out);
int select_masks_launch(sd_ctx* ctx, const double* d_binarized, int C, int F, int K, double min_num_frames,
                        float* d_out);

// Small host -> device parameter uploads go through a ring of pinned slots so that they are truly
// asynchronous and the caller's buffer can be reused immediately.
constexpr int kRingSlots = 16;
constexpr size_t kRingSlotBytes = 256 * 1024;
struct Ring {
    char* base = nullptr;
    cudaEvent_t ev[kRingSlots] = {};
    int next = 0;
};
static Ring* ring_of(sd_ctx* ctx);

int upload_small(sd_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
    Ring* r = ring_of(ctx);
    if (!r || bytes > kRingSlotBytes) {  // large or no ring: plain (staged) copy
        SD_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        if (!r) SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return SD_OK;
    }
    const int s = r->next;
    r->next = (s + 1) % kRingSlots;
    SD_CUDA(ctx, cudaEventSynchronize(r->ev[s]));
    std::memcpy(r->base + (size_t)s * kRingSlotBytes, h_src, bytes);
    SD_CUDA(ctx, cudaMemcpyAsync(d_dst, r->base + (size_t)s * kRingSlotBytes, bytes, cudaMemcpyHostToDevice,
                                 ctx->stream));
    SD_CUDA(ctx, cudaEventRecord(r->ev[s], ctx->stream));
    return SD_OK;
}
}  // namespace sdb

namespace sdb {
// the one place a latched device status becomes a message
int status_message(sd_ctx* ctx, int st) {
    switch (st) {
        case SD_OK: return SD_OK;
        case SD_ERR_ZERO_MAGNITUDE: return ctx->fail(st, "Vectors have zero magnitude.");  // SD:494
        case SD_ERR_CAPACITY:
            return ctx->fail(st, "more raw clusters than the asynchronous clustering path sizes its buffers for (1024)");
        default: return ctx->fail(st, "device-side failure %d", st);
    }
}
}  // namespace sdb

using namespace sdb;

struct CtxExtra {
    Ring ring;
    // copy streams / events of the pipelined host-pointer STFT
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {}, ev_comp[2] = {}, ev_out[2] = {};
    bool pipe_ready = false;
};

// lives in the context itself (no process-wide registry: contexts are created and destroyed from any thread)
static CtxExtra* extra_of(sd_ctx* ctx) { return static_cast<CtxExtra*>(ctx->extra); }

namespace sdb {
static Ring* ring_of(sd_ctx* ctx) {
    CtxExtra* ex = extra_of(ctx);
    return ex && ex->ring.base ? &ex->ring : nullptr;
}
}  // namespace sdb

extern "C" {

int sd_version(void) { return SDB200_VERSION; }

int sd_ctx_create(int device, sd_ctx** out) {
    if (!out) return SD_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return SD_ERR_CUDA;  // no CPU fallback by design
    }
    if (cudaSetDevice(device) != cudaSuccess) return SD_ERR_CUDA;
    sd_ctx* ctx = new sd_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        ctx->num_sms = prop.multiProcessorCount;
        ctx->l2_bytes = (size_t)prop.l2CacheSize;
    }
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return SD_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    for (int i = 0; i < 16; ++i) {
        cudaEventCreate(&ctx->ev_start[i]);
        cudaEventCreate(&ctx->ev_stop[i]);
    }
    if (cudaMalloc(&ctx->d_status, sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->d_status, 0, sizeof(int)) != cudaSuccess ||
        cudaHostAlloc(&ctx->h_status, sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        // [8], [9]: STFT tile counters; [16..]: scratch row for the frames past the end of an item (stft400_kernel)
        cudaMalloc(&ctx->d_stats, (16 + 256) * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(ctx->d_stats, 0, (16 + 256) * sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError();
        sd_ctx_destroy(ctx);
        return SD_ERR_CUDA;
    }
    CtxExtra* ex = new CtxExtra();
    if (cudaHostAlloc(&ex->ring.base, kRingSlots * kRingSlotBytes, cudaHostAllocDefault) == cudaSuccess) {
        for (int i = 0; i < kRingSlots; ++i) cudaEventCreateWithFlags(&ex->ring.ev[i], cudaEventDisableTiming);
    } else {
        ex->ring.base = nullptr;
        cudaGetLastError();
    }
    ctx->extra = ex;
    *out = ctx;
    return SD_OK;
}

void sd_ctx_destroy(sd_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& b : ctx->bufs)
        if (b.p) cudaFree(b.p);
    for (auto& e : ctx->ev_lk)
        if (e) cudaEventDestroy(e);
    if (ctx->d_window) cudaFree(ctx->d_window);
    if (ctx->d_twiddle) cudaFree(ctx->d_twiddle);
    if (ctx->d_mel) cudaFree(ctx->d_mel);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    if (ctx->d_status) cudaFree(ctx->d_status);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    for (int i = 0; i < 16; ++i) {
        cudaEventDestroy(ctx->ev_start[i]);
        cudaEventDestroy(ctx->ev_stop[i]);
    }
    if (CtxExtra* ex = extra_of(ctx)) {
            if (ex->ring.base) {
                for (int k = 0; k < kRingSlots; ++k) cudaEventDestroy(ex->ring.ev[k]);
                cudaFreeHost(ex->ring.base);
            }
            if (ex->pipe_ready) {
                for (int k = 0; k < 2; ++k) {
                    cudaEventDestroy(ex->ev_in[k]);
                    cudaEventDestroy(ex->ev_comp[k]);
                    cudaEventDestroy(ex->ev_out[k]);
                }
                cudaStreamDestroy(ex->s_in);
                cudaStreamDestroy(ex->s_out);
            }
            delete ex;
            ctx->extra = nullptr;
        }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* sd_last_error(const sd_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sd_ctx_set_stream(sd_ctx* ctx, void* cuda_stream) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SD_OK;
}
void* sd_ctx_stream(sd_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int sd_sync(sd_ctx* ctx) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}
int sd_malloc(sd_ctx* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(SD_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return SD_OK;
}
int sd_free(sd_ctx* ctx, void* dptr) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_CUDA(ctx, cudaFree(dptr));
    return SD_OK;
}
int sd_host_alloc(sd_ctx* ctx, size_t bytes, void** hptr) {
    if (!ctx || !hptr) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(SD_ERR_NOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return SD_OK;
}
int sd_host_free(sd_ctx* ctx, void* hptr) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_CUDA(ctx, cudaFreeHost(hptr));
    return SD_OK;
}
int sd_memcpy_h2d(sd_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SD_OK;
}
int sd_memcpy_d2h(sd_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SD_OK;
}
int sd_memset(sd_ctx* ctx, void* dst, int value, size_t bytes) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_CUDA(ctx, cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return SD_OK;
}
int sd_timer_start(sd_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 16) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    SD_CUDA(ctx, cudaEventRecord(ctx->ev_start[slot], ctx->stream));
    return SD_OK;
}
int sd_timer_stop(sd_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 16) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    SD_CUDA(ctx, cudaEventRecord(ctx->ev_stop[slot], ctx->stream));
    return SD_OK;
}
int sd_timer_elapsed_ms(sd_ctx* ctx, int slot, float* ms) {
    if (!ctx || !ms || slot < 0 || slot >= 16) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    SD_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop[slot]));
    SD_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_start[slot], ctx->ev_stop[slot]));
    return SD_OK;
}
int64_t sd_launch_count(const sd_ctx* ctx) { return ctx ? ctx->launches : 0; }

int sd_ctx_set_option(sd_ctx* ctx, int option, int value) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    switch (option) {
        case SD_OPT_FORCE_EXACT_LINKAGE:
            ctx->force_exact_linkage = value != 0;
            return SD_OK;
        case SD_OPT_STFT_VARIANT:
            ctx->stft_variant = value;
            return SD_OK;
        case SD_OPT_LINKAGE_CLUSTER:
            ctx->linkage_cluster = value != 0;
            return SD_OK;
        case SD_OPT_STFT_WAVES:
            if (value < 1 || value > 64) return ctx->fail(SD_ERR_INVALID, "SD_OPT_STFT_WAVES must be 1..64");
            ctx->stft_waves = value;
            return SD_OK;
        case SD_OPT_LINKAGE_WIDE:
            if (value < 0 || value > 2) return ctx->fail(SD_ERR_INVALID, "SD_OPT_LINKAGE_WIDE must be 0, 1 or 2");
            ctx->linkage_wide = value;
            return SD_OK;
        case SD_OPT_LINKAGE_THREADS:
            if (value != 0 && value != 128 && value != 256 && value != 512 && value != 1024)
                return ctx->fail(SD_ERR_INVALID, "SD_OPT_LINKAGE_THREADS must be 0, 128, 256, 512 or 1024");
            ctx->linkage_threads = value;
            return SD_OK;
        default:
            return ctx->fail(SD_ERR_INVALID, "unknown option %d", option);
    }
}

int sd_debug_counters(sd_ctx* ctx, int64_t* out8, int reset) {
    if (!ctx || !out8) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SD_CUDA(ctx, cudaMemcpy(out8, ctx->d_stats, 8 * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (reset) SD_CUDA(ctx, cudaMemset(ctx->d_stats, 0, 8 * sizeof(int64_t)));
    return SD_OK;
}

int sd_flush_l2(sd_ctx* ctx) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    if (!ctx->flush_buf) {
        ctx->flush_bytes = std::max<size_t>(2 * ctx->l2_bytes, (size_t)256 << 20);
        SD_CUDA(ctx, cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
    }
    SD_CUDA(ctx, cudaMemsetAsync(ctx->flush_buf, 0, ctx->flush_bytes, ctx->stream));
    return SD_OK;
}

/* ---------------------------------------------------------------- STFT / fbank */

void sd_stft_default_params(sd_stft_params* p) {
    if (!p) return;
    p->n_fft = 400;
    p->hop = 160;
    p->window_kind = SD_WINDOW_HAMMING_PERIODIC;
    p->window = nullptr;
    p->preemph = 0.f;
    p->pad_batch_to = 0;
    p->frame_mode = SD_FRAMES_CENTER_ZERO;
    p->remove_dc_offset = 0;
}

void sd_stft_kaldi_params(sd_stft_params* p, int snip_edges) {
    if (!p) return;
    sd_stft_default_params(p);
    p->window_kind = SD_WINDOW_POVEY;
    p->preemph = 0.97f;
    p->remove_dc_offset = 1;
    p->frame_mode = snip_edges ? SD_FRAMES_KALDI_SNIP : SD_FRAMES_KALDI_REFLECT;
}

int64_t sd_stft_num_frames(int L, int hop) { return hop > 0 ? 1 + L / hop : 0; }

int64_t sd_stft_num_frames_mode(int L, int n_fft, int hop, int frame_mode) {
    if (hop <= 0 || n_fft <= 0) return 0;
    switch (frame_mode) {
        case SD_FRAMES_KALDI_REFLECT: return (L + hop / 2) / hop;
        case SD_FRAMES_KALDI_SNIP: return L < n_fft ? 0 : 1 + (L - n_fft) / hop;
        default: return 1 + L / hop;
    }
}

int sd_stft_dev(sd_ctx* ctx, const float* d_wav, int B, int L, const sd_stft_params* p, float* d_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_wav && d_out && p, "sd_stft_dev: null pointer");
    SD_REQUIRE(ctx, B > 0 && L > 0, "sd_stft_dev: B and L must be positive");
    return stft_launch(ctx, d_wav, B, L, p, d_out);
}

// Host-pointer STFT.  The batch is cut into chunks that flow through a three-stage pipeline -- H2D of chunk
// i+1, the kernel on chunk i and D2H of chunk i-1 run concurrently on three streams with double-buffered device
// chunks -- so a PCIe-bound call costs about max(H2D, D2H) instead of their sum.  (Pinned host buffers are needed
// for the copies to be truly asynchronous; pageable ones still work, just serialised by the driver.)
int sd_stft(sd_ctx* ctx, const float* wav, int B, int L, const sd_stft_params* p, float* out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, wav && out && p, "sd_stft: null pointer");
    SD_REQUIRE(ctx, B > 0 && L > 0, "sd_stft: B and L must be positive");
    const int64_t T = sd_stft_num_frames_mode(L, p->n_fft, p->hop, p->frame_mode);
    const size_t in_item = sizeof(float) * (size_t)L;
    const size_t out_item = sizeof(float) * (size_t)T * (p->n_fft / 2 + 1) * 2;
    CtxExtra* ex = extra_of(ctx);
    if (!ex) return ctx->fail(SD_ERR_INVALID, "sd_stft: unknown context");
    if (!ex->pipe_ready) {
        SD_CUDA(ctx, cudaStreamCreateWithFlags(&ex->s_in, cudaStreamNonBlocking));
        SD_CUDA(ctx, cudaStreamCreateWithFlags(&ex->s_out, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            SD_CUDA(ctx, cudaEventCreateWithFlags(&ex->ev_in[k], cudaEventDisableTiming));
            SD_CUDA(ctx, cudaEventCreateWithFlags(&ex->ev_comp[k], cudaEventDisableTiming));
            SD_CUDA(ctx, cudaEventCreateWithFlags(&ex->ev_out[k], cudaEventDisableTiming));
        }
        ex->pipe_ready = true;
    }
    // ~192 MB of output per chunk keeps every copy long enough to run at full PCIe rate
    int chunk = (int)std::max<size_t>(1, ((size_t)192 << 20) / out_item);
    if (chunk > B) chunk = B;
    float* d_in = (float*)ctx->scratch(BUF_STFT_IN, 2 * in_item * chunk);
    float* d_out = (float*)ctx->scratch(BUF_STFT_OUT, 2 * out_item * chunk);
    if (!d_in || !d_out) return SD_ERR_NOMEM;
    sd_stft_params q = *p;
    q.pad_batch_to = 0;
    // work queued on the context stream before this call must finish before the buffers are reused
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int idx = 0;
    for (int b0 = 0; b0 < B; b0 += chunk, ++idx) {
        const int nb = std::min(chunk, B - b0);
        const int slot = idx & 1;
        float* din = d_in + (size_t)slot * chunk * L;
        float* dout = d_out + (size_t)slot * chunk * (out_item / sizeof(float));
        if (idx >= 2) SD_CUDA(ctx, cudaStreamWaitEvent(ex->s_in, ex->ev_comp[slot], 0));  // input slot consumed
        SD_CUDA(ctx, cudaMemcpyAsync(din, wav + (size_t)b0 * L, in_item * nb, cudaMemcpyHostToDevice, ex->s_in));
        SD_CUDA(ctx, cudaEventRecord(ex->ev_in[slot], ex->s_in));
        SD_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ex->ev_in[slot], 0));
        if (idx >= 2) SD_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ex->ev_out[slot], 0));  // output slot drained
        int rc = stft_launch(ctx, din, nb, L, &q, dout);
        if (rc) return rc;
        SD_CUDA(ctx, cudaEventRecord(ex->ev_comp[slot], ctx->stream));
        SD_CUDA(ctx, cudaStreamWaitEvent(ex->s_out, ex->ev_comp[slot], 0));
        SD_CUDA(ctx, cudaMemcpyAsync(out + (size_t)b0 * (out_item / sizeof(float)), dout, out_item * nb,
                                     cudaMemcpyDeviceToHost, ex->s_out));
        SD_CUDA(ctx, cudaEventRecord(ex->ev_out[slot], ex->s_out));
    }
    if (p->pad_batch_to > B)  // _infer: rows beyond the real batch are zeros (speakerDiarizer.cpp:1904)
        std::memset(out + (size_t)B * (out_item / sizeof(float)), 0, out_item * (size_t)(p->pad_batch_to - B));
    SD_CUDA(ctx, cudaStreamSynchronize(ex->s_out));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

int sd_pack_wav_lens(const float* lens, int n_lens, int batch, float* out) {
    if (!out || batch <= 0 || n_lens < 0 || n_lens > batch || (n_lens && !lens)) return SD_ERR_INVALID;
    for (int i = 0; i < batch; ++i) out[i] = i < n_lens ? lens[i] : 1.0f;  // speakerDiarizer.cpp:1899-1900
    return SD_OK;
}

void sd_fbank_default_params(sd_fbank_params* p) {
    if (!p) return;
    sd_stft_default_params(&p->stft);
    p->n_mels = 80;
    p->f_min = 0.f;
    p->f_max = 8000.f;
    p->sample_rate = 16000;
    p->top_db = 80.f;
    p->amin = 1e-10f;
    p->mean_norm = 1;
    p->mel_kind = 0;
    p->log_kind = 0;
}

void sd_fbank_kaldi_params(sd_fbank_params* p, int snip_edges) {
    if (!p) return;
    sd_fbank_default_params(p);
    sd_stft_kaldi_params(&p->stft, snip_edges);
    p->f_min = 20.f;           /* kaldi::MelBanksOptions low_freq */
    p->f_max = 0.f;            /* high_freq 0 = Nyquist */
    p->amin = 1.1920929e-07f;  /* std::numeric_limits<float>::epsilon(), the floor before the log */
    p->mean_norm = 0;
    p->mel_kind = 1;
    p->log_kind = 1;
}

int sd_fbank_dev(sd_ctx* ctx, const float* d_wav, int B, int L, const float* d_wav_lens, const sd_fbank_params* p,
                 float* d_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_wav && d_out && p && d_wav_lens, "sd_fbank_dev: null pointer");
    SD_REQUIRE(ctx, B > 0 && L > 0, "sd_fbank_dev: B and L must be positive");
    return fbank_launch(ctx, d_wav, B, L, d_wav_lens, p, d_out);
}

int sd_fbank(sd_ctx* ctx, const float* wav, int B, int L, const float* wav_lens, const sd_fbank_params* p,
             float* out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, wav && out && p && wav_lens, "sd_fbank: null pointer");
    SD_REQUIRE(ctx, B > 0 && L > 0, "sd_fbank: B and L must be positive");
    const int64_t T = sd_stft_num_frames_mode(L, p->stft.n_fft, p->stft.hop, p->stft.frame_mode);
    const size_t in_bytes = sizeof(float) * (size_t)B * L;
    const size_t out_bytes = sizeof(float) * (size_t)B * T * p->n_mels;
    float* d_in = (float*)ctx->scratch(BUF_STFT_IN, in_bytes);
    float* d_out = (float*)ctx->scratch(BUF_FB_OUT, out_bytes);
    float* d_lens = (float*)ctx->scratch(BUF_FB_LENS, sizeof(float) * (size_t)B);
    if (!d_in || !d_out || !d_lens) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_in, wav, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(d_lens, wav_lens, sizeof(float) * (size_t)B, cudaMemcpyHostToDevice, ctx->stream));
    int rc = fbank_launch(ctx, d_in, B, L, d_lens, p, d_out);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

/* ---------------------------------------------------------------- aggregation */

int sd_np_rint(double v) { return np_rint_host(v); }

int64_t sd_closest_frame(const sd_window* w, double t) {
    return w ? closest_frame_host(w->start, w->step, w->duration, t) : -1;
}

int64_t sd_aggregate_num_frames(int C, const sd_window* chunks, const sd_window* frames) {
    if (!chunks || !frames || C <= 0) return -1;
    const double target = chunks->start + chunks->duration + (double)(size_t)(C - 1) * chunks->step;
    return closest_frame_host(chunks->start, frames->step, frames->duration, target) + 1;
}

static void fill_post(const sd_window* chunks, const sd_window* frames, sd_window* post) {
    if (!post) return;  // speakerDiarizer.cpp:1278-1281
    post->start = chunks->start;
    post->step = frames->step;
    post->duration = frames->duration;
    post->num_samples = chunks->num_samples;
}

int sd_aggregate_dev(sd_ctx* ctx, const double* d_scores, int C, int F, int K, const sd_window* chunks,
                     const sd_window* frames, int hamming, double missing, int skip_average, double epsilon,
                     double* d_out, int64_t cap_rows, int64_t* num_frames, sd_window* post_frames, double* d_count_out,
                     double* d_mask_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_scores && d_out && chunks && frames, "sd_aggregate_dev: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0, "sd_aggregate_dev: C, F, K must be positive");
    SD_REQUIRE(ctx, chunks->num_samples > 0, "sd_aggregate: scores_frames.num_samples must be > 0 (SD:1181)");
    const int64_t NF = sd_aggregate_num_frames(C, chunks, frames);
    if (num_frames) *num_frames = NF;
    if (NF > cap_rows) return ctx->fail(SD_ERR_CAPACITY, "sd_aggregate: need %lld rows, have %lld", (long long)NF,
                                        (long long)cap_rows);
    fill_post(chunks, frames, post_frames);
    return aggregate_launch(ctx, d_scores, C, F, K, chunks, frames, hamming, missing, skip_average, epsilon, d_out, NF,
                            d_count_out, d_mask_out);
}

int sd_aggregate(sd_ctx* ctx, const double* scores, int C, int F, int K, const sd_window* chunks,
                 const sd_window* frames, int hamming, double missing, int skip_average, double epsilon, double* out,
                 int64_t cap_rows, int64_t* num_frames, sd_window* post_frames, double* count_out, double* mask_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, scores && out && chunks && frames, "sd_aggregate: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0, "sd_aggregate: C, F, K must be positive");
    const int64_t NF = sd_aggregate_num_frames(C, chunks, frames);
    if (num_frames) *num_frames = NF;
    if (NF > cap_rows) return ctx->fail(SD_ERR_CAPACITY, "sd_aggregate: need %lld rows, have %lld", (long long)NF,
                                        (long long)cap_rows);
    const size_t in_bytes = sizeof(double) * (size_t)C * F * K, out_bytes = sizeof(double) * (size_t)NF * K;
    double* d_in = (double*)ctx->scratch(BUF_AGG_IN, in_bytes);
    double* d_out = (double*)ctx->scratch(BUF_AGG_OUT, out_bytes);
    double* d_cnt = count_out ? (double*)ctx->scratch(BUF_AGG_AUX, out_bytes) : nullptr;
    double* d_msk = mask_out ? (double*)ctx->scratch(BUF_AGG_AUX2, out_bytes) : nullptr;
    if (!d_in || !d_out || (count_out && !d_cnt) || (mask_out && !d_msk)) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_in, scores, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    int64_t nf2 = 0;
    int rc = sd_aggregate_dev(ctx, d_in, C, F, K, chunks, frames, hamming, missing, skip_average, epsilon, d_out, NF,
                              &nf2, post_frames, d_cnt, d_msk);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (count_out) SD_CUDA(ctx, cudaMemcpyAsync(count_out, d_cnt, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (mask_out) SD_CUDA(ctx, cudaMemcpyAsync(mask_out, d_msk, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

/* ---------------------------------------------------------------- binarize / trim / count / clean */

int sd_binarize_dev(sd_ctx* ctx, const float* d_scores, int C, int F, int K, double onset, int initial_state,
                    double* d_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_scores && d_out, "sd_binarize_dev: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0, "sd_binarize_dev: C, F, K must be positive");
    return binarize_launch(ctx, d_scores, C, F, K, onset, initial_state, d_out);
}

int sd_binarize(sd_ctx* ctx, const float* scores, int C, int F, int K, double onset, int initial_state, double* out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, scores && out, "sd_binarize: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0, "sd_binarize: C, F, K must be positive");
    const size_t n = (size_t)C * F * K;
    float* d_in = (float*)ctx->scratch(BUF_BIN_IN, sizeof(float) * n);
    double* d_out = (double*)ctx->scratch(BUF_BIN_OUT, sizeof(double) * n);
    if (!d_in || !d_out) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_in, scores, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = binarize_launch(ctx, d_in, C, F, K, onset, initial_state, d_out);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

int sd_binarize_rows(sd_ctx* ctx, const double* scores, int R, int F, double onset, int initial_state, uint8_t* out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, scores && out, "sd_binarize_rows: null pointer");
    SD_REQUIRE(ctx, R > 0 && F > 0, "sd_binarize_rows: R, F must be positive");
    const size_t n = (size_t)R * F;
    double* d_in = (double*)ctx->scratch(BUF_BIN_OUT, sizeof(double) * n);
    uint8_t* d_out = (uint8_t*)ctx->scratch(BUF_BIN_IN, n);
    if (!d_in || !d_out) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_in, scores, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = binarize_rows_launch(ctx, d_in, R, F, onset, initial_state, d_out);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

int64_t sd_trim_num_frames(int F, double left, double right) {
    // floor, not round: speakerDiarizer.cpp:1755-1758
    return (int64_t)F - (int64_t)std::floor((double)F * left) - (int64_t)std::floor((double)F * right);
}

static void trimmed_window(int F, double left, double right, const sd_window* before, sd_window* tw) {
    tw->start = before->start + left * before->duration;  // speakerDiarizer.cpp:1776-1779
    tw->step = before->step;
    tw->duration = (1 - left - right) * before->duration;
    tw->num_samples = sd_trim_num_frames(F, left, right);
}

int sd_trim(sd_ctx* ctx, const double* binarized, int C, int F, int K, double left, double right,
            const sd_window* before, double* out, sd_window* trimmed_frames) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, binarized && out && before, "sd_trim: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0, "sd_trim: C, F, K must be positive");
    const int nl = (int)std::floor((double)F * left);
    const int Ft = (int)sd_trim_num_frames(F, left, right);
    SD_REQUIRE(ctx, Ft > 0, "sd_trim: nothing left after trimming");
    if (trimmed_frames) trimmed_window(F, left, right, before, trimmed_frames);
    const size_t in_bytes = sizeof(double) * (size_t)C * F * K, out_bytes = sizeof(double) * (size_t)C * Ft * K;
    double* d_in = (double*)ctx->scratch(BUF_BIN_OUT, in_bytes);
    double* d_out = (double*)ctx->scratch(BUF_CNT_TMP, out_bytes);
    if (!d_in || !d_out) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_in, binarized, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = trim_launch(ctx, d_in, C, F, K, nl, Ft, d_out);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

int sd_speaker_count_dev(sd_ctx* ctx, const double* d_binarized, int C, int F, int K, const sd_window* chunks,
                         const sd_window* frames, int32_t* d_out, int64_t cap, int64_t* n_out, sd_window* count_frames) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_binarized && d_out && chunks && frames, "sd_speaker_count_dev: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0, "sd_speaker_count_dev: C, F, K must be positive");
    // trim 10% / 10% of the chunk window (speakerDiarizer.cpp:1691-1693; the reference anchors it at 0.0)
    sd_window before = *chunks, tw;
    before.start = 0.0;
    trimmed_window(F, 0.1, 0.1, &before, &tw);
    const int nl = (int)std::floor((double)F * 0.1);
    const int Ft = (int)tw.num_samples;
    SD_REQUIRE(ctx, Ft > 0, "sd_speaker_count: nothing left after trimming");
    const int64_t NF = sd_aggregate_num_frames(C, &tw, frames);
    if (n_out) *n_out = NF;
    if (NF > cap) return ctx->fail(SD_ERR_CAPACITY, "sd_speaker_count: need %lld entries, have %lld", (long long)NF,
                                   (long long)cap);
    // one kernel: sum over the classes of the trimmed frames -> aggregate(hamming=false, missing=0.0,
    // skip_average=false) (SD:1719) -> np.rint (SD:1731-1735); the stage-dump entry points still run the three steps
    // separately (stages.cu)
    const int rc = speaker_count_launch(ctx, d_binarized, C, F, K, nl, Ft, &tw, frames, NF, d_out);
    if (rc) return rc;
    fill_post(&tw, frames, count_frames);
    return SD_OK;
}

int sd_speaker_count(sd_ctx* ctx, const double* binarized, int C, int F, int K, const sd_window* chunks,
                     const sd_window* frames, int32_t* out, int64_t cap, int64_t* n_out, sd_window* count_frames) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, binarized && out && chunks && frames, "sd_speaker_count: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0, "sd_speaker_count: C, F, K must be positive");
    const size_t in_bytes = sizeof(double) * (size_t)C * F * K;
    double* d_in = (double*)ctx->scratch(BUF_BIN_OUT, in_bytes);
    int32_t* d_out = (int32_t*)ctx->scratch(BUF_GENERIC_A, sizeof(int32_t) * (size_t)std::max<int64_t>(cap, 1));
    if (!d_in || !d_out) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_in, binarized, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    int64_t n = 0;
    int rc = sd_speaker_count_dev(ctx, d_in, C, F, K, chunks, frames, d_out, cap, &n, count_frames);
    if (n_out) *n_out = n;
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, d_out, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

int sd_clean_segmentations(sd_ctx* ctx, const double* binarized, int C, int F, int K, double* out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, binarized && out, "sd_clean_segmentations: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0, "sd_clean_segmentations: C, F, K must be positive");
    const size_t bytes = sizeof(double) * (size_t)C * F * K;
    double* d_in = (double*)ctx->scratch(BUF_BIN_OUT, bytes);
    double* d_out = (double*)ctx->scratch(BUF_CNT_TMP, bytes);
    if (!d_in || !d_out) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_in, binarized, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = clean_launch(ctx, d_in, (int64_t)C * F, K, d_out);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

/* ---------------------------------------------------------------- clustering library */

static int check_status(sd_ctx* ctx) {
    SD_CUDA(ctx, cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return sdb::status_message(ctx, *ctx->h_status);
}

static int reset_status(sd_ctx* ctx) {
    SD_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->stream));
    return SD_OK;
}

int sd_normalize(sd_ctx* ctx, double* x, int N, int D) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, x, "sd_normalize: null pointer");
    SD_REQUIRE(ctx, N > 0 && D > 0, "sd_normalize: N, D must be positive");
    const size_t bytes = sizeof(double) * (size_t)N * D;
    double* d_x = (double*)ctx->scratch(BUF_CL_X, bytes);
    double* d_xn = (double*)ctx->scratch(BUF_CL_XN, bytes);
    if (!d_x || !d_xn) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_x, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = normalize_launch(ctx, d_x, N, D, d_xn);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(x, d_xn, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

int sd_pdist(sd_ctx* ctx, const double* x, int N, int D, int mode, double* condensed) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, x && condensed, "sd_pdist: null pointer");
    SD_REQUIRE(ctx, N > 1 && D > 0, "sd_pdist: need N > 1, D > 0");
    const size_t bytes = sizeof(double) * (size_t)N * D;
    const size_t cbytes = sizeof(double) * ((size_t)N * (N - 1) / 2);
    double* d_x = (double*)ctx->scratch(BUF_CL_X, bytes);
    double* d_c = (double*)ctx->scratch(BUF_GENERIC_B, cbytes);
    if (!d_x || !d_c) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_x, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = pdist_condensed_launch(ctx, d_x, N, D, mode, d_c);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(condensed, d_c, cbytes, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

int sd_linkage_dev(sd_ctx* ctx, const double* d_x, int N, int D, double* d_Z) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_x && d_Z, "sd_linkage_dev: null pointer");
    SD_REQUIRE(ctx, N > 1 && D > 0, "sd_linkage_dev: need N > 1, D > 0");
    // enqueue only: the status word belongs to the caller (sd_status_reset / sd_status_check)
    return linkage_launch(ctx, d_x, N, D, d_Z, SD_PDIST_EXACT_F64);
}

int sd_linkage_stage_ms(sd_ctx* ctx, float* pdist_ms, float* merges_ms) {
    if (!ctx || !pdist_ms || !merges_ms) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    SD_REQUIRE(ctx, ctx->ev_lk[0] != nullptr, "sd_linkage_stage_ms: no linkage has run on this context yet");
    SD_CUDA(ctx, cudaEventSynchronize(ctx->ev_lk[2]));
    SD_CUDA(ctx, cudaEventElapsedTime(pdist_ms, ctx->ev_lk[0], ctx->ev_lk[1]));
    SD_CUDA(ctx, cudaEventElapsedTime(merges_ms, ctx->ev_lk[1], ctx->ev_lk[2]));
    return SD_OK;
}

int sd_linkage(sd_ctx* ctx, const double* x, int N, int D, double* Z) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, x && Z, "sd_linkage: null pointer");
    SD_REQUIRE(ctx, N > 1 && D > 0, "sd_linkage: need N > 1, D > 0");
    const size_t bytes = sizeof(double) * (size_t)N * D;
    double* d_x = (double*)ctx->scratch(BUF_CL_X, bytes);
    double* d_Z = (double*)ctx->scratch(BUF_CL_Z, sizeof(double) * 4 * (size_t)N);
    if (!d_x || !d_Z) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_x, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = sd_linkage_dev(ctx, d_x, N, D, d_Z);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(Z, d_Z, sizeof(double) * 4 * (size_t)(N - 1), cudaMemcpyDeviceToHost, ctx->stream));
    return check_status(ctx);
}

int sd_fcluster(sd_ctx* ctx, const double* Z, int N, double cutoff, int32_t* T) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, Z && T, "sd_fcluster: null pointer");
    SD_REQUIRE(ctx, N > 1, "sd_fcluster: need N > 1");
    double* d_Z = (double*)ctx->scratch(BUF_CL_Z, sizeof(double) * 4 * (size_t)N);
    int* d_T = (int*)ctx->scratch(BUF_CL_LABELS, sizeof(int) * (size_t)N + 256);
    if (!d_Z || !d_T) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_Z, Z, sizeof(double) * 4 * (size_t)(N - 1), cudaMemcpyHostToDevice, ctx->stream));
    int rc = fcluster_launch(ctx, d_Z, N, cutoff, d_T, d_T + N);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(T, d_T, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

int sd_cluster(sd_ctx* ctx, const double* x, int N, int D, double cutoff, int32_t* T) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, x && T, "sd_cluster: null pointer");
    SD_REQUIRE(ctx, N > 1 && D > 0, "sd_cluster: need N > 1, D > 0");
    const size_t bytes = sizeof(double) * (size_t)N * D;
    double* d_x = (double*)ctx->scratch(BUF_CL_X, bytes);
    double* d_Z = (double*)ctx->scratch(BUF_CL_Z, sizeof(double) * 4 * (size_t)N);
    int* d_T = (int*)ctx->scratch(BUF_CL_LABELS, sizeof(int) * (size_t)N + 256);
    if (!d_x || !d_Z || !d_T) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_x, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = reset_status(ctx);
    if (rc) return rc;
    rc = linkage_launch(ctx, d_x, N, D, d_Z, SD_PDIST_EXACT_F64);
    if (rc) return rc;
    rc = fcluster_launch(ctx, d_Z, N, cutoff, d_T, d_T + N);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(T, d_T, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    return check_status(ctx);
}

int sd_cosine_cdist(sd_ctx* ctx, const double* a, int na, const double* b, int nb, int D, double* out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, a && b && out, "sd_cosine_cdist: null pointer");
    SD_REQUIRE(ctx, na > 0 && nb > 0 && D > 0, "sd_cosine_cdist: sizes must be positive");
    double* d_a = (double*)ctx->scratch(BUF_CL_X, sizeof(double) * (size_t)na * D);
    double* d_b = (double*)ctx->scratch(BUF_CL_XN, sizeof(double) * (size_t)nb * D);
    double* d_o = (double*)ctx->scratch(BUF_CL_SOFT, sizeof(double) * (size_t)na * nb);
    if (!d_a || !d_b || !d_o) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_a, a, sizeof(double) * (size_t)na * D, cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(d_b, b, sizeof(double) * (size_t)nb * D, cudaMemcpyHostToDevice, ctx->stream));
    int rc = reset_status(ctx);
    if (rc) return rc;
    rc = cosine_cdist_launch(ctx, d_a, na, d_b, nb, D, d_o);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, d_o, sizeof(double) * (size_t)na * nb, cudaMemcpyDeviceToHost, ctx->stream));
    return check_status(ctx);
}

/* ---------------------------------------------------------------- clustering driver */

void sd_cluster_default_params(sd_cluster_params* p) {
    if (!p) return;
    p->threshold = 0.7153814381597874f;  // stored as float by the reference (speakerDiarizer.cpp:2049)
    p->min_cluster_size = 15;
    p->num_clusters = -1;
    p->min_clusters = -1;
    p->max_clusters = -1;
    p->pdist_mode = SD_PDIST_EXACT_F64;
}

int sd_cluster_labels(sd_ctx* ctx, const double* x, int N, int D, const sd_cluster_params* p, int32_t* labels) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, x && labels && p, "sd_cluster_labels: null pointer");
    SD_REQUIRE(ctx, N > 0 && D > 0, "sd_cluster_labels: N, D must be positive");
    if (N == 1) {
        labels[0] = 0;
        return SD_OK;
    }
    const size_t bytes = sizeof(double) * (size_t)N * D;
    double* d_x = (double*)ctx->scratch(BUF_CL_X, bytes);
    int* d_lab = (int*)ctx->scratch(BUF_CL_LABELS, sizeof(int) * (size_t)N + 256);
    if (!d_x || !d_lab) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_x, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = reset_status(ctx);
    if (rc) return rc;
    rc = cluster_labels_launch(ctx, d_x, N, D, p, d_lab, d_lab + N, 0);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(labels, d_lab, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    return check_status(ctx);
}

int sd_clustering_dev(sd_ctx* ctx, const double* d_embeddings, int C, int S, int D, const sd_cluster_params* p,
                      const double* d_binarized, int F, int32_t* d_hard, double* d_soft, int soft_k_cap,
                      int* num_clusters) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_embeddings && d_hard && p, "sd_clustering_dev: null pointer");
    SD_REQUIRE(ctx, C > 0 && S > 0 && D > 0, "sd_clustering_dev: C, S, D must be positive");
    const int R = C * S;
    unsigned char* d_valid = (unsigned char*)ctx->scratch(BUF_GENERIC_A, (size_t)R);
    if (!d_valid) return SD_ERR_NOMEM;
    int rc = row_valid_launch(ctx, d_embeddings, R, D, d_valid);
    if (rc) return rc;
    std::vector<unsigned char> valid((size_t)R);
    SD_CUDA(ctx, cudaMemcpyAsync(valid.data(), d_valid, (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int> keep;
    keep.reserve((size_t)R);
    for (int r = 0; r < R; ++r)
        if (valid[r]) keep.push_back(r);
    rc = reset_status(ctx);
    if (rc) return rc;
    rc = clustering_launch(ctx, d_embeddings, C, S, D, keep.data(), (int)keep.size(), p, d_binarized, F, d_hard, d_soft,
                           soft_k_cap, num_clusters, nullptr);
    if (rc) return rc;
    return check_status(ctx);
}

int sd_clustering_async_dev(sd_ctx* ctx, const double* d_embeddings, int C, int S, int D, const sd_cluster_params* p,
                            const int32_t* keep_rows, int n_keep, const double* d_binarized, int F, int32_t* d_hard,
                            double* d_soft, int soft_k_cap, int32_t* d_num_clusters) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    SD_REQUIRE(ctx, d_embeddings && d_hard && p && d_num_clusters, "sd_clustering_async_dev: null pointer");
    SD_REQUIRE(ctx, C > 0 && S > 0 && D > 0, "sd_clustering_async_dev: C, S, D must be positive");
    SD_REQUIRE(ctx, n_keep >= 0 && n_keep <= C * S && (n_keep == 0 || keep_rows),
               "sd_clustering_async_dev: keep_rows must list 0..C*S valid rows");
    for (int i = 0; i < n_keep; ++i)
        SD_REQUIRE(ctx, keep_rows[i] >= 0 && keep_rows[i] < C * S && (i == 0 || keep_rows[i] > keep_rows[i - 1]),
                   "sd_clustering_async_dev: keep_rows must be ascending row indices");
    return clustering_launch(ctx, d_embeddings, C, S, D, keep_rows, n_keep, p, d_binarized, F, d_hard, d_soft, soft_k_cap,
                             nullptr, d_num_clusters);
}

int sd_status_reset(sd_ctx* ctx) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return reset_status(ctx);
}

int sd_status_check(sd_ctx* ctx) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return check_status(ctx);
}

int sd_clustering(sd_ctx* ctx, const double* embeddings, int C, int S, int D, const sd_cluster_params* p,
                  const double* binarized, int F, int32_t* hard, double* soft, int soft_k_cap, int* num_clusters) {
    return sd_clustering_ex(ctx, embeddings, C, S, D, p, binarized, F, hard, soft, nullptr, soft_k_cap, num_clusters);
}

int sd_clustering_ex(sd_ctx* ctx, const double* embeddings, int C, int S, int D, const sd_cluster_params* p,
                     const double* binarized, int F, int32_t* hard, double* soft, double* dist, int soft_k_cap,
                     int* num_clusters) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, embeddings && hard && p, "sd_clustering: null pointer");
    SD_REQUIRE(ctx, C > 0 && S > 0 && D > 0, "sd_clustering: C, S, D must be positive");
    SD_REQUIRE(ctx, !binarized || F > 0, "sd_clustering: F must be positive when binarized is given");
    const int R = C * S;
    const size_t ebytes = sizeof(double) * (size_t)R * D;
    double* d_emb = (double*)ctx->scratch(BUF_CL_EMB, ebytes);
    int* d_hard = (int*)ctx->scratch(BUF_GENERIC_B, sizeof(int) * (size_t)R);
    double* d_bin = nullptr;
    double* d_soft = nullptr;
    if (!d_emb || !d_hard) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_emb, embeddings, ebytes, cudaMemcpyHostToDevice, ctx->stream));
    if (binarized) {
        const size_t bbytes = sizeof(double) * (size_t)C * F * S;
        d_bin = (double*)ctx->scratch(BUF_CL_BIN, bbytes);
        if (!d_bin) return SD_ERR_NOMEM;
        SD_CUDA(ctx, cudaMemcpyAsync(d_bin, binarized, bbytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    double* d_dist = nullptr;
    if ((soft || dist) && soft_k_cap > 0) {
        const size_t sb = sizeof(double) * (size_t)R * soft_k_cap;
        double* d_both = (double*)ctx->scratch(BUF_CL_SOFT, 2 * sb);
        if (!d_both) return SD_ERR_NOMEM;
        // entries beyond the number of clusters stay NaN
        SD_CUDA(ctx, cudaMemsetAsync(d_both, 0xff, 2 * sb, ctx->stream));
        if (soft) d_soft = d_both;
        if (dist) d_dist = d_both + (size_t)R * soft_k_cap;
    }
    // filter_embeddings (speakerDiarizer.cpp:2222-2229): the host already holds the rows
    std::vector<int> keep;
    keep.reserve((size_t)R);
    for (int r = 0; r < R; ++r)
        if (!std::isnan(embeddings[(size_t)r * D])) keep.push_back(r);
    int rc = reset_status(ctx);
    if (rc) return rc;
    rc = clustering_launch(ctx, d_emb, C, S, D, keep.data(), (int)keep.size(), p, d_bin, F, d_hard, d_soft, soft_k_cap,
                           num_clusters, nullptr, d_dist);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(hard, d_hard, sizeof(int) * (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_dist)
        SD_CUDA(ctx, cudaMemcpyAsync(dist, d_dist, sizeof(double) * (size_t)R * soft_k_cap, cudaMemcpyDeviceToHost,
                                     ctx->stream));
    if (d_soft)
        SD_CUDA(ctx, cudaMemcpyAsync(soft, d_soft, sizeof(double) * (size_t)R * soft_k_cap, cudaMemcpyDeviceToHost,
                                     ctx->stream));
    return check_status(ctx);
}

/* ---------------------------------------------------------------- next rows: masking */

int sd_mask_compact(sd_ctx* ctx, const float* wav, const float* masks, int B, int L, int F, int min_num_samples,
                    float* signals, float* wav_lens, uint8_t* too_short, int* all_too_short) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, wav && masks && signals && wav_lens && too_short, "sd_mask_compact: null pointer");
    SD_REQUIRE(ctx, B > 0 && L > 0 && F > 0 && L > F, "sd_mask_compact: need B > 0 and L > F > 0 (SD:753)");
    const size_t wb = sizeof(float) * (size_t)B * L, mb = sizeof(float) * (size_t)B * F;
    float* d_wav = (float*)ctx->scratch(BUF_STFT_IN, wb);
    float* d_sig = (float*)ctx->scratch(BUF_STFT_OUT, wb);
    char* d_small = (char*)ctx->scratch(BUF_GENERIC_B, mb + sizeof(float) * (size_t)B + 2 * (size_t)B + 64);
    if (!d_wav || !d_sig || !d_small) return SD_ERR_NOMEM;
    float* d_masks = reinterpret_cast<float*>(d_small);
    float* d_lens = reinterpret_cast<float*>(d_small + mb);
    unsigned char* d_ts = reinterpret_cast<unsigned char*>(d_small + mb + sizeof(float) * (size_t)B);
    unsigned char* d_inv = d_ts + B;
    SD_CUDA(ctx, cudaMemcpyAsync(d_wav, wav, wb, cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(d_masks, masks, mb, cudaMemcpyHostToDevice, ctx->stream));
    int rc = mask_compact_launch(ctx, d_wav, nullptr, L, (long)B * L, d_masks, B, L, F, B, min_num_samples, d_sig, d_lens,
                                 d_ts, d_inv);
    if (rc) return rc;
    unsigned char inv = 0;
    SD_CUDA(ctx, cudaMemcpyAsync(signals, d_sig, wb, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(wav_lens, d_lens, sizeof(float) * (size_t)B, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(too_short, d_ts, (size_t)B, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(&inv, d_inv, 1, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (all_too_short) *all_too_short = inv;
    return SD_OK;
}

int sd_select_masks_dev(sd_ctx* ctx, const double* d_binarized, int C, int F, int K, double min_num_frames,
                        float* d_masks) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_binarized && d_masks, "sd_select_masks_dev: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0, "sd_select_masks_dev: C, F, K must be positive");
    return select_masks_launch(ctx, d_binarized, C, F, K, min_num_frames, d_masks);
}

int sd_mask_compact_file_dev(sd_ctx* ctx, const float* d_wave, int64_t num_samples, int C, int K, int L,
                             int step_samples, const float* d_masks, int F, int batch, int min_num_samples,
                             float* d_signals, float* d_wav_lens, uint8_t* d_too_short, uint8_t* d_batch_invalid) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_wave && d_masks && d_signals && d_wav_lens && d_too_short, "sd_mask_compact_file_dev: null pointer");
    SD_REQUIRE(ctx, C > 0 && K > 0 && L > F && F > 0 && batch > 0 && step_samples > 0, "sd_mask_compact_file_dev: bad sizes");
    const int R = C * K;
    std::vector<long> base((size_t)R);
    for (int c = 0; c < C; ++c)
        for (int k = 0; k < K; ++k) base[(size_t)c * K + k] = (long)c * step_samples;
    long* d_base = (long*)ctx->scratch(BUF_CL_MISC, sizeof(long) * (size_t)R);
    if (!d_base) return SD_ERR_NOMEM;
    int rc = upload_small(ctx, d_base, base.data(), sizeof(long) * (size_t)R);
    if (rc) return rc;
    return mask_compact_launch(ctx, d_wave, d_base, 0, (long)num_samples, d_masks, R, L, F, batch, min_num_samples,
                               d_signals, d_wav_lens, d_too_short, d_batch_invalid);
}

/* ---------------------------------------------------------------- next rows: reconstruct / to_annotation */

int sd_reconstruct_rows(int C, const sd_window* chunks, int64_t n_count, const sd_window* count_frames,
                        int64_t* rows, sd_window* frames_out) {
    if (!chunks || !count_frames || C <= 0 || n_count < 0) return SD_ERR_INVALID;
    return reconstruct_rows(C, chunks, n_count, count_frames, rows, frames_out);
}

int sd_reconstruct_dev(sd_ctx* ctx, const float* d_seg, int C, int F, int K, const sd_window* chunks,
                       const int32_t* d_hard, int cols, const int32_t* d_count, int64_t n_count,
                       const sd_window* count_frames, double* d_out, int64_t cap_elems, int64_t* rows_out,
                       sd_window* frames_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_seg && d_hard && d_count && d_out && chunks && count_frames, "sd_reconstruct_dev: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0 && cols > 0 && n_count > 0, "sd_reconstruct_dev: sizes must be positive");
    SD_REQUIRE(ctx, chunks->num_samples > 0, "sd_reconstruct: chunk window num_samples must be > 0 (SD:1181)");
    return reconstruct_launch(ctx, d_seg, C, F, K, chunks, d_hard, cols, d_count, n_count, count_frames, d_out, cap_elems,
                              rows_out, frames_out);
}

int sd_reconstruct(sd_ctx* ctx, const float* segmentations, int C, int F, int K, const sd_window* chunks,
                   const int32_t* hard_clusters, const int32_t* count, int64_t n_count,
                   const sd_window* count_frames, double* out, int64_t cap_elems, int64_t* rows_out, int* cols_out,
                   sd_window* frames_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, segmentations && hard_clusters && count && out && chunks && count_frames,
               "sd_reconstruct: null pointer");
    SD_REQUIRE(ctx, C > 0 && F > 0 && K > 0 && n_count > 0, "sd_reconstruct: sizes must be positive");
    int Kc = 0;  // SD:2803-2812: max label (floor 0) + 1
    for (long i = 0; i < (long)C * K; ++i) Kc = hard_clusters[i] > Kc ? hard_clusters[i] : Kc;
    Kc += 1;
    if (cols_out) *cols_out = Kc;
    const size_t seg_b = sizeof(float) * (size_t)C * F * K, hard_b = sizeof(int) * (size_t)C * K,
                 cnt_b = sizeof(int) * (size_t)n_count;
    float* d_seg = (float*)ctx->scratch(BUF_BIN_IN, seg_b);
    char* d_io = (char*)ctx->scratch(BUF_DZ_IO, ((hard_b + 15) / 16) * 16 + cnt_b);
    int64_t rows = 0;
    reconstruct_rows(C, chunks, n_count, count_frames, &rows, nullptr);
    if (rows_out) *rows_out = rows;
    if (rows * Kc > cap_elems)
        return ctx->fail(SD_ERR_CAPACITY, "sd_reconstruct: need %lld elements, have %lld", (long long)(rows * Kc),
                         (long long)cap_elems);
    double* d_out = (double*)ctx->scratch(BUF_DZ_IO2, sizeof(double) * (size_t)(rows > 0 ? rows : 1) * Kc);
    if (!d_seg || !d_io || !d_out) return SD_ERR_NOMEM;
    int* d_hard = reinterpret_cast<int*>(d_io);
    int* d_count = reinterpret_cast<int*>(d_io + ((hard_b + 15) / 16) * 16);
    SD_CUDA(ctx, cudaMemcpyAsync(d_seg, segmentations, seg_b, cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(d_hard, hard_clusters, hard_b, cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(d_count, count, cnt_b, cudaMemcpyHostToDevice, ctx->stream));
    int rc = reconstruct_launch(ctx, d_seg, C, F, K, chunks, d_hard, Kc, d_count, n_count, count_frames, d_out,
                                rows * Kc, nullptr, frames_out);
    if (rc) return rc;
    if (rows > 0)
        SD_CUDA(ctx, cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)rows * Kc, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

int sd_to_annotation_dev(sd_ctx* ctx, const double* d_scores, int64_t rows, int cols, const sd_window* frames,
                         double onset, double offset, double min_duration_on, double min_duration_off,
                         double* d_segments, int32_t* d_labels, int64_t cap, int64_t* n_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_scores && frames && d_segments && d_labels && n_out, "sd_to_annotation_dev: null pointer");
    SD_REQUIRE(ctx, rows > 0 && cols > 0 && cap >= 0, "sd_to_annotation_dev: rows and cols must be positive");
    long* d_n = (long*)ctx->scratch(BUF_GENERIC_A, sizeof(long));
    if (!d_n) return SD_ERR_NOMEM;
    int rc = to_annotation_launch(ctx, d_scores, rows, cols, frames, onset, offset, min_duration_on, min_duration_off,
                                  d_segments, d_labels, cap, d_n);
    if (rc) return rc;
    long n = 0;
    SD_CUDA(ctx, cudaMemcpyAsync(&n, d_n, sizeof(long), cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    if (n > cap) return ctx->fail(SD_ERR_CAPACITY, "sd_to_annotation: %ld segments, capacity %lld", n, (long long)cap);
    return SD_OK;
}

int sd_to_annotation(sd_ctx* ctx, const double* scores, int64_t rows, int cols, const sd_window* frames, double onset,
                     double offset, double min_duration_on, double min_duration_off, double* segments,
                     int32_t* labels, int64_t cap, int64_t* n_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, scores && frames && segments && labels && n_out, "sd_to_annotation: null pointer");
    SD_REQUIRE(ctx, rows > 0 && cols > 0 && cap >= 0, "sd_to_annotation: rows and cols must be positive");
    const size_t in_b = sizeof(double) * (size_t)rows * cols;
    double* d_in = (double*)ctx->scratch(BUF_DZ_IO2, in_b);
    char* d_o = (char*)ctx->scratch(BUF_DZ_IO, (sizeof(double) * 2 + sizeof(int)) * (size_t)(cap > 0 ? cap : 1));
    if (!d_in || !d_o) return SD_ERR_NOMEM;
    double* d_seg = reinterpret_cast<double*>(d_o);
    int* d_lab = reinterpret_cast<int*>(d_o + sizeof(double) * 2 * (size_t)(cap > 0 ? cap : 1));
    SD_CUDA(ctx, cudaMemcpyAsync(d_in, scores, in_b, cudaMemcpyHostToDevice, ctx->stream));
    int rc = sd_to_annotation_dev(ctx, d_in, rows, cols, frames, onset, offset, min_duration_on, min_duration_off, d_seg,
                                  d_lab, cap, n_out);
    if (rc) return rc;
    const int64_t n = *n_out;
    if (n > 0) {
        SD_CUDA(ctx, cudaMemcpyAsync(segments, d_seg, sizeof(double) * 2 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        SD_CUDA(ctx, cudaMemcpyAsync(labels, d_lab, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SD_OK;
}

/* ---------------------------------------------------------------- next rows: ingest */

int sd_ingest_pcm16_dev(sd_ctx* ctx, const int16_t* d_pcm, int64_t n, float* d_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_pcm && d_out && n > 0, "sd_ingest_pcm16_dev: null pointer or empty input");
    return ingest_pcm16_launch(ctx, d_pcm, (long)n, d_out);
}

int sd_ingest_pcm16(sd_ctx* ctx, const int16_t* pcm, int64_t n, float* out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, pcm && out && n > 0, "sd_ingest_pcm16: null pointer or empty input");
    short* d_in = (short*)ctx->scratch(BUF_GENERIC_A, sizeof(short) * (size_t)n);
    float* d_out = (float*)ctx->scratch(BUF_GENERIC_B, sizeof(float) * (size_t)n);
    if (!d_in || !d_out) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_in, pcm, sizeof(short) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = ingest_pcm16_launch(ctx, d_in, (long)n, d_out);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, d_out, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

int sd_slide_geometry(int64_t num_samples, double duration, double step, int64_t* full_chunks, int64_t* tail_start,
                      int64_t* tail_len) {
    if (num_samples < 0 || !(duration > 0) || !(step > 0)) return SD_ERR_INVALID;
    const int64_t window = (int64_t)std::round(duration * 16000), hop = (int64_t)std::round(step * 16000);  // SD:1411
    if (hop <= 0) return SD_ERR_INVALID;
    // while (i + window < num_samples) i += hop  (SD:1419) in closed form
    const int64_t full = num_samples > window ? (num_samples - window - 1) / hop + 1 : 0;
    const int64_t i = full * hop;
    if (full_chunks) *full_chunks = full;
    const bool tail = i + 1 < num_samples;  // SD:1451
    if (tail_start) *tail_start = tail ? i : -1;
    if (tail_len) *tail_len = tail ? num_samples - i : 0;
    return SD_OK;
}

int sd_crop_chunks_dev(sd_ctx* ctx, const float* d_wave, int64_t num_samples, const double* starts_s, int n_chunks,
                       double duration, int sample_rate, float* d_out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, d_wave && starts_s && d_out, "sd_crop_chunks_dev: null pointer");
    SD_REQUIRE(ctx, num_samples > 0 && n_chunks > 0 && n_chunks <= 65535 && duration > 0 && sample_rate > 0,
               "sd_crop_chunks_dev: bad sizes (1..65535 chunks per call)");
    return crop_chunks_launch(ctx, d_wave, (long)num_samples, starts_s, n_chunks, duration, sample_rate, d_out);
}

int sd_crop_chunks(sd_ctx* ctx, const float* wave, int64_t num_samples, const double* starts_s, int n_chunks,
                   double duration, int sample_rate, float* out) {
    if (!ctx) return SD_ERR_INVALID;
    cudaSetDevice(ctx->device);  // the current device is per host thread: calls may come from any thread
    SD_REQUIRE(ctx, wave && starts_s && out, "sd_crop_chunks: null pointer");
    SD_REQUIRE(ctx, num_samples > 0 && n_chunks > 0 && duration > 0 && sample_rate > 0, "sd_crop_chunks: bad sizes");
    const size_t L = (size_t)std::floor(duration * sample_rate);
    float* d_w = (float*)ctx->scratch(BUF_STFT_IN, sizeof(float) * (size_t)num_samples);
    float* d_o = (float*)ctx->scratch(BUF_STFT_OUT, sizeof(float) * L * (size_t)n_chunks);
    if (!d_w || !d_o) return SD_ERR_NOMEM;
    SD_CUDA(ctx, cudaMemcpyAsync(d_w, wave, sizeof(float) * (size_t)num_samples, cudaMemcpyHostToDevice, ctx->stream));
    int rc = sd_crop_chunks_dev(ctx, d_w, num_samples, starts_s, n_chunks, duration, sample_rate, d_o);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, d_o, sizeof(float) * L * (size_t)n_chunks, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SD_OK;
}

}  // extern "C"
