// Shared context / error plumbing of libsdb200.so (no torch, no third-party dependency).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sdb200.h"

namespace sdb {

// grow-only device scratch buffers owned by the context
enum BufTag {
    BUF_STFT_IN = 0,
    BUF_STFT_OUT,
    BUF_AGG_IN,
    BUF_AGG_OUT,
    BUF_AGG_AUX,
    BUF_AGG_AUX2,
    BUF_AGG_STARTS,
    BUF_BIN_IN,
    BUF_BIN_OUT,
    BUF_CNT_TMP,
    BUF_CNT_OUT,
    BUF_CL_EMB,
    BUF_CL_X,
    BUF_CL_XN,
    BUF_CL_DIST,
    BUF_CL_Z,
    BUF_CL_WORK,
    BUF_CL_LABELS,
    BUF_CL_CENT,
    BUF_CL_OUT,
    BUF_CL_SOFT,
    BUF_CL_MISC,
    BUF_CL_BIN,
    BUF_FB_TMP,
    BUF_FB_OUT,
    BUF_FB_LENS,
    BUF_GENERIC_A,
    BUF_GENERIC_B,
    BUF_TC_OPERANDS,
    BUF_DZ_CS,
    BUF_DZ_ACT,
    BUF_DZ_IO,
    BUF_DZ_IO2,
    BUF_AN_FLAGS,
    BUF_AN_LISTS,
    BUF_AN_META,
    BUF_COUNT
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace sdb

struct sd_ctx {
    int device = 0;
    int num_sms = 148;
    size_t l2_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    std::string err;
    int64_t launches = 0;
    cudaEvent_t ev_start[16] = {};
    cudaEvent_t ev_stop[16] = {};
    sdb::DevBuf bufs[sdb::BUF_COUNT];
    // cached constant tables
    float* d_window = nullptr;     // 400 floats of the last window uploaded
    int window_kind = -1;
    std::vector<float> h_window;
    float* d_twiddle = nullptr;    // [20][20] float2, tw[r][k1] = exp(-2 pi i r k1 / 400)
    float* d_mel = nullptr;        // mel matrix [201][n_mels]
    int mel_key = 0;
    int mel_parts = 0;             // d_mel holds the (frame, part) table (mel_table.h) instead of the per-filter one
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    int* d_status = nullptr;       // device-side status word (zero-magnitude etc.)
    int* h_status = nullptr;       // pinned mirror
    unsigned long long* d_stats = nullptr;  // diagnostic counters (sd_debug_counters)
    int force_exact_linkage = 0;            // test hook: skip the heap-free fast path
    int linkage_threads = 0;                // tuning hook: 0 = auto, 512 or 1024
    int linkage_cluster = 1;                // 1 = spread the merge loop over an 8-CTA cluster when the state fits
    int linkage_wide = 0;                   // whole-GPU merge loop: 0 never (default), 1 for N >= 32768, 2 always
    int stft_variant = 0;                   // tuning hook: 0 = 4 CTAs/SM (<=102 regs), 1 = 3 CTAs/SM
    int stft_waves = 1;                     // CTAs per resident slot of the STFT grid (1 = persistent)
    cudaEvent_t ev_lk[3] = {};              // around pdist / the merge loop of the last linkage (sd_linkage_stage_ms)
    void* extra = nullptr;                  // api.cu's CtxExtra (pinned upload ring, copy streams), owned by the context
    // batch timeline (SDB_BATCH_TRACE, batch.cu): %globaltimer stamps written by one-thread kernels on the stream
    unsigned long long* d_trace = nullptr;
    int trace_cap = 0;
    std::vector<int> trace_tags;            // tag of stamp i (host side)

    int fail(int code, const char* fmt, ...) {
        char b[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(b, sizeof(b), fmt, ap);
        va_end(ap);
        err = b;
        return code;
    }
    // returns nullptr on allocation failure (err set)
    void* scratch(int tag, size_t bytes) {
        sdb::DevBuf& b = bufs[tag];
        if (bytes <= b.cap && b.p) return b.p;
        if (b.p) {
            cudaStreamSynchronize(stream);
            cudaFree(b.p);
            b.p = nullptr;
            b.cap = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            fail(SD_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            b.p = nullptr;
            return nullptr;
        }
        b.cap = want;
        return b.p;
    }
};

#define SD_CUDA(ctx, call)                                                                             \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            cudaGetLastError();                                                                        \
            return (ctx)->fail(SD_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
        }                                                                                              \
    } while (0)

#define SD_LAUNCH_CHECK(ctx)                                                                           \
    do {                                                                                               \
        (ctx)->launches++;                                                                             \
        cudaError_t e__ = cudaGetLastError();                                                          \
        if (e__ != cudaSuccess)                                                                        \
            return (ctx)->fail(SD_ERR_CUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

#define SD_REQUIRE(ctx, cond, msg)                                     \
    do {                                                               \
        if (!(cond)) return (ctx)->fail(SD_ERR_INVALID, "%s", msg);   \
    } while (0)

namespace sdb {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and occupancy are per DEVICE and per kernel: remember what was
// done for every (device, kernel) pair behind a mutex, so that contexts on several GPUs and on several host threads
// of one process all get the opt-in (a process-wide `static bool` would skip it on the second GPU).
// Returns the kernel's resident blocks per SM for (threads, smem) when `threads` > 0, else 0.
template <typename Kernel>
inline int kernel_setup(sd_ctx* ctx, Kernel kernel, int max_dyn_smem, int threads = 0, size_t smem = 0) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, std::pair<int, int>> done;  // -> (smem opted in, blocks per SM)
    const void* fn = reinterpret_cast<const void*>(kernel);
    std::lock_guard<std::mutex> lock(mu);
    auto& e = done[std::make_pair(ctx->device, fn)];
    if (max_dyn_smem > e.first) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem);
        if (err != cudaSuccess) {
            cudaGetLastError();
            ctx->fail(SD_ERR_CUDA, "cudaFuncSetAttribute(%d B dynamic shared memory): %s", max_dyn_smem,
                      cudaGetErrorString(err));
            return -1;
        }
        e.first = max_dyn_smem;
        e.second = 0;
    }
    if (threads > 0 && e.second == 0) {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, threads, smem) != cudaSuccess) {
            cudaGetLastError();
            b = 1;
        }
        e.second = b < 1 ? 1 : b;
    }
    return e.second;
}

// ---- internal launch entry points (device pointers, enqueue only) ----
int stft_launch(sd_ctx* ctx, const float* d_wav, int B, int L, const sd_stft_params* p, float* d_out);
int fbank_launch(sd_ctx* ctx, const float* d_wav, int B, int L, const float* d_lens, const sd_fbank_params* p,
                 float* d_out);
int aggregate_launch(sd_ctx* ctx, const double* d_scores, int C, int F, int K, const sd_window* chunks,
                     const sd_window* frames, int hamming, double missing, int skip_average, double epsilon,
                     double* d_out, int64_t NF, double* d_count, double* d_mask);
int speaker_count_launch(sd_ctx* ctx, const double* d_bin, int C, int F, int K, int nl, int Ft, const sd_window* tw,
                         const sd_window* frames, int64_t NF, int32_t* d_out);
int binarize_launch(sd_ctx* ctx, const float* d_scores, int C, int F, int K, double onset, int initial_state,
                    double* d_out);
int binarize_rows_launch(sd_ctx* ctx, const double* d_scores, int R, int F, double onset, int initial_state,
                         uint8_t* d_out);
int trim_launch(sd_ctx* ctx, const double* d_bin, int C, int F, int K, int nl, int Ft, double* d_out);
int trim_sum_launch(sd_ctx* ctx, const double* d_bin, int C, int F, int K, int nl, int Ft, double* d_out);
int rint_launch(sd_ctx* ctx, const double* d_in, int64_t n, int32_t* d_out);
int clean_launch(sd_ctx* ctx, const double* d_bin, int64_t rows, int K, double* d_out);

// latched device status word -> sd_status + message (api.cu)
int status_message(sd_ctx* ctx, int st);

// batch timeline: no-op unless the context has a trace buffer (batch.cu)
void trace_stamp(sd_ctx* ctx, int tag);

// host-side scalar helpers shared by several translation units
int np_rint_host(double v);
int64_t closest_frame_host(double sw_start, double sw_step, double sw_duration, double t);

}  // namespace sdb
