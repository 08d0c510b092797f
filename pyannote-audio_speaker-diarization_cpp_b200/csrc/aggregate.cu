// Sliding-window aggregation.
// Replaces PipelineHelper::aggregate (speakerDiarizer.cpp:1167-1311): overlap-add of per-chunk frame scores
// onto the global frame grid with NaN masking, chunk counts, optional averaging and `missing` fill.
//
// Gather formulation: one thread per output element (frame, class) walks the <= ceil(F / frames_per_step)+1
// chunks that cover the frame in increasing chunk order, so every fp64 sum is formed in exactly the order
// of the reference's chunk-major scatter loop (speakerDiarizer.cpp:1249-1268) -- bit-identical, no atomics.
// HBM-bound: 8*C*F*K bytes read + 8*NF*K written (+ the optional count/mask planes).
#include "common.cuh"

#include <cfloat>
#include <cmath>

namespace sdb {

// Helper::np_rint (speakerDiarizer.cpp:260-272)
__host__ __device__ inline int np_rint_impl(double v) {
    const double sgn = v > 0 ? 1.0 : -1.0;
    const double off = v - (double)(int)v - 0.5 * sgn;
    if (fabs(off) < DBL_EPSILON) {
        const int r = (int)round(v);
        return (r % 2 == 0) ? r : r - (v > 0 ? 1 : -1);
    }
    return (int)round(v);
}

int np_rint_host(double v) { return np_rint_impl(v); }

// SlidingWindow::closest_frame (speakerDiarizer.cpp:1084-1090)
int64_t closest_frame_host(double sw_start, double sw_step, double sw_duration, double t) {
    double pos = (t - sw_start - .5 * sw_duration) / sw_step;
    if (pos < 0.0) pos = 0.0;
    return (int64_t)(size_t)np_rint_impl(pos);
}

template <bool HAMMING>
__global__ void __launch_bounds__(256)
    aggregate_kernel(const double* __restrict__ scores, const int* __restrict__ starts,
                     const double* __restrict__ weights, int C, int F, int K, long NF, double missing,
                     int skip_average, double epsilon, double* __restrict__ out, double* __restrict__ count_out,
                     double* __restrict__ mask_out) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= NF * K) return;
    const long f = idx / K;
    const int k = (int)(idx - f * K);
    // first chunk whose window [start, start + F) reaches frame f (starts[] is non-decreasing)
    int lo = 0, hi = C;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long)starts[mid] + F > f)
            hi = mid;
        else
            lo = mid + 1;
    }
    double acc = 0.0, cnt = 0.0, msk = 0.0;
    for (int i = lo; i < C; ++i) {
        const long s = starts[i];
        if (s > f) break;
        const int j = (int)(f - s);
        double v = scores[((size_t)i * F + j) * K + k];
        double m = 1.0;
        if (isnan(v)) {  // speakerDiarizer.cpp:1191-1204
            m = 0.0;
            v = 0.0;
        }
        if (HAMMING) {
            const double w = weights[j];
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(v, m), w));
            cnt = __dadd_rn(cnt, __dmul_rn(m, w));
        } else {
            acc = __dadd_rn(acc, __dmul_rn(v, m));
            cnt = __dadd_rn(cnt, m);
        }
        if (m > msk) msk = m;
    }
    if (count_out) count_out[idx] = cnt;
    if (mask_out) mask_out[idx] = msk;
    if (!skip_average) acc = __ddiv_rn(acc, cnt > epsilon ? cnt : epsilon);  // speakerDiarizer.cpp:1288
    if (fabs(msk) < DBL_EPSILON) acc = missing;                              // speakerDiarizer.cpp:1298-1307
    out[idx] = acc;
}

// speaker_count in one pass (speakerDiarizer.cpp:1691-1735): trim -> sum over the classes -> aggregate(hamming = false,
// missing = 0, average) -> np.rint, without the two intermediates.  One thread per output frame walks the covering
// chunks in increasing order like aggregate_kernel; a chunk's value is sum_k bin[c][j + nl][k] formed k-ascending from
// 0.0 exactly as the separate trim_sum pass formed it, so every bit of the chain is the unfused one's.
__global__ void __launch_bounds__(256)
    speaker_count_kernel(const double* __restrict__ bin, const int* __restrict__ starts, int C, int F, int K, int nl,
                         int Ft, long NF, double epsilon, int32_t* __restrict__ out) {
    const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= NF) return;
    int lo = 0, hi = C;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long)starts[mid] + Ft > f)
            hi = mid;
        else
            lo = mid + 1;
    }
    double acc = 0.0, cnt = 0.0, msk = 0.0;
    for (int i = lo; i < C; ++i) {
        const long s = starts[i];
        if (s > f) break;
        const int j = (int)(f - s);
        const double* p = bin + ((size_t)i * F + (size_t)(j + nl)) * K;
        double v = 0.0;
        for (int k = 0; k < K; ++k) v = __dadd_rn(v, p[k]);  // speakerDiarizer.cpp:1701-1714
        double m = 1.0;
        if (isnan(v)) {
            m = 0.0;
            v = 0.0;
        }
        acc = __dadd_rn(acc, __dmul_rn(v, m));
        cnt = __dadd_rn(cnt, m);
        if (m > msk) msk = m;
    }
    acc = __ddiv_rn(acc, cnt > epsilon ? cnt : epsilon);
    if (fabs(msk) < DBL_EPSILON) acc = 0.0;
    out[f] = np_rint_impl(acc);
}

int upload_small(sd_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);

// tw: the trimmed chunk window (its frames are the Ft kept frames of every chunk)
int speaker_count_launch(sd_ctx* ctx, const double* d_bin, int C, int F, int K, int nl, int Ft, const sd_window* tw,
                         const sd_window* frames, int64_t NF, int32_t* d_out) {
    std::vector<int> starts((size_t)C);
    double start = tw->start;
    for (int i = 0; i < C; ++i) {  // as aggregate_launch: running sum of the chunk step, closest frame
        starts[i] = (int)closest_frame_host(tw->start, frames->step, frames->duration, start);
        start += tw->step;
    }
    int* d_starts = (int*)ctx->scratch(BUF_AGG_STARTS, sizeof(int) * (size_t)C + 8);
    if (!d_starts) return SD_ERR_NOMEM;
    const int rc = upload_small(ctx, d_starts, starts.data(), sizeof(int) * (size_t)C);
    if (rc) return rc;
    speaker_count_kernel<<<(unsigned)((NF + 255) / 256), 256, 0, ctx->stream>>>(d_bin, d_starts, C, F, K, nl, Ft, (long)NF,
                                                                            DBL_EPSILON, d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

int aggregate_launch(sd_ctx* ctx, const double* d_scores, int C, int F, int K, const sd_window* chunks,
                     const sd_window* frames, int hamming, double missing, int skip_average, double epsilon,
                     double* d_out, int64_t NF, double* d_count, double* d_mask) {
    // chunk start frames: running sum of the chunk step (speakerDiarizer.cpp:1248-1253), grid anchored at
    // the chunk window start (speakerDiarizer.cpp:1233)
    std::vector<int> starts((size_t)C);
    double start = chunks->start;
    for (int i = 0; i < C; ++i) {
        starts[i] = (int)closest_frame_host(chunks->start, frames->step, frames->duration, start);
        start += chunks->step;
    }
    const size_t off_w = ((sizeof(int) * (size_t)C + 7) / 8) * 8;
    char* base = (char*)ctx->scratch(BUF_AGG_STARTS, off_w + sizeof(double) * (size_t)F);
    if (!base) return SD_ERR_NOMEM;
    int* d_starts = reinterpret_cast<int*>(base);
    double* d_w = reinterpret_cast<double*>(base + off_w);
    int rc = upload_small(ctx, d_starts, starts.data(), sizeof(int) * (size_t)C);
    if (rc) return rc;
    if (hamming) {
        // np.hamming(F) (symmetric) -- pyannote Inference.aggregate; the reference asserts here
        std::vector<double> w((size_t)F);
        for (int j = 0; j < F; ++j) w[j] = F > 1 ? 0.54 - 0.46 * std::cos(2.0 * M_PI * j / (double)(F - 1)) : 1.0;
        rc = upload_small(ctx, d_w, w.data(), sizeof(double) * (size_t)F);
        if (rc) return rc;
    }
    const long total = (long)NF * K;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (hamming)
        aggregate_kernel<true><<<grid, 256, 0, ctx->stream>>>(d_scores, d_starts, d_w, C, F, K, (long)NF, missing,
                                                             skip_average, epsilon, d_out, d_count, d_mask);
    else
        aggregate_kernel<false><<<grid, 256, 0, ctx->stream>>>(d_scores, d_starts, nullptr, C, F, K, (long)NF, missing,
                                                              skip_average, epsilon, d_out, d_count, d_mask);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

}  // namespace sdb
