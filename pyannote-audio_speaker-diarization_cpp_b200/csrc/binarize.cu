// Hysteresis binarisation, trim, per-frame sums and clean-up.
// Replaces SegmentModel::binarize_swf / binarize_ndarray (speakerDiarizer.cpp:1506-1639, with
// Helper::wellDefinedIndex / cumulativeSum / numpy_where, 623-708), SegmentModel::trim (1742-1782), the
// sum over classes of speaker_count (1701-1714), np.rint (1731-1735) and Helper::cleanSegmentations (710-743).
//
// binarize: the six row-sized temporaries of the reference collapse to "value of (s > onset) at the last
// frame <= t whose score is not within DBL_EPSILON of onset, else initial_state".  One warp owns one
// (chunk, class) row and scans it 32 frames at a time with two ballots and a count-leading-zeros.
#include "common.cuh"

#include <cfloat>

namespace sdb {

__host__ __device__ inline int np_rint_dev(double v) {  // Helper::np_rint, speakerDiarizer.cpp:260-272
    const double sgn = v > 0 ? 1.0 : -1.0;
    const double off = v - (double)(int)v - 0.5 * sgn;
    if (fabs(off) < DBL_EPSILON) {
        const int r = (int)round(v);
        return (r % 2 == 0) ? r : r - (v > 0 ? 1 : -1);
    }
    return (int)round(v);
}

// rows are (c, k) with element stride `estride` and row base c*cstride + k*kstride
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
    binarize_kernel(const TIn* __restrict__ scores, long rows, int F, int K, double onset, int initial_state,
                    TOut* __restrict__ out) {
    const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long c = row / K;
    const int k = (int)(row - c * K);
    const size_t base = (size_t)c * F * K + k;
    unsigned carry = initial_state ? 1u : 0u;
    const unsigned le_mask = lane == 31 ? 0xffffffffu : ((2u << lane) - 1u);
    for (int t0 = 0; t0 < F; t0 += 32) {
        const int t = t0 + lane;
        const bool valid = t < F;
        double s = 0.0;
        if (valid) s = (double)scores[base + (size_t)t * K];  // float -> double, speakerDiarizer.cpp:1526
        const bool on = valid && (s > onset);
        const bool defined = valid && !(fabs(s - onset) < DBL_EPSILON);  // speakerDiarizer.cpp:1595
        const unsigned onb = __ballot_sync(0xffffffffu, on);
        const unsigned defb = __ballot_sync(0xffffffffu, defined);
        const unsigned m = defb & le_mask;
        const unsigned res = m ? ((onb >> (31 - __clz(m))) & 1u) : carry;
        if (valid) out[base + (size_t)t * K] = (TOut)res;
        if (defb) carry = (onb >> (31 - __clz(defb))) & 1u;
    }
}

int binarize_launch(sd_ctx* ctx, const float* d_scores, int C, int F, int K, double onset, int initial_state,
                    double* d_out) {
    const long rows = (long)C * K;
    const unsigned grid = (unsigned)((rows * 32 + 255) / 256);
    binarize_kernel<float, double><<<grid, 256, 0, ctx->stream>>>(d_scores, rows, F, K, onset, initial_state, d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

int binarize_rows_launch(sd_ctx* ctx, const double* d_scores, int R, int F, double onset, int initial_state,
                         uint8_t* d_out) {
    const unsigned grid = (unsigned)(((long)R * 32 + 255) / 256);
    binarize_kernel<double, uint8_t><<<grid, 256, 0, ctx->stream>>>(d_scores, (long)R, F, 1, onset, initial_state,
                                                                    d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

// out[c][j][k] = bin[c][j + nl][k]
__global__ void __launch_bounds__(256)
    trim_kernel(const double* __restrict__ bin, long total, int F, int Ft, int K, int nl, double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long per = (long)Ft * K;
    const long c = i / per;
    const long rem = i - c * per;
    out[i] = bin[(size_t)c * F * K + (size_t)nl * K + rem];
}

// out[c][j] = sum_k bin[c][j + nl][k]   (speakerDiarizer.cpp:1701-1714, k ascending)
__global__ void __launch_bounds__(256) trim_sum_kernel(const double* __restrict__ bin, long total, int F, int Ft,
                                                        int K, int nl, double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long c = i / Ft;
    const long j = i - c * Ft;
    const double* p = bin + ((size_t)c * F + (size_t)(j + nl)) * K;
    double s = 0.0;
    for (int k = 0; k < K; ++k) s = __dadd_rn(s, p[k]);
    out[i] = s;
}

__global__ void __launch_bounds__(256) rint_kernel(const double* __restrict__ in, long n, int32_t* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = np_rint_dev(in[i]);
}

// keep frames where fewer than two classes are active (speakerDiarizer.cpp:720-740)
__global__ void __launch_bounds__(256)
    clean_kernel(const double* __restrict__ bin, long rows, int K, double* __restrict__ out) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double* p = bin + (size_t)r * K;
    double s = 0.0;
    for (int k = 0; k < K; ++k) s = __dadd_rn(s, p[k]);
    const bool keep = s < 2.0;
    for (int k = 0; k < K; ++k) out[(size_t)r * K + k] = keep ? p[k] : 0.0;
}

int trim_launch(sd_ctx* ctx, const double* d_bin, int C, int F, int K, int nl, int Ft, double* d_out) {
    const long total = (long)C * Ft * K;
    trim_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_bin, total, F, Ft, K, nl, d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}
int trim_sum_launch(sd_ctx* ctx, const double* d_bin, int C, int F, int K, int nl, int Ft, double* d_out) {
    const long total = (long)C * Ft;
    trim_sum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_bin, total, F, Ft, K, nl, d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}
int rint_launch(sd_ctx* ctx, const double* d_in, int64_t n, int32_t* d_out) {
    rint_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_in, (long)n, d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}
int clean_launch(sd_ctx* ctx, const double* d_bin, int64_t rows, int K, double* d_out) {
    clean_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(d_bin, (long)rows, K, d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

}  // namespace sdb
