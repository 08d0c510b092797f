// Post-clustering rows:
//   reconstruct      (speakerDiarizer.cpp:2789-2848)  per-cluster max over the local speakers of every chunk
//   to_diarization   (2638-2764) + crop_segment (2568-2635)  skip-average overlap-add, crop to the common extent,
//                    keep the count[t] most active clusters of every frame
//   to_annotation    (2852-2935) + Track::support (911-941) + removeShort (943-953) + finalResult (962-978)
// Everything is exact fp64 / integer work and must be bit-identical to the reference, so products and sums that
// the reference rounds separately use the _rn intrinsics (no FMA contraction).
#include "common.cuh"

#include <cfloat>
#include <cmath>
#include <vector>

namespace sdb {

int aggregate_launch(sd_ctx* ctx, const double* d_scores, int C, int F, int K, const sd_window* chunks,
                     const sd_window* frames, int hamming, double missing, int skip_average, double epsilon,
                     double* d_out, int64_t NF, double* d_count, double* d_mask);

// ---------------------------------------------------------------------------------------------- reconstruct

// One thread per (chunk, frame, cluster).  NaN where no local speaker of the chunk maps to the cluster.
__global__ void __launch_bounds__(256)
    clustered_scores_kernel(const float* __restrict__ seg, const int* __restrict__ hard, int C, int F, int K, int Kc,
                            double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)C * F * Kc;
    if (i >= total) return;
    const int k = (int)(i % Kc);
    const long cf = i / Kc;
    const int c = (int)(cf / F);
    const float* row = seg + cf * K;
    const int* h = hard + (size_t)c * K;
    bool any = false;
    float m = -INFINITY;
    for (int s = 0; s < K; ++s)
        if (h[s] == k) {
            any = true;
            const float v = row[s];
            m = (m < v) ? v : m;  // std::max(maxValue, v): NaN never replaces the running value
        }
    out[i] = any ? (double)m : __longlong_as_double(0x7ff8000000000000LL);
}

// One thread per output frame: stable descending order of the activations (insertion sort, ties keep the
// cluster order, SD:2724-2730), the first min(count, Kc) clusters are set to 1.
__global__ void __launch_bounds__(128)
    top_count_kernel(const double* __restrict__ act, long a0, const int* __restrict__ count, long c0, long rows,
                     long crow, int Kc, int* __restrict__ order_scratch, double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double* o = out + i * Kc;
    for (int k = 0; k < Kc; ++k) o[k] = 0.0;
    const double* a = act + (a0 + i) * Kc;
    int* order = order_scratch + i * Kc;
    for (int k = 0; k < Kc; ++k) {
        int j = k;
        const double key = -a[k];
        while (j > 0 && key < (-a[order[j - 1]])) {
            order[j] = order[j - 1];
            --j;
        }
        order[j] = k;
    }
    if (i >= crow) return;  // sorted for every row (the reference's sorted_speakers), ones only where a count exists
    int cnt = count[c0 + i];
    if (cnt > Kc) cnt = Kc;
    for (int j = 0; j < cnt; ++j) o[order[j]] = 1.0;
}

// SlidingWindow::operator[] (SD:1092-1115): start of window #pos by repeated += step, 0.0 once the window would
// run past num_samples.
static double window_start(double step, double duration, long num_samples, int pos) {
    const int wsize = (int)std::round(duration * 16000.0), ssize = (int)std::round(step * 16000.0);
    double start = 0.0;
    size_t cur = 0;
    for (int idx = 0;; ++idx) {
        if (idx == pos) return start;
        if (cur + (size_t)wsize >= (size_t)num_samples) return 0.0;
        start += step;
        cur += (size_t)ssize;
    }
}

struct CropRange {
    long r0, r1;
    double new_start;
};

// crop_segment, loose mode (SD:2568-2635): the index arithmetic is done in float.
static CropRange crop_range(long n_rows, const sd_window& src, double focus_start, double focus_end) {
    CropRange r{0, 0, 0.0};
    int a = (int)std::ceil((float)((focus_start - src.duration - src.start) / src.step));
    if (a < 0) a = 0;
    const int b = (int)std::floor((float)((focus_end - src.start) / src.step)) + 1;
    r.new_start = (double)(float)window_start(src.step, src.duration, (long)src.num_samples, a);
    if ((long)a >= n_rows) return r;
    r.r0 = a;
    r.r1 = (long)b < n_rows ? (long)b : n_rows;
    if (r.r1 < r.r0) r.r1 = r.r0;
    return r;
}

struct ReconGeom {
    int64_t NF;
    sd_window post;
    CropRange act, cnt;
    long rows, crow;
};

// crop of to_diarization (SD:2686-2713) for activations[NF] on window `post` and count[n_count] on window `cf`
static ReconGeom crop_geometry(int64_t NF, const sd_window& post, int64_t n_count, const sd_window* cf) {
    ReconGeom g;
    g.NF = NF;
    g.post = post;
    // extents (SD:2691-2706)
    const double a_end = (g.post.start + (0 - .5) * g.post.step + .5 * g.post.duration) + (double)g.NF * g.post.step;
    const double c_end = (cf->start + (0 - .5) * cf->step + .5 * cf->duration) + (double)n_count * cf->step;
    const double f0 = g.post.start > cf->start ? g.post.start : cf->start;
    const double f1 = a_end < c_end ? a_end : c_end;
    g.act = crop_range((long)g.NF, g.post, f0, f1);
    g.cnt = crop_range((long)n_count, *cf, f0, f1);
    g.rows = g.act.r1 - g.act.r0;
    g.crow = g.cnt.r1 - g.cnt.r0;
    return g;
}

static ReconGeom recon_geometry(int C, const sd_window* chunks, int64_t n_count, const sd_window* cf) {
    const double target = chunks->start + chunks->duration + (double)(size_t)(C - 1) * chunks->step;
    const int64_t NF = closest_frame_host(chunks->start, cf->step, cf->duration, target) + 1;
    return crop_geometry(NF, sd_window{chunks->start, cf->step, cf->duration, chunks->num_samples}, n_count, cf);
}

// The pieces of reconstruct on their own (stage dumps of verifyEveryStepResult.py: clustered_segmentations,
// to_diarization_activations, cropped_activations, cropped_count, sorted_speakers).
int clustered_scores_launch(sd_ctx* ctx, const float* d_seg, int C, int F, int K, const int* d_hard, int Kc,
                            double* d_out) {
    const size_t n_cs = (size_t)C * F * Kc;
    clustered_scores_kernel<<<(unsigned)((n_cs + 255) / 256), 256, 0, ctx->stream>>>(d_seg, d_hard, C, F, K, Kc, d_out);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

// to_diarization without its aggregate (SD:2672-2764): d_act[NF][Kc] on window `act_frames` (the post_frames of the
// aggregate, num_samples included), count on `cf`.  crop4 = {first activation row, rows, first count row, count rows}.
int to_diarization_launch(sd_ctx* ctx, const double* d_act, int64_t NF, int Kc, const sd_window* act_frames,
                          const int* d_count, int64_t n_count, const sd_window* cf, double* d_out, int64_t cap_elems,
                          int* d_order, int64_t* rows_out, sd_window* frames_out, int64_t* crop4) {
    const ReconGeom g = crop_geometry(NF, *act_frames, n_count, cf);
    if (rows_out) *rows_out = g.rows;
    if (frames_out) *frames_out = sd_window{g.act.new_start, g.post.step, g.post.duration, 0};
    const long crow = g.crow < g.rows ? g.crow : g.rows;
    if (crop4) {
        crop4[0] = g.act.r0;
        crop4[1] = g.rows;
        crop4[2] = g.cnt.r0;
        crop4[3] = g.crow;
    }
    if (g.rows * Kc > cap_elems)
        return ctx->fail(SD_ERR_CAPACITY, "sd_to_diarization: need %lld elements, have %lld", (long long)(g.rows * Kc),
                         (long long)cap_elems);
    if (g.rows > 0) {
        top_count_kernel<<<(unsigned)((g.rows + 127) / 128), 128, 0, ctx->stream>>>(d_act, g.act.r0, d_count, g.cnt.r0,
                                                                                   g.rows, crow, Kc, d_order, d_out);
        SD_LAUNCH_CHECK(ctx);
    }
    return SD_OK;
}

int reconstruct_rows(int C, const sd_window* chunks, int64_t n_count, const sd_window* cf, int64_t* rows,
                     sd_window* frames_out) {
    const ReconGeom g = recon_geometry(C, chunks, n_count, cf);
    if (rows) *rows = g.rows;
    if (frames_out) *frames_out = sd_window{g.act.new_start, g.post.step, g.post.duration, 0};
    return SD_OK;
}

int reconstruct_launch(sd_ctx* ctx, const float* d_seg, int C, int F, int K, const sd_window* chunks, const int* d_hard,
                       int Kc, const int* d_count, int64_t n_count, const sd_window* cf, double* d_out,
                       int64_t cap_elems, int64_t* rows_out, sd_window* frames_out) {
    const ReconGeom g = recon_geometry(C, chunks, n_count, cf);
    if (rows_out) *rows_out = g.rows;
    if (frames_out) *frames_out = sd_window{g.act.new_start, g.post.step, g.post.duration, 0};
    if (g.rows * Kc > cap_elems)
        return ctx->fail(SD_ERR_CAPACITY, "sd_reconstruct: need %lld elements, have %lld", (long long)(g.rows * Kc),
                         (long long)cap_elems);
    const size_t n_cs = (size_t)C * F * Kc;
    double* d_cs = (double*)ctx->scratch(BUF_DZ_CS, sizeof(double) * n_cs);
    double* d_act = (double*)ctx->scratch(BUF_DZ_ACT, sizeof(double) * (size_t)g.NF * Kc);
    if (!d_cs || !d_act) return SD_ERR_NOMEM;
    clustered_scores_kernel<<<(unsigned)((n_cs + 255) / 256), 256, 0, ctx->stream>>>(d_seg, d_hard, C, F, K, Kc, d_cs);
    SD_LAUNCH_CHECK(ctx);
    const sd_window fr{0.0, cf->step, cf->duration, 0};
    int rc = aggregate_launch(ctx, d_cs, C, F, Kc, chunks, &fr, 0, 0.0, 1, DBL_EPSILON, d_act, g.NF, nullptr, nullptr);
    if (rc) return rc;
    if (g.rows > 0) {
        int* d_order = (int*)ctx->scratch(BUF_CL_MISC, sizeof(int) * (size_t)g.rows * Kc);
        if (!d_order) return SD_ERR_NOMEM;
        const long crow = g.crow < g.rows ? g.crow : g.rows;
        top_count_kernel<<<(unsigned)((g.rows + 127) / 128), 128, 0, ctx->stream>>>(d_act, g.act.r0, d_count, g.cnt.r0,
                                                                                   g.rows, crow, Kc, d_order, d_out);
        SD_LAUNCH_CHECK(ctx);
    }
    return SD_OK;
}

// ---------------------------------------------------------------------------------------------- to_annotation

// middle of frame i (SlidingWindow::operator[] + Segment::middle, SD:2876-2881)
__device__ __forceinline__ double frame_middle(double f_start, double f_step, double f_duration, long i) {
    const double s = __dadd_rn(f_start, __dmul_rn((double)i, f_step));
    return __ddiv_rn(__dadd_rn(s, __dadd_rn(s, f_duration)), 2.0);
}

// bit 0: v > onset, bit 1: v < offset; stored class-major so the per-class walk reads consecutive bytes
__global__ void __launch_bounds__(256)
    annot_flags_kernel(const double* __restrict__ scores, long rows, int cols, double onset, double offset,
                       unsigned char* __restrict__ flags) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long t = i / cols;
    const int k = (int)(i - t * cols);
    const double v = scores[i];
    flags[(size_t)k * rows + t] = (unsigned char)((v > onset ? 1 : 0) | (v < offset ? 2 : 0));
}

// A 2-state transition function packed in 2 bits: bit 0 = next state when inactive, bit 1 = next state when active.
__device__ __forceinline__ unsigned step_fn(unsigned char fl) {
    return ((fl & 1) ? 1u : 0u) | ((fl & 2) ? 0u : 2u);  // inactive -> on? ; active -> !off
}
__device__ __forceinline__ unsigned apply_fn(unsigned fn, unsigned state) { return (fn >> state) & 1u; }
// (g o f)(s) = g(f(s))
__device__ __forceinline__ unsigned compose_fn(unsigned f, unsigned g) {
    return apply_fn(g, apply_fn(f, 0)) | (apply_fn(g, apply_fn(f, 1)) << 1);
}

constexpr int kAnThreads = 1024;

// inclusive block scan of ints (1024 threads): warp shuffles + one pass over the 32 warp totals
__device__ __forceinline__ int an_scan_incl(int v, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += u;
        }
        warp_tot[lane] = t;
    }
    __syncthreads();
    if (warp > 0) v += warp_tot[warp - 1];
    if (total) *total = warp_tot[31];
    __syncthreads();  // warp_tot may be reused
    return v;
}

// One CTA per class.  Every thread owns a contiguous slice of frames: (1) compose the slice's transition
// function, block-scan the functions to get the state entering each slice, (2) count segment starts / ends,
// block-scan the counts, (3) emit (start, end) pairs in time order into list A, (4) support(): consecutive
// segments are disjoint and ordered, so "gap to the running segment < min_duration_off" only involves neighbours:
// chain heads are flagged, ranked with a block scan and the merged segments go to list B, (5) removeShort(): the
// kept segments are ranked and copied back to A.  (4) and (5) degenerate to copies when disabled.
__global__ void __launch_bounds__(kAnThreads)
    annot_runs_kernel(const unsigned char* __restrict__ flags, long rows, int cols, double f_start, double f_step,
                      double f_duration, double min_on, double min_off, double* __restrict__ lists,
                      double* __restrict__ lists_b, long maxseg, int* __restrict__ nseg) {
    __shared__ unsigned s_fn[kAnThreads];
    __shared__ int warp_tot[32];
    const int k = blockIdx.x, tid = threadIdx.x;
    const unsigned char* fl = flags + (size_t)k * rows;
    double* segs = lists + (size_t)k * maxseg * 2;
    double* segb = lists_b + (size_t)k * maxseg * 2;
    // frame 0 only sets the initial state (v > onset); slices cover frames 1..rows-1
    const long n = rows - 1;
    const long per = n > 0 ? (n + kAnThreads - 1) / kAnThreads : 1;
    const bool owns_last = n > 0 && tid == (int)((n - 1) / per);  // the slice holding frame rows-1
    const long lo = 1 + (long)tid * per, hi = min(rows, lo + per);
    unsigned fn = 2u;  // identity
    for (long t = lo; t < hi; ++t) fn = compose_fn(fn, step_fn(fl[t]));
    s_fn[tid] = fn;
    __syncthreads();
    for (int o = 1; o < kAnThreads; o <<= 1) {  // inclusive scan of function composition (earlier o later)
        unsigned mine = s_fn[tid], prev = tid >= o ? s_fn[tid - o] : 2u;
        __syncthreads();
        s_fn[tid] = compose_fn(prev, mine);
        __syncthreads();
    }
    const unsigned init = rows > 0 ? (fl[0] & 1u) : 0u;
    const unsigned state = tid == 0 ? init : apply_fn(s_fn[tid - 1], init);
    // a segment is counted where it ends (deactivation, or the last frame while active) and where it starts
    int ends = 0, starts = 0;
    {
        unsigned s = state;
        for (long t = lo; t < hi; ++t) {
            const unsigned ns = apply_fn(step_fn(fl[t]), s);
            ends += (s == 1u && ns == 0u);
            starts += (s == 0u && ns == 1u);
            s = ns;
        }
        if (owns_last && s == 1u) ++ends;  // still active at the last frame
        if (tid == 0 && init) ++starts;
    }
    if (n <= 0 && tid == 0) ends = init ? 1 : 0;  // a single frame: active -> (TS(0), TS(0))
    int total = 0;
    int w = an_scan_incl(ends, warp_tot, &total) - ends;
    int ws = an_scan_incl(starts, warp_tot, nullptr) - starts;
    {
        // the j-th start of the class pairs with the j-th end
        if (tid == 0 && init) segs[2 * (size_t)ws++] = frame_middle(f_start, f_step, f_duration, 0);
        unsigned s = state;
        for (long t = lo; t < hi; ++t) {
            const unsigned ns = apply_fn(step_fn(fl[t]), s);
            if (s == 1u && ns == 0u) segs[2 * (size_t)w++ + 1] = frame_middle(f_start, f_step, f_duration, t);
            if (s == 0u && ns == 1u) segs[2 * (size_t)ws++] = frame_middle(f_start, f_step, f_duration, t);
            s = ns;
        }
        if (owns_last && s == 1u) segs[2 * (size_t)w++ + 1] = frame_middle(f_start, f_step, f_duration, rows - 1);
    }
    if (n <= 0 && tid == 0 && init) segs[1] = frame_middle(f_start, f_step, f_duration, 0);
    __syncthreads();
    // ---- Track::support (SD:911-941): A -> B
    int cnt = total;
    {
        int carry = 0;
        for (int base = 0; base < cnt; base += kAnThreads) {
            const int i = base + tid;
            bool head = false, tail = false;
            if (i < cnt) {
                auto chained = [&](int a) {  // does raw segment a+1 extend the chain that contains raw segment a?
                    if (!(min_off > 0.0)) return false;
                    const double ce = segs[2 * a + 1], ns = segs[2 * (a + 1)];
                    const double gap = ce >= ns ? 0.0 : __dsub_rn(ns, ce);  // cur.start < next.start always holds
                    return gap < min_off;
                };
                head = i == 0 || !chained(i - 1);
                tail = i == cnt - 1 || !chained(i);
            }
            int chunk_total = 0;
            const int rank = carry + an_scan_incl(head ? 1 : 0, warp_tot, &chunk_total) - 1;  // chain index of i
            if (head) segb[2 * (size_t)rank] = segs[2 * i];
            if (tail) segb[2 * (size_t)rank + 1] = segs[2 * i + 1];
            carry += chunk_total;
        }
        cnt = carry;
    }
    __syncthreads();
    // ---- Track::removeShort never examines the first segment (SD:943-953): B -> A
    {
        int carry = 0;
        for (int base = 0; base < cnt; base += kAnThreads) {
            const int i = base + tid;
            bool keep = false;
            double s0 = 0.0, e0 = 0.0;
            if (i < cnt) {
                s0 = segb[2 * i];
                e0 = segb[2 * i + 1];
                keep = i == 0 || !(min_on > 0.0) || !(__dsub_rn(e0, s0) < min_on);
            }
            int chunk_total = 0;
            const int rank = carry + an_scan_incl(keep ? 1 : 0, warp_tot, &chunk_total) - 1;
            if (keep) {
                segs[2 * (size_t)rank] = s0;
                segs[2 * (size_t)rank + 1] = e0;
            }
            carry += chunk_total;
        }
        cnt = carry;
    }
    if (tid == 0) nseg[k] = cnt;
}

// finalResult (SD:962-978): all segments ordered by start.  Every class list is already sorted, so the final
// position of a segment is the number of segments that precede it: per other class a binary search (ties are
// broken by class index, i.e. a stable merge; std::sort leaves the order of equal starts unspecified).
__global__ void __launch_bounds__(256)
    annot_merge_kernel(const double* __restrict__ lists, long maxseg, const int* __restrict__ nseg, int cols,
                       double* __restrict__ seg_out, int* __restrict__ label_out, long cap, long* __restrict__ n_out) {
    long total = 0;
    for (int k = 0; k < cols; ++k) total += nseg[k];
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = total;
    if (total > cap) return;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        int k = 0;
        long idx = g;
        while (idx >= nseg[k]) idx -= nseg[k++];
        const double* mine = lists + (size_t)k * maxseg * 2;
        const double s = mine[2 * idx];
        long rank = idx;
        for (int q = 0; q < cols; ++q) {
            if (q == k) continue;
            const double* other = lists + (size_t)q * maxseg * 2;
            long lo = 0, hi = nseg[q];
            while (lo < hi) {  // count of starts < s (q > k) or <= s (q < k)
                const long mid = (lo + hi) >> 1;
                const double v = other[2 * mid];
                if (q < k ? (v <= s) : (v < s))
                    lo = mid + 1;
                else
                    hi = mid;
            }
            rank += lo;
        }
        seg_out[2 * rank] = s;
        seg_out[2 * rank + 1] = mine[2 * idx + 1];
        label_out[rank] = k;
    }
}

int to_annotation_launch(sd_ctx* ctx, const double* d_scores, int64_t rows, int cols, const sd_window* frames,
                         double onset, double offset, double min_on, double min_off, double* d_seg, int* d_label,
                         int64_t cap, long* d_n) {
    const long maxseg = rows / 2 + 2;
    unsigned char* d_flags = (unsigned char*)ctx->scratch(BUF_AN_FLAGS, (size_t)rows * cols);
    double* d_lists = (double*)ctx->scratch(BUF_AN_LISTS, sizeof(double) * 4 * (size_t)maxseg * cols);  // lists A and B
    double* d_lists_b = d_lists ? d_lists + 2 * (size_t)maxseg * cols : nullptr;
    int* d_nseg = (int*)ctx->scratch(BUF_AN_META, sizeof(int) * (size_t)cols);
    if (!d_flags || !d_lists || !d_nseg) return SD_ERR_NOMEM;
    const long total = (long)rows * cols;
    annot_flags_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_scores, (long)rows, cols, onset, offset,
                                                                               d_flags);
    SD_LAUNCH_CHECK(ctx);
    annot_runs_kernel<<<cols, kAnThreads, 0, ctx->stream>>>(d_flags, (long)rows, cols, frames->start, frames->step,
                                                            frames->duration, min_on, min_off, d_lists, d_lists_b, maxseg,
                                                            d_nseg);
    SD_LAUNCH_CHECK(ctx);
    annot_merge_kernel<<<32, 256, 0, ctx->stream>>>(d_lists, maxseg, d_nseg, cols, d_seg, d_label, (long)cap, d_n);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

}  // namespace sdb
