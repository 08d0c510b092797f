// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Builds the reference's own implementation of the hot path into
// oracle/_ref/libsdref.so.  The reference translation unit is compiled
// UNMODIFIED from where it lies (-I/root/reference/pipeline/src); nothing from
// it is copied into this repository.  This file only adds flat-buffer
// extern "C" entry points in front of the reference's nested-std::vector
// functions so that pytest (ctypes) and the C++ shim test can drive the
// reference and the CUDA library with identical inputs.
//
// Reference functions reached (file:line in /root/reference/pipeline/src):
//   SlidingWindow::closest_frame          speakerDiarizer.cpp:1084
//   Helper::np_rint                       speakerDiarizer.cpp:261
//   EmbeddingModel1::infer / _infer       speakerDiarizer.cpp:1977 / 1889
//   PipelineHelper::aggregate             speakerDiarizer.cpp:1167
//   SegmentModel::binarize_swf/_ndarray   speakerDiarizer.cpp:1506 / 1565
//   SegmentModel::trim / speaker_count    speakerDiarizer.cpp:1742 / 1665
//   Helper::cleanSegmentations            speakerDiarizer.cpp:710
//   Helper::normalizeEmbeddings           speakerDiarizer.cpp:344
//   Helper::cosineSimilarity              speakerDiarizer.cpp:502
//   Cluster::clustering / cluster         speakerDiarizer.cpp:2063 / 2300
//   Clustering::linkage/fcluster/cluster  clustering/clustering.cpp:417/442/459
//   Helper::interpolate / padSequence     speakerDiarizer.cpp:746 / 770
//   reconstruct / to_diarization          speakerDiarizer.cpp:2789 / 2638
//   to_annotation                         speakerDiarizer.cpp:2852
#define main ref_main
#include "speakerDiarizer.cpp"
#undef main

#include <streambuf>

namespace {

// The reference prints from inside its hot loops (speakerDiarizer.cpp:1252,
// 2414, 1918...).  Silence std::cout for the duration of a shim call.
struct Quiet {
    std::streambuf* old;
    Quiet() : old(std::cout.rdbuf(nullptr)) {}
    ~Quiet() {
        std::cout.rdbuf(old);
        std::cout.clear();
    }
};

template <typename T>
std::vector<std::vector<std::vector<T>>> to3d(const T* p, int a, int b, int c) {
    std::vector<std::vector<std::vector<T>>> v(a, std::vector<std::vector<T>>(b, std::vector<T>(c)));
    for (int i = 0; i < a; ++i)
        for (int j = 0; j < b; ++j)
            for (int k = 0; k < c; ++k) v[i][j][k] = p[((size_t)i * b + j) * c + k];
    return v;
}

template <typename T>
std::vector<std::vector<T>> to2d(const T* p, int a, int b) {
    std::vector<std::vector<T>> v(a, std::vector<T>(b));
    for (int i = 0; i < a; ++i)
        for (int j = 0; j < b; ++j) v[i][j] = p[(size_t)i * b + j];
    return v;
}

template <typename T, typename U>
void from3d(const std::vector<std::vector<std::vector<T>>>& v, U* p) {
    size_t n = 0;
    for (auto& a : v)
        for (auto& b : a)
            for (auto c : b) p[n++] = static_cast<U>(c);
}

template <typename T, typename U>
void from2d(const std::vector<std::vector<T>>& v, U* p) {
    size_t n = 0;
    for (auto& a : v)
        for (auto c : a) p[n++] = static_cast<U>(c);
}

// the OnnxModel constructor prints node names: build the instances under a scoped Quiet
SegmentModel& seg_model() {
    static SegmentModel* m = [] {
        Quiet q;
        return new SegmentModel("segment2.onnx");
    }();
    return *m;
}

EmbeddingModel1& emb_model() {
    static EmbeddingModel1* m = [] {
        Quiet q;
        return new EmbeddingModel1("emd4.onnx");
    }();
    return *m;
}

}  // namespace

// defined in the reference's clustering/clustering.cpp:408 (external linkage, C++ mangling)
double euclideanDistance(const std::vector<double>& vec1, const std::vector<double>& vec2);

extern "C" {

int ref_np_rint(double v) { return Helper::np_rint(v); }

long ref_closest_frame(double sw_start, double sw_step, double sw_duration, double t) {
    SlidingWindow sw(sw_start, sw_step, sw_duration);
    return (long)sw.closest_frame(t);
}

// EmbeddingModel1::infer -> captured ORT input.  out must hold 32*T*(n_fft/2+1)*2
// floats, wav_lens_out 32 floats, shape_out 4 int64.  Returns 0 on success.
int ref_stft(const float* wav, int B, int L, const float* lens, int n_lens, float* out, float* wav_lens_out,
             int64_t* shape_out) {
    Quiet q;
    auto data = to2d(wav, B, L);
    std::vector<float> l(lens, lens + n_lens);
    emb_model().infer(data, l);
    auto& cap = ort_stub::captured();
    if (cap.size() != 2) return 1;
    for (size_t i = 0; i < cap[0].shape.size() && i < 4; ++i) shape_out[i] = cap[0].shape[i];
    std::memcpy(out, cap[0].data.data(), cap[0].data.size() * sizeof(float));
    std::memcpy(wav_lens_out, cap[1].data.data(), cap[1].data.size() * sizeof(float));
    return 0;
}

// torch::stft alone, exactly as issued at speakerDiarizer.cpp:1997-2013 (fp64
// input, fp32 periodic Hamming window), bulk-copied out instead of the
// per-element .item<float>() loop.  Used to time "FFT only" and to validate
// long inputs where the as-written path would take minutes.  out: [B][T][F][2] fp32.
int ref_stft_fft_only(const float* wav, int B, int L, float* out) {
    std::vector<double> x((size_t)B * L);
    for (size_t i = 0; i < x.size(); ++i) x[i] = wav[i];
    torch::Tensor input = torch::from_blob(x.data(), {B, L}, torch::kFloat64);
    torch::Tensor window = torch::hamming_window(400);
    auto y = torch::stft(input, 400, 160, 400, window, true, "constant", false, true, false);
    auto s = y.transpose(2, 1).contiguous().to(torch::kFloat32);
    std::memcpy(out, s.data_ptr<float>(), (size_t)s.numel() * sizeof(float));
    return 0;
}

// PipelineHelper::aggregate.  post_out = {start, step, duration, num_samples}.
// Returns number of frames (rows of out), or -1 if cap is too small.
long ref_aggregate(const double* scores, int C, int F, int K, double sf_start, double sf_step, double sf_duration,
                   long sf_num_samples, double pf_step, double pf_duration, double missing, int skip_average,
                   double* out, long cap_rows, double* post_out) {
    Quiet q;
    auto s = to3d(scores, C, F, K);
    SlidingWindow sf(sf_start, sf_step, sf_duration, (size_t)sf_num_samples);
    SlidingWindow pf(0.0, pf_step, pf_duration);
    SlidingWindow post;
    auto r = PipelineHelper::aggregate(s, sf, pf, post, false, missing, skip_average != 0);
    if ((long)r.size() > cap_rows) return -1;
    from2d(r, out);
    post_out[0] = post.start;
    post_out[1] = post.step;
    post_out[2] = post.duration;
    post_out[3] = (double)post.num_samples;
    return (long)r.size();
}

int ref_binarize_swf(const float* scores, int C, int F, int K, int initial_state, double* out) {
    Quiet q;
    auto s = to3d(scores, C, F, K);
    auto r = seg_model().binarize_swf(s, initial_state != 0);
    from3d(r, out);
    return 0;
}

int ref_binarize_ndarray(const double* scores, int R, int F, double onset, int initial_state, unsigned char* out) {
    Quiet q;
    auto s = to2d(scores, R, F);
    auto r = seg_model().binarize_ndarray(s, onset, initial_state != 0);
    size_t n = 0;
    for (auto& a : r)
        for (bool b : a) out[n++] = b ? 1 : 0;
    return 0;
}

// SegmentModel::trim.  Returns trimmed frame count; tw_out = {start, step, duration, num_samples}.
long ref_trim(const double* binarized, int C, int F, int K, double left, double right, double bt_start,
              double bt_step, double bt_duration, double* out, double* tw_out) {
    Quiet q;
    auto b = to3d(binarized, C, F, K);
    SlidingWindow before(bt_start, bt_step, bt_duration), tw;
    auto r = seg_model().trim(b, left, right, before, tw);
    from3d(r, out);
    tw_out[0] = tw.start;
    tw_out[1] = tw.step;
    tw_out[2] = tw.duration;
    tw_out[3] = (double)tw.num_samples;
    return (long)r[0].size();
}

// SegmentModel::speaker_count.  Returns number of frames, -1 if cap too small.
long ref_speaker_count(const float* segmentations, const double* binarized, int C, int F, int K, double pf_start,
                       double pf_step, double pf_duration, int num_samples, int* out, long cap,
                       double* count_frames_out) {
    Quiet q;
    auto s = to3d(segmentations, C, F, K);
    auto b = to3d(binarized, C, F, K);
    SlidingWindow pre(pf_start, pf_step, pf_duration);
    SlidingWindow cf((size_t)num_samples);
    auto r = seg_model().speaker_count(s, b, pre, cf, num_samples);
    if ((long)r.size() > cap) return -1;
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    count_frames_out[0] = cf.start;
    count_frames_out[1] = cf.step;
    count_frames_out[2] = cf.duration;
    count_frames_out[3] = (double)cf.num_samples;
    return (long)r.size();
}

int ref_clean_segmentations(const double* binarized, int C, int F, int K, double* out) {
    auto b = to3d(binarized, C, F, K);
    auto r = Helper::cleanSegmentations(b);
    from3d(r, out);
    return 0;
}

int ref_normalize_embeddings(double* x, int N, int D) {
    auto v = to2d((const double*)x, N, D);
    Helper::normalizeEmbeddings(v);
    from2d(v, x);
    return 0;
}

// euclideanDistance as used by Clustering::linkage (clustering.cpp:408-431): condensed pdist.
int ref_pdist(const double* x, int N, int D, double* out) {
    auto v = to2d(x, N, D);
    size_t n = 0;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) out[n++] = euclideanDistance(v[i], v[j]);
    return 0;
}

int ref_linkage(const double* x, int N, int D, double* Z) {
    auto v = to2d(x, N, D);
    std::vector<std::vector<double>> z;
    Clustering::linkage(v, z);
    from2d(z, Z);
    return 0;
}

int ref_fcluster(const double* Z, int N, double cutoff, int* T) {
    auto z = to2d(Z, N - 1, 4);
    std::vector<int> t;
    Clustering::fcluster(z, cutoff, t);
    for (int i = 0; i < N; ++i) T[i] = t[i];
    return 0;
}

int ref_clustering_cluster(const double* x, int N, int D, double cutoff, int* T) {
    auto v = to2d(x, N, D);
    auto t = Clustering::cluster(v, cutoff);
    for (int i = 0; i < N; ++i) T[i] = t[i];
    return 0;
}

// Helper::cosineSimilarity (really cosine *distance*) A[na][D] x B[nb][D] -> out[na][nb].
// Returns 2 when the reference throws (zero magnitude / size mismatch).
int ref_cosine_cdist(const double* a, int na, const double* b, int nb, int D, double* out) {
    auto va = to2d(a, na, D);
    auto vb = to2d(b, nb, D);
    try {
        auto r = Helper::cosineSimilarity(va, vb);
        from2d(r, out);
    } catch (const std::runtime_error&) {
        return 2;
    }
    return 0;
}

// Cluster::cluster on already-filtered embeddings (labels after small->large reassignment).
int ref_cluster_labels(const double* x, int N, int D, int* labels) {
    Quiet q;
    auto v = to2d(x, N, D);
    Cluster c;
    int num_clusters = -1, min_clusters = -1, max_clusters = -1;
    c.set_num_clusters(N, num_clusters, min_clusters, max_clusters);
    try {
        auto r = c.cluster(v, min_clusters, max_clusters, num_clusters);
        for (int i = 0; i < N; ++i) labels[i] = r[i];
    } catch (const std::runtime_error&) {
        return 2;
    }
    return 0;
}

// Cluster::clustering on embeddings[C][S][D] (NaN rows = absent) -> hard[C][S]; then, when
// binarized != NULL, the inactive-speaker mask of speakerDiarizer.cpp:3166-3191 (-2).
int ref_clustering_stage(const double* emb, int C, int S, int D, const double* binarized, int F, int* hard) {
    Quiet q;
    auto e = to3d(emb, C, S, D);
    std::vector<std::vector<std::vector<double>>> segs;  // unused by the reference body
    std::vector<std::vector<int>> hc;
    Cluster c;
    try {
        c.clustering(e, segs, hc);
    } catch (const std::runtime_error&) {
        return 2;
    }
    if (binarized) {
        for (int i = 0; i < C; ++i)
            for (int k = 0; k < S; ++k) {
                float acc = 0.0f;  // reference accumulates in float (3172-3183)
                for (int j = 0; j < F; ++j) acc += binarized[((size_t)i * F + j) * S + k];
                if (abs(acc) < std::numeric_limits<double>::epsilon()) hc[i][k] = -2;
            }
    }
    from2d(hc, hard);
    return 0;
}

// ---- "next" rows ----

// Helper::interpolate + Helper::padSequence + wav_lens logic of getEmbedding (2436-2510).
// masks[B][F] float, wav[B][L] float -> signals[B][L], wav_lens[B] (normalised), too_short[B].
// Returns 1 when max_len < min_num_samples (reference returns all-NaN embeddings), else 0.
int ref_mask_compact(const float* wav, const float* masks, int B, int L, int F, float* signals, float* wav_lens,
                     unsigned char* too_short) {
    auto w = to2d(wav, B, L);
    auto m = to2d(masks, B, F);
    auto imasks = Helper::interpolate(m, L, 0.5);
    auto sig = Helper::padSequence(w, imasks);
    from2d(sig, signals);
    float max_len = 0;
    for (int i = 0; i < B; ++i) {
        float tmp = std::accumulate(imasks[i].begin(), imasks[i].end(), 0.0);
        wav_lens[i] = tmp;
        if (tmp > max_len) max_len = tmp;
    }
    if (max_len < min_num_samples) return 1;
    for (int i = 0; i < B; ++i) {
        if (wav_lens[i] < min_num_samples) {
            wav_lens[i] = 1.0;
            too_short[i] = 1;
        } else {
            wav_lens[i] /= max_len;
            too_short[i] = 0;
        }
    }
    return 0;
}

// reconstruct (2789) -> discrete diarization [rows][num_clusters]; returns rows, writes cols and frames.
long ref_reconstruct(const float* segmentations, int C, int F, int K, double sf_start, double sf_step,
                     double sf_duration, long sf_num_samples, const int* hard, const int* count, long n_count,
                     double cf_start, double cf_step, double cf_duration, long cf_num_samples, double* out,
                     long cap_elems, int* cols_out, double* frames_out) {
    Quiet q;
    auto s = to3d(segmentations, C, F, K);
    auto hc = to2d(hard, C, K);
    std::vector<int> cnt(count, count + n_count);
    SlidingWindow sf(sf_start, sf_step, sf_duration, (size_t)sf_num_samples);
    SlidingWindow cf(cf_start, cf_step, cf_duration, (size_t)cf_num_samples);
    SlidingWindow af;
    auto r = reconstruct(s, sf, hc, cnt, cf, af);
    if (r.empty()) return 0;
    if ((long)(r.size() * r[0].size()) > cap_elems) return -1;
    from2d(r, out);
    *cols_out = (int)r[0].size();
    frames_out[0] = af.start;
    frames_out[1] = af.step;
    frames_out[2] = af.duration;
    return (long)r.size();
}

// to_annotation (2852) + Annotation::finalResult (962).  Returns number of segments.
long ref_to_annotation(const double* scores, long rows, int cols, double f_start, double f_step, double f_duration,
                       double onset, double offset, double min_duration_on, double min_duration_off, double* seg_out,
                       int* label_out, long cap) {
    Quiet q;
    std::vector<std::vector<double>> sc(rows, std::vector<double>(cols));
    for (long i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) sc[i][j] = scores[i * cols + j];
    SlidingWindow fr(f_start, f_step, f_duration);
    // to_annotation returns by value through a move constructor that drops the
    // tracks (speakerDiarizer.cpp:1004-1007) unless the copy is elided, so build in place.
    Annotation ann = to_annotation(sc, fr, onset, offset, min_duration_on, min_duration_off);
    auto res = ann.finalResult();
    if ((long)res.size() > cap) return -1;
    for (size_t i = 0; i < res.size(); ++i) {
        seg_out[2 * i] = res[i].start;
        seg_out[2 * i + 1] = res[i].end;
        label_out[i] = res[i].label;
    }
    return (long)res.size();
}

// ---- f4: ingest ----

// wav::WavReader (frontend/wav.h:62-126) + the /32768 scaling of speakerDiarization() (2939-2951).
// meta = {num_channels, bits_per_sample, sample_rate}; returns num_samples (or -1 when cap is too small).
long ref_wav_load(const char* path, float* out, long cap, int* meta) {
    wav::WavReader wav_reader(path);
    meta[0] = wav_reader.num_channels();
    meta[1] = wav_reader.bits_per_sample();
    meta[2] = wav_reader.sample_rate();
    const float* audio = wav_reader.data();
    int num_samples = wav_reader.num_samples();
    if (num_samples > cap) return -1;
    std::vector<float> input_wav{audio, audio + num_samples};
    for (int i = 0; i < num_samples; ++i) input_wav[i] = input_wav[i] * 1.0f / 32768.0;
    std::copy(input_wav.begin(), input_wav.end(), out);
    return num_samples;
}

// SegmentModel::crop (1641-1662); the model's duration is 5.0 s.
long ref_crop(const float* wave, long n, double start, float* out) {
    SegmentModel* mm;
    {
        Quiet q;
        static SegmentModel model("segment2.onnx");
        mm = &model;
    }
    std::vector<float> w(wave, wave + n);
    auto c = mm->crop(w, std::make_pair(start, start + 5.0));
    std::copy(c.begin(), c.end(), out);
    return (long)c.size();
}

}  // extern "C"
