// sdb200_host.hpp -- C++ host shim: the reference's function-level API on top of the libsdb200 C-ABI.
//
// The reference (leohuang2013/pyannote-audio_speaker-diarization_cpp) has no plugin interface; its hot path is
// a set of C++ functions taking and returning nested std::vector by value.  Every function below
// keeps the name, argument order, defaults and error behaviour of the reference function it replaces
// (file:line relative to the reference checkout, SD = pipeline/src/speakerDiarizer.cpp,
// CL = pipeline/src/clustering/clustering.cpp), converts nested vectors <-> flat row-major buffers and calls the
// C-ABI.  There is no CPU implementation behind it: without a CUDA device every call throws.
//
// `SlidingWindow`-typed parameters are templates: any type with public members start, step, duration,
// num_samples works, in particular the reference's own class (SD:1029), so the shim can be included in the
// reference translation unit and called from speakerDiarization() (SD:2937) unchanged -- see INTEGRATION.md.
//
// Errors: the reference asserts or throws std::runtime_error; the shim throws std::runtime_error carrying
// sd_last_error().  SD_ERR_ZERO_MAGNITUDE maps to the reference's message "Vectors have zero magnitude.".
//
// Stage dumps.  Built with WRITE_DATA, the reference writes /tmp/cpp_<stage>.txt files from INSIDE several of the
// bodies replaced here (SD:1271-1275, 1627-1636, 1696-1723, 2074, 2186, 2206, 2330-2331, 2453-2454, 2491, 2654,
// 2716-2717, 2732, 2841; consumer pipeline/script/verifyEveryStepResult.py).  To keep them, include this header
// AFTER the reference's debugWrite* templates (SD:87-245) with
//     #define SDB200_DUMP1(data, ...) debugWrite(data, __VA_ARGS__)
//     #define SDB200_DUMP2(data, ...) debugWrite2d(data, __VA_ARGS__)
//     #define SDB200_DUMP3(data, ...) debugWrite3d(data, __VA_ARGS__)
// (tests/dropin/patch_reference.py does exactly that).  The shim then takes the decomposed route -- the same call
// structure as the reference body, every piece computed by a libsdb200 call -- and hands each intermediate to the
// reference's own writer, so the files are byte-identical.  Without the macros the fused calls are used.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <utility>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sdb200.h"

#if defined(SDB200_DUMP1) && defined(SDB200_DUMP2) && defined(SDB200_DUMP3)
#define SDB200_STAGE_DUMPS 1
#else
#define SDB200_STAGE_DUMPS 0
#endif

namespace sdb200 {

using vec1f = std::vector<float>;
using vec2f = std::vector<std::vector<float>>;
using vec3f = std::vector<std::vector<std::vector<float>>>;
using vec4f = std::vector<std::vector<std::vector<std::vector<float>>>>;
using vec1d = std::vector<double>;
using vec2d = std::vector<std::vector<double>>;
using vec3d = std::vector<std::vector<std::vector<double>>>;

// One context per host thread (the reference is single-threaded and non-re-entrant).
class Context {
public:
    explicit Context(int device = 0) {
        if (sd_ctx_create(device, &ctx_) != SD_OK)
            throw std::runtime_error("sdb200: no usable CUDA device (this library has no CPU fallback)");
    }
    ~Context() { sd_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    sd_ctx* get() const { return ctx_; }
    void check(int rc) const {
        if (rc == SD_OK) return;
        if (rc == SD_ERR_ZERO_MAGNITUDE) throw std::runtime_error("Vectors have zero magnitude.");  // SD:494
        throw std::runtime_error(std::string("sdb200: ") + sd_last_error(ctx_));
    }

private:
    sd_ctx* ctx_ = nullptr;
};

inline Context& context() {
    static thread_local Context c(0);
    return c;
}

namespace detail {
template <typename T, typename U>
std::vector<U> flatten3(const std::vector<std::vector<std::vector<T>>>& v) {
    std::vector<U> out;
    if (v.empty() || v[0].empty()) return out;
    out.reserve(v.size() * v[0].size() * v[0][0].size());
    for (const auto& a : v)
        for (const auto& b : a)
            for (T c : b) out.push_back(static_cast<U>(c));
    return out;
}
template <typename T, typename U>
std::vector<U> flatten2(const std::vector<std::vector<T>>& v) {
    std::vector<U> out;
    if (v.empty()) return out;
    out.reserve(v.size() * v[0].size());
    for (const auto& a : v)
        for (T c : a) out.push_back(static_cast<U>(c));
    return out;
}
template <typename U, typename T>
std::vector<std::vector<std::vector<T>>> unflatten3(const std::vector<U>& f, size_t a, size_t b, size_t c) {
    std::vector<std::vector<std::vector<T>>> v(a, std::vector<std::vector<T>>(b, std::vector<T>(c)));
    size_t n = 0;
    for (size_t i = 0; i < a; ++i)
        for (size_t j = 0; j < b; ++j)
            for (size_t k = 0; k < c; ++k) v[i][j][k] = static_cast<T>(f[n++]);
    return v;
}
template <typename U, typename T>
std::vector<std::vector<T>> unflatten2(const std::vector<U>& f, size_t a, size_t b) {
    std::vector<std::vector<T>> v(a, std::vector<T>(b));
    size_t n = 0;
    for (size_t i = 0; i < a; ++i)
        for (size_t j = 0; j < b; ++j) v[i][j] = static_cast<T>(f[n++]);
    return v;
}
template <class SW>
sd_window to_window(const SW& w) {
    sd_window r;
    r.start = w.start;
    r.step = w.step;
    r.duration = w.duration;
    r.num_samples = static_cast<int64_t>(w.num_samples);
    return r;
}
template <class SW>
void from_window(const sd_window& r, SW& w) {
    w.start = r.start;
    w.step = r.step;
    w.duration = r.duration;
    w.num_samples = static_cast<decltype(w.num_samples)>(r.num_samples);
}
}  // namespace detail

// ---------------------------------------------------------------------------------------------------
// Embedding stage front-end
// ---------------------------------------------------------------------------------------------------

// What EmbeddingModel1::infer (SD:1977-2036) + the packing of _infer (SD:1889-1917) hand to emd4.onnx:
// `audio` is the flat [32][T][201][2] tensor (rows beyond data.size() are zero), `wav_lens` the 32 relative
// lengths (1.0 beyond lens.size()), `dims` the four tensor dimensions for Ort::Value::CreateTensor.
struct EmbeddingInput {
    std::vector<float> audio;
    std::vector<float> wav_lens;
    int64_t dims[4];
};

inline EmbeddingInput embedding_input(const vec2f& data, const vec1f& lens, int batch_size = 32) {
    Context& c = context();
    const int B = static_cast<int>(data.size());
    const int L = static_cast<int>(data[0].size());
    sd_stft_params p;
    sd_stft_default_params(&p);
    p.pad_batch_to = batch_size;
    const int64_t T = sd_stft_num_frames(L, p.hop);
    const int rows = B > batch_size ? B : batch_size;
    EmbeddingInput out;
    out.audio.resize(static_cast<size_t>(rows) * T * (p.n_fft / 2 + 1) * 2);
    out.wav_lens.resize(batch_size);
    std::vector<float> flat = detail::flatten2<float, float>(data);
    c.check(sd_stft(c.get(), flat.data(), B, L, &p, out.audio.data()));
    if (sd_pack_wav_lens(lens.data(), static_cast<int>(lens.size()), batch_size, out.wav_lens.data()) != SD_OK)
        throw std::runtime_error("sdb200: more wav_lens than batch rows");
    out.dims[0] = rows;
    out.dims[1] = T;
    out.dims[2] = p.n_fft / 2 + 1;
    out.dims[3] = 2;
    return out;
}

// The rest of EmbeddingModel1::_infer (SD:1930-1975) on that input: build the two tensors, run the session, unpack
// `rows` embeddings.  Templated on the ONNX Runtime types so that this header needs no ORT include; with it the body
// of EmbeddingModel1::infer (SD:1977) becomes
//     auto in = sdb200::embedding_input( data, lens, m_batchSize );
//     return sdb200::run_embedding_model<Ort::Value, Ort::RunOptions>( *session_, memory_info_, input_node_names_,
//                                                                      output_node_names_, in, data.size());
// and the signature, the batch-of-32 padding and the output layout stay exactly the reference's.
template <class Value, class RunOptions, class Session, class MemoryInfo>
vec2f run_embedding_model(Session& session, MemoryInfo& memory_info, const std::vector<const char*>& input_names,
                          const std::vector<const char*>& output_names, EmbeddingInput& in, size_t rows) {
    std::vector<Value> inputs;
    inputs.emplace_back(Value::template CreateTensor<float>(memory_info, in.audio.data(), in.audio.size(), in.dims, 4));
    int64_t lens_dims[1] = {static_cast<int64_t>(in.wav_lens.size())};
    inputs.emplace_back(
        Value::template CreateTensor<float>(memory_info, in.wav_lens.data(), in.wav_lens.size(), lens_dims, 1));
    auto outputs = session.Run(RunOptions{nullptr}, input_names.data(), inputs.data(), inputs.size(),
                               output_names.data(), output_names.size());
    const float* o = outputs[0].template GetTensorData<float>();
    const auto shape = outputs[0].GetTensorTypeAndShapeInfo().GetShape();
    const size_t dim = static_cast<size_t>(shape[2]);  // [batch][1][192]
    vec2f res(rows, vec1f(dim));
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < dim; ++j) res[i][j] = o[i * dim + j];
    return res;
}

// The 4-D vector the reference builds at SD:2018-2036 ([B][T][201][2]); kept for callers that want that shape.
inline vec4f stft(const vec2f& data) {
    Context& c = context();
    const int B = static_cast<int>(data.size());
    const int L = static_cast<int>(data[0].size());
    sd_stft_params p;
    sd_stft_default_params(&p);
    const int64_t T = sd_stft_num_frames(L, p.hop);
    const int bins = p.n_fft / 2 + 1;
    std::vector<float> flat = detail::flatten2<float, float>(data), out(static_cast<size_t>(B) * T * bins * 2);
    c.check(sd_stft(c.get(), flat.data(), B, L, &p, out.data()));
    vec4f r(B, vec3f(T, vec2f(bins, vec1f(2))));
    size_t n = 0;
    for (int b = 0; b < B; ++b)
        for (int64_t t = 0; t < T; ++t)
            for (int f = 0; f < bins; ++f) {
                r[b][t][f][0] = out[n++];
                r[b][t][f][1] = out[n++];
            }
    return r;
}

// ---------------------------------------------------------------------------------------------------
// PipelineHelper::aggregate, SD:1167-1311
// ---------------------------------------------------------------------------------------------------
template <class SW>
vec2d aggregate(const vec3d& scoreData, const SW& scores_frames, const SW& pre_frames, SW& post_frames,
                bool hamming = false, double missing = NAN, bool skip_average = false,
                double epsilon = std::numeric_limits<double>::epsilon()) {
    Context& c = context();
    const int C = static_cast<int>(scoreData.size());
    const int F = static_cast<int>(scoreData[0].size());
    const int K = static_cast<int>(scoreData[0][0].size());
    const sd_window cw = detail::to_window(scores_frames), fw = detail::to_window(pre_frames);
    const int64_t NF = sd_aggregate_num_frames(C, &cw, &fw);
    std::vector<double> flat = detail::flatten3<double, double>(scoreData), out(static_cast<size_t>(NF) * K);
    int64_t n = 0;
    sd_window post;
#if SDB200_STAGE_DUMPS
    {   // SD:1271-1275: masks / scores (NaN -> 0), and the sums before the division and the `missing` fill
        vec3d masks(C, vec2d(F, vec1d(K, 1.0))), scores = scoreData;
        for (int i = 0; i < C; ++i)
            for (int j = 0; j < F; ++j)
                for (int k = 0; k < K; ++k)
                    if (std::isnan(scoreData[i][j][k])) masks[i][j][k] = scores[i][j][k] = 0.0;
        std::vector<double> sum(out.size()), cnt(out.size()), msk(out.size());
        c.check(sd_aggregate(c.get(), flat.data(), C, F, K, &cw, &fw, hamming ? 1 : 0, 0.0, 1, epsilon, sum.data(), NF,
                             &n, &post, cnt.data(), msk.data()));
        SDB200_DUMP3(masks, "cpp_masks_in_aggregate");
        SDB200_DUMP3(scores, "cpp_scores_in_aggregate");
        SDB200_DUMP2((detail::unflatten2<double, double>(sum, NF, K)), "cpp_aggregated_output");
        SDB200_DUMP2((detail::unflatten2<double, double>(msk, NF, K)), "cpp_aggregated_mask");
        SDB200_DUMP2((detail::unflatten2<double, double>(cnt, NF, K)), "cpp_overlapping_chunk_count");
    }
#endif
    c.check(sd_aggregate(c.get(), flat.data(), C, F, K, &cw, &fw, hamming ? 1 : 0, missing, skip_average ? 1 : 0,
                         epsilon, out.data(), NF, &n, &post, nullptr, nullptr));
    detail::from_window(post, post_frames);
    return detail::unflatten2<double, double>(out, static_cast<size_t>(NF), static_cast<size_t>(K));
}

// ---------------------------------------------------------------------------------------------------
// SegmentModel::binarize_swf / binarize_ndarray / trim / speaker_count, Helper::cleanSegmentations
// ---------------------------------------------------------------------------------------------------
constexpr double kOnset = 0.4442333667381752;  // SegmentModel::m_diarization_segmentation_threashold, SD:1339

inline std::vector<std::vector<bool>> binarize_ndarray(const vec2d& scores, double onset = 0.5,
                                                       bool initialState = false);

inline vec3d binarize_swf(const vec3f& scores, bool initial_state = false, double onset = kOnset) {  // SD:1506
    Context& c = context();
    const int C = static_cast<int>(scores.size()), F = static_cast<int>(scores[0].size()),
              K = static_cast<int>(scores[0][0].size());
#if SDB200_STAGE_DUMPS
    {   // the reference's route, "c f k -> (c k) f" / binarize_ndarray / back (SD:1512-1562), so that its dumps appear
        vec2d rows(static_cast<size_t>(C) * K, vec1d(F));
        for (int i = 0; i < C; ++i)
            for (int j = 0; j < F; ++j)
                for (int k = 0; k < K; ++k) rows[static_cast<size_t>(i) * K + k][j] = scores[i][j][k];
        const auto b = binarize_ndarray(rows, onset, initial_state);
        vec3d r(C, vec2d(F, vec1d(K)));
        for (int i = 0; i < C; ++i)
            for (int j = 0; j < F; ++j)
                for (int k = 0; k < K; ++k) r[i][j][k] = b[static_cast<size_t>(i) * K + k][j];
        return r;
    }
#endif
    std::vector<float> flat = detail::flatten3<float, float>(scores);
    std::vector<double> out(flat.size());
    c.check(sd_binarize(c.get(), flat.data(), C, F, K, onset, initial_state ? 1 : 0, out.data()));
    return detail::unflatten3<double, double>(out, C, F, K);
}

// SD:1565
inline std::vector<std::vector<bool>> binarize_ndarray(const vec2d& scores, double onset, bool initialState) {
    Context& c = context();
    const int R = static_cast<int>(scores.size()), F = static_cast<int>(scores[0].size());
    std::vector<double> flat = detail::flatten2<double, double>(scores);
    std::vector<uint8_t> out(flat.size());
    c.check(sd_binarize_rows(c.get(), flat.data(), R, F, onset, initialState ? 1 : 0, out.data()));
    std::vector<std::vector<bool>> r(R, std::vector<bool>(F));
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < F; ++j) r[i][j] = out[static_cast<size_t>(i) * F + j] != 0;
#if SDB200_STAGE_DUMPS
    {   // SD:1626-1636
        std::vector<uint8_t> on(flat.size());
        std::vector<int32_t> same(flat.size()), wdi(flat.size());
        int cols = 0;
        c.check(sd_binarize_rows_stages(c.get(), flat.data(), R, F, onset, on.data(), same.data(), wdi.data(), &cols));
        std::vector<std::vector<bool>> on2(R, std::vector<bool>(F)), init(R, std::vector<bool>(F, initialState));
        std::vector<std::vector<int>> same2(R, std::vector<int>(F)), wdi2(R, std::vector<int>(cols)),
            samples(R, std::vector<int>(F));
        for (int i = 0; i < R; ++i) {
            for (int j = 0; j < F; ++j) {
                on2[i][j] = on[static_cast<size_t>(i) * F + j] != 0;
                same2[i][j] = same[static_cast<size_t>(i) * F + j];
                samples[i][j] = i;
            }
            for (int j = 0; j < cols; ++j) wdi2[i][j] = wdi[static_cast<size_t>(i) * F + j];
        }
        SDB200_DUMP2(scores, "cpp_binarize_score");
        SDB200_DUMP2(same2, "cpp_same_as");
        SDB200_DUMP2(on2, "cpp_on");
        SDB200_DUMP2(wdi2, "cpp_well_defined_idx");
        SDB200_DUMP2(init, "cpp_initial_state");
        SDB200_DUMP2(samples, "cpp_samples");
        SDB200_DUMP2(r, "cpp_binary_ndarray");
    }
#endif
    return r;
}

template <class SW>
vec3d trim(const vec3d& binarized, double left, double right, const SW& before_trim, SW& trimmed_frames) {  // SD:1742
    Context& c = context();
    const int C = static_cast<int>(binarized.size()), F = static_cast<int>(binarized[0].size()),
              K = static_cast<int>(binarized[0][0].size());
    const int64_t Ft = sd_trim_num_frames(F, left, right);
    std::vector<double> flat = detail::flatten3<double, double>(binarized), out(static_cast<size_t>(C) * Ft * K);
    const sd_window bw = detail::to_window(before_trim);
    sd_window tw;
    c.check(sd_trim(c.get(), flat.data(), C, F, K, left, right, &bw, out.data(), &tw));
    detail::from_window(tw, trimmed_frames);
    return detail::unflatten3<double, double>(out, C, static_cast<size_t>(Ft), K);
}

// SD:1665-1738.  `segmentations` and `num_samples` are unused by the reference body as well; the chunk window is
// SegmentModel's (0.0, m_step = 0.5, m_duration = 5.0) unless overridden.
template <class SW>
std::vector<int> speaker_count(const vec3f& /*segmentations*/, const vec3d& binarized, const SW& pre_frame,
                               SW& count_frames, int /*num_samples*/, double chunk_step = 0.5,
                               double chunk_duration = 5.0) {
    Context& c = context();
    const int C = static_cast<int>(binarized.size()), F = static_cast<int>(binarized[0].size()),
              K = static_cast<int>(binarized[0][0].size());
#if SDB200_STAGE_DUMPS
    {   // the reference's route trim -> sum over classes -> aggregate -> np.rint with its dumps (SD:1688-1737)
        SW trimmed_frames = pre_frame, chunk_frames = pre_frame;
        chunk_frames.start = 0.0;
        chunk_frames.step = chunk_step;
        chunk_frames.duration = chunk_duration;
        const vec3d trimmed = trim(binarized, 0.1, 0.1, chunk_frames, trimmed_frames);
        SDB200_DUMP3(trimmed, "cpp_trimmed");
        const size_t Ft = trimmed[0].size();
        std::vector<double> flat = detail::flatten3<double, double>(binarized), sum(static_cast<size_t>(C) * Ft);
        c.check(sd_trim_sum(c.get(), flat.data(), C, F, K, 0.1, 0.1, sum.data()));
        const vec3d sum_trimmed = detail::unflatten3<double, double>(sum, C, Ft, 1);
        SDB200_DUMP3(sum_trimmed, "cpp_sum_trimmed");
        const vec2d count_data = aggregate(sum_trimmed, trimmed_frames, pre_frame, count_frames, false, 0.0, false);
        SDB200_DUMP2(count_data, "cpp_count_data");
        std::vector<int> res(count_data.size());
        for (size_t i = 0; i < res.size(); ++i) res[i] = sd_np_rint(count_data[i][0]);
        return res;
    }
#endif
    sd_window cw;
    cw.start = 0.0;
    cw.step = chunk_step;
    cw.duration = chunk_duration;
    cw.num_samples = 1;
    const sd_window fw = detail::to_window(pre_frame);
    std::vector<double> flat = detail::flatten3<double, double>(binarized);
    const int64_t cap = static_cast<int64_t>((C * chunk_step + chunk_duration) / fw.step) + F + 64;
    std::vector<int32_t> out(static_cast<size_t>(cap));
    int64_t n = 0;
    sd_window cf;
    c.check(sd_speaker_count(c.get(), flat.data(), C, F, K, &cw, &fw, out.data(), cap, &n, &cf));
    detail::from_window(cf, count_frames);
    return std::vector<int>(out.begin(), out.begin() + n);
}

inline vec3d cleanSegmentations(const vec3d& data) {  // Helper::cleanSegmentations, SD:710
    Context& c = context();
    const int C = static_cast<int>(data.size()), F = static_cast<int>(data[0].size()),
              K = static_cast<int>(data[0][0].size());
    std::vector<double> flat = detail::flatten3<double, double>(data), out(flat.size());
    c.check(sd_clean_segmentations(c.get(), flat.data(), C, F, K, out.data()));
    return detail::unflatten3<double, double>(out, C, F, K);
}

// ---------------------------------------------------------------------------------------------------
// Clustering library (clustering.h:4-12) and Helper::normalizeEmbeddings (SD:344)
// ---------------------------------------------------------------------------------------------------
inline void normalizeEmbeddings(vec2d& embeddings) {
    Context& c = context();
    const int N = static_cast<int>(embeddings.size()), D = static_cast<int>(embeddings[0].size());
    std::vector<double> flat = detail::flatten2<double, double>(embeddings);
    c.check(sd_normalize(c.get(), flat.data(), N, D));
    embeddings = detail::unflatten2<double, double>(flat, N, D);
}

struct Clustering {
    static void linkage(const vec2d& input, vec2d& dendrogram) {  // CL:417
        Context& c = context();
        const int N = static_cast<int>(input.size()), D = static_cast<int>(input[0].size());
        std::vector<double> flat = detail::flatten2<double, double>(input), Z(static_cast<size_t>(N - 1) * 4);
        c.check(sd_linkage(c.get(), flat.data(), N, D, Z.data()));
        dendrogram = detail::unflatten2<double, double>(Z, static_cast<size_t>(N - 1), 4);
    }
    static void fcluster(const vec2d& Z, double cutoff, std::vector<int>& clusters) {  // CL:442
        Context& c = context();
        const int N = static_cast<int>(Z.size()) + 1;
        std::vector<double> flat = detail::flatten2<double, double>(Z);
        std::vector<int32_t> T(static_cast<size_t>(N));
        c.check(sd_fcluster(c.get(), flat.data(), N, cutoff, T.data()));
        clusters.assign(T.begin(), T.end());
    }
    static std::vector<int> cluster(const vec2d& input, double cutoff) {  // CL:459
        Context& c = context();
        const int N = static_cast<int>(input.size()), D = static_cast<int>(input[0].size());
        std::vector<double> flat = detail::flatten2<double, double>(input);
        std::vector<int32_t> T(static_cast<size_t>(N));
        c.check(sd_cluster(c.get(), flat.data(), N, D, cutoff, T.data()));
        return std::vector<int>(T.begin(), T.end());
    }
};

// Helper::cosineSimilarity (cosine *distance*), SD:502
inline vec2d cosineSimilarity(const vec2d& a, const vec2d& b) {
    Context& c = context();
    const int na = static_cast<int>(a.size()), nb = static_cast<int>(b.size()), D = static_cast<int>(a[0].size());
    if (b[0].size() != a[0].size()) throw std::runtime_error("Vector sizes must be equal.");  // SD:480
    std::vector<double> fa = detail::flatten2<double, double>(a), fb = detail::flatten2<double, double>(b),
                        out(static_cast<size_t>(na) * nb);
    c.check(sd_cosine_cdist(c.get(), fa.data(), na, fb.data(), nb, D, out.data()));
    return detail::unflatten2<double, double>(out, na, nb);
}

// ---------------------------------------------------------------------------------------------------
// Cluster (SD:2044-2425)
// ---------------------------------------------------------------------------------------------------
class Cluster {
public:
    // Cluster::clustering, SD:2063.  `segmentations` is unused by the reference body too.  When `binarized` is
    // given, the inactive-speaker pass of speakerDiarization() (SD:3166-3191, hard = -2) is applied as well.
    void clustering(const vec3d& embeddings, const vec3d& /*segmentations*/, std::vector<std::vector<int>>& hard_clusters,
                    int num_clusters = -1, int min_clusters = -1, int max_clusters = -1,
                    const vec3d* binarized = nullptr) {
        Context& c = context();
        const int C = static_cast<int>(embeddings.size()), S = static_cast<int>(embeddings[0].size()),
                  D = static_cast<int>(embeddings[0][0].size());
        sd_cluster_params p;
        sd_cluster_default_params(&p);
        p.num_clusters = num_clusters;
        p.min_clusters = min_clusters;
        p.max_clusters = max_clusters;
        std::vector<double> flat = detail::flatten3<double, double>(embeddings), bin;
        int F = 0;
        if (binarized) {
            bin = detail::flatten3<double, double>(*binarized);
            F = static_cast<int>((*binarized)[0].size());
        }
        std::vector<int32_t> hard(static_cast<size_t>(C) * S);
        int k = 0;
#if SDB200_STAGE_DUMPS
        bool assigned = true;
        {   // SD:2073-2075, and the dumps of cluster() (SD:2329-2332) which the reference reaches from here
            vec2d filtered;
            for (const auto& chunk : embeddings)
                for (const auto& e : chunk)
                    if (!std::isnan(e[0])) filtered.push_back(e);
            SDB200_DUMP2(filtered, "cpp_filtered_embeddings");
            const int n = static_cast<int>(filtered.size());
            int nc = num_clusters, lo = min_clusters, hi = max_clusters;  // set_num_clusters, SD:2261-2296
            if (nc != -1) lo = nc; else if (lo == -1) lo = 1;
            lo = std::max(1, std::min(n, lo));
            if (nc == -1 && hi == -1) hi = n;  // (`max_clusters == num_clusters;` at SD:2278 is a no-op comparison)
            hi = std::max(1, std::min(n, hi));
            if (lo > hi) lo = hi;
            if (lo == hi) nc = lo;
            assigned = hi >= 2;  // SD:2081-2088: otherwise everything is cluster 0 and assign_embeddings is skipped
            if (assigned) cluster(filtered, lo, hi, nc);
        }
#endif
        c.check(sd_clustering(c.get(), flat.data(), C, S, D, &p, binarized ? bin.data() : nullptr, F, hard.data(),
                              nullptr, 0, &k));
#if SDB200_STAGE_DUMPS
        if (assigned && k > 0) {  // SD:2185-2207 (assign_embeddings): distances to the k centroids, soft = 2 - dist
            std::vector<double> soft(static_cast<size_t>(C) * S * k), dist(soft.size());
            c.check(sd_clustering_ex(c.get(), flat.data(), C, S, D, &p, binarized ? bin.data() : nullptr, F, hard.data(),
                                     soft.data(), dist.data(), k, &k));
            SDB200_DUMP2((detail::unflatten2<double, double>(dist, static_cast<size_t>(C) * S, k)), "cpp_dist", true);
            SDB200_DUMP3((detail::unflatten3<double, double>(soft, C, S, k)), "cpp_soft_clusters", true);
        }
#endif
        hard_clusters = detail::unflatten2<int32_t, int>(hard, C, S);
    }

    // Cluster::cluster, SD:2300 (labels of already-filtered embeddings)
    std::vector<int> cluster(const vec2d& embeddings, int min_clusters, int max_clusters, int num_clusters) {
        Context& c = context();
        const int N = static_cast<int>(embeddings.size()), D = static_cast<int>(embeddings[0].size());
        sd_cluster_params p;
        sd_cluster_default_params(&p);
        p.num_clusters = num_clusters;
        p.min_clusters = min_clusters;
        p.max_clusters = max_clusters;
        std::vector<double> flat = detail::flatten2<double, double>(embeddings);
        std::vector<int32_t> lab(static_cast<size_t>(N));
#if SDB200_STAGE_DUMPS
        {   // SD:2319-2332: normalised copy and the raw fcluster labels - 1
            vec2d normalized = embeddings;
            normalizeEmbeddings(normalized);
            std::vector<int> clusters = Clustering::cluster(normalized, static_cast<double>(p.threshold));
            for (auto& v : clusters) v -= 1;
            SDB200_DUMP2(normalized, "cpp_norm_embeddings");
            SDB200_DUMP1(clusters, "cpp_clusters");
        }
#endif
        c.check(sd_cluster_labels(c.get(), flat.data(), N, D, &p, lab.data()));
        return std::vector<int>(lab.begin(), lab.end());
    }
};

// ---------------------------------------------------------------------------------------------------
// Rows either side of the hot path
// ---------------------------------------------------------------------------------------------------

// The masking prologue of getEmbedding (SD:2466-2510): Helper::interpolate(masks, num_samples, 0.5) (SD:746),
// Helper::padSequence(waveforms, imasks) (SD:770), wav_lens = imasks.sum / max and the too-short bookkeeping.
// Returns false when even the longest item is shorter than min_num_samples -- the reference then fills the
// embeddings with NaN and skips the model (SD:2479-2486).  `too_short[i]` marks items whose embedding the
// reference overwrites with NaN after inference (SD:2545-2556).
// `dump_number`: the running batch number getEmbedding appends to its dump names (SD:2442-2454, 2489-2491).
inline bool masked_signals(const vec2f& waveforms, const vec2f& masks, int min_num_samples, vec2f& signals,
                           vec1f& wav_lens, std::vector<bool>& too_short, int dump_number = 0) {
    Context& c = context();
    const int B = static_cast<int>(waveforms.size()), L = static_cast<int>(waveforms[0].size()),
              F = static_cast<int>(masks[0].size());
    std::vector<float> w = detail::flatten2<float, float>(waveforms), m = detail::flatten2<float, float>(masks);
    std::vector<float> sig(w.size());
    std::vector<uint8_t> ts(static_cast<size_t>(B));
    wav_lens.assign(static_cast<size_t>(B), 0.f);
    int all_short = 0;
    c.check(sd_mask_compact(c.get(), w.data(), m.data(), B, L, F, min_num_samples, sig.data(), wav_lens.data(), ts.data(),
                            &all_short));
#if SDB200_STAGE_DUMPS
    {   // SD:2452-2455 (masks, imasks) and, unless the whole batch is too short, SD:2488-2492 (raw wav_lens)
        std::vector<uint8_t> im(w.size());
        std::vector<int32_t> counts(static_cast<size_t>(B));
        c.check(sd_mask_interpolate(c.get(), m.data(), B, F, L, 0.5f, im.data(), counts.data()));
        std::vector<std::vector<bool>> imasks(B, std::vector<bool>(L));
        for (int i = 0; i < B; ++i)
            for (int j = 0; j < L; ++j) imasks[i][j] = im[static_cast<size_t>(i) * L + j] != 0;
        SDB200_DUMP2(masks, std::string("cpp_masks") + std::to_string(dump_number), true);
        SDB200_DUMP2(imasks, std::string("cpp_imasks") + std::to_string(dump_number));
        if (!all_short) {
            vec1f raw(counts.begin(), counts.end());
            SDB200_DUMP1(raw, std::string("cpp_wav_lens") + std::to_string(dump_number));
        }
    }
#else
    (void)dump_number;
#endif
    signals = detail::unflatten2<float, float>(sig, static_cast<size_t>(B), static_cast<size_t>(L));
    too_short.assign(ts.begin(), ts.end());
    return all_short == 0;
}

// reconstruct (SD:2789-2848), including to_diarization (SD:2638) and crop_segment (SD:2568).
template <class SW>
vec2d reconstruct(const vec3f& segmentations, const SW& segmentations_frames,
                  const std::vector<std::vector<int>>& hard_clusters, const std::vector<int>& count_data,
                  const SW& count_frames, SW& activations_frames) {
    Context& c = context();
    const int C = static_cast<int>(segmentations.size()), F = static_cast<int>(segmentations[0].size()),
              K = static_cast<int>(segmentations[0][0].size());
    const sd_window cw = detail::to_window(segmentations_frames), cf = detail::to_window(count_frames);
    std::vector<float> seg = detail::flatten3<float, float>(segmentations);
    std::vector<int32_t> hard = detail::flatten2<int, int32_t>(hard_clusters);
    std::vector<int32_t> count(count_data.begin(), count_data.end());
    int kc = 0;
    for (int32_t h : hard) kc = h > kc ? h : kc;
    kc += 1;
    int64_t rows = 0;
#if SDB200_STAGE_DUMPS
    {   // the reference's route with its dumps: clusteredSegmentations (SD:2840-2842) -> to_diarization: aggregate
        // (whose own dumps appear too), crop_segment, sorted_speakers (SD:2653-2654, 2715-2718, 2731-2733)
        std::vector<double> cs(static_cast<size_t>(C) * F * kc);
        c.check(sd_clustered_segmentations(c.get(), seg.data(), C, F, K, hard.data(), kc, cs.data()));
        const vec3d clustered = detail::unflatten3<double, double>(cs, C, F, kc);
        SDB200_DUMP3(clustered, "cpp_clustered_segmentations");
        SW act_frames = count_frames;
        const vec2d activations = aggregate(clustered, segmentations_frames, count_frames, act_frames, false, 0.0, true);
        SDB200_DUMP2(activations, "cpp_to_diarization_activations");
        const sd_window aw = detail::to_window(act_frames);
        std::vector<double> act = detail::flatten2<double, double>(activations), bin(act.size() + 1);
        std::vector<int32_t> order(act.size() + 1);
        int64_t crop[4] = {0, 0, 0, 0};
        sd_window fr;
        c.check(sd_to_diarization(c.get(), act.data(), static_cast<int64_t>(activations.size()), kc, &aw, count.data(),
                                  static_cast<int64_t>(count.size()), &cf, bin.data(), static_cast<int64_t>(bin.size()),
                                  &rows, &fr, order.data(), crop));
        const vec2d cropped_activations(activations.begin() + crop[0], activations.begin() + crop[0] + crop[1]);
        std::vector<std::vector<int>> cropped_count(static_cast<size_t>(crop[3]), std::vector<int>(1)),
            sorted_speakers = detail::unflatten2<int32_t, int>(order, static_cast<size_t>(rows), kc);
        for (int64_t i = 0; i < crop[3]; ++i) cropped_count[i][0] = std::min<int>(count[crop[2] + i], kc);  // SD:2672-2678
        SDB200_DUMP2(cropped_activations, "cpp_cropped_activations");
        SDB200_DUMP2(cropped_count, "cpp_cropped_count");
        SDB200_DUMP2(sorted_speakers, "cpp_sorted_speakers");
        activations_frames.start = fr.start;
        activations_frames.step = fr.step;
        activations_frames.duration = fr.duration;
        bin.resize(static_cast<size_t>(rows) * kc);
        return detail::unflatten2<double, double>(bin, static_cast<size_t>(rows), static_cast<size_t>(kc));
    }
#endif
    c.check(sd_reconstruct_rows(C, &cw, static_cast<int64_t>(count.size()), &cf, &rows, nullptr));
    std::vector<double> out(static_cast<size_t>(rows > 0 ? rows : 0) * kc + 1);
    int cols = 0;
    sd_window fr;
    c.check(sd_reconstruct(c.get(), seg.data(), C, F, K, &cw, hard.data(), count.data(),
                           static_cast<int64_t>(count.size()), &cf, out.data(), static_cast<int64_t>(out.size()), &rows,
                           &cols, &fr));
    activations_frames.start = fr.start;
    activations_frames.step = fr.step;
    activations_frames.duration = fr.duration;
    out.resize(static_cast<size_t>(rows) * cols);
    return detail::unflatten2<double, double>(out, static_cast<size_t>(rows), static_cast<size_t>(cols));
}

// What Annotation::finalResult() (SD:962-978) yields for to_annotation (SD:2852-2935): the speech turns of all
// clusters ordered by start.  `Result` is any type constructible from (double start, double end, int label),
// e.g. the reference's Annotation::Result (SD:866).
struct Turn {
    double start, end;
    int label;
    Turn(double s, double e, int l) : start(s), end(e), label(l) {}
};
template <class Result = Turn, class SW>
std::vector<Result> to_annotation(const vec2d& scores, const SW& frames, double onset, double offset,
                                  double min_duration_on, double min_duration_off) {
    Context& c = context();
    const int64_t rows = static_cast<int64_t>(scores.size());
    const int cols = static_cast<int>(scores[0].size());
    const sd_window fw = detail::to_window(frames);
    std::vector<double> flat = detail::flatten2<double, double>(scores);
    const int64_t cap = (rows / 2 + 2) * cols;
    std::vector<double> seg(static_cast<size_t>(cap) * 2);
    std::vector<int32_t> lab(static_cast<size_t>(cap));
    int64_t n = 0;
    c.check(sd_to_annotation(c.get(), flat.data(), rows, cols, &fw, onset, offset, min_duration_on, min_duration_off,
                             seg.data(), lab.data(), cap, &n));
    std::vector<Result> out;
    out.reserve(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) out.emplace_back(seg[2 * i], seg[2 * i + 1], static_cast<int>(lab[i]));
    return out;
}

// Ingest: wav::WavReader (frontend/wav.h:62-126) + the /32768 scaling at SD:2948-2951 in one call.  Only the RIFF
// chunk walk runs on the host; the sample conversion is the device kernel.  16-bit PCM, mono (the reference reads
// the first num_samples interleaved values of a multi-channel file, i.e. it does not support them either).
inline std::vector<float> read_wav(const std::string& path, int* sample_rate = nullptr) {
    FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) throw std::runtime_error("sdb200: cannot open " + path);
    auto fail = [&](const char* why) {
        std::fclose(fp);
        throw std::runtime_error("sdb200: " + path + ": " + why);
    };
    unsigned char h[12];
    if (std::fread(h, 1, 12, fp) != 12 || std::memcmp(h, "RIFF", 4) || std::memcmp(h + 8, "WAVE", 4)) fail("not a RIFF/WAVE file");
    uint16_t channels = 0, bits = 0;
    uint32_t rate = 0, data_size = 0;
    bool have_fmt = false;
    for (;;) {  // walk sub-chunks until "data" (LIST / fact chunks are skipped, wav.h:84-91)
        unsigned char ck[8];
        if (std::fread(ck, 1, 8, fp) != 8) fail("no data chunk");
        const uint32_t size = ck[4] | (ck[5] << 8) | (ck[6] << 16) | ((uint32_t)ck[7] << 24);
        if (!std::memcmp(ck, "fmt ", 4)) {
            unsigned char f[16];
            if (size < 16 || std::fread(f, 1, 16, fp) != 16) fail("fmt chunk shorter than 16 bytes");
            channels = f[2] | (f[3] << 8);
            rate = f[4] | (f[5] << 8) | (f[6] << 16) | ((uint32_t)f[7] << 24);
            bits = f[14] | (f[15] << 8);
            std::fseek(fp, size - 16, SEEK_CUR);
            have_fmt = true;
        } else if (!std::memcmp(ck, "data", 4)) {
            data_size = size;
            break;
        } else {
            std::fseek(fp, size, SEEK_CUR);
        }
    }
    if (!have_fmt || bits != 16 || channels != 1) fail("only 16-bit mono PCM is supported");
    std::vector<int16_t> pcm(data_size / 2);
    if (std::fread(pcm.data(), 2, pcm.size(), fp) != pcm.size()) fail("truncated data chunk");
    std::fclose(fp);
    if (sample_rate) *sample_rate = static_cast<int>(rate);
    std::vector<float> out(pcm.size());
    Context& c = context();
    if (!pcm.empty()) c.check(sd_ingest_pcm16(c.get(), pcm.data(), static_cast<int64_t>(pcm.size()), out.data()));
    return out;
}

// SegmentModel::crop (SD:1641-1662) for a batch of windows: one zero-padded chunk per segment start.
inline vec2f crop(const std::vector<float>& waveform, const std::vector<std::pair<double, double>>& segments,
                  double duration = 5.0, int sample_rate = 16000) {
    Context& c = context();
    std::vector<double> starts;
    for (const auto& s : segments) starts.push_back(s.first);
    const size_t L = static_cast<size_t>(std::floor(duration * sample_rate));
    std::vector<float> flat(L * starts.size());
    c.check(sd_crop_chunks(c.get(), waveform.data(), static_cast<int64_t>(waveform.size()), starts.data(),
                           static_cast<int>(starts.size()), duration, sample_rate, flat.data()));
    return detail::unflatten2<float, float>(flat, starts.size(), L);
}

}  // namespace sdb200
