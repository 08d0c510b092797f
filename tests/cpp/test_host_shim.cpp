// TEST INFRASTRUCTURE.  The reference translation unit (unmodified, compiled from /root/reference) and the
// sdb200 host shim side by side: every hot-path function is called through both with the same nested-vector
// inputs and the results are compared (exact for integer / fp64 results, 1e-4 abs for the STFT).
// Built here (needs /root/reference); the binary travels to the GPU box and is run by tests/test_gpu_host_shim.py.
#define main ref_main
#include "speakerDiarizer.cpp"
#undef main

#include <random>
#include <tuple>

#include "../../pyannote-audio_speaker-diarization_cpp_b200/host/sdb200_host.hpp"

static int g_fail = 0;
#define CHECK(cond, what)                                            \
    do {                                                             \
        if (!(cond)) {                                               \
            std::fprintf(stderr, "FAIL %s (%s:%d)\n", what, __FILE__, __LINE__); \
            ++g_fail;                                                \
        } else                                                       \
            std::fprintf(stderr, "ok   %s\n", what);                 \
    } while (0)

struct Quiet {
    std::streambuf* old;
    Quiet() : old(std::cout.rdbuf(nullptr)) {}
    ~Quiet() {
        std::cout.rdbuf(old);
        std::cout.clear();
    }
};

template <typename T>
static bool same(const T& a, const T& b) {
    return a == b;
}
static bool same_nan(const std::vector<std::vector<double>>& a, const std::vector<std::vector<double>>& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) {
        if (a[i].size() != b[i].size()) return false;
        for (size_t j = 0; j < a[i].size(); ++j)
            if (!(a[i][j] == b[i][j] || (std::isnan(a[i][j]) && std::isnan(b[i][j])))) return false;
    }
    return true;
}

static const char* g_wav_path = nullptr;

int main(int argc, char** argv) {
    if (argc > 1) g_wav_path = argv[1];
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::normal_distribution<double> G(0.0, 1.0);
    const int C = 40, F = 293, K = 3, D = 192;

    // segmentation scores: smooth-ish random tracks
    std::vector<std::vector<std::vector<float>>> seg(C, std::vector<std::vector<float>>(F, std::vector<float>(K)));
    for (int c = 0; c < C; ++c)
        for (int k = 0; k < K; ++k) {
            float v = U(rng);
            for (int f = 0; f < F; ++f) {
                v = 0.95f * v + 0.05f * U(rng) + 0.02f * (U(rng) - 0.5f);
                seg[c][f][k] = std::min(1.f, std::max(0.f, v + 0.3f * std::sin(0.05f * f + c)));
            }
        }
    Quiet q;
    SegmentModel mm("segment2.onnx");

    auto ref_bin = mm.binarize_swf(seg, false);
    auto got_bin = sdb200::binarize_swf(seg, false);
    CHECK(same(ref_bin, got_bin), "binarize_swf");

    SlidingWindow before(0.0, 0.5, 5.0), tw_ref, tw_got;
    auto ref_trim = mm.trim(ref_bin, 0.1, 0.1, before, tw_ref);
    auto got_trim = sdb200::trim(ref_bin, 0.1, 0.1, before, tw_got);
    CHECK(same(ref_trim, got_trim) && tw_ref.start == tw_got.start && tw_ref.duration == tw_got.duration &&
              tw_ref.num_samples == tw_got.num_samples,
          "trim");

    SlidingWindow pre(0.0, 0.016875, 0.016875), cf_ref(16000 * 30), cf_got(16000 * 30);
    auto ref_cnt = mm.speaker_count(seg, ref_bin, pre, cf_ref, 16000 * 30);
    auto got_cnt = sdb200::speaker_count(seg, ref_bin, pre, cf_got, 16000 * 30);
    CHECK(same(ref_cnt, got_cnt) && cf_ref.start == cf_got.start && cf_ref.step == cf_got.step &&
              cf_ref.num_samples == cf_got.num_samples,
          "speaker_count");

    CHECK(same(Helper::cleanSegmentations(ref_bin), sdb200::cleanSegmentations(ref_bin)), "cleanSegmentations");

    // aggregate with NaN columns, both averaging modes
    std::vector<std::vector<std::vector<double>>> sc(C, std::vector<std::vector<double>>(F, std::vector<double>(4)));
    for (int c = 0; c < C; ++c)
        for (int f = 0; f < F; ++f)
            for (int k = 0; k < 4; ++k) sc[c][f][k] = ((c + k) % 3 == 0) ? NAN : (double)seg[c][f][k % K];
    SlidingWindow sf(0.0, 0.5, 5.0, 16000 * 30), post_ref, post_got;
    auto ref_agg = PipelineHelper::aggregate(sc, sf, pre, post_ref, false, 0.0, true);
    auto got_agg = sdb200::aggregate(sc, sf, pre, post_got, false, 0.0, true);
    CHECK(same_nan(ref_agg, got_agg) && post_ref.start == post_got.start && post_ref.step == post_got.step, "aggregate(skip_average)");
    ref_agg = PipelineHelper::aggregate(sc, sf, pre, post_ref, false, NAN, false);
    got_agg = sdb200::aggregate(sc, sf, pre, post_got, false, NAN, false);
    CHECK(same_nan(ref_agg, got_agg), "aggregate(average, missing=NaN)");

    // binarize_ndarray with exact ties
    std::vector<std::vector<double>> rows(17, std::vector<double>(F));
    for (auto& r : rows)
        for (auto& v : r) v = (U(rng) < 0.2f) ? 0.5 : (double)U(rng);
    // the reference asserts (SD:685) unless at least one row has no frame equal to onset
    for (auto& v : rows[5]) v = (v == 0.5) ? 0.75 : v;
    CHECK(same(mm.binarize_ndarray(rows, 0.5, false), sdb200::binarize_ndarray(rows, 0.5, false)), "binarize_ndarray");
    CHECK(same(mm.binarize_ndarray(rows, 0.5, true), sdb200::binarize_ndarray(rows, 0.5, true)), "binarize_ndarray(initial on)");

    // embeddings: 4 speakers + one tiny cluster, some NaN rows
    std::vector<std::vector<double>> cen(5, std::vector<double>(D));
    for (auto& c : cen) {
        double n = 0;
        for (auto& v : c) {
            v = G(rng);
            n += v * v;
        }
        for (auto& v : c) v /= std::sqrt(n);
    }
    std::vector<std::vector<std::vector<double>>> emb(C, std::vector<std::vector<double>>(K, std::vector<double>(D)));
    int idx = 0;
    for (int c = 0; c < C; ++c)
        for (int s = 0; s < K; ++s, ++idx) {
            int spk = (idx < 4) ? 4 : (int)(rng() % 4);
            double gain = 5.0 + 25.0 * U(rng);
            bool nan = (rng() % 17) == 0;
            for (int d = 0; d < D; ++d)
                emb[c][s][d] = nan ? NAN : (double)(float)((cen[spk][d] + 0.023 * G(rng)) * gain);
        }
    std::vector<std::vector<int>> hard_ref, hard_got;
    Cluster cst;
    cst.clustering(emb, ref_bin, hard_ref);
    sdb200::Cluster cst2;
    cst2.clustering(emb, ref_bin, hard_got);
    CHECK(same(hard_ref, hard_got), "Cluster::clustering");

    std::vector<std::vector<double>> filt;
    for (auto& a : emb)
        for (auto& e : a)
            if (!std::isnan(e[0])) filt.push_back(e);
    auto n1 = filt, n2 = filt;
    Helper::normalizeEmbeddings(n1);
    sdb200::normalizeEmbeddings(n2);
    CHECK(same(n1, n2), "normalizeEmbeddings");
    std::vector<std::vector<double>> Zr, Zg;
    ::Clustering::linkage(n1, Zr);
    sdb200::Clustering::linkage(n1, Zg);
    CHECK(same(Zr, Zg), "Clustering::linkage");
    std::vector<int> Tr, Tg;
    ::Clustering::fcluster(Zr, 0.7153814435005188, Tr);
    sdb200::Clustering::fcluster(Zr, 0.7153814435005188, Tg);
    CHECK(same(Tr, Tg), "Clustering::fcluster");
    CHECK(same(::Clustering::cluster(n1, 1.1), sdb200::Clustering::cluster(n1, 1.1)), "Clustering::cluster");
    // the reference's own toy (pipeline/src/clustering/cluster.cpp:8-13)
    std::vector<std::vector<double>> toy = {{0, 0}, {0, 1}, {1, 0}, {0, 4}, {0, 3}, {1, 4},
                                            {4, 0}, {3, 0}, {4, 1}, {4, 4}, {3, 4}, {4, 3}};
    CHECK(same(sdb200::Clustering::cluster(toy, 1.1), std::vector<int>({5, 5, 6, 7, 7, 8, 1, 1, 2, 3, 3, 4})), "toy clusters");
    auto lc = std::vector<std::vector<double>>(filt.begin(), filt.begin() + 9);
    auto scn = std::vector<std::vector<double>>(filt.begin() + 9, filt.begin() + 13);
    CHECK(same(Helper::cosineSimilarity(lc, scn), sdb200::cosineSimilarity(lc, scn)), "cosineSimilarity");
    bool threw = false;
    scn[1].assign(D, 0.0);
    try {
        sdb200::cosineSimilarity(lc, scn);
    } catch (const std::runtime_error& e) {
        threw = std::string(e.what()) == "Vectors have zero magnitude.";
    }
    CHECK(threw, "zero magnitude throws like the reference");

    // rows after clustering: reconstruct (+ to_diarization, crop_segment) and to_annotation
    {
        SlidingWindow act_ref, act_got;
        auto rec_ref = reconstruct(seg, sf, hard_ref, ref_cnt, cf_ref, act_ref);
        auto rec_got = sdb200::reconstruct(seg, sf, hard_ref, ref_cnt, cf_ref, act_got);
        CHECK(same(rec_ref, rec_got) && act_ref.start == act_got.start && act_ref.step == act_got.step &&
                  act_ref.duration == act_got.duration,
              "reconstruct");
        auto ann = to_annotation(rec_ref, act_ref, 0.5, 0.5, 0.0, 0.5817029604921046f).finalResult();
        auto turns = sdb200::to_annotation(rec_ref, act_ref, 0.5, 0.5, 0.0, 0.5817029604921046f);
        auto key = [](double s, double e, int l) { return std::make_tuple(s, e, l); };
        std::vector<std::tuple<double, double, int>> ka, kb;
        for (auto& r : ann) ka.push_back(key(r.start, r.end, r.label));
        for (auto& r : turns) kb.push_back(key(r.start, r.end, r.label));
        bool sorted = std::is_sorted(turns.begin(), turns.end(), [](const sdb200::Turn& a, const sdb200::Turn& b) { return a.start < b.start; });
        std::sort(ka.begin(), ka.end());  // std::sort by start leaves equal starts in unspecified order
        std::sort(kb.begin(), kb.end());
        CHECK(!ka.empty() && ka == kb && sorted, "to_annotation");
    }

    // masking prologue of getEmbedding: Helper::interpolate + Helper::padSequence
    {
        const int B = 5, L = 80000;
        std::vector<std::vector<float>> wv(B, std::vector<float>(L)), mk(B, std::vector<float>(F));
        for (auto& r : wv)
            for (auto& v : r) v = U(rng) - 0.5f;
        for (int b = 0; b < B; ++b)
            for (int f = 0; f < F; ++f) mk[b][f] = (float)ref_bin[b * 3][f][b % K];
        std::fill(mk[3].begin(), mk[3].end(), 0.f);
        mk[3][7] = 1.f;  // shorter than min_num_samples -> too short
        auto imasks = Helper::interpolate(mk, L, 0.5);
        auto sig_ref = Helper::padSequence(wv, imasks);
        std::vector<std::vector<float>> sig_got;
        std::vector<float> wl;
        std::vector<bool> ts;
        bool ok = sdb200::masked_signals(wv, mk, 640, sig_got, wl, ts);
        float mx = 0;
        std::vector<float> cnt(B);
        for (int b = 0; b < B; ++b) {
            cnt[b] = (float)std::count(imasks[b].begin(), imasks[b].end(), true);
            mx = std::max(mx, cnt[b]);
        }
        bool lens_ok = true;
        for (int b = 0; b < B; ++b) lens_ok = lens_ok && (cnt[b] < 640 ? (wl[b] == 1.0f && ts[b]) : (wl[b] == cnt[b] / mx && !ts[b]));
        CHECK(ok && same(sig_ref, sig_got) && lens_ok, "masked_signals (interpolate + padSequence + wav_lens)");
    }

    // ingest: WavReader + scaling, SegmentModel::crop.  The fixture path comes from argv[1] (tests/golden/tiny_list.wav).
    if (g_wav_path) {
        wav::WavReader rd(g_wav_path);
        std::vector<float> ref_w(rd.data(), rd.data() + rd.num_samples());
        for (auto& v : ref_w) v = v * 1.0f / 32768.0;
        int sr = 0;
        auto got_w = sdb200::read_wav(g_wav_path, &sr);
        CHECK(same(ref_w, got_w) && sr == rd.sample_rate() && !got_w.empty(), "read_wav (WavReader + /32768)");
        std::vector<float> longw(16000 * 7 + 333);
        for (auto& v : longw) v = U(rng) - 0.5f;
        std::vector<std::pair<double, double>> segs = {{0.0, 5.0}, {0.5, 5.5}, {2.25, 7.25}, {6.99, 11.99}};
        auto got_c = sdb200::crop(longw, segs);
        bool ok = got_c.size() == segs.size();
        for (size_t i = 0; ok && i < segs.size(); ++i) ok = same(mm.crop(longw, segs[i]), got_c[i]);
        CHECK(ok, "crop (SegmentModel::crop)");
    }

    // STFT front-end: what reaches emd4.onnx (captured by the ORT stub) vs the shim
    std::vector<std::vector<float>> wav(3, std::vector<float>(16000));
    for (auto& r : wav)
        for (size_t i = 0; i < r.size(); ++i) r[i] = 0.3f * std::sin(0.01f * i * (1 + (&r - &wav[0]))) + 0.1f * (U(rng) - 0.5f);
    std::vector<float> lens = {1.0f, 0.5f, 0.25f};
    EmbeddingModel1 em("emd4.onnx");
    em.infer(wav, lens);
    auto& cap = ort_stub::captured();
    auto got = sdb200::embedding_input(wav, lens);
    bool shape_ok = cap.size() == 2 && cap[0].data.size() == got.audio.size() && cap[0].shape[0] == got.dims[0] &&
                    cap[0].shape[1] == got.dims[1] && cap[0].shape[2] == got.dims[2] && cap[0].shape[3] == got.dims[3];
    double err = 0;
    if (shape_ok)
        for (size_t i = 0; i < got.audio.size(); ++i) err = std::max(err, (double)std::fabs(got.audio[i] - cap[0].data[i]));
    std::fprintf(stderr, "stft max abs err vs EmbeddingModel1::infer: %.3g\n", err);
    CHECK(shape_ok && err < 1e-4, "embedding_input (STFT tensor handed to emd4.onnx)");
    CHECK(cap.size() == 2 && cap[1].data == got.wav_lens, "wav_lens packing");

    std::fprintf(stderr, "%s: %d failure(s)\n", g_fail ? "FAILED" : "PASSED", g_fail);
    return g_fail ? 1 : 0;
}
