// TEST INFRASTRUCTURE: CPU replay of the fused fbank kernel's mel projection (csrc/mel_table.h: ten 20-bin parts, an
// even and an odd accumulator per (frame, part), flush columns, extra columns for filters that straddle a part
// boundary) against the plain per-filter sum, for the speechbrain and the Kaldi filterbanks.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pyannote-audio_speaker-diarization_cpp_b200/csrc/mel_table.h"

using namespace sdb;

static int run(const char* name, const MelTable& t, int n_mels, bool expect_supported) {
    static MelParts mp;
    const int rc = build_mel_parts(t, n_mels, mp);
    if (rc != SD_OK) {
        printf("%s: part table not available (rc %d)%s\n", name, rc, expect_supported ? " -- UNEXPECTED" : " (expected)");
        return expect_supported ? 1 : 0;
    }
    const int frames = 16, stride = n_mels + kMelExtraCols + 1;
    std::vector<float> pw(frames * kBins), out(frames * stride, -12345.f);
    srand(7);
    for (auto& v : pw) v = (float)rand() / RAND_MAX * 100.f;
    for (int f = 0; f < frames; ++f)
        for (int q = 0; q < kMelParts; ++q) {  // one GPU thread
            float accE = 0.f, accO = 0.f;
            const float* prow = &pw[f * kBins + q * kMelPartBins];
            float* orow = &out[f * stride];
            for (int i = 0; i < kMelEntries; ++i) {
                const MelEntry& en = mp.e[q][i];
                if (en.fE >= 0) { orow[en.fE] = accE; accE = 0.f; }
                if (en.fO >= 0) { orow[en.fO] = accO; accO = 0.f; }
                if (i < kMelEntries - 1) {
                    const int k = q * kMelPartBins + i;
                    if (k >= kBins + 0 && (en.wE != 0.f || en.wO != 0.f)) { printf("%s: weight past the last bin\n", name); return 2; }
                    const float p = k < kBins ? prow[i] : 0.f;  // the kernel reads a valid (finite) float there
                    accE = fmaf(en.wE, p, accE);
                    accO = fmaf(en.wO, p, accO);
                }
            }
            if (accE != 0.f || accO != 0.f) { printf("%s: accumulator not flushed (part %d)\n", name, q); return 3; }
        }
    double maxrel = 0;
    int straddlers = 0;
    for (int m = 0; m < n_mels; ++m) straddlers += mp.extra[m] >= 0;
    for (int f = 0; f < frames; ++f)
        for (int m = 0; m < n_mels; ++m) {
            double want = 0;
            for (int i = 0; i < t.cnt[m]; ++i) want += (double)t.w[t.off[m] + i] * pw[f * kBins + t.lo[m] + i];
            float got = mp.has[m] ? out[f * stride + m] : 0.f;
            if (mp.extra[m] >= 0) got += out[f * stride + n_mels + mp.extra[m]];
            if (mp.has[m] && out[f * stride + m] == -12345.f) { printf("%s: filter %d never written\n", name, m); return 4; }
            const double rel = fabs(got - want) / (fabs(want) + 1e-3);
            if (rel > maxrel) maxrel = rel;
        }
    printf("%s: %d filters, %d straddle a part boundary, max rel err %.2e\n", name, n_mels, straddlers, maxrel);
    return maxrel < 1e-5 ? 0 : 5;
}

int main() {
    sd_fbank_params p;
    memset(&p, 0, sizeof(p));
    p.n_mels = 80; p.f_min = 0.f; p.f_max = 8000.f; p.sample_rate = 16000;
    static MelTable t;
    int bad = 0;
    memset(&t, 0, sizeof(t));
    if (build_mel_table(&p, t) != SD_OK) return 10;
    bad |= run("speechbrain 80", t, 80, true);
    memset(&t, 0, sizeof(t));
    p.f_min = 20.f; p.f_max = 0.f;
    if (build_mel_table_kaldi(&p, t) != SD_OK) return 11;
    bad |= run("kaldi 80", t, 80, true);
    memset(&t, 0, sizeof(t));
    p.n_mels = 23;
    if (build_mel_table_kaldi(&p, t) != SD_OK) return 12;
    bad |= run("kaldi 23", t, 23, false);  // wide filters span three parts: the per-filter projection stays
    memset(&t, 0, sizeof(t));
    p.n_mels = 128; p.f_min = 0.f; p.f_max = 8000.f;
    if (build_mel_table(&p, t) != SD_OK) return 13;
    bad |= run("speechbrain 128", t, 128, true);
    return bad;
}
