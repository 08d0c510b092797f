// TEST INFRASTRUCTURE: thread-by-thread CPU replay of the three STFT kernel phases in csrc/fft400.cuh,
// checked against a direct fp64 DFT.  Validates the index maps / exchange layouts without a GPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pyannote-audio_speaker-diarization_cpp_b200/csrc/fft400.cuh"

int main() {
    using namespace sdb;
    const int groups = 8, frames = 16, nthreads = groups * kRadix;
    const int nsig = (frames - 1) * kHop + kNfft;
    std::vector<float> sig(nsig), w(kNfft);
    srand(1);
    for (auto& s : sig) s = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (int n = 0; n < kNfft; ++n) w[n] = (float)(0.54 - 0.46 * cos(2.0 * M_PI * n / kNfft));
    std::vector<float2> xchg(groups * kGroupStride), zbuf(groups * kGroupStride);
    std::vector<float> out(frames * kBins * 2, 0.f);
    struct TS { float win[20]; float2 tw[20]; };
    std::vector<TS> ts(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        int r = t % kRadix;
        for (int n1 = 0; n1 < 20; ++n1) ts[t].win[n1] = w[20 * n1 + r];
        for (int k1 = 0; k1 < 20; ++k1) {
            double a = -2.0 * M_PI * (double)(r * k1) / kNfft;
            ts[t].tw[k1] = make_float2((float)cos(a), (float)sin(a));
        }
    }
    // variant A: per-thread window / twiddle registers, unpadded samples, two exchange buffers
    for (int t = 0; t < nthreads; ++t) {
        int g = t / kRadix, r = t % kRadix;
        stft_phase1(sig.data(), (2 * g) * kHop, (2 * g + 1) * kHop, ts[t].win, ts[t].tw, g, r, xchg.data());
    }
    for (int t = 0; t < nthreads; ++t) stft_phase2(xchg.data(), t / kRadix, t % kRadix, zbuf.data());
    for (int t = 0; t < nthreads; ++t) {
        int g = t / kRadix;
        stft_phase3(zbuf.data(), g, t % kRadix, &out[(2 * g) * kBins * 2], &out[(2 * g + 1) * kBins * 2]);
    }
    // variant B (the kernel's): padded staging, shared window / twiddle tables, one transpose through shared memory, the
    // real-pair split between adjacent lanes
    std::vector<float> sigp(sig_padded_size(nsig) + 8, 0.f), out2(frames * kBins * 2, 0.f);
    std::vector<wtab_t> wtab(w.size());  // w/2: the kernel folds the 1/2 of the real-pair split into the window
    for (size_t i = 0; i < w.size(); ++i) wtab[i] = wtab_make(0.5f * w[i]);
    for (int j = 0; j < nsig; ++j) sigp[sig_pos(j)] = sig[j];
    std::vector<float2> twT(kTwTableUnits), xb(groups * kGroupStride);
    for (int e = 0; e < kTwTableUnits; ++e) {
        const int src = tw_table_source(e), r = src / 20, k1 = src % 20;
        double a = -2.0 * M_PI * (double)(r * k1) / kNfft;
        twT[e] = make_float2((float)cos(a), (float)sin(a));
    }
    // bank check of the twiddle layout: every half-warp must hit 16 distinct 8-byte banks (or equal addresses)
    for (int h = 0; h < nthreads / 16; ++h) {
        int owner[16];
        for (int b = 0; b < 16; ++b) owner[b] = -1;
        for (int t = 16 * h; t < 16 * h + 16; ++t) {
            const int u = tw_thread_offset(t), bank = u % 16;
            if (owner[bank] != -1 && owner[bank] != u) { printf("twiddle bank conflict in half-warp %d\n", h); return 3; }
            owner[bank] = u;
        }
    }
    // bank check of the padded sample layout: every warp must hit 32 distinct 4-byte banks
    for (int wp = 0; wp < nthreads / 32; ++wp)
        for (int n1 = 0; n1 < 20; ++n1)
            for (int fr = 0; fr < 2; ++fr) {
                int owner[32];
                for (int b = 0; b < 32; ++b) owner[b] = -1;
                for (int t = 32 * wp; t < 32 * wp + 32; ++t) {
                    const int g = t / kRadix, r = t % kRadix;
                    const int a = sig_frame_off(2 * g + fr) + 20 * n1 + r + (fr ? sig_pad_odd(n1) : sig_pad_even(n1));
                    if (a != sig_pos((2 * g + fr) * kHop + 20 * n1 + r)) { printf("sig_pos mismatch\n"); return 4; }
                    if (owner[a % 32] != -1 && owner[a % 32] != a) { printf("sample bank conflict in warp %d\n", wp); return 4; }
                    owner[a % 32] = a;
                }
            }
    std::vector<std::vector<float2>> regs(nthreads, std::vector<float2>(20));
    for (int t = 0; t < nthreads; ++t) {
        int g = t / kRadix, r = t % kRadix;
        stft_phase1_tab(sigp.data(), sig_frame_off(2 * g), sig_frame_off(2 * g + 1), wtab.data(), twT.data() + tw_thread_offset(t), g, r, xb.data());
    }
    for (int t = 0; t < nthreads; ++t) {
        float2 v[20];
        stft_phase2_load(xb.data(), t / kRadix, t % kRadix, v);
        for (int i = 0; i < 20; ++i) regs[t][i] = v[i];
    }
    // second exchange: slot s plays role pair_role(s) in phase 2 and Z[400 - k] sits in the pair lane t ^ 1 (on the
    // device: shfl.xor 1; here: the other thread's register array).  Pairs must not straddle a warp.
    for (int t = 0; t < nthreads; ++t) {
        if ((t ^ 1) / 32 != t / 32 || (t ^ 1) / kRadix != t / kRadix) { printf("pair lanes straddle at %d\n", t); return 5; }
        const int s = t % kRadix;
        if (pair_slot(pair_role(s)) != s) { printf("pair_slot is not the inverse of pair_role at %d\n", s); return 5; }
        if (s >= 2 && pair_role(s) + pair_role(s ^ 1) != 20) { printf("lanes %d and %d are not partners\n", s, s ^ 1); return 5; }
    }
    for (int t = 0; t < nthreads; ++t) {
        int g = t / kRadix;
        float2 v[20];
        for (int i = 0; i < 20; ++i) v[i] = regs[t][i];
        const std::vector<float2>& pv = regs[t ^ 1];
        stft_split_store_pair(v, t % kRadix, &out2[(2 * g) * kBins * 2], &out2[(2 * g + 1) * kBins * 2],
                              [&](float2 mine, int idx) { return t % kRadix < 2 ? mine : pv[idx]; });
    }
    // the fbank front-end's variant of the same exchange: |A|^2, |B|^2
    {
        std::vector<float> pw(frames * kBins, -1.f);
        for (int t = 0; t < nthreads; ++t) {
            int g = t / kRadix;
            float2 v[20];
            for (int i = 0; i < 20; ++i) v[i] = regs[t][i];
            const std::vector<float2>& pv = regs[t ^ 1];
            stft_split_power_pair(v, t % kRadix, &pw[(2 * g) * kBins], &pw[(2 * g + 1) * kBins],
                                  [&](float2 mine, int idx) { return t % kRadix < 2 ? mine : pv[idx]; });
        }
        for (int f = 0; f < frames; ++f)
            for (int k = 0; k < kBins; ++k) {
                const double re = out2[(f * kBins + k) * 2], im = out2[(f * kBins + k) * 2 + 1];
                if (fabs(re * re + im * im - (double)pw[f * kBins + k]) > 1e-4 * (1.0 + re * re + im * im)) {
                    printf("power mismatch at frame %d bin %d\n", f, k);
                    return 6;
                }
            }
    }
    // the kernel's phase 1 composes fourteen of its nineteen twiddles from five table entries (twiddle_store): the
    // two variants agree to rounding, not bit for bit
    for (size_t i = 0; i < out.size(); ++i)
        if (fabs((double)out[i] - (double)out2[i]) > 2e-5) {
            printf("variant mismatch at %zu: %g vs %g\n", i, out[i], out2[i]);
            return 2;
        }
    double maxerr = 0, maxabs = 0;
    for (int f = 0; f < frames; ++f)
        for (int k = 0; k < kBins; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < kNfft; ++n) {
                double x = (double)sig[f * kHop + n] * (double)w[n];
                double a = -2.0 * M_PI * (double)((long)k * n % kNfft) / kNfft;
                re += x * cos(a);
                im += x * sin(a);
            }
            maxerr = fmax(maxerr, fmax(fabs(re - out[(f * kBins + k) * 2]), fabs(im - out[(f * kBins + k) * 2 + 1])));
            maxerr = fmax(maxerr, fmax(fabs(re - out2[(f * kBins + k) * 2]), fabs(im - out2[(f * kBins + k) * 2 + 1])));
            maxabs = fmax(maxabs, fmax(fabs(re), fabs(im)));
        }
    printf("max|X| %.3f  max abs err %.3e\n", maxabs, maxerr);
    return maxerr < 2e-5 ? 0 : 1;
}
