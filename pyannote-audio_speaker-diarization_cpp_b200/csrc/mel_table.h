// Mel filterbank tables of the fused log-mel front-end (host side; plain C++ so that tests/cpp/emulate_mel.cpp can
// replay the kernel's projection on the CPU -- test infrastructure only).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/sdb200.h"
#include "fft400.cuh"

namespace sdb {

struct MelTable {       // device copy lives in ctx->d_mel
    int lo[128];        // first bin of filter m
    int cnt[128];       // number of bins with non-zero weight
    int off[128];       // offset of its weights in w[]
    float w[1024];      // packed non-zero weights
};

// speechbrain 0.5.14 Filterbank (triangular, fp32 arithmetic like the torch module): mel points
// linspace(mel(f_min), mel(f_max), n_mels + 2); centre = hz[1..n_mels]; band = hz[m+1] - hz[m] for both slopes.
inline int build_mel_table(const sd_fbank_params* p, MelTable& t) {
    const int n_bins = kBins, n_mels = p->n_mels, np = n_mels + 2;
    if (n_mels < 1 || n_mels > 128) return SD_ERR_UNSUPPORTED;
    std::vector<float> hz(np);
    const float mlo = (float)(2595.0 * std::log10(1.0 + (double)p->f_min / 700.0));
    const float mhi = (float)(2595.0 * std::log10(1.0 + (double)p->f_max / 700.0));
    for (int i = 0; i < np; ++i) {
        const float mel = mlo + (mhi - mlo) * (float)i / (float)(np - 1);
        hz[i] = 700.0f * (std::pow(10.0f, mel / 2595.0f) - 1.0f);
    }
    int used = 0;
    for (int m = 0; m < n_mels; ++m) {
        const float fc = hz[m + 1], band = hz[m + 1] - hz[m];
        int lo = -1, hi = -1;
        std::vector<float> wts(n_bins);
        for (int f = 0; f < n_bins; ++f) {
            const float freq = (float)(p->sample_rate / 2) * (float)f / (float)(n_bins - 1);
            const float slope = (freq - fc) / band;
            const float l = slope + 1.0f, r = -slope + 1.0f;
            const float v = std::max(0.0f, std::min(l, r));
            wts[f] = v;
            if (v > 0.0f) {
                if (lo < 0) lo = f;
                hi = f;
            }
        }
        t.lo[m] = lo < 0 ? 0 : lo;
        t.cnt[m] = lo < 0 ? 0 : hi - lo + 1;
        t.off[m] = used;
        if (used + t.cnt[m] > 1024) return SD_ERR_UNSUPPORTED;
        for (int i = 0; i < t.cnt[m]; ++i) t.w[used + i] = wts[t.lo[m] + i];
        used += t.cnt[m];
    }
    return SD_OK;
}

// Kaldi mel banks (kaldi::MelBanks, no VTLN; torchaudio.compliance.kaldi.get_mel_banks): n_mels triangles equally
// spaced on mel = 1127 ln(1 + f/700) between f_min and f_max (0 = Nyquist), evaluated at the centres of the n_fft/2
// lower FFT bins (the Nyquist bin gets no weight), slopes linear in mel.
inline int build_mel_table_kaldi(const sd_fbank_params* p, MelTable& t) {
    const int n_mels = p->n_mels, n_fft_bins = kNfft / 2;
    if (n_mels < 1 || n_mels > 128) return SD_ERR_UNSUPPORTED;
    auto mel = [](double f) { return 1127.0 * std::log(1.0 + f / 700.0); };
    const double nyquist = 0.5 * p->sample_rate, bin_width = (double)p->sample_rate / kNfft;
    const double high = p->f_max <= 0.f ? nyquist + p->f_max : (double)p->f_max;
    const double mel_lo = mel(p->f_min), mel_hi = mel(high), delta = (mel_hi - mel_lo) / (n_mels + 1);
    int used = 0;
    for (int m = 0; m < n_mels; ++m) {
        const double left = mel_lo + m * delta, center = left + delta, right = center + delta;
        int lo = -1, hi = -1;
        std::vector<float> wts(n_fft_bins, 0.f);
        for (int i = 0; i < n_fft_bins; ++i) {
            const double mf = mel(bin_width * i);
            const double up = (mf - left) / (center - left), down = (right - mf) / (right - center);
            const double v = std::max(0.0, std::min(up, down));
            wts[i] = (float)v;
            if (v > 0.0) {
                if (lo < 0) lo = i;
                hi = i;
            }
        }
        t.lo[m] = lo < 0 ? 0 : lo;
        t.cnt[m] = lo < 0 ? 0 : hi - lo + 1;
        t.off[m] = used;
        if (used + t.cnt[m] > 1024) return SD_ERR_UNSUPPORTED;
        for (int i = 0; i < t.cnt[m]; ++i) t.w[used + i] = wts[t.lo[m] + i];
        used += t.cnt[m];
    }
    return SD_OK;
}


// ---- the projection as the kernel runs it ---------------------------------------------------------------------------
// out[f][m] = sum_k W[m][k] P[f][k] with triangular filters: a bin feeds at most two filters, and they are consecutive,
// so one of them is even and one is odd.  The 201 bins are cut into ten parts of 20 (the last one 21); a thread owns one
// (frame, part): it walks the part's bins once, loads P[f][k] once, and feeds an "even" and an "odd" accumulator with the
// two weights of the bin.  When the filter behind an accumulator changes (supports are contiguous, so the old one is
// finished inside this part) the accumulator is flushed to the filter's output column and cleared.  A filter whose bins
// lie in two parts gets its first partial sum in column m and the second one in an extra column n_mels + j; the
// final pass adds the two.  Every part has the same number of entries, so the two half-warps of a warp (16 frames x 2
// parts) run in lockstep, the 16 lanes of a half-warp read P with the frame stride (201 = 9 mod 32: conflict-free) and
// all read the same table entry (broadcast).
constexpr int kMelParts = 10;
constexpr int kMelPartBins = 20;
constexpr int kMelEntries = 22;   // 21 bin entries (parts 0..8: the 21st has zero weights) + the final flush
constexpr int kMelExtraCols = 32; // second partial sums of the filters that straddle a part boundary
struct alignas(16) MelEntry {
    float wE, wO;  // weights of this bin for the even / odd filter it feeds (0 if none)
    int fE, fO;    // output column to flush the even / odd accumulator to BEFORE this bin is added, or -1
};
struct MelParts {
    MelEntry e[kMelParts][kMelEntries];
    int extra[128];  // extra column of filter m (its second partial sum), or -1
    int has[128];    // filter m has at least one bin
};
inline int mel_part_of_bin(int k) { return k / kMelPartBins < kMelParts ? k / kMelPartBins : kMelParts - 1; }

// SD_ERR_UNSUPPORTED when the filterbank does not have the structure above (a bin under three filters, two filters of
// the same parity, a filter in three parts, too many straddlers): the caller then keeps the per-filter projection.
inline int build_mel_parts(const MelTable& t, int n_mels, MelParts& mp) {
    std::memset(&mp, 0, sizeof(mp));
    for (int m = 0; m < 128; ++m) mp.extra[m] = -1;
    int n_extra = 0;
    for (int m = 0; m < n_mels; ++m) {
        mp.has[m] = t.cnt[m] > 0;
        if (t.cnt[m] > 0 && mel_part_of_bin(t.lo[m] + t.cnt[m] - 1) - mel_part_of_bin(t.lo[m]) > 1) return SD_ERR_UNSUPPORTED;
    }
    auto weight = [&](int m, int k) -> float {
        return (t.cnt[m] > 0 && k >= t.lo[m] && k < t.lo[m] + t.cnt[m]) ? t.w[t.off[m] + k - t.lo[m]] : 0.f;
    };
    auto column = [&](int m, int q) -> int {  // where part q's partial sum of filter m goes
        if (mel_part_of_bin(t.lo[m]) == q) return m;
        if (mp.extra[m] < 0) {
            if (n_extra >= kMelExtraCols) return -2;
            mp.extra[m] = n_extra++;
        }
        return n_mels + mp.extra[m];
    };
    for (int q = 0; q < kMelParts; ++q) {
        const int kb = q * kMelPartBins, ke = q == kMelParts - 1 ? kBins : kb + kMelPartBins;
        int cur[2] = {-1, -1};  // filter behind the even / odd accumulator
        for (int i = 0; i < kMelEntries; ++i) {
            MelEntry& en = mp.e[q][i];
            en.wE = en.wO = 0.f;
            en.fE = en.fO = -1;
            const int k = kb + i;
            int need[2] = {-1, -1};
            if (i < kMelEntries - 1 && k < ke) {
                for (int m = 0; m < n_mels; ++m) {
                    if (t.cnt[m] > 0 && k >= t.lo[m] && k < t.lo[m] + t.cnt[m]) {
                        if (need[m & 1] >= 0) return SD_ERR_UNSUPPORTED;  // two filters of one parity on a bin
                        need[m & 1] = m;
                    }
                }
            }
            const bool last = i == kMelEntries - 1, in_part = !last && k < ke;
            if (!last && !in_part) continue;  // padding entry of a 20-bin part: zero weights, nothing flushed
            for (int par = 0; par < 2; ++par) {
                // supports are contiguous: a filter is finished in this part as soon as one of its bins does not feed it
                if (cur[par] >= 0 && (last || cur[par] != need[par])) {
                    const int c = column(cur[par], q);
                    if (c == -2) return SD_ERR_UNSUPPORTED;
                    (par == 0 ? en.fE : en.fO) = c;
                    cur[par] = -1;
                }
                if (in_part && need[par] >= 0) {
                    cur[par] = need[par];
                    (par == 0 ? en.wE : en.wO) = weight(need[par], k);
                }
            }
        }
    }
    return SD_OK;
}

}  // namespace sdb
