// 400-point DFT building blocks for the STFT front-end (replaces torch::stft at SD:2008).
//
// Two real frames a, b are transformed together as one complex sequence z = a + i*b with a 20 x 20
// Cooley-Tukey split; each 20-point DFT is a 4 x 5 prime-factor (Good-Thomas) transform, so there are no
// twiddles inside it.  A "group" of 20 threads owns one frame pair:
//
//   phase 1  thread r (= n2): 20-point DFT over n1 of z[20*n1 + r] * w[20*n1 + r], times W400^(r*k1),
//            written to the exchange buffer at [k1][r]
//   phase 2  thread r (= k1): 20-point DFT over n2 of exchange[r][n2]  ->  Z[r + 20*k2], written to zbuf
//   phase 3  thread r: bins k = r, r+20, ... <= 200:  A[k] = (Z[k] + conj Z[400-k]) / 2,
//                                                    B[k] = (Z[k] - conj Z[400-k]) / (2i)
//
// Everything is __host__ __device__ so the exact index maps can be replayed on the CPU by
// tests/cpp/emulate_stft.cpp (a thread-by-thread emulation of the kernel; test infrastructure only).
#pragma once

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define SD_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define SD_HD inline
struct float2 {
    float x, y;
};
static inline float2 make_float2(float x, float y) {
    float2 r;
    r.x = x;
    r.y = y;
    return r;
}
#endif

namespace sdb {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kRadix = 20;          // threads per frame pair, points per thread
constexpr int kXchgRow = 21;        // padded row stride (float2 units) of the phase-1 -> phase-2 transpose
constexpr int kGroupStride = 436;   // float2 units per group in either exchange buffer (436 - 20 = 26 * 16)

SD_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SD_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SD_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// multiply by -i (forward-transform rotation)
SD_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

SD_HD void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = mul_mi(csub(x1, x3));
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = cadd(t1, t3);
    x3 = csub(t1, t3);
}

SD_HD void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
    const float c = 0.55901699437494742f;   // (cos(2pi/5) - cos(4pi/5)) / 2
    const float s1 = 0.95105651629515357f;  // sin(2pi/5)
    const float s2 = 0.58778525229247313f;  // sin(4pi/5)
    float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    float2 t5 = cadd(t1, t2);
    float2 m1 = make_float2(x0.x - 0.25f * t5.x, x0.y - 0.25f * t5.y);
    float2 d = csub(t1, t2);
    float2 m2 = make_float2(c * d.x, c * d.y);
    float2 a1 = cadd(m1, m2), a2 = csub(m1, m2);
    float2 u1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    float2 u2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    u1 = mul_mi(u1);
    u2 = mul_mi(u2);
    x0 = cadd(x0, t5);
    x1 = cadd(a1, u1);
    x4 = csub(a1, u1);
    x2 = cadd(a2, u2);
    x3 = csub(a2, u2);
}

// Slot of the register array holding input sample n (natural order) == n; after dft20() output bin k is
// found at slot dft20_slot(k).
SD_HD constexpr int dft20_slot(int k) { return (5 * (k % 4) + 4 * (k % 5)) % 20; }

// In-place 20-point forward DFT, v[n] natural-order input.
SD_HD void dft20(float2 (&v)[20]) {
#define SDB_S(n1, n2) v[(5 * (n1) + 4 * (n2)) % 20]
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) dft4(SDB_S(0, n2), SDB_S(1, n2), SDB_S(2, n2), SDB_S(3, n2));
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft5(SDB_S(k1, 0), SDB_S(k1, 1), SDB_S(k1, 2), SDB_S(k1, 3), SDB_S(k1, 4));
#undef SDB_S
}

// ---- the three phases, for thread (group g, role r) ------------------------------------------------
// sig   : staged samples of the tile; frame f of the tile starts at sig[f * kHop]
// win   : window values of this thread, win[n1] = w[20*n1 + r]
// tw    : twiddles of this thread,     tw[k1]  = exp(-2*pi*i * r*k1 / 400)
// xchg  : transpose buffer, zbuf: spectrum buffer (both kGroupStride float2 per group)
SD_HD void stft_phase1(const float* sig, int fa_off, int fb_off, const float (&win)[20], const float2 (&tw)[20], int g,
                       int r, float2* xchg) {
    float2 v[20];
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
        const int o = 20 * n1 + r;
        v[n1] = make_float2(sig[fa_off + o] * win[n1], sig[fb_off + o] * win[n1]);
    }
    dft20(v);
    float2* dst = xchg + g * kGroupStride + r;
#pragma unroll
    for (int k1 = 0; k1 < 20; ++k1) {
        float2 y = v[dft20_slot(k1)];
        if (k1 > 0) y = cmul(y, tw[k1]);
        dst[k1 * kXchgRow] = y;
    }
}

SD_HD void stft_phase2(const float2* xchg, int g, int r, float2* zbuf) {
    float2 v[20];
    const float2* src = xchg + g * kGroupStride + r * kXchgRow;
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = src[n2];
    dft20(v);
    float2* dst = zbuf + g * kGroupStride + r;
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) dst[20 * k2] = v[dft20_slot(k2)];
}

// Writes bins of frame A to outA[2k..2k+1] and of frame B to outB (either may be null when the frame is
// past the end of the item).
SD_HD void stft_phase3(const float2* zbuf, int g, int r, float* outA, float* outB) {
    const float2* z = zbuf + g * kGroupStride;
#pragma unroll
    for (int m = 0; m <= 10; ++m) {
        const int k = r + 20 * m;
        if (k <= 200) {
            const float2 zk = z[k];
            const float2 zm = z[k == 0 ? 0 : kNfft - k];
            if (outA) reinterpret_cast<float2*>(outA)[k] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            if (outB) reinterpret_cast<float2*>(outB)[k] = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
        }
    }
}

}  // namespace sdb
