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

// Phase-2 roles.  The real-pair split needs Z[k] and Z[400 - k]; with k = k1 + 20 k2 the partner of role k1 is role
// 20 - k1 of the same group (roles 0 and 10 are their own partners).  Phase 2 therefore hands the roles out so that
// partners sit in ADJACENT lanes -- slot s = tid % 20 plays role pair_role(s): (0, 10), (1, 19), (2, 18), ... -- and a
// group starts at an even thread index, so a pair never straddles a warp: the second exchange is one shfl.xor(1) per
// value instead of a trip through shared memory behind two block barriers.  Phase 1 keeps slot == n2 (its sample and
// twiddle accesses stay as they were); only the ROW it writes for output k1 moves to pair_slot(k1), the slot that reads
// it back, so both transposes keep their lane -> address patterns (and their conflict-free paddings).
SD_HD constexpr int pair_role(int s) { return (s & 1) == 0 ? s / 2 : (s == 1 ? 10 : 20 - s / 2); }
SD_HD constexpr int pair_slot(int k1) { return k1 == 0 ? 0 : k1 == 10 ? 1 : k1 < 10 ? 2 * k1 : 2 * (20 - k1) + 1; }

SD_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SD_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SD_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// multiply by -i (forward-transform rotation)
SD_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

// Blackwell (sm_100a) has packed fp32 arithmetic -- add / mul / fma on a register pair in ONE issue slot (SASS FADD2 /
// FMUL2 / FFMA2).  A complex value is exactly such a pair, so the butterflies below issue half the floating-point
// instructions of the scalar form; the kernel is issue-bound (profiles/r01_stft_v5_ncu_summary.json: 578 M
// warp-instructions, 57 % of them FADD/FFMA/FMUL), so that is where the time goes.  Multiplication by -i is folded
// into an FFMA2 with the constant pair (1, -1) / (-1, 1) applied to the swapped operand (x * +-1 is exact, so the
// results are the scalar ones bit for bit).  The host build (tests/cpp/emulate_stft.cpp) keeps the scalar form.
#if defined(__CUDA_ARCH__) && !defined(SD_NO_PACKED_F32)
#define SD_PACKED_F32 1
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 upk2(unsigned long long u) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(u));
    return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(r);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float bx, float by) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(bx, by)));
    return upk2(r);
}
// a * (bx, by) + c, element-wise
__device__ __forceinline__ float2 fma2(float2 a, float bx, float by, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(bx, by)), "l"(pk2(c.x, c.y)));
    return upk2(r);
}
// t + (-i) d  and  t - (-i) d
__device__ __forceinline__ float2 add_mi(float2 t, float2 d) { return fma2(make_float2(d.y, d.x), 1.f, -1.f, t); }
__device__ __forceinline__ float2 sub_mi(float2 t, float2 d) { return fma2(make_float2(d.y, d.x), -1.f, 1.f, t); }

__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 t0 = add2(x0, x2), t1 = sub2(x0, x2), t2 = add2(x1, x3), d = sub2(x1, x3);
    x0 = add2(t0, t2);
    x2 = sub2(t0, t2);
    x1 = add_mi(t1, d);
    x3 = sub_mi(t1, d);
}

__device__ __forceinline__ void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
    const float c = 0.55901699437494742f;   // (cos(2pi/5) - cos(4pi/5)) / 2
    const float s1 = 0.95105651629515357f;  // sin(2pi/5)
    const float s2 = 0.58778525229247313f;  // sin(4pi/5)
    const float2 t1 = add2(x1, x4), t2 = add2(x2, x3), t3 = sub2(x1, x4), t4 = sub2(x2, x3);
    const float2 t5 = add2(t1, t2);
    const float2 m1 = fma2(t5, -0.25f, -0.25f, x0);
    const float2 d = sub2(t1, t2);
    const float2 a1 = fma2(d, c, c, m1), a2 = fma2(d, -c, -c, m1);
    const float2 u1 = fma2(t4, s2, s2, mul2(t3, s1, s1));
    const float2 u2 = fma2(t4, -s1, -s1, mul2(t3, s2, s2));
    x0 = add2(x0, t5);
    x1 = add_mi(a1, u1);
    x4 = sub_mi(a1, u1);
    x2 = add_mi(a2, u2);
    x3 = sub_mi(a2, u2);
}
#else
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
#endif

// element-wise product of two pairs; (a.x + b.x, a.y - b.y) and (a.y + b.y, b.x - a.x) of the real-pair split
#if defined(SD_PACKED_F32)
__device__ __forceinline__ float2 cmul_elem(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(r);
}
__device__ __forceinline__ float2 split_a(float2 a, float2 b) { return fma2(b, 1.f, -1.f, a); }
__device__ __forceinline__ float2 split_b(float2 a, float2 b) {
    return fma2(make_float2(a.y, a.x), 1.f, -1.f, make_float2(b.y, b.x));
}
#else
SD_HD float2 cmul_elem(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
SD_HD float2 split_a(float2 a, float2 b) { return make_float2(a.x + b.x, a.y - b.y); }
SD_HD float2 split_b(float2 a, float2 b) { return make_float2(a.y + b.y, b.x - a.x); }
#endif

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

// Padded staging: 20 extra floats are inserted after every second hop segment of the tile, so frame pair g starts at
// g * (2*kHop + 20) = 340 g = 20 g (mod 32) banks: the lanes of the two or three 20-thread groups that share a warp
// then cover distinct banks -> conflict-free 4-byte reads.  An even segment and the odd one after it stay contiguous
// (320 floats = 1 280 bytes, 16-byte aligned), so a tile is fetched with one bulk (TMA) copy per segment PAIR: nine
// copies instead of eighteen -- the copy instruction takes its operands from uniform registers, so the lanes of the
// issuing warp are served one after the other (ELECT loop in SASS) and every copy costs that warp ~9 instructions
// which the other warps of the CTA wait for at the end-of-tile barrier.  Tiles start at an even frame, so segment
// parity == frame parity.
constexpr int kPadEven = 0;    // after an even hop segment
constexpr int kPadOdd = 20;    // after an odd hop segment
constexpr int kPairStride = 2 * kHop + kPadEven + kPadOdd;  // floats between the first frames of consecutive pairs
// padded position of sample j of the tile
SD_HD constexpr int sig_pos(int j) {
    return j + (kPadEven + kPadOdd) * ((j / kHop) / 2) + kPadEven * ((j / kHop) & 1);
}
SD_HD constexpr int sig_frame_off(int f) { return sig_pos(f * kHop); }  // padded position of frame f's first sample
// extra offset of sample 20*n1 + r of an even / odd frame (the frame crosses two segment boundaries)
SD_HD constexpr int sig_pad_even(int n1) { return (20 * n1) / kHop == 0 ? 0 : (20 * n1) / kHop == 1 ? kPadEven : kPadEven + kPadOdd; }
SD_HD constexpr int sig_pad_odd(int n1) { return (20 * n1) / kHop == 0 ? 0 : (20 * n1) / kHop == 1 ? kPadOdd : kPadEven + kPadOdd; }
// floats needed to stage n samples
SD_HD constexpr int sig_padded_size(int n) { return sig_pos(n - 1) + 1; }

// Twiddle table layout.  Only the five rows k1 = 1, 2, 4, 8, 16 are stored (twiddle_store composes the rest).  A row
// of 20 float2 does not fit the 16 eight-byte banks: in a half-warp that holds lanes r = a..19 of one group and
// r' = 0..a-5 of the next, roles 16..19 would alias roles 0..3.  So roles 0..15 read T0[row*16 + r] and roles 16..19
// read one of four copies T1[row*16 + 4c + (r-16)], c = (a-4)/4, placed on exactly the four banks their half-warp
// leaves free.  kTwRow = 16 for both, so each thread just keeps a base pointer and indexes it with row * kTwRow.
constexpr int kTwRow = 16;
constexpr int kTwRows = 5;                            // row j holds k1 = 1 << j
constexpr int kTwTableUnits = 2 * kTwRows * kTwRow;   // T0 then T1: 160 float2 = 1 280 bytes
SD_HD int tw_thread_offset(int tid) {  // float2 units from the start of the table for thread `tid` of the CTA
    const int g = tid / kRadix, r = tid - g * kRadix;
    if (r < 16) return r;
    const int a = 16 * (tid >> 4) - 20 * g;  // first role of this group inside the thread's half-warp (4, 8, 12, 16)
    return kTwRows * kTwRow + (a - 4) + (r - 16);
}
// fills the table from tw[r*20 + k1] = exp(-2 pi i r k1 / 400); entry index e in [0, kTwTableUnits)
SD_HD int tw_table_source(int e) {  // returns r*20 + k1 of the value stored at table entry e
    const int half = e / (kTwRows * kTwRow), rem = e - half * kTwRows * kTwRow;
    const int row = rem / kTwRow, c = rem - row * kTwRow;
    const int r = half == 0 ? c : 16 + (c & 3);
    return r * 20 + (1 << row);
}

// Multiply the phase-1 outputs by W400^(r*k1), k1 = 0..19, and write them to the transpose buffer.
// The shared-memory pipe is the kernel's busiest unit, so only the five twiddles W^(r*{1,2,4,8,16}) are read from the
// table and the other fourteen are products of at most three of them (<= 3 roundings on a unit-modulus value:
// ~2e-7 relative, far inside the 1e-4 bar): 5 instead of 19 eight-byte shared loads per thread and tile.
// Each product is formed when its bin is reached, while the registers of the bins already stored are free again.
SD_HD void twiddle_store(const float2 (&v)[20], const float2* twp, float2* dst) {
    dst[pair_slot(0) * kXchgRow] = v[dft20_slot(0)];
    const float2 t1 = twp[0 * kTwRow], t2 = twp[1 * kTwRow];
    const float2 t3 = cmul(t1, t2);
    dst[pair_slot(1) * kXchgRow] = cmul(v[dft20_slot(1)], t1);
    dst[pair_slot(2) * kXchgRow] = cmul(v[dft20_slot(2)], t2);
    dst[pair_slot(3) * kXchgRow] = cmul(v[dft20_slot(3)], t3);
    const float2 t4 = twp[2 * kTwRow];
    dst[pair_slot(4) * kXchgRow] = cmul(v[dft20_slot(4)], t4);
    dst[pair_slot(5) * kXchgRow] = cmul(v[dft20_slot(5)], cmul(t4, t1));
    dst[pair_slot(6) * kXchgRow] = cmul(v[dft20_slot(6)], cmul(t4, t2));
    dst[pair_slot(7) * kXchgRow] = cmul(v[dft20_slot(7)], cmul(t4, t3));
    const float2 t8 = twp[3 * kTwRow];
    const float2 t12 = cmul(t8, t4);
    dst[pair_slot(8) * kXchgRow] = cmul(v[dft20_slot(8)], t8);
    dst[pair_slot(9) * kXchgRow] = cmul(v[dft20_slot(9)], cmul(t8, t1));
    dst[pair_slot(10) * kXchgRow] = cmul(v[dft20_slot(10)], cmul(t8, t2));
    dst[pair_slot(11) * kXchgRow] = cmul(v[dft20_slot(11)], cmul(t8, t3));
    dst[pair_slot(12) * kXchgRow] = cmul(v[dft20_slot(12)], t12);
    dst[pair_slot(13) * kXchgRow] = cmul(v[dft20_slot(13)], cmul(t12, t1));
    dst[pair_slot(14) * kXchgRow] = cmul(v[dft20_slot(14)], cmul(t12, t2));
    dst[pair_slot(15) * kXchgRow] = cmul(v[dft20_slot(15)], cmul(t12, t3));
    const float2 t16 = twp[4 * kTwRow];
    dst[pair_slot(16) * kXchgRow] = cmul(v[dft20_slot(16)], t16);
    dst[pair_slot(17) * kXchgRow] = cmul(v[dft20_slot(17)], cmul(t16, t1));
    dst[pair_slot(18) * kXchgRow] = cmul(v[dft20_slot(18)], cmul(t16, t2));
    dst[pair_slot(19) * kXchgRow] = cmul(v[dft20_slot(19)], cmul(t16, t3));
}

// phase 1 with the window (wtab[20*n1 + r]) in a shared table and the twiddles behind a per-thread base pointer
// Window table element: the plain value, or (SD_WINDOW_PAIRS) the pair (w, w) so that the two frames are windowed by
// one packed multiply at the price of 8-byte table reads.
#if defined(SD_WINDOW_PAIRS)
typedef float2 wtab_t;
SD_HD wtab_t wtab_make(float w) { return make_float2(w, w); }
SD_HD float wtab_value(wtab_t w) { return w.x; }
SD_HD float2 wtab_apply(float a, float b, wtab_t w) { return cmul_elem(make_float2(a, b), w); }
#else
typedef float wtab_t;
SD_HD wtab_t wtab_make(float w) { return w; }
SD_HD float wtab_value(wtab_t w) { return w; }
SD_HD float2 wtab_apply(float a, float b, wtab_t w) { return make_float2(a * w, b * w); }
#endif

SD_HD void stft_phase1_tab(const float* sig, int fa_off, int fb_off, const wtab_t* wtab, const float2* twp, int g, int r,
                           float2* xchg) {
    // Frame B starts one hop (160 = 8 * 20 samples) after frame A, so B's sample n1 IS A's sample n1 + 8 (the same
    // staged float) for n1 < 12: 28 shared-memory loads instead of 40.
    float a[20], b[8];
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) a[n1] = sig[fa_off + 20 * n1 + r + sig_pad_even(n1)];
#pragma unroll
    for (int n1 = 12; n1 < 20; ++n1) b[n1 - 12] = sig[fb_off + 20 * n1 + r + sig_pad_odd(n1)];
    float2 v[20];
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) v[n1] = wtab_apply(a[n1], n1 < 12 ? a[n1 + 8] : b[n1 - 12], wtab[20 * n1 + r]);
    dft20(v);
    twiddle_store(v, twp, xchg + g * kGroupStride + r);
}

// phase 1 for the reference's window (periodic Hamming, pre-scaled by 1/2) without a window table:
//   w[n] / 2 = 0.27 - 0.23 cos(2 pi n / 400),  n = 20 n1 + r  =>  cos(theta_r + n1 pi/10) = cr C[n1] - sr S[n1]
// with (cr, sr) = (cos, sin)(2 pi r / 400) in two registers of the thread and C, S compile-time constants: two FFMA per
// sample instead of a shared-memory load (the shared-memory pipe is the kernel's limiter, the FMA pipe has room).
// The values differ from at::hamming_window's fp32 table by a few ulp (~1e-7 relative), far inside the 1e-4 bar.
SD_HD void stft_phase1_hamming(const float* sig, int fa_off, int fb_off, float cr, float sr, const float2* twp, int g,
                               int r, float2* xchg) {
    // -0.23 cos(n1 pi / 10) and 0.23 sin(n1 pi / 10), n1 = 0..19
    constexpr float K1[20] = {-0.23f, -0.21874299874788533f, -0.1860739087062379f, -0.13519060802726882f,
                              -0.0710739087062379f, 0.0f, 0.0710739087062379f, 0.13519060802726882f,
                              0.1860739087062379f, 0.21874299874788533f, 0.23f, 0.21874299874788533f,
                              0.1860739087062379f, 0.13519060802726882f, 0.0710739087062379f, 0.0f,
                              -0.0710739087062379f, -0.13519060802726882f, -0.1860739087062379f,
                              -0.21874299874788533f};
    constexpr float K2[20] = {0.0f, 0.0710739087062379f, 0.13519060802726882f, 0.1860739087062379f,
                              0.21874299874788533f, 0.23f, 0.21874299874788533f, 0.1860739087062379f,
                              0.13519060802726882f, 0.0710739087062379f, 0.0f, -0.0710739087062379f,
                              -0.13519060802726882f, -0.1860739087062379f, -0.21874299874788533f, -0.23f,
                              -0.21874299874788533f, -0.1860739087062379f, -0.13519060802726882f,
                              -0.0710739087062379f};
    float a[20], b[8];
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) a[n1] = sig[fa_off + 20 * n1 + r + sig_pad_even(n1)];
#pragma unroll
    for (int n1 = 12; n1 < 20; ++n1) b[n1 - 12] = sig[fb_off + 20 * n1 + r + sig_pad_odd(n1)];
    float2 v[20];
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
        const float w = fmaf(K1[n1], cr, fmaf(K2[n1], sr, 0.27f));
#if defined(SD_PACKED_F32) && defined(SD_WINDOW_MUL2)
        v[n1] = mul2(make_float2(a[n1], n1 < 12 ? a[n1 + 8] : b[n1 - 12]), w, w);
#else
        v[n1] = make_float2(a[n1] * w, (n1 < 12 ? a[n1 + 8] : b[n1 - 12]) * w);
#endif
    }
    dft20(v);
    twiddle_store(v, twp, xchg + g * kGroupStride + r);
}

// ---- Kaldi-compatible per-frame conditioning (north_star bullet 1) -------------------------------------------------
// kaldi::ProcessWindow (feature-window.cc) as restated by torchaudio.compliance.kaldi._get_window: optional DC removal
// (subtract the mean of the frame), pre-emphasis y[n] = x[n] - c x[n-1] with y[0] = x[0] - c x[0], then the window
// (povey).  With DC removal the two steps collapse to y[n] = (x[n] - c x[n-1]) - (1 - c) * mean for every n.
// extra padded offset of sample o (0..399) of an even / odd frame
SD_HD constexpr int sig_pad_even_at(int o) { return o / kHop == 0 ? 0 : o / kHop == 1 ? kPadEven : kPadEven + kPadOdd; }
SD_HD constexpr int sig_pad_odd_at(int o) { return o / kHop == 0 ? 0 : o / kHop == 1 ? kPadOdd : kPadEven + kPadOdd; }

// sum of this thread's 20 samples of frame A (x) and frame B (y); the group adds its 20 partials for the frame mean
SD_HD float2 stft_frame_partial_sums(const float* sig, int fa_off, int fb_off, int r) {
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
        const int o = 20 * n1 + r;
        sa += sig[fa_off + o + sig_pad_even(n1)];
        sb += sig[fb_off + o + sig_pad_odd(n1)];
    }
    return make_float2(sa, sb);
}

// phase 1 with pre-emphasis coefficient c and the per-frame offsets dc = (1 - c) * mean (zero when DC removal is off)
SD_HD void stft_phase1_kaldi(const float* sig, int fa_off, int fb_off, const wtab_t* wtab, const float2* twp, int g, int r,
                             float2* xchg, float c, float2 dc) {
    float2 v[20];
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
        const int o = 20 * n1 + r;
        const float w = wtab_value(wtab[o]);
        const int pa = fa_off + o + sig_pad_even(n1), pb = fb_off + o + sig_pad_odd(n1);
        // previous sample: one float to the left, except across a hop-segment boundary (r == 0 at n1 = 8, 16) and
        // at the start of the frame, where Kaldi uses the first sample itself
        int qa = pa - 1, qb = pb - 1;
        if (r == 0) {
            qa = n1 > 0 ? fa_off + (20 * n1 - 1) + sig_pad_even_at(n1 > 0 ? 20 * n1 - 1 : 0) : pa;
            qb = n1 > 0 ? fb_off + (20 * n1 - 1) + sig_pad_odd_at(n1 > 0 ? 20 * n1 - 1 : 0) : pb;
        }
        const float xa = sig[pa] - c * sig[qa] - dc.x, xb = sig[pb] - c * sig[qb] - dc.y;
        v[n1] = make_float2(xa * w, xb * w);
    }
    dft20(v);
    twiddle_store(v, twp, xchg + g * kGroupStride + r);
}

// phase 2 of the kernels: slot r reads the row that phase 1 wrote for output k1 = pair_role(r) and transforms it:
// afterwards v[dft20_slot(k2)] = Z[pair_role(r) + 20 k2]
SD_HD void stft_phase2_load(const float2* xchg, int g, int r, float2 (&v)[20]) {
    const float2* src = xchg + g * kGroupStride + r * kXchgRow;
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) v[n2] = src[n2];
    dft20(v);
}
// ... (barrier) ... then write Z[r + 20*k2] into the same buffer
SD_HD void stft_phase2_store(const float2 (&v)[20], int g, int r, float2* zbuf) {
    float2* dst = zbuf + g * kGroupStride + r;
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) dst[20 * k2] = v[dft20_slot(k2)];
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

// Phase 3 for a spectrum that is already scaled by 1/2 (the factor is folded into the window table):
//   A[k] = Z[k] + conj Z[400-k],  B[k] = -i (Z[k] - conj Z[400-k]).
// Thread r handles k = r + 20 m, m = 0..9 (k = 0 and k = 200 are the two extra bins of thread 0), so the partner
// index 400 - k is a compile-time offset from a per-thread base and there is no per-bin branch.
template <bool HAS_A, bool HAS_B>
SD_HD void stft_phase3_fast(const float2* zbuf, int g, int r, float* outA, float* outB) {
    const float2* z = zbuf + g * kGroupStride;
    float2* oa = reinterpret_cast<float2*>(outA);
    float2* ob = reinterpret_cast<float2*>(outB);
    if (r == 0) {  // purely real bins: partner of Z[0] is Z[0], of Z[200] is Z[200]
        const float2 z0 = z[0], zn = z[200];
        if (HAS_A) {
            oa[0] = make_float2(z0.x + z0.x, 0.f);
            oa[200] = make_float2(zn.x + zn.x, 0.f);
        }
        if (HAS_B) {
            ob[0] = make_float2(z0.y + z0.y, 0.f);
            ob[200] = make_float2(zn.y + zn.y, 0.f);
        }
    }
    const float2* zk = z + r;
    const float2* zm = z + kNfft - r;
#pragma unroll
    for (int m = 0; m < 10; ++m) {
        if (m == 0 && r == 0) continue;
        const float2 a = zk[20 * m];
        const float2 b = zm[-20 * m];
        if (HAS_A) oa[r + 20 * m] = make_float2(a.x + b.x, a.y - b.y);
        if (HAS_B) ob[r + 20 * m] = make_float2(a.y + b.y, b.x - a.x);
    }
}

// ---- compact second exchange -----------------------------------------------------------------------------
// After phase 2 thread r holds Z[r + 20*k2] (k2 = 0..19) in registers.  The one-sided bins k <= 200 of the two
// real frames need Z[k] (own registers, k2 = m <= 9) and Z[400 - k] (held by thread 20 - r as k2 = 19 - m >= 10).
// So only the upper half Z[j], j >= 200, ever crosses threads: thread r publishes its k2 = 10..19 values
// (and Z[0] for the k = 0 bin) into a 201-entry buffer indexed by j - 200.
constexpr int kZStride = 212;  // float2 units per group (201 used); 212 - 20 = 12 * 16 keeps the stores conflict-free

SD_HD void stft_publish_upper(const float2 (&v)[20], int g, int r, float2* zup, int zstride = kZStride) {
    float2* dst = zup + g * zstride + r;
#pragma unroll
    for (int k2 = 10; k2 < 20; ++k2) dst[20 * (k2 - 10)] = v[dft20_slot(k2)];  // Z[200 + r + 20 (k2 - 10)]
    if (r == 0) dst[200] = v[dft20_slot(0)];                                   // Z[400] == Z[0] (periodicity)
}

// Spectrum scaled by 1/2 (folded into the window): A[k] = Z[k] + conj Z[400-k], B[k] = -i (Z[k] - conj Z[400-k]).
template <bool HAS_A, bool HAS_B>
SD_HD void stft_split_store(const float2 (&v)[20], const float2* zup, int g, int r, float* outA, float* outB,
                            int zstride = kZStride) {
    const float2* zu = zup + g * zstride;  // zu[j - 200] = Z[j]
    float2* oa = reinterpret_cast<float2*>(outA);
    float2* ob = reinterpret_cast<float2*>(outB);
    const float2* zm = zu + 200 - r;  // Z[400 - (r + 20 m)] = zu[200 - r - 20 m]
#pragma unroll
    for (int m = 0; m < 10; ++m) {
        const float2 a = v[dft20_slot(m)];
        const float2 b = zm[-20 * m];
        if (HAS_A) oa[r + 20 * m] = split_a(a, b);
        if (HAS_B) ob[r + 20 * m] = split_b(a, b);
    }
    if (r == 0) {  // k = 200: partner of Z[200] is itself
        const float2 a = v[dft20_slot(10)];
        if (HAS_A) oa[200] = make_float2(a.x + a.x, 0.f);
        if (HAS_B) ob[200] = make_float2(a.y + a.y, 0.f);
    }
}

// ---- second exchange between pair lanes -------------------------------------------------------------------
// Slot s (role k1 = pair_role(s)) owns bins k = k1 + 20 m, m = 0..9 (plus k = 200 for role 0).  Z[400 - k] is the
// pair lane's k2 = 19 - m value -- the same register index on both sides, so both lanes send v[dft20_slot(19 - m)] and
// receive the other's.  Roles 0 and 10 are their own partners: role 10 needs its own k2 = 19 - m value, role 0 its own
// k2 = 20 - m (Z[400 - 20 m]); both "receive from themselves" (role 0 after selecting the other register).
// `partner(mine, index)` returns `mine` in the self-paired slots (s < 2) and the pair lane's v[index] otherwise: one
// indexed shuffle on the device (source lane = own lane or lane ^ 1), an array access in the host emulation.
// Spectrum scaled by 1/2 (folded into the window): A[k] = Z[k] + conj Z[400-k], B[k] = -i (Z[k] - conj Z[400-k]).
// outA / outB must both be writable (frames past the end of the item go to a scratch row): the exchange has to run in
// every lane, and unconditional stores keep the loop free of branches.
template <typename Partner>
SD_HD void stft_split_store_pair(const float2 (&v)[20], int s, float* outA, float* outB, Partner partner) {
    const int k1 = pair_role(s);
    float2* oa = reinterpret_cast<float2*>(outA);
    float2* ob = reinterpret_cast<float2*>(outB);
#pragma unroll
    for (int m = 0; m < 10; ++m) {
        const float2 a = v[dft20_slot(m)];
        const float2 mine = s == 0 ? v[dft20_slot((20 - m) % 20)] : v[dft20_slot(19 - m)];
        const float2 b = partner(mine, dft20_slot(19 - m));
        oa[k1 + 20 * m] = split_a(a, b);
        ob[k1 + 20 * m] = split_b(a, b);
    }
    if (s == 0) {  // k = 200: partner of Z[200] is itself
        const float2 a = v[dft20_slot(10)];
        oa[200] = make_float2(a.x + a.x, 0.f);
        ob[200] = make_float2(a.y + a.y, 0.f);
    }
}

// |A[k]|^2 and |B[k]|^2 of the slot's bins into pa / pb (fbank front-end), same exchange
template <typename Partner>
SD_HD void stft_split_power_pair(const float2 (&v)[20], int s, float* pa, float* pb, Partner partner) {
    const int k1 = pair_role(s);
#pragma unroll
    for (int m = 0; m < 10; ++m) {
        const float2 a = v[dft20_slot(m)];
        const float2 mine = s == 0 ? v[dft20_slot((20 - m) % 20)] : v[dft20_slot(19 - m)];
        const float2 c = partner(mine, dft20_slot(19 - m));
        const float ar = a.x + c.x, ai = a.y - c.y, br = a.y + c.y, bi = c.x - a.x;
        pa[k1 + 20 * m] = ar * ar + ai * ai;
        pb[k1 + 20 * m] = br * br + bi * bi;
    }
    if (s == 0) {
        const float2 a = v[dft20_slot(10)];
        pa[200] = 4.f * a.x * a.x;
        pb[200] = 4.f * a.y * a.y;
    }
}

#if defined(__CUDACC__)
struct PairShuffle {  // the pair lane's copy of a register through the warp; self-paired slots read their own lane
    int src;
    __device__ __forceinline__ explicit PairShuffle(int s) {
        const int lane = (int)(threadIdx.x & 31);
        src = s < 2 ? lane : lane ^ 1;
    }
    __device__ __forceinline__ float2 operator()(float2 mine, int) const {
        return make_float2(__shfl_sync(0xffffffffu, mine.x, src), __shfl_sync(0xffffffffu, mine.y, src));
    }
};
#endif

}  // namespace sdb
