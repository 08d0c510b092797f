/* TEST INFRASTRUCTURE ONLY -- plain-C CPU restatement of the reference hot path.
 * See sd_oracle.h for scope, citations and the parity-pinning statement.
 * Build: gcc -O2 -std=c11 -ffp-contract=off (no FMA contraction: the reference is built for
 * baseline x86-64, which has none, and the linkage compares doubles for exact equality).
 */
#include "sd_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ a5 */

/* SD:260-272.  "Exact half" (within DBL_EPSILON) rounds to even, everything else rounds half away. */
int sdo_np_rint(double v) {
    double sgn = v > 0 ? 1.0 : -1.0;
    double frac_off = v - (double)(int)v - 0.5 * sgn;
    if (fabs(frac_off) < DBL_EPSILON) {
        int r = (int)round(v);
        return (r % 2 == 0) ? r : r - (v > 0 ? 1 : -1);
    }
    return (int)round(v);
}

/* SD:1084-1090: negative positions clamp to frame 0 before rounding. */
long sdo_closest_frame(double sw_start, double sw_step, double sw_duration, double t) {
    double pos = (t - sw_start - .5 * sw_duration) / sw_step;
    if (pos < 0.0) pos = 0.0;
    return (long)(size_t)sdo_np_rint(pos);
}

/* ------------------------------------------------------------------ a1/a2 */

/* at::hamming_window(n) with periodic=true in fp32: arange(n+1) * (2pi/n) -> cos -> *(-0.46) -> +0.54,
 * every step rounded to float, first n entries kept (call site SD:2007). */
void sdo_hamming_window_f32(int n, float* w) {
    const float scale = (float)(M_PI * 2.0 / (double)n);
    const float nbeta = (float)(-0.46);
    const float alpha = (float)0.54;
    for (int i = 0; i < n; ++i) {
        float a = (float)i * scale;
        float c = cosf(a);
        float m = c * nbeta;
        w[i] = m + alpha;
    }
}

static void dft_frame(const float* x, int L, int n_fft, int hop, const float* window, const double* cs,
                      const double* sn, int t, float* out /*[bins][2]*/) {
    const int half = n_fft / 2;
    const int bins = half + 1;
    double* fr = (double*)malloc(sizeof(double) * (size_t)n_fft);
    for (int n = 0; n < n_fft; ++n) {
        long src = (long)t * hop - half + n; /* centre padding with zeros, SD:2008 (center=true, "constant") */
        double s = (src >= 0 && src < L) ? (double)x[src] : 0.0;
        fr[n] = s * (double)window[n];
    }
    for (int f = 0; f < bins; ++f) {
        double re = 0.0, im = 0.0;
        for (int n = 0; n < n_fft; ++n) {
            int k = (int)(((long)f * n) % n_fft);
            re += fr[n] * cs[k];
            im -= fr[n] * sn[k];
        }
        out[2 * f] = (float)re; /* fp64 result stored as fp32, SD:2031 */
        out[2 * f + 1] = (float)im;
    }
    free(fr);
}

static void twiddles(int n_fft, double** cs, double** sn) {
    *cs = (double*)malloc(sizeof(double) * (size_t)n_fft);
    *sn = (double*)malloc(sizeof(double) * (size_t)n_fft);
    for (int k = 0; k < n_fft; ++k) {
        (*cs)[k] = cos(2.0 * M_PI * (double)k / (double)n_fft);
        (*sn)[k] = sin(2.0 * M_PI * (double)k / (double)n_fft);
    }
}

int sdo_stft_frames(const float* x, int L, int n_fft, int hop, const float* window, int t0, int t1, float* out) {
    double *cs, *sn;
    const int bins = n_fft / 2 + 1;
    twiddles(n_fft, &cs, &sn);
    for (int t = t0; t < t1; ++t) dft_frame(x, L, n_fft, hop, window, cs, sn, t, out + (size_t)(t - t0) * bins * 2);
    free(cs);
    free(sn);
    return 0;
}

int sdo_stft(const float* wav, int B, int L, int n_fft, int hop, const float* window, float* out) {
    const int T = 1 + L / hop;
    const int bins = n_fft / 2 + 1;
    for (int b = 0; b < B; ++b)
        sdo_stft_frames(wav + (size_t)b * L, L, n_fft, hop, window, 0, T, out + (size_t)b * T * bins * 2);
    return 0;
}

/* ------------------------------------------------------------------ a3 (PARITY UNPINNED) */

/* speechbrain 0.5.14 Filterbank (triangular, f_min 0, f_max sr/2): mel points linspace(mel(fmin), mel(fmax),
 * n_mels+2); centre = hz[1..n_mels]; band = hz[i+1]-hz[i] (left spacing for both slopes);
 * W[f][m] = max(0, min((f-fc)/b + 1, -(f-fc)/b + 1)), all in fp32 like the torch module. */
void sdo_mel_matrix(int n_bins, int n_mels, double f_min, double f_max, int sample_rate, float* W) {
    int np = n_mels + 2;
    float* hz = (float*)malloc(sizeof(float) * (size_t)np);
    float mlo = (float)(2595.0 * log10(1.0 + f_min / 700.0));
    float mhi = (float)(2595.0 * log10(1.0 + f_max / 700.0));
    for (int i = 0; i < np; ++i) {
        float mel = mlo + (mhi - mlo) * (float)i / (float)(np - 1);
        hz[i] = 700.0f * (powf(10.0f, mel / 2595.0f) - 1.0f);
    }
    for (int f = 0; f < n_bins; ++f) {
        float freq = (float)(sample_rate / 2) * (float)f / (float)(n_bins - 1);
        for (int m = 0; m < n_mels; ++m) {
            float fc = hz[m + 1];
            float band = hz[m + 1] - hz[m];
            float slope = (freq - fc) / band;
            float l = slope + 1.0f, r = -slope + 1.0f;
            float v = l < r ? l : r;
            W[(size_t)f * n_mels + m] = v > 0.0f ? v : 0.0f;
        }
    }
    free(hz);
}

int sdo_fbank_tail(const float* stft, int B, int T, int n_bins, int n_mels, const float* W, const float* wav_lens,
                   float* out) {
    double* acc = (double*)malloc(sizeof(double) * (size_t)n_mels);
    float* pw = (float*)malloc(sizeof(float) * (size_t)n_bins);
    for (int b = 0; b < B; ++b) {
        float* o = out + (size_t)b * T * n_mels;
        float mx = -INFINITY;
        for (int t = 0; t < T; ++t) {
            const float* s = stft + ((size_t)b * T + t) * n_bins * 2;
            for (int f = 0; f < n_bins; ++f) pw[f] = s[2 * f] * s[2 * f] + s[2 * f + 1] * s[2 * f + 1];
            for (int m = 0; m < n_mels; ++m) {
                double a = 0.0; /* fp64 accumulate: the oracle is the accuracy yardstick, torch.matmul is fp32 */
                for (int f = 0; f < n_bins; ++f) a += (double)pw[f] * (double)W[(size_t)f * n_mels + m];
                float v = (float)a;
                if (v < 1e-10f) v = 1e-10f;
                float db = 10.0f * log10f(v); /* - 10*log10(max(amin, ref=1.0)) == 0 */
                o[(size_t)t * n_mels + m] = db;
                if (db > mx) mx = db;
            }
        }
        float floor_db = mx - 80.0f; /* top_db over the whole utterance (amax over time and mel) */
        for (size_t i = 0; i < (size_t)T * n_mels; ++i)
            if (o[i] < floor_db) o[i] = floor_db;
        /* MyNormalization, threeModel.py:333-369: mean over the first round(len*T) frames, std-norm off */
        long n = lrintf(wav_lens[b] * (float)T); /* torch.round = half to even */
        if (n > T) n = T;
        for (int m = 0; m < n_mels; ++m) acc[m] = 0.0;
        for (long t = 0; t < n; ++t)
            for (int m = 0; m < n_mels; ++m) acc[m] += o[(size_t)t * n_mels + m];
        for (int m = 0; m < n_mels; ++m) {
            float mean = n > 0 ? (float)(acc[m] / (double)n) : NAN;
            for (int t = 0; t < T; ++t) o[(size_t)t * n_mels + m] -= mean;
        }
    }
    free(acc);
    free(pw);
    return 0;
}

/* ------------------------------------------------------------------ a4 */

long sdo_aggregate(const double* scores, int C, int F, int K, double sf_start, double sf_step, double sf_duration,
                   long sf_num_samples, double pf_step, double pf_duration, int hamming, double missing,
                   int skip_average, double epsilon, double* out, long cap_rows, double* count_out, double* mask_out,
                   double* post_out) {
    /* SD:1232-1234: grid of output frames starts at the chunk window start, step/duration of pre_frames */
    double target = sf_start + sf_duration + (double)(size_t)(C - 1) * sf_step;
    long NF = sdo_closest_frame(sf_start, pf_step, pf_duration, target) + 1;
    if (NF > cap_rows) return -1;
    size_t tot = (size_t)NF * K;
    double* cnt = (double*)calloc(tot, sizeof(double));
    double* msk = (double*)calloc(tot, sizeof(double));
    double* wgt = NULL;
    for (size_t i = 0; i < tot; ++i) out[i] = 0.0;
    if (hamming) { /* no reference implementation (SD:1214 asserts); np.hamming(F) per pyannote inference.py */
        wgt = (double*)malloc(sizeof(double) * (size_t)F);
        for (int j = 0; j < F; ++j) wgt[j] = F > 1 ? 0.54 - 0.46 * cos(2.0 * M_PI * j / (double)(F - 1)) : 1.0;
    }
    double start = sf_start; /* SD:1248-1253: running sum, not i*step */
    for (int i = 0; i < C; ++i) {
        long s0 = sdo_closest_frame(sf_start, pf_step, pf_duration, start);
        start += sf_step;
        for (int j = 0; j < F; ++j) {
            long row = s0 + j;
            if (row >= NF) break; /* reference indexes past the end here (UB); numpy slicing clips */
            for (int k = 0; k < K; ++k) {
                double v = scores[((size_t)i * F + j) * K + k];
                double m = 1.0;
                if (isnan(v)) { /* SD:1191-1204 */
                    m = 0.0;
                    v = 0.0;
                }
                double w = wgt ? wgt[j] : 1.0;
                size_t o = (size_t)row * K + k;
                if (wgt) {
                    out[o] += v * m * w;
                    cnt[o] += m * w;
                } else {
                    out[o] += v * m;
                    cnt[o] += m;
                }
                if (m > msk[o]) msk[o] = m;
            }
        }
    }
    if (count_out) memcpy(count_out, cnt, tot * sizeof(double));
    if (mask_out) memcpy(mask_out, msk, tot * sizeof(double));
    if (!skip_average)
        for (size_t o = 0; o < tot; ++o) out[o] /= (cnt[o] > epsilon ? cnt[o] : epsilon); /* SD:1288 */
    for (size_t o = 0; o < tot; ++o)
        if (fabs(msk[o]) < DBL_EPSILON) out[o] = missing; /* SD:1298-1307 */
    if (post_out) { /* SD:1278-1281 */
        post_out[0] = sf_start;
        post_out[1] = pf_step;
        post_out[2] = pf_duration;
        post_out[3] = (double)sf_num_samples;
    }
    free(cnt);
    free(msk);
    free(wgt);
    return NF;
}

/* ------------------------------------------------------------------ a6 */

int sdo_binarize(const float* scores, int C, int F, int K, double onset, int initial_state, double* out) {
    /* SD:1565-1639 collapses to: value of (s > onset) at the last frame <= t whose score is not within
     * DBL_EPSILON of onset; initial_state before the first such frame.  Rows are (chunk, class). */
    for (int c = 0; c < C; ++c)
        for (int k = 0; k < K; ++k) {
            int have = 0, state = initial_state ? 1 : 0;
            for (int t = 0; t < F; ++t) {
                double s = (double)scores[((size_t)c * F + t) * K + k]; /* float -> double, SD:1526 */
                int on = s > onset;
                int defined = !(fabs(s - onset) < DBL_EPSILON);
                if (defined) {
                    have = 1;
                    state = on;
                }
                (void)have;
                out[((size_t)c * F + t) * K + k] = state ? 1.0 : 0.0;
            }
        }
    return 0;
}

/* ------------------------------------------------------------------ a7 */

long sdo_trim(const double* binarized, int C, int F, int K, double left, double right, double bt_start, double bt_step,
              double bt_duration, double* out, double* tw_out) {
    long nl = (long)floor((double)F * left); /* floor, not round: SD:1755-1758 */
    long nr = (long)floor((double)F * right);
    long Ft = F - nl - nr;
    if (out)
        for (int c = 0; c < C; ++c)
            for (long j = 0; j < Ft; ++j)
                for (int k = 0; k < K; ++k)
                    out[((size_t)c * Ft + j) * K + k] = binarized[((size_t)c * F + (j + nl)) * K + k];
    if (tw_out) { /* SD:1776-1779 */
        tw_out[0] = bt_start + left * bt_duration;
        tw_out[1] = bt_step;
        tw_out[2] = (1 - left - right) * bt_duration;
        tw_out[3] = (double)Ft;
    }
    return Ft;
}

long sdo_speaker_count(const double* binarized, int C, int F, int K, double chunk_step, double chunk_duration,
                       double pf_step, double pf_duration, int* out, long cap, double* count_frames_out) {
    double tw[4];
    long Ft = sdo_trim(binarized, C, F, K, 0.1, 0.1, 0.0, chunk_step, chunk_duration, NULL, tw);
    double* trimmed = (double*)malloc(sizeof(double) * (size_t)C * Ft * K);
    sdo_trim(binarized, C, F, K, 0.1, 0.1, 0.0, chunk_step, chunk_duration, trimmed, tw);
    double* sum = (double*)malloc(sizeof(double) * (size_t)C * Ft);
    for (size_t r = 0; r < (size_t)C * Ft; ++r) { /* SD:1701-1714 */
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += trimmed[r * K + k];
        sum[r] = s;
    }
    double target = tw[0] + tw[2] + (double)(size_t)(C - 1) * tw[1];
    long NF = sdo_closest_frame(tw[0], pf_step, pf_duration, target) + 1;
    long res = -1;
    if (NF <= cap) {
        double* agg = (double*)malloc(sizeof(double) * (size_t)NF);
        res = sdo_aggregate(sum, C, (int)Ft, 1, tw[0], tw[1], tw[2], (long)tw[3], pf_step, pf_duration, 0, 0.0, 0,
                            DBL_EPSILON, agg, NF, NULL, NULL, count_frames_out); /* SD:1719-1720 */
        for (long i = 0; i < NF; ++i) out[i] = sdo_np_rint(agg[i]); /* SD:1731-1735 */
        free(agg);
    }
    free(sum);
    free(trimmed);
    return res;
}

int sdo_clean_segmentations(const double* binarized, int C, int F, int K, double* out) {
    for (size_t r = 0; r < (size_t)C * F; ++r) { /* SD:720-740: keep frames where fewer than two are active */
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += binarized[r * K + k];
        for (int k = 0; k < K; ++k) out[r * K + k] = s < 2.0 ? binarized[r * K + k] : 0.0;
    }
    return 0;
}

/* ------------------------------------------------------------------ a9/a10 */

int sdo_normalize_embeddings(double* x, int N, int D) {
    for (int i = 0; i < N; ++i) {
        double* r = x + (size_t)i * D;
        double ss = 0.0;
        for (int k = 0; k < D; ++k) ss += r[k] * r[k];
        double norm = (double)(float)sqrt(ss); /* L2Norm returns float, SD:332-339 */
        if (norm != 0.0)
            for (int k = 0; k < D; ++k) r[k] /= norm;
    }
    return 0;
}

static double euclid(const double* a, const double* b, int D) {
    double s = 0.0;
    for (int k = 0; k < D; ++k) { /* CL:408-415: sequential, mul then add */
        double d = a[k] - b[k];
        s += d * d;
    }
    return sqrt(s);
}

int sdo_pdist(const double* x, int N, int D, double* out) {
    size_t p = 0;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) out[p++] = euclid(x + (size_t)i * D, x + (size_t)j * D, D);
    return 0;
}

/* ------------------------------------------------------------------ a11 */

/* Indexed binary min-heap with the exact comparison rules of CL:28-119 (ties never swap). */
typedef struct {
    int* pos_of;  /* key -> slot */
    int* key_at;  /* slot -> key */
    double* val;  /* slot -> value */
    int n;
} mheap;

static void hswap(mheap* h, int a, int b) {
    double tv = h->val[a];
    h->val[a] = h->val[b];
    h->val[b] = tv;
    int ka = h->key_at[a], kb = h->key_at[b];
    h->key_at[a] = kb;
    h->key_at[b] = ka;
    h->pos_of[ka] = b;
    h->pos_of[kb] = a;
}
static void hdown(mheap* h, int i) {
    for (int c = 2 * i + 1; c < h->n; c = 2 * i + 1) {
        if (c + 1 < h->n && h->val[c + 1] < h->val[c]) ++c;
        if (!(h->val[i] > h->val[c])) break;
        hswap(h, i, c);
        i = c;
    }
}
static void hup(mheap* h, int i) {
    while (i > 0) {
        int p = (i - 1) >> 1;
        if (!(h->val[p] > h->val[i])) break;
        hswap(h, i, p);
        i = p;
    }
}
static void hset(mheap* h, int key, double v) { /* CL:109-118 */
    int i = h->pos_of[key];
    double old = h->val[i];
    h->val[i] = v;
    if (v < old)
        hup(h, i);
    else
        hdown(h, i);
}

static int64_t cidx(int64_t n, int64_t i, int64_t j) { /* CL:236-242, in 64-bit (the reference's int overflows) */
    if (i > j) {
        int64_t t = i;
        i = j;
        j = t;
    }
    return n * i - (i * (i + 1) / 2) + (j - i - 1);
}

static double centroid_update(double dxi, double dyi, double dxy, int sx, int sy) { /* CL:250-256 */
    return sqrt((((sx * dxi * dxi) + (sy * dyi * dyi)) - (sx * sy * dxy * dxy) / (sx + sy)) / (sx + sy));
}

static void row_min(int n, const double* D, const int* size, int x, int* arg, double* val) { /* CL:259-276 */
    double best = INFINITY;
    int y = -1;
    for (int i = x + 1; i < n; ++i) {
        if (!size[i]) continue;
        double d = D[cidx(n, x, i)];
        if (d < best) {
            best = d;
            y = i;
        }
    }
    *arg = y;
    *val = best;
}

int sdo_linkage_condensed(const double* dists, int n, double* Z) {
    if (n < 2) return 0;
    size_t np = (size_t)n * (n - 1) / 2;
    double* D = (double*)malloc(sizeof(double) * np);
    memcpy(D, dists, sizeof(double) * np);
    int* size = (int*)malloc(sizeof(int) * (size_t)n);
    int* cid = (int*)malloc(sizeof(int) * (size_t)n);
    int* nbr = (int*)malloc(sizeof(int) * (size_t)n);
    double* lb = (double*)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        size[i] = 1;
        cid[i] = i;
    }
    for (int x = 0; x < n - 1; ++x) row_min(n, D, size, x, &nbr[x], &lb[x]);

    mheap h;
    h.n = n - 1;
    h.pos_of = (int*)malloc(sizeof(int) * (size_t)n);
    h.key_at = (int*)malloc(sizeof(int) * (size_t)n);
    h.val = (double*)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < h.n; ++i) {
        h.pos_of[i] = h.key_at[i] = i;
        h.val[i] = lb[i];
    }
    for (int i = h.n / 2; i >= 0; --i) hdown(&h, i); /* CL:94-96 */

    for (int k = 0; k < n - 1; ++k) {
        int x = 0, y = 0;
        double dist = 0.0;
        for (int tries = 0; tries < n - k; ++tries) { /* CL:323-339: pop with lazy revalidation */
            x = h.key_at[0];
            dist = h.val[0];
            y = nbr[x];
            if (dist == D[cidx(n, x, y)]) break;
            row_min(n, D, size, x, &y, &dist);
            nbr[x] = y;
            lb[x] = dist;
            hset(&h, x, dist);
        }
        hswap(&h, 0, h.n - 1); /* remove_min, CL:103-107 */
        h.n -= 1;
        hdown(&h, 0);

        int ix = cid[x], iy = cid[y], nx = size[x], ny = size[y];
        if (ix > iy) {
            int t = ix;
            ix = iy;
            iy = t;
        }
        Z[4 * (size_t)k + 0] = ix;
        Z[4 * (size_t)k + 1] = iy;
        Z[4 * (size_t)k + 2] = dist;
        Z[4 * (size_t)k + 3] = nx + ny;
        size[x] = 0;
        size[y] = nx + ny;
        cid[y] = n + k;

        for (int z = 0; z < n; ++z) { /* CL:361-370 */
            if (!size[z] || z == y) continue;
            int64_t zy = cidx(n, z, y);
            D[zy] = centroid_update(D[cidx(n, z, x)], D[zy], dist, nx, ny);
        }
        for (int z = 0; z < x; ++z) /* CL:374-378 */
            if (size[z] > 0 && nbr[z] == x) nbr[z] = y;
        for (int z = 0; z < y; ++z) { /* CL:381-392 */
            if (!size[z]) continue;
            double d = D[cidx(n, z, y)];
            if (d < lb[z]) {
                nbr[z] = y;
                lb[z] = d;
                hset(&h, z, d);
            }
        }
        if (y < n - 1) { /* CL:395-404 */
            int z;
            double d;
            row_min(n, D, size, y, &z, &d);
            if (z != -1) {
                nbr[y] = z;
                lb[y] = d;
                hset(&h, y, d);
            }
        }
    }
    free(D);
    free(size);
    free(cid);
    free(nbr);
    free(lb);
    free(h.pos_of);
    free(h.key_at);
    free(h.val);
    return 0;
}

int sdo_linkage(const double* x, int N, int D, double* Z) {
    if (N < 2) return 0;
    size_t np = (size_t)N * (N - 1) / 2;
    double* d = (double*)malloc(sizeof(double) * np);
    sdo_pdist(x, N, D, d);
    sdo_linkage_condensed(d, N, Z);
    free(d);
    return 0;
}

/* ------------------------------------------------------------------ a12 */

int sdo_fcluster(const double* Z, int n, double cutoff, int* T) {
    if (n < 2) {
        if (n == 1) T[0] = 1;
        return 0;
    }
    /* CL:121-172: MD[node] = largest merge distance inside the subtree.  Children always precede
     * their parent in Z, so one forward sweep performs the same comparisons as the reference's
     * post-order walk (own distance first, then left, then right, strict '>'). */
    double* MD = (double*)malloc(sizeof(double) * (size_t)(n - 1));
    for (int k = 0; k < n - 1; ++k) {
        int lc = (int)Z[4 * (size_t)k], rc = (int)Z[4 * (size_t)k + 1];
        double m = Z[4 * (size_t)k + 2];
        if (lc >= n && MD[lc - n] > m) m = MD[lc - n];
        if (rc >= n && MD[rc - n] > m) m = MD[rc - n];
        MD[k] = m;
    }
    /* CL:174-232: depth-first from the root, left internal child, right internal child, then leaf
     * children; the first node on a path with MD <= cutoff leads a flat cluster. */
    int* stack = (int*)malloc(sizeof(int) * (size_t)n);
    unsigned char* stage = (unsigned char*)calloc((size_t)n, 1); /* per internal node: 0 new,1 left done,2 right done */
    int sp = 0, ncl = 0, leader = -1;
    stack[0] = n - 2;
    while (sp >= 0) {
        int r = stack[sp];
        int lc = (int)Z[4 * (size_t)r], rc = (int)Z[4 * (size_t)r + 1];
        if (leader == -1 && MD[r] <= cutoff) {
            leader = r;
            ++ncl;
        }
        if (stage[r] == 0) {
            stage[r] = 1;
            if (lc >= n) {
                stack[++sp] = lc - n;
                continue;
            }
        }
        if (stage[r] == 1) {
            stage[r] = 2;
            if (rc >= n) {
                stack[++sp] = rc - n;
                continue;
            }
        }
        if (lc < n) {
            if (leader == -1) ++ncl;
            T[lc] = ncl;
        }
        if (rc < n) {
            if (leader == -1) ++ncl;
            T[rc] = ncl;
        }
        if (leader == r) leader = -1;
        --sp;
    }
    free(stack);
    free(stage);
    free(MD);
    return 0;
}

/* ------------------------------------------------------------------ a13/a14 */

static int cos_dist(const double* a, const double* b, int D, double* out) { /* SD:476-498 */
    double dot = 0.0, ma = 0.0, mb = 0.0;
    for (int k = 0; k < D; ++k) {
        dot += a[k] * b[k];
        ma += a[k] * a[k];
        mb += b[k] * b[k];
    }
    if (ma == 0.0 || mb == 0.0) return 2; /* reference throws "Vectors have zero magnitude." */
    *out = 1.0 - (dot / (sqrt(ma) * sqrt(mb)));
    return 0;
}

int sdo_cosine_cdist(const double* a, int na, const double* b, int nb, int D, double* out) {
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) {
            int rc = cos_dist(a + (size_t)i * D, b + (size_t)j * D, D, &out[(size_t)i * nb + j]);
            if (rc) return rc;
        }
    return 0;
}

static void label_mean(const double* x, const int* labels, int N, int D, int label, double* mean) { /* SD:443-473 */
    int cnt = 0;
    for (int j = 0; j < D; ++j) mean[j] = 0.0;
    for (int i = 0; i < N; ++i)
        if (labels[i] == label) {
            for (int j = 0; j < D; ++j) mean[j] += x[(size_t)i * D + j];
            ++cnt;
        }
    if (cnt > 0)
        for (int j = 0; j < D; ++j) mean[j] /= (double)cnt;
}

static int cmp_int(const void* a, const void* b) { return (*(const int*)a > *(const int*)b) - (*(const int*)a < *(const int*)b); }

int sdo_cluster_labels(const double* x, int N, int D, float threshold, int min_cluster_size, int* labels) {
    if (N <= 0) return 0;
    if (N == 1) {
        labels[0] = 0;
        return 0;
    }
    /* SD:2308-2309 */
    long tenth = (long)round(0.1 * (double)N);
    long mcs = tenth > 1 ? tenth : 1;
    if (mcs > min_cluster_size) mcs = min_cluster_size;

    double* nx = (double*)malloc(sizeof(double) * (size_t)N * D);
    memcpy(nx, x, sizeof(double) * (size_t)N * D);
    sdo_normalize_embeddings(nx, N, D);
    double* Z = (double*)malloc(sizeof(double) * 4 * (size_t)(N - 1));
    sdo_linkage(nx, N, D, Z);
    sdo_fcluster(Z, N, (double)threshold, labels); /* float threshold promoted, SD:2049/2323 */
    free(Z);
    free(nx);
    int maxl = 0;
    for (int i = 0; i < N; ++i) {
        labels[i] -= 1;
        if (labels[i] > maxl) maxl = labels[i];
    }
    int nlab = maxl + 1;
    int* cnt = (int*)calloc((size_t)nlab, sizeof(int));
    for (int i = 0; i < N; ++i) cnt[labels[i]]++;
    int* large = (int*)malloc(sizeof(int) * (size_t)nlab);
    int* small = (int*)malloc(sizeof(int) * (size_t)nlab);
    int nl = 0, ns = 0;
    for (int l = 0; l < nlab; ++l) { /* ascending == the reference's sorted lists, SD:2382-2383 */
        if (!cnt[l]) continue;
        if (cnt[l] >= mcs)
            large[nl++] = l;
        else
            small[ns++] = l;
    }
    int rc = 0;
    if (nl == 0) { /* SD:2371-2375 (with NDEBUG; a live assert at 2369 fires first in the reference build) */
        for (int i = 0; i < N; ++i) labels[i] = 0;
    } else if (ns > 0) {
        double* lc = (double*)malloc(sizeof(double) * (size_t)nl * D);
        double* sc = (double*)malloc(sizeof(double) * (size_t)ns * D);
        for (int a = 0; a < nl; ++a) label_mean(x, labels, N, D, large[a], lc + (size_t)a * D); /* un-normalised */
        for (int b = 0; b < ns; ++b) label_mean(x, labels, N, D, small[b], sc + (size_t)b * D);
        double* cd = (double*)malloc(sizeof(double) * (size_t)nl * ns);
        rc = sdo_cosine_cdist(lc, nl, sc, ns, D, cd);
        if (!rc) {
            for (int b = 0; b < ns; ++b) {
                float best = FLT_MAX; /* float running minimum, SD:2396-2403 */
                int arg = -1;
                for (int a = 0; a < nl; ++a)
                    if (cd[(size_t)a * ns + b] < best) {
                        best = (float)cd[(size_t)a * ns + b];
                        arg = a;
                    }
                if (arg >= 0)
                    for (int i = 0; i < N; ++i)
                        if (labels[i] == small[b]) labels[i] = large[arg];
            }
            /* SD:519-548: rank within sorted unique labels */
            int* uniq = (int*)malloc(sizeof(int) * (size_t)N);
            memcpy(uniq, labels, sizeof(int) * (size_t)N);
            qsort(uniq, (size_t)N, sizeof(int), cmp_int);
            int nu = 0;
            for (int i = 0; i < N; ++i)
                if (i == 0 || uniq[i] != uniq[nu - 1]) uniq[nu++] = uniq[i];
            for (int i = 0; i < N; ++i) {
                int lo = 0;
                while (uniq[lo] != labels[i]) ++lo;
                labels[i] = lo;
            }
            free(uniq);
        }
        free(cd);
        free(lc);
        free(sc);
    }
    free(cnt);
    free(large);
    free(small);
    return rc;
}

int sdo_clustering_stage(const double* emb, int C, int S, int D, float threshold, int min_cluster_size,
                         const double* binarized, int F, int* hard, double* soft_out, int soft_k_cap,
                         int* num_clusters_out) {
    int R = C * S;
    int* keep = (int*)malloc(sizeof(int) * (size_t)R);
    int N = 0;
    for (int r = 0; r < R; ++r) /* SD:2222-2229: first element decides */
        if (!isnan(emb[(size_t)r * D])) keep[N++] = r;
    int rc = 0;
    if (num_clusters_out) *num_clusters_out = 1;
    if (N < 2) { /* set_num_clusters -> max_clusters < 2, SD:2081-2088 */
        for (int r = 0; r < R; ++r) hard[r] = 0;
    } else {
        double* fx = (double*)malloc(sizeof(double) * (size_t)N * D);
        for (int i = 0; i < N; ++i) memcpy(fx + (size_t)i * D, emb + (size_t)keep[i] * D, sizeof(double) * (size_t)D);
        int* lab = (int*)malloc(sizeof(int) * (size_t)N);
        rc = sdo_cluster_labels(fx, N, D, threshold, min_cluster_size, lab);
        if (!rc) {
            int Kc = 0;
            for (int i = 0; i < N; ++i)
                if (lab[i] + 1 > Kc) Kc = lab[i] + 1;
            if (num_clusters_out) *num_clusters_out = Kc;
            double* cen = (double*)malloc(sizeof(double) * (size_t)Kc * D);
            for (int k = 0; k < Kc; ++k) { /* SD:2147-2167: sum in index order then divide (0/0 -> NaN if empty) */
                double* c = cen + (size_t)k * D;
                size_t cnt = 0;
                for (int j = 0; j < D; ++j) c[j] = 0.0;
                for (int i = 0; i < N; ++i)
                    if (lab[i] == k) {
                        ++cnt;
                        for (int j = 0; j < D; ++j) c[j] += fx[(size_t)i * D + j];
                    }
                for (int j = 0; j < D; ++j) c[j] /= (double)cnt;
            }
            for (int r = 0; r < R && !rc; ++r) { /* SD:2180-2211 */
                int arg = 0;
                double best = -DBL_MAX;
                for (int k = 0; k < Kc; ++k) {
                    double d;
                    rc = cos_dist(emb + (size_t)r * D, cen + (size_t)k * D, D, &d);
                    if (rc) break;
                    double soft = 2.0 - d;
                    if (soft_out && k < soft_k_cap) soft_out[(size_t)r * soft_k_cap + k] = soft;
                    if (soft > best) {
                        best = soft;
                        arg = k;
                    }
                }
                hard[r] = arg;
            }
            free(cen);
        }
        free(lab);
        free(fx);
    }
    if (!rc && binarized) /* SD:3172-3191: speakers never active in a chunk */
        for (int c = 0; c < C; ++c)
            for (int s = 0; s < S; ++s) {
                float acc = 0.0f;
                for (int f = 0; f < F; ++f) acc += (float)binarized[((size_t)c * F + f) * S + s];
                if (fabsf(acc) < DBL_EPSILON) hard[c * S + s] = -2;
            }
    free(keep);
    return rc;
}

/* ------------------------------------------------------------------ f1: masking / compaction */

int sdo_mask_compact(const float* wav, const float* masks, int B, int L, int F, int min_num_samples, float* signals,
                     float* wav_lens, unsigned char* too_short) {
    float max_len = 0;
    for (int b = 0; b < B; ++b) {
        const float* w = wav + (size_t)b * L;
        float* o = signals + (size_t)b * L;
        size_t n = 0;
        for (int j = 0; j < L; ++j) {
            int src = (int)((long)j * F / L); /* nearest-neighbour upsample, SD:760 */
            if (masks[(size_t)b * F + src] > 0.5f) o[n++] = w[j]; /* stream compaction, SD:784-794 */
        }
        for (size_t j = n; j < (size_t)L; ++j) o[j] = 0.0f;
        wav_lens[b] = (float)n;
        if ((float)n > max_len) max_len = (float)n;
    }
    if (max_len < (float)min_num_samples) return 1; /* SD:2479-2486 */
    for (int b = 0; b < B; ++b) { /* SD:2498-2510 */
        if (wav_lens[b] < (float)min_num_samples) {
            wav_lens[b] = 1.0f;
            too_short[b] = 1;
        } else {
            wav_lens[b] /= max_len;
            too_short[b] = 0;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ f2: reconstruct / to_diarization */

/* SlidingWindow::operator[] (SD:1092-1115): start of window #pos, counted from 0.0 by repeated += step,
 * or 0.0 when the window would run past num_samples. */
static double window_start(double step, double duration, long num_samples, int pos) {
    int wsize = (int)round(duration * 16000.0), ssize = (int)round(step * 16000.0);
    double start = 0.0;
    size_t cur = 0;
    int idx = 0;
    for (;;) {
        if (idx == pos) return start;
        if (cur + (size_t)wsize >= (size_t)num_samples) break;
        start += step;
        cur += (size_t)ssize;
        ++idx;
    }
    return 0.0;
}

/* crop_segment (SD:2568-2635), mode loose: row range [r0, r1) and start time of the cropped window. */
static void crop_range(long n_rows, double src_start, double src_step, double src_duration, long src_num_samples,
                       double focus_start, double focus_end, long* r0, long* r1, double* new_start) {
    float i_ = (float)((focus_start - src_duration - src_start) / src_step); /* float index math, SD:2577-2587 */
    int a = (int)ceilf(i_);
    if (a < 0) a = 0;
    float j_ = (float)((focus_end - src_start) / src_step);
    int b = (int)floorf(j_) + 1;
    *new_start = (double)(float)window_start(src_step, src_duration, src_num_samples, a);
    long s = a, e = b;
    if (s >= n_rows) {
        *r0 = *r1 = 0;
        return;
    }
    *r0 = s;
    *r1 = e < n_rows ? e : n_rows;
    if (*r1 < *r0) *r1 = *r0;
}

long sdo_reconstruct(const float* segmentations, int C, int F, int K, double sf_start, double sf_step,
                     double sf_duration, long sf_num_samples, const int* hard, const int* count, long n_count,
                     double cf_start, double cf_step, double cf_duration, long cf_num_samples, double* out,
                     long cap_elems, int* cols_out, double* frames_out) {
    int Kc = 0; /* SD:2803-2812: max label (floor 0) + 1 */
    for (int i = 0; i < C * K; ++i)
        if (hard[i] > Kc) Kc = hard[i];
    Kc += 1;
    size_t per = (size_t)F * Kc;
    double* cs = (double*)malloc(sizeof(double) * (size_t)C * per);
    for (size_t i = 0; i < (size_t)C * per; ++i) cs[i] = NAN;
    for (int c = 0; c < C; ++c) /* SD:2818-2838: per cluster, max over the local speakers assigned to it */
        for (int s = 0; s < K; ++s) {
            int k = hard[c * K + s];
            if (k == -2) continue;
            if (k < 0) continue; /* reference would index out of range; never produced by the pipeline */
            for (int f = 0; f < F; ++f) {
                float m = -INFINITY;
                for (int s2 = 0; s2 < K; ++s2)
                    if (hard[c * K + s2] == k) {
                        float v = segmentations[((size_t)c * F + f) * K + s2];
                        m = (m < v) ? v : m; /* std::max(maxValue, x): keeps maxValue when x is NaN */
                    }
                cs[(size_t)c * per + (size_t)f * Kc + k] = (double)m;
            }
        }
    /* to_diarization, SD:2638-2764 */
    double target = sf_start + sf_duration + (double)(size_t)(C - 1) * sf_step;
    long NF = sdo_closest_frame(sf_start, cf_step, cf_duration, target) + 1;
    double* act = (double*)malloc(sizeof(double) * (size_t)NF * Kc);
    double post[4];
    sdo_aggregate(cs, C, F, Kc, sf_start, sf_step, sf_duration, sf_num_samples, cf_step, cf_duration, 0, 0.0, 1,
                  DBL_EPSILON, act, NF, NULL, NULL, post);
    free(cs);
    /* extents (SD:2691-2706) */
    double a_end = (post[0] + (0 - .5) * post[1] + .5 * post[2]) + (double)NF * post[1];
    double c_end = (cf_start + (0 - .5) * cf_step + .5 * cf_duration) + (double)n_count * cf_step;
    double f0 = post[0] > cf_start ? post[0] : cf_start;
    double f1 = a_end < c_end ? a_end : c_end;
    long a0, a1, c0, c1;
    double a_start, c_start_unused;
    crop_range(NF, post[0], post[1], post[2], (long)post[3], f0, f1, &a0, &a1, &a_start);
    /* count_frames.num_samples is the trimmed frame count set at SD:1779/1281; it only feeds operator[] */
    crop_range(n_count, cf_start, cf_step, cf_duration, cf_num_samples, f0, f1, &c0, &c1, &c_start_unused);
    long rows = a1 - a0, crow = c1 - c0;
    if (rows * Kc > cap_elems) {
        free(act);
        return -1;
    }
    for (long i = 0; i < rows * Kc; ++i) out[i] = 0.0;
    int* order = (int*)malloc(sizeof(int) * (size_t)Kc);
    for (long i = 0; i < crow && i < rows; ++i) {
        const double* a = act + (size_t)(a0 + i) * Kc;
        /* stable argsort of -activation (SD:2724-2730): insertion sort keeps ties in index order */
        for (int k = 0; k < Kc; ++k) {
            int j = k;
            while (j > 0 && (-a[k]) < (-a[order[j - 1]])) {
                order[j] = order[j - 1];
                --j;
            }
            order[j] = k;
        }
        int cnt = count[c0 + i];
        if (cnt > Kc) cnt = Kc; /* SD:2678-2685 */
        for (int j = 0; j < cnt; ++j) out[(size_t)i * Kc + order[j]] = 1.0;
    }
    free(order);
    free(act);
    *cols_out = Kc;
    frames_out[0] = a_start;
    frames_out[1] = post[1];
    frames_out[2] = post[2];
    return rows;
}

/* ------------------------------------------------------------------ f3: to_annotation */

typedef struct {
    double s, e;
    int label;
} seg_t;

long sdo_to_annotation(const double* scores, long rows, int cols, double f_start, double f_step, double f_duration,
                       double onset, double offset, double min_duration_on, double min_duration_off, double* seg_out,
                       int* label_out, long cap) {
    size_t capn = 16, n = 0;
    seg_t* segs = (seg_t*)malloc(sizeof(seg_t) * capn);
#define PUSH(S, E, Lb)                                                   \
    do {                                                                 \
        if (n == capn) {                                                 \
            capn *= 2;                                                   \
            segs = (seg_t*)realloc(segs, sizeof(seg_t) * capn);          \
        }                                                                \
        segs[n].s = (S);                                                 \
        segs[n].e = (E);                                                 \
        segs[n].label = (Lb);                                            \
        ++n;                                                             \
    } while (0)
#define TS(i) ((((f_start + (double)(i) * f_step)) + ((f_start + (double)(i) * f_step) + f_duration)) / 2)
    for (int k = 0; k < cols; ++k) { /* SD:2883-2921 */
        size_t first = n;
        double start = TS(0);
        int active = scores[k] > onset;
        for (long t = 1; t < rows; ++t) {
            double v = scores[(size_t)t * cols + k];
            if (active) {
                if (v < offset) {
                    PUSH(start, TS(t), k);
                    start = TS(t);
                    active = 0;
                }
            } else if (v > onset) {
                start = TS(t);
                active = 1;
            }
        }
        if (active) PUSH(start, TS(rows - 1), k);
        /* Track::support (SD:911-941): segments of one label are emitted in time order already */
        if (min_duration_off > 0.0 && n > first) {
            size_t w = first;
            seg_t cur = segs[first];
            for (size_t i = first + 1; i < n; ++i) {
                seg_t nx = segs[i];
                double gap;
                if (cur.s < nx.s)
                    gap = cur.e >= nx.s ? 0.0 : nx.s - cur.e;
                else
                    gap = cur.s <= nx.e ? 0.0 : cur.s - nx.e;
                if (gap < min_duration_off) {
                    if (nx.s < cur.s) cur.s = nx.s;
                    if (nx.e > cur.e) cur.e = nx.e;
                } else {
                    segs[w++] = cur;
                    cur = nx;
                }
            }
            segs[w++] = cur;
            n = w;
        }
        if (min_duration_on > 0) { /* Track::removeShort never examines index 0 (SD:943-953) */
            size_t w = first + (n > first ? 1 : 0);
            for (size_t i = first + 1; i < n; ++i)
                if (!(segs[i].e - segs[i].s < min_duration_on)) segs[w++] = segs[i];
            n = w;
        }
    }
#undef TS
#undef PUSH
    /* finalResult (SD:962-978): sort by start.  std::sort's order among equal starts is unspecified;
     * this restatement is stable (track order), tests compare ties as sets. */
    for (size_t i = 1; i < n; ++i) {
        seg_t v = segs[i];
        size_t j = i;
        while (j > 0 && v.s < segs[j - 1].s) {
            segs[j] = segs[j - 1];
            --j;
        }
        segs[j] = v;
    }
    long res = (long)n;
    if (res > cap)
        res = -1;
    else
        for (size_t i = 0; i < n; ++i) {
            seg_out[2 * i] = segs[i].s;
            seg_out[2 * i + 1] = segs[i].e;
            label_out[i] = segs[i].label;
        }
    free(segs);
    return res;
}

/* ------------------------------------------------------------------ f4: ingest / chunker */

void sdo_ingest_pcm16(const short* pcm, long n, float* out) {
    for (long i = 0; i < n; ++i) {
        float v = (float)pcm[i];                 /* wav.h:100-104 */
        out[i] = (float)((double)(v * 1.0f) / 32768.0); /* SD:2948-2951 */
    }
}

long sdo_crop(const float* wave, long n, double start, double duration, int sample_rate, float* out) {
    int start_frame = (int)floor(start * sample_rate);
    int frames = (int)n;
    int num_frames = (int)floor(duration * sample_rate);
    int end_frame = start_frame + num_frames;
    int pad_start = -(start_frame < 0 ? start_frame : 0);
    int pad_end = (end_frame > frames ? end_frame : frames) - frames;
    if (start_frame < 0) start_frame = 0;
    if (end_frame > frames) end_frame = frames;
    long w = 0;
    for (int i = 0; i < pad_start; ++i) out[w++] = 0.0f;
    for (int i = start_frame; i < end_frame; ++i) out[w++] = wave[i];
    for (int i = 0; i < pad_end; ++i) out[w++] = 0.0f;
    return w;
}

void sdo_slide_geometry(long num_samples, double duration, double step, long* full_chunks, long* tail_start,
                        long* tail_len) {
    int window_size = (int)round(duration * 16000), step_size = (int)round(step * 16000);
    size_t i = 0, n = 0;
    while (i + (size_t)window_size < (size_t)num_samples) { /* SD:1419 */
        ++n;
        i += (size_t)step_size;
    }
    *full_chunks = (long)n;
    if (i + 1 < (size_t)num_samples) { /* SD:1451: at least one sample remains */
        *tail_start = (long)i;
        *tail_len = num_samples - (long)i;
    } else {
        *tail_start = -1;
        *tail_len = 0;
    }
}
