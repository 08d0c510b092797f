// Clustering stage.
//
//   normalise          Helper::normalizeEmbeddings            speakerDiarizer.cpp:330-357
//   pdist (fp64)       Clustering::linkage, euclideanDistance  clustering/clustering.cpp:408-431
//   linkage            fast_linkage (scipy generic AHC)        clustering/clustering.cpp:289-406 (+Heap 28-119)
//   fcluster           get_max_dist_for_each_cluster / cluster_monocrit  clustering.cpp:121-232, 442-457
//   cluster driver     Cluster::cluster                        speakerDiarizer.cpp:2300-2422
//   assignment         Cluster::assign_embeddings (+ -2 mask)  speakerDiarizer.cpp:2120-2212, 3166-3191
//
// Parity rules: every fp64 operation that feeds an exact comparison in the reference is issued
// with the round-to-nearest intrinsics (__dadd_rn / __dmul_rn / __ddiv_rn / __dsqrt_rn), which nvcc never
// contracts into FMA, in the reference's association order.  The merge sequence is driven by the same
// indexed binary heap with the same strict comparisons, so ties resolve as in the reference.
//
// Data layout: the distance matrix lives in HBM as a full symmetric fp64 square D[N][ld] (ld = N rounded up
// to 16), so a cluster's row is one contiguous, coalesced stream for the Lance-Williams update and for the
// nearest-neighbour rescans; the merged cluster's column is patched with strided 8-byte stores.
#include "common.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace sdb {

// ------------------------------------------------------------------------------------------------
// gather + normalise
// ------------------------------------------------------------------------------------------------

// x[i][:] = emb[keep[i]][:];  xn[i][:] = x[i][:] / (double)(float)sqrt(sum_k x^2)
// One warp per row: coalesced loads (lane l holds elements l, l+32, ...), and the squared norm is accumulated
// strictly in k order (the reference's sequential fp64 sum) by broadcasting one element at a time.
constexpr int GN_MAX_PER_LANE = 16;  // D <= 512
__global__ void __launch_bounds__(256)
    gather_normalize_kernel(const double* __restrict__ emb, const int* __restrict__ keep, int N, int D,
                            double* __restrict__ x, double* __restrict__ xn) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= N) return;
    const double* src = emb + (size_t)(keep ? keep[i] : i) * D;
    if (D <= 32 * GN_MAX_PER_LANE) {
        double v[GN_MAX_PER_LANE];
#pragma unroll
        for (int q = 0; q < GN_MAX_PER_LANE; ++q) {
            const int k = q * 32 + lane;
            v[q] = k < D ? src[k] : 0.0;
        }
        double ss = 0.0;
#pragma unroll
        for (int q = 0; q < GN_MAX_PER_LANE; ++q) {
            if (q * 32 < D) {
                for (int l = 0; l < 32; ++l) {
                    const double e = __shfl_sync(0xffffffffu, v[q], l);
                    if (q * 32 + l < D) ss = __dadd_rn(ss, __dmul_rn(e, e));
                }
            }
        }
        const double norm = (double)(float)__dsqrt_rn(ss);  // L2Norm returns float, speakerDiarizer.cpp:332
#pragma unroll
        for (int q = 0; q < GN_MAX_PER_LANE; ++q) {
            const int k = q * 32 + lane;
            if (k < D) {
                if (x) x[(size_t)i * D + k] = v[q];
                if (xn) xn[(size_t)i * D + k] = norm != 0.0 ? __ddiv_rn(v[q], norm) : v[q];
            }
        }
    } else if (lane == 0) {  // very wide rows: plain sequential fallback
        double ss = 0.0;
        for (int k = 0; k < D; ++k) {
            const double e = src[k];
            if (x) x[(size_t)i * D + k] = e;
            ss = __dadd_rn(ss, __dmul_rn(e, e));
        }
        const double norm = (double)(float)__dsqrt_rn(ss);
        if (xn)
            for (int k = 0; k < D; ++k) xn[(size_t)i * D + k] = norm != 0.0 ? __ddiv_rn(src[k], norm) : src[k];
    }
}

__global__ void row_valid_kernel(const double* __restrict__ emb, int R, int D, unsigned char* __restrict__ valid) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) valid[r] = isnan(emb[(size_t)r * D]) ? 0 : 1;  // speakerDiarizer.cpp:2224
}

// ------------------------------------------------------------------------------------------------
// pdist, exact fp64 (parity mode)
// ------------------------------------------------------------------------------------------------

constexpr int PD_TILE = 64;
constexpr int PD_KC = 16;

// Dm[i][j] = sqrt(sum_k (x_ik - x_jk)^2), k ascending, mul then add, no FMA.  The output is the full symmetric
// square, but only the tiles on and above the diagonal are computed: (a-b)^2 == (b-a)^2 bit for bit, so a tile below
// the diagonal would repeat the arithmetic of its mirror image -- each value is written to both places instead
// (the mirrored stores are whole 32-byte sectors: a thread holds four consecutive rows of a column).
__global__ void __launch_bounds__(256)
    pdist_f64_kernel(const double* __restrict__ x, int N, int D, double* __restrict__ Dm, long ld,
                     const int* __restrict__ run_flag) {
    if (blockIdx.x < blockIdx.y) return;  // mirror image of tile (x, y)
    if (run_flag && !*run_flag) return;
    __shared__ double A[PD_KC][PD_TILE + 1];
    __shared__ double B[PD_KC][PD_TILE + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int bi = blockIdx.y * PD_TILE, bj = blockIdx.x * PD_TILE;
    double acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;
    for (int k0 = 0; k0 < D; k0 += PD_KC) {
#pragma unroll
        for (int e = threadIdx.x; e < PD_TILE * PD_KC; e += 256) {
            const int row = e / PD_KC, kk = e % PD_KC;
            const int k = k0 + kk;
            const int gi = bi + row, gj = bj + row;
            A[kk][row] = (gi < N && k < D) ? x[(size_t)gi * D + k] : 0.0;
            B[kk][row] = (gj < N && k < D) ? x[(size_t)gj * D + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < PD_KC; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) a[p] = A[kk][ty * 4 + p];
#pragma unroll
            for (int q = 0; q < 4; ++q) b[q] = B[kk][tx * 4 + q];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double d = __dsub_rn(a[p], b[q]);
                    acc[p][q] = __dadd_rn(acc[p][q], __dmul_rn(d, d));
                }
        }
        __syncthreads();
    }
    const bool mirror = blockIdx.x != blockIdx.y;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int gi = bi + ty * 4 + p;
        if (gi >= N) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int gj = bj + tx * 4 + q;
            if (gj < N) {
                const double d = __dsqrt_rn(acc[p][q]);
                Dm[(size_t)gi * ld + gj] = d;
                if (mirror) Dm[(size_t)gj * ld + gi] = d;
            }
        }
    }
}

// square -> condensed (row-major upper triangle), clustering.cpp:423-431 ordering
__global__ void __launch_bounds__(256)
    condense_kernel(const double* __restrict__ Dm, long ld, int N, double* __restrict__ out) {
    const int i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j <= i || j >= N) return;
    const long long p = (long long)N * i - ((long long)i * (i + 1) / 2) + (j - i - 1);
    out[p] = Dm[(size_t)i * ld + j];
}

// condensed -> square (for sd_linkage variants fed with a condensed matrix)
__global__ void __launch_bounds__(256)
    expand_kernel(const double* __restrict__ cond, long ld, int N, double* __restrict__ Dm) {
    const int i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    double v = 0.0;
    if (i != j) {
        const int a = i < j ? i : j, b = i < j ? j : i;
        v = cond[(long long)N * a - ((long long)a * (a + 1) / 2) + (b - a - 1)];
    }
    Dm[(size_t)i * ld + j] = v;
}

// ------------------------------------------------------------------------------------------------
// linkage
// ------------------------------------------------------------------------------------------------

struct LinkWork {
    double* D;      // [N][ld]
    long ld;
    int* size;      // [N]
    int* cid;       // [N]
    int* nbr;       // [N]   nearest-neighbour candidate among higher indices
    double* lb;     // [N]   lower bound of the distance to it
    double* cur;    // [N]   mirror of D[z][nbr[z]]
    int* pos_of;    // heap: key -> slot
    int* key_at;    // heap: slot -> key
    double* hval;   // heap: slot -> value
    unsigned* bitmap;  // [ceil(N/32)] rows whose bound dropped in this merge
    double* Z;      // [N-1][4]
    int* status;    // device status word
    unsigned long long* stats;  // [0] stale revalidations, [1] heap updates replayed after sweeps, [2] fallbacks
    void* fast_scratch;         // state of the heap-free kernel when it does not fit in shared memory
};

struct MinIdx {
    double v;
    int i;
};

__device__ __forceinline__ MinIdx better(MinIdx a, MinIdx b) {
    // smaller value wins; equal values keep the lower index (what a sequential strict-'<' scan returns)
    if (b.i >= 0 && (a.i < 0 || b.v < a.v || (b.v == a.v && b.i < a.i))) return b;
    return a;
}

__device__ __forceinline__ MinIdx warp_min(MinIdx m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MinIdx t;
        t.v = __shfl_xor_sync(0xffffffffu, m.v, o);
        t.i = __shfl_xor_sync(0xffffffffu, m.i, o);
        m = better(m, t);
    }
    return m;
}

// find_min_dist for all rows at start (clustering.cpp:314-318): one warp per row
__global__ void __launch_bounds__(256) rowmin_init_kernel(LinkWork w, int n, const int* __restrict__ run_flag) {
    if (run_flag && !*run_flag) return;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    if (lane == 0) {
        w.size[row] = 1;
        w.cid[row] = row;
    }
    if (row >= n - 1) return;
    MinIdx m;
    m.v = INFINITY;
    m.i = -1;
    const double* r = w.D + (size_t)row * w.ld;
    for (int i = row + 1 + lane; i < n; i += 32) {
        const double d = r[i];
        if (d < m.v) {
            m.v = d;
            m.i = i;
        }
    }
    m = warp_min(m);
    if (lane == 0) {
        w.nbr[row] = m.i;
        w.lb[row] = m.i >= 0 ? m.v : INFINITY;
        w.pos_of[row] = row;
        w.key_at[row] = row;
        w.hval[row] = m.i >= 0 ? m.v : INFINITY;
    }
}

// ---- indexed binary min-heap, comparison rules of clustering.cpp:28-119 ----
// The reference sifts with pairwise swaps; moving a "hole" instead performs the same comparisons in the
// same order and leaves the same arrangement, with one dependent shared-memory load per level.
struct Heap {
    int* pos_of;
    int* key_at;
    double* val;
    int n;
    __device__ __forceinline__ void place(int i, int key, double v) {
        val[i] = v;
        key_at[i] = key;
        pos_of[key] = i;
    }
    __device__ __forceinline__ void down(int i, int key, double v) {
        for (int c = 2 * i + 1; c < n; c = 2 * i + 1) {
            double cv = val[c];
            if (c + 1 < n) {
                const double rv = val[c + 1];
                if (rv < cv) {
                    cv = rv;
                    ++c;
                }
            }
            if (!(v > cv)) break;
            place(i, key_at[c], cv);
            i = c;
        }
        place(i, key, v);
    }
    __device__ __forceinline__ void up(int i, int key, double v) {
        while (i > 0) {
            const int p = (i - 1) >> 1;
            const double pv = val[p];
            if (!(pv > v)) break;
            place(i, key_at[p], pv);
            i = p;
        }
        place(i, key, v);
    }
    // sift_down(i) of the element already stored at slot i
    __device__ __forceinline__ void down_at(int i) { down(i, key_at[i], val[i]); }
    __device__ __forceinline__ void set(int key, double v) {  // change_value, clustering.cpp:109-118
        const int i = pos_of[key];
        const double old = val[i];
        if (v < old)
            up(i, key, v);
        else
            down(i, key, v);
    }
    __device__ __forceinline__ void remove_min() {  // clustering.cpp:103-107: swap(0, n-1); --n; sift_down(0)
        const int last = n - 1;
        const int k0 = key_at[0], kl = key_at[last];
        const double v0 = val[0], vl = val[last];
        place(last, k0, v0);
        n = last;
        if (n > 0) down(0, kl, vl);
    }
};

__device__ __forceinline__ double centroid_update(double dxi, double dyi, double dxy, int nx, int ny) {
    // clustering.cpp:250-256, same association, no contraction
    const double t1 = __dmul_rn(__dmul_rn((double)nx, dxi), dxi);
    const double t2 = __dmul_rn(__dmul_rn((double)ny, dyi), dyi);
    const double t3 = __ddiv_rn(__dmul_rn(__dmul_rn((double)(nx * ny), dxy), dxy), (double)(nx + ny));
    return __dsqrt_rn(__ddiv_rn(__dsub_rn(__dadd_rn(t1, t2), t3), (double)(nx + ny)));
}

constexpr int LK_THREADS = 1024;
constexpr int LK_WORKERS = LK_THREADS - 32;  // warp 0 owns the heap; warps 1..31 sweep rows

__device__ __forceinline__ MinIdx block_min(MinIdx m, MinIdx* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    m = warp_min(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (warp == 0) {
        MinIdx t = red[lane];  // LK_THREADS / 32 == 32 entries
        t = warp_min(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// Where the per-cluster state lives.  ALL: everything in shared memory (N <= ~5 000); HEAP: only the heap in
// shared memory (N <= ~14 000); GLOBAL: everything in the (L1/L2-cached) global workspace.
enum LinkMode { LK_SMEM_ALL = 0, LK_SMEM_HEAP = 1, LK_GLOBAL = 2 };

__host__ __device__ inline size_t link_smem_bytes(int mode, int n) {
    const size_t N = (size_t)((n + 1) / 2 * 2);
    if (mode == LK_SMEM_ALL) return N * (8 + 8 + 8 + 4 + 4 + 4 + 4 + 4) + ((size_t)n / 32 + 2) * 4 + 64;
    if (mode == LK_SMEM_HEAP) return N * (8 + 4 + 4) + 64;
    return 64;
}

// One persistent CTA performs all N-1 merges of one problem (grid.x = number of independent problems).
//
// Per merge, fast path = two block barriers:
//   thread 0   : read heap top, compare with the cached current distance of its candidate pair
//   ---- barrier ----
//   warp 0     : remove_min + dendrogram row           | warps 1..31 : Lance-Williams sweep over all live z
//   ---- barrier ----
//   warp 0     : reduce y's new nearest neighbour, replay heap updates for rows whose bound dropped
// `cur[z]` mirrors D[z][nbr[z]] exactly (it is refreshed whenever that entry or nbr[z] changes), so the
// reference's "dist == D[x][y]" validity test needs no global-memory round trip.
template <int MODE>
__global__ void __launch_bounds__(LK_THREADS)
    linkage_kernel(const LinkWork* __restrict__ works, const int* ns, const int* __restrict__ run_flags) {
    if (run_flags && !run_flags[blockIdx.x]) return;
    const LinkWork w = works[blockIdx.x];
    const int n = ns[blockIdx.x];
    if (threadIdx.x == 0 && run_flags && w.stats) w.stats[2] += 1;  // fell back from the heap-free path
    extern __shared__ __align__(16) unsigned char lk_smem[];
    __shared__ MinIdx red[33];
    __shared__ int s_x, s_y, s_nx, s_ny, s_stale, s_abort;
    __shared__ double s_dist;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n < 2) return;

    // ---- state pointers ----
    double *hval, *lb, *cur;
    int *pos_of, *key_at, *nbr, *size, *cid;
    unsigned* bitmap;
    {
        const size_t N = (size_t)((n + 1) / 2 * 2);
        unsigned char* p = lk_smem;
        if (MODE == LK_SMEM_ALL || MODE == LK_SMEM_HEAP) {
            hval = reinterpret_cast<double*>(p);
            p += N * 8;
        } else
            hval = w.hval;
        if (MODE == LK_SMEM_ALL) {
            lb = reinterpret_cast<double*>(p);
            p += N * 8;
            cur = reinterpret_cast<double*>(p);
            p += N * 8;
        } else {
            lb = w.lb;
            cur = w.cur;
        }
        if (MODE == LK_SMEM_ALL || MODE == LK_SMEM_HEAP) {
            pos_of = reinterpret_cast<int*>(p);
            p += N * 4;
            key_at = reinterpret_cast<int*>(p);
            p += N * 4;
        } else {
            pos_of = w.pos_of;
            key_at = w.key_at;
        }
        if (MODE == LK_SMEM_ALL) {
            nbr = reinterpret_cast<int*>(p);
            p += N * 4;
            size = reinterpret_cast<int*>(p);
            p += N * 4;
            cid = reinterpret_cast<int*>(p);
            p += N * 4;
            bitmap = reinterpret_cast<unsigned*>(p);
        } else {
            nbr = w.nbr;
            size = w.size;
            cid = w.cid;
            bitmap = w.bitmap;
        }
    }
    // rowmin_init_kernel left the initial state in the global workspace
    for (int i = tid; i < n; i += LK_THREADS) {
        if (MODE != LK_GLOBAL && i < n - 1) {
            hval[i] = w.hval[i];
            pos_of[i] = i;
            key_at[i] = i;
        }
        if (MODE == LK_SMEM_ALL) {
            size[i] = 1;
            cid[i] = i;
            if (i < n - 1) {
                nbr[i] = w.nbr[i];
                lb[i] = w.lb[i];
            }
        }
        if (i < n - 1) cur[i] = w.lb[i];  // D[i][nbr[i]] == the row minimum right after initialisation
    }
    if (tid == 0) s_abort = 0;
    __syncthreads();

    Heap h;
    h.pos_of = pos_of;
    h.key_at = key_at;
    h.val = hval;
    h.n = n - 1;

    // heapify: the reference sifts down i = size/2 .. 0 (clustering.cpp:94-96).  Sift-downs of nodes on one
    // level touch disjoint subtrees, and every deeper level is finished before a shallower one in the
    // sequential order too, so processing level by level (deepest first) yields the same heap.
    {
        const int last = h.n / 2;
        int top = 0;
        while (((2 << top) - 1) <= last) ++top;
        for (int lev = top; lev >= 0; --lev) {
            const int first = (1 << lev) - 1;
            int end = (2 << lev) - 2;
            if (end > last) end = last;
            for (int i = first + tid; i <= end; i += LK_THREADS)
                if (i < h.n) h.down_at(i);
            __syncthreads();
        }
    }

    const int nwords = (n + 31) >> 5;
    for (int k = 0; k < n - 1; ++k) {
        // ---- closest pair, revalidating stale candidates (clustering.cpp:323-339) ----
        int tries = 0;
        bool forced = false;
        for (;;) {
            if (tid == 0) {
                int x, y;
                double dist;
                if (!forced) {
                    x = h.key_at[0];
                    dist = h.val[0];
                    y = nbr[x];
                    s_stale = (y < 0) || !(dist == cur[x]);
                } else {  // loop bound of the reference reached: continue with the recomputed pair
                    x = s_x;
                    y = s_y;
                    dist = s_dist;
                    s_stale = 0;
                }
                s_x = x;
                s_y = y;
                s_dist = dist;
                if (!s_stale) {
                    if (y < 0) {
                        s_abort = 1;
                    } else {
                        const int nx = size[x], ny = size[y];
                        s_nx = nx;
                        s_ny = ny;
                        size[x] = 0;  // clustering.cpp:356-357 (visible to the sweep after the barrier)
                        size[y] = nx + ny;
                    }
                }
            }
            __syncthreads();
            if (!s_stale) break;
            const int x = s_x;
            MinIdx m;
            m.v = INFINITY;
            m.i = -1;
            const double* r = w.D + (size_t)x * w.ld;
            for (int i = x + 1 + tid; i < n; i += LK_THREADS) {
                if (size[i] == 0) continue;
                const double d = r[i];
                if (d < m.v) {
                    m.v = d;
                    m.i = i;
                }
            }
            m = block_min(m, red);
            ++tries;
            if (tid == 0) {
                if (w.stats) w.stats[0] += 1;
                const double v = m.i >= 0 ? m.v : INFINITY;
                nbr[x] = m.i;
                lb[x] = v;
                cur[x] = v;
                h.set(x, v);
                s_y = m.i;
                s_dist = v;
            }
            forced = tries >= n - k;
        }
        if (s_abort) {  // no live partner: only reachable through NaN distances (the reference indexes out of range)
            if (tid == 0) atomicExch(w.status, SD_ERR_INVALID);
            return;
        }
        const int x = s_x, y = s_y, nx = s_nx, ny = s_ny;
        const double dist = s_dist;
        MinIdx ym;
        ym.v = INFINITY;
        ym.i = -1;
        if (warp == 0) {
            if (lane == 0) {
                h.remove_min();
                int ix = cid[x], iy = cid[y];
                if (ix > iy) {
                    const int t = ix;
                    ix = iy;
                    iy = t;
                }
                double* z = w.Z + 4 * (size_t)k;  // clustering.cpp:347-354
                z[0] = ix;
                z[1] = iy;
                z[2] = dist;
                z[3] = nx + ny;
                cid[y] = n + k;
            }
        } else {
            // ---- Lance-Williams update of row/column y, neighbour fix-ups and y's own nearest neighbour
            //      (clustering.cpp:361-404) fused into one sweep over z ----
            const double* rowx = w.D + (size_t)x * w.ld;
            double* rowy = w.D + (size_t)y * w.ld;
            const int wt = tid - 32;
            for (int z0 = 0; z0 < n; z0 += LK_WORKERS) {
                const int z = z0 + wt;
                bool changed = false;
                if (z < n && z != y && size[z] != 0) {
                    const double nd = centroid_update(rowx[z], rowy[z], dist, nx, ny);
                    rowy[z] = nd;
                    w.D[(size_t)z * w.ld + y] = nd;
                    if (z < y) {
                        int nb = nbr[z];
                        if (z < x && nb == x) nb = y;  // clustering.cpp:374-378
                        if (nd < lb[z]) {              // clustering.cpp:381-392
                            nb = y;
                            lb[z] = nd;
                            changed = true;
                        }
                        if (nb == y) cur[z] = nd;  // keep cur[z] == D[z][nbr[z]]
                        nbr[z] = nb;
                    } else if (nd < ym.v) {  // z > y, ascending within a thread: first strict minimum
                        ym.v = nd;
                        ym.i = z;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, changed);
                const int word = (z0 >> 5) + (warp - 1);
                if (lane == 0 && word < nwords) bitmap[word] = bal;
            }
            ym = warp_min(ym);
            if (lane == 0) red[warp] = ym;
        }
        __syncthreads();
        // ---- warp 0: y's nearest neighbour, then the heap updates in increasing z (sequential order) ----
        if (warp == 0) {
            MinIdx t;
            t.v = INFINITY;
            t.i = -1;
            if (lane > 0) t = red[lane];
            t = warp_min(t);
            for (int wb = 0; wb < nwords; wb += 32) {
                const unsigned word = (wb + lane < nwords) ? bitmap[wb + lane] : 0u;
                unsigned nz = __ballot_sync(0xffffffffu, word != 0u);
                while (nz) {
                    const int src = __ffs(nz) - 1;
                    nz &= nz - 1;
                    unsigned bits = __shfl_sync(0xffffffffu, word, src);
                    if (lane == 0) {
                        while (bits) {
                            const int b = __ffs(bits) - 1;
                            bits &= bits - 1;
                            const int z = ((wb + src) << 5) + b;
                            h.set(z, lb[z]);
                            if (w.stats) w.stats[1] += 1;
                        }
                    }
                }
            }
            if (lane == 0 && y < n - 1 && t.i != -1) {  // clustering.cpp:395-404
                nbr[y] = t.i;
                lb[y] = t.v;
                cur[y] = t.v;
                h.set(y, t.v);
            }
        }
        // no barrier: only thread 0 touches the heap / s_* until the next one, the other warps wait there
    }
}

// ---- fast path: heap-free merges with a uniqueness proof ----------------------------------------------
//
// The binary heap of the reference only decides *which* row is examined next when several rows share the
// smallest bound.  Whenever the smallest bound is attained by exactly one row, every valid heap -- whatever its
// internal arrangement -- returns that row, so the whole state sequence (revalidations, merges, updates) is
// the one the reference goes through.  This kernel therefore keeps no heap: the (value, row, multiplicity) of
// the smallest bound is a block-wide reduction fused into the Lance-Williams sweep, and the ~40 dependent
// decrease-key operations per merge of the heap version disappear.  If the minimum is ever tied (or the
// reference's pop-loop bound is reached) the kernel raises `*need_exact` and stops; the exact heap kernel
// then redoes the problem from a fresh distance matrix.  Both paths give the reference's Z bit for bit.
struct Top {
    double v;  // smallest value (never NaN, never negative: distances and +inf)
    int i;     // lowest row attaining it (-1: none)
    int c;     // how many rows attain it
};

// Warp argmin of non-negative doubles with multiplicity.  For v >= 0 the IEEE bit pattern is monotone, so the
// 64-bit minimum is two 32-bit redux.sync steps; ties resolve to the lowest index.  A lane with c == 0
// contributes nothing (it must pass v = +inf).
__device__ __forceinline__ Top warp_top(double v, int i, int c) {
    const unsigned full = 0xffffffffu;
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(full, hi);
    const bool cand = (hi == mh) && c > 0;
    const unsigned b = __ballot_sync(full, cand);
    Top t;
    if (b == 0u) {  // nothing to reduce
        t.v = INFINITY;
        t.i = -1;
        t.c = 0;
    } else if ((b & (b - 1u)) == 0u) {  // common case: the high words already single out one lane
        const int src = __ffs(b) - 1;
        t.v = __shfl_sync(full, v, src);
        t.i = __shfl_sync(full, i, src);
        t.c = __shfl_sync(full, c, src);
    } else {
        const unsigned ml = __reduce_min_sync(full, cand ? lo : 0xffffffffu);
        const bool is_min = cand && (lo == ml);
        t.c = (int)__reduce_add_sync(full, is_min ? (unsigned)c : 0u);
        t.i = (int)__reduce_min_sync(full, is_min ? (unsigned)i : 0xffffffffu);
        t.v = __hiloint2double((int)mh, (int)ml);
    }
    return t;
}

// s / d for a divisor whose correctly rounded reciprocal r = RN(1/d) is known (Markstein): q = RN(s r),
// rem = s - q d exactly (FMA), result RN(q + rem r) == RN(s / d) whenever d's significand is not all ones --
// d is a cluster size here.  Non-finite inputs take the library division.
__device__ __forceinline__ double div_by(double s, double d, double r) {
    const double q = __dmul_rn(s, r);
    const double rem = __fma_rn(-q, d, s);
    const double q1 = __fma_rn(rem, r, q);
    return isfinite(s) ? q1 : __ddiv_rn(s, d);
}

enum FastMode { LF_SMEM = 0, LF_GLOBAL = 1 };

enum FastOp { OP_MERGE = 0, OP_RESCAN = 1, OP_ABORT = 2 };

struct MergeRec {
    int op, x, y, nx, ny;
    double dist;
};

struct GroupMin {  // smallest bound of a 32-row group: value, lowest row attaining it, multiplicity
    double v;
    int i;
    int c;
};

__host__ __device__ inline size_t linkfast_smem_bytes(int n) {
    const size_t N = (size_t)((n + 31) / 32 * 32);
    return N * (8 + 8 + 4 + 4 + 4) + (N / 32) * sizeof(GroupMin) + 64;
}

// Heap-free linkage, one persistent CTA per problem.
//
//   warp 0 ("control")  keeps the smallest bound through a two-level structure -- per 32-row group
//                        (min value, lowest row, multiplicity) in shared memory -- pops the closest pair and
//                        publishes either the merge or a request to revalidate a stale candidate;
//   all warps           execute the request: rescan one row (find_min_dist) or sweep the live rows
//                        (Lance-Williams update of row/column y, neighbour fix-ups, group minima, y's next
//                        nearest neighbour).  Global loads are issued before any dependent work so that each
//                        phase pays the L2 latency once.
// Two block barriers per request.
template <int MODE, int T>
__global__ void __launch_bounds__(T)
    linkage_fast_kernel(const LinkWork* __restrict__ works, const int* ns, int* __restrict__ need_exact) {
    const LinkWork w = works[blockIdx.x];
    const int n = ns[blockIdx.x];
    extern __shared__ __align__(16) unsigned char lk_smem[];
    constexpr int NW = T / 32;
    constexpr int PF = 4;  // groups (sweep) / strides (rescan) whose loads are in flight together
    __shared__ MergeRec rec;
    __shared__ double part_v[NW];
    __shared__ int part_i[NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) need_exact[blockIdx.x] = 0;
    if (n < 2) return;

    const int NG = (n + 31) / 32;        // groups of 32 rows
    const int HG = (n - 1 + 31) / 32;    // groups that contain heap rows (0 .. n-2)
    double *lb, *cur;
    GroupMin* gm;
    int *nbr, *size, *cid;
    {
        const size_t N = (size_t)NG * 32;
        unsigned char* p = MODE == LF_SMEM ? lk_smem : reinterpret_cast<unsigned char*>(w.fast_scratch);
        gm = reinterpret_cast<GroupMin*>(p);
        p += (size_t)NG * sizeof(GroupMin);
        lb = reinterpret_cast<double*>(p);
        p += N * 8;
        cur = reinterpret_cast<double*>(p);
        p += N * 8;
        nbr = reinterpret_cast<int*>(p);
        p += N * 4;
        size = reinterpret_cast<int*>(p);
        p += N * 4;
        cid = reinterpret_cast<int*>(p);
    }
    for (int i = tid; i < NG * 32; i += T) {
        size[i] = i < n ? 1 : 0;
        cid[i] = i;
        nbr[i] = i < n - 1 ? w.nbr[i] : -1;
        lb[i] = i < n - 1 ? w.lb[i] : INFINITY;
        cur[i] = lb[i];  // D[i][nbr[i]] equals the row minimum right after initialisation
    }
    __syncthreads();

    auto store_group = [&](int g, const Top& t) {
        if (lane == 0) {
            GroupMin e;
            e.v = t.v;
            e.i = t.i;
            e.c = t.c;
            gm[g] = e;
        }
    };
    // group minimum over live heap rows; all lanes of the calling warp
    auto group_min = [&](int g) {
        const int z = g * 32 + lane;
        const bool in = z < n - 1 && size[z] != 0;
        store_group(g, warp_top(in ? lb[z] : INFINITY, z, in ? 1 : 0));
    };
    // smallest bound over all groups (control warp)
    auto select_top = [&]() -> Top {
        Top m;
        m.v = INFINITY;
        m.i = -1;
        m.c = 0;
        for (int g = lane; g < HG; g += 32) {
            const GroupMin e = gm[g];
            if (e.c == 0) continue;
            if (m.c == 0 || e.v < m.v) {
                m.v = e.v;
                m.i = e.i;
                m.c = e.c;
            } else if (e.v == m.v)
                m.c += e.c;  // g ascending: the stored row stays the lowest
        }
        return warp_top(m.v, m.i, m.c);
    };
    for (int g = warp; g < HG; g += NW) group_min(g);
    __syncthreads();

    int k = 0;
    int tries = 0;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, ta, tb;
    for (;;) {
        ta = clock64();
        // ================= control warp: pop, decide, publish =================
        if (warp == 0) {
            int op = OP_MERGE, x = -1, y = -1;
            double dist = 0.0;
            if (k >= n - 1) {
                op = OP_ABORT;  // all merges done
            } else {
                const Top top = select_top();
                x = top.i;
                dist = top.v;
                if (top.c != 1 || x < 0 || tries >= n - k) {  // tied minimum: the heap order would matter
                    if (lane == 0) need_exact[blockIdx.x] = 1;
                    op = OP_ABORT;
                } else {
                    y = nbr[x];
                    if (!(y >= 0 && dist == cur[x])) op = OP_RESCAN;  // stale candidate (clustering.cpp:329)
                }
            }
            if (lane == 0) {
                rec.op = op;
                rec.x = x;
                rec.y = y;
                rec.dist = dist;
                if (op == OP_MERGE) {
                    const int nx = size[x], ny = size[y];
                    const int ix = cid[x], iy = cid[y];
                    rec.nx = nx;
                    rec.ny = ny;
                    double* z = w.Z + 4 * (size_t)k;  // clustering.cpp:347-358
                    z[0] = ix < iy ? ix : iy;
                    z[1] = ix < iy ? iy : ix;
                    z[2] = dist;
                    z[3] = nx + ny;
                    size[x] = 0;
                    size[y] = nx + ny;
                    cid[y] = n + k;
                }
            }
        }
        tb = clock64();
        c0 += tb - ta;  // select + publish (warp 0)
        __syncthreads();  // B1: the request is published
        ta = clock64();
        c1 += ta - tb;  // B1
        const int op = rec.op;
        if (op == OP_ABORT) {
            if (tid == 0 && w.stats) {  // cycle breakdown as seen by the control warp (sd_debug_counters)
                w.stats[3] += c0;       // pop + publish
                w.stats[4] += c2 + c3;  // revalidation (row rescans)
                w.stats[5] += c4 + c5;  // Lance-Williams sweeps
                w.stats[6] += c1;       // waiting for the request barrier
                w.stats[7] += (unsigned long long)(n - 1);
            }
            return;
        }
        const int x = rec.x;
        if (op == OP_RESCAN) {
            // find_min_dist(x), clustering.cpp:259-276: first minimum in index order over live i > x
            const double* r = w.D + (size_t)x * w.ld;
            double bv = INFINITY;
            int bi = -1;
            for (int i0 = x + 1 + tid; i0 < n; i0 += T * PF) {
                double d[PF];
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int i = i0 + u * T;
                    d[u] = i < n ? r[i] : INFINITY;
                }
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int i = i0 + u * T;
                    if (i < n && size[i] != 0 && d[u] < bv) {
                        bv = d[u];
                        bi = i;
                    }
                }
            }
            const Top m = warp_top(bi >= 0 ? bv : INFINITY, bi, bi >= 0 ? 1 : 0);
            if (lane == 0) {
                part_v[warp] = m.v;
                part_i[warp] = m.i;
            }
            tb = clock64();
            c2 += tb - ta;  // rescan work
            __syncthreads();  // B2
            if (warp == 0) {
                const bool has = lane < NW && part_i[lane] >= 0;
                const Top t = warp_top(has ? part_v[lane] : INFINITY, has ? part_i[lane] : -1, has ? 1 : 0);
                if (lane == 0) {
                    const double v = t.i >= 0 ? t.v : INFINITY;
                    nbr[x] = t.i;
                    lb[x] = v;
                    cur[x] = v;
                    if (w.stats) w.stats[0] += 1;
                }
                __syncwarp();
                group_min(x >> 5);
                __syncwarp();
            }
            c3 += clock64() - tb;  // rescan B2 + post
            ++tries;
            continue;
        }
        // ================= all warps: Lance-Williams sweep (clustering.cpp:361-404) =================
        const int y = rec.y, nx = rec.nx, ny = rec.ny;
        const double dist = rec.dist;
        const double* rowx = w.D + (size_t)x * w.ld;
        double* rowy = w.D + (size_t)y * w.ld;
        double ymv = INFINITY;  // y's new nearest neighbour among z > y
        int ymi = -1;
        double fx = 0.0, fy = 0.0, fs = 1.0, t3 = 0.0, rs = 1.0;
        bool have_terms = false;
        for (int g0 = warp; g0 < NG; g0 += NW * PF) {
            double dx[PF], dy[PF];
            bool live[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int z = (g0 + u * NW) * 32 + lane;
                live[u] = z < n && z != x && z != y && size[z] != 0;
                dx[u] = live[u] ? rowx[z] : 0.0;
                dy[u] = live[u] ? rowy[z] : 0.0;
            }
            if (!have_terms) {
                // centroid update, clustering.cpp:250-256: the third term and the divisor do not depend on z;
                // computed here, under the shadow of the loads just issued
                fx = (double)nx;
                fy = (double)ny;
                fs = (double)(nx + ny);
                t3 = __ddiv_rn(__dmul_rn(__dmul_rn((double)(nx * ny), dist), dist), fs);
                rs = __drcp_rn(fs);
                have_terms = true;
            }
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int g = g0 + u * NW;
                if (g >= NG) break;
                const int z = g * 32 + lane;
                double lbz = INFINITY;
                if (live[u]) {
                    const double t1 = __dmul_rn(__dmul_rn(fx, dx[u]), dx[u]);
                    const double t2 = __dmul_rn(__dmul_rn(fy, dy[u]), dy[u]);
                    const double nd = __dsqrt_rn(div_by(__dsub_rn(__dadd_rn(t1, t2), t3), fs, rs));
                    rowy[z] = nd;
                    w.D[(size_t)z * w.ld + y] = nd;
                    if (z < y) {
                        int nb = nbr[z];
                        lbz = lb[z];
                        if (z < x && nb == x) nb = y;  // clustering.cpp:374-378
                        if (nd < lbz) {                // clustering.cpp:381-392
                            nb = y;
                            lbz = nd;
                            lb[z] = nd;
                        }
                        if (nb == y) cur[z] = nd;  // cur[z] mirrors D[z][nbr[z]]
                        nbr[z] = nb;
                    } else {
                        if (nd < ymv) {  // groups ascend within a warp: first strict minimum
                            ymv = nd;
                            ymi = z;
                        }
                        if (z < n - 1) lbz = lb[z];
                    }
                }
                if (g < HG) {
                    const bool in = live[u] && z < n - 1;
                    store_group(g, warp_top(in ? lbz : INFINITY, z, in ? 1 : 0));
                }
            }
        }
        {
            const Top t = warp_top(ymi >= 0 ? ymv : INFINITY, ymi, ymi >= 0 ? 1 : 0);
            if (lane == 0) {
                part_v[warp] = t.v;
                part_i[warp] = t.i;
            }
        }
        tb = clock64();
        c4 += tb - ta;  // sweep work
        __syncthreads();  // B2: sweep results are visible
        if (warp == 0 && y < n - 1) {
            const bool has = lane < NW && part_i[lane] >= 0;
            const Top t = warp_top(has ? part_v[lane] : INFINITY, has ? part_i[lane] : -1, has ? 1 : 0);
            if (lane == 0 && t.i != -1) {  // clustering.cpp:395-404
                nbr[y] = t.i;
                lb[y] = t.v;
                cur[y] = t.v;
            }
            __syncwarp();
            group_min(y >> 5);  // the sweep left y out of its group
            __syncwarp();
        }
        c5 += clock64() - tb;  // sweep B2 + post
        ++k;
        tries = 0;
    }
}

// ---- cluster path: the heap-free merge loop spread over a thread-block cluster ---------------------------
//
// Same algorithm and the same uniqueness proof as linkage_fast_kernel, but the rows are dealt out group by
// group (32 rows) to the 8 CTAs of a cluster and, inside a CTA, to its warps; a row's state (bound, candidate,
// cluster size and id, size and id of the candidate) is only ever touched by the one thread that owns the row,
// so there is no block barrier anywhere.  Every request (merge sweep or row rescan) ends with each warp
// publishing its partial results -- the smallest bound among its rows and its share of the new nearest
// neighbour of the row being recomputed -- into the shared memory of all 8 CTAs with st.async (DSMEM stores
// that complete a transaction count on the destination's mbarrier), so a request costs one DSMEM flight and
// one mbarrier wake-up instead of a cluster barrier with a GPU-scope fence.  After the wait every warp of every
// CTA reduces the same 8*NW entries and reaches the same decision: nothing has to be broadcast.
// The distance matrix stays in global memory (L2-resident).  Entries written by one CTA and read by another
// are ordered by a __threadfence() between the sweep's stores and the publish (merge requests only) and are
// read with ld.global.cg after a cluster-scope acquire on the mbarrier (measured: the weaker CTA-scope wait,
// which drops the CCTL.IVALL the compiler adds, does not change the run time).
constexpr int kLcCtas = 8;

// own rows (lb, cur, nbr, alive) + -- unless they live in global memory -- the replicas of cluster size / id
__host__ __device__ inline size_t linkcluster_smem_bytes(int n, bool global_replicas) {
    const size_t ng = (size_t)(n + 31) / 32;
    const size_t ngl = (ng + kLcCtas - 1) / kLcCtas;
    return ngl * 32 * (8 + 8 + 4 + 4) + (global_replicas ? 0 : ng * 32 * 8) + 64;
}
__host__ __device__ inline size_t linkcluster_global_bytes(int n) {  // replicas of all 8 CTAs
    return (size_t)kLcCtas * ((size_t)(n + 31) / 32 * 32) * 8 + 256;
}

namespace lc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async16(uint32_t raddr, uint32_t rbar, uint4 v) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LC_DONE_%=;\n"
        "bra LC_WAIT_%=;\n"
        "LC_DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// fire-and-forget fetch of one 128-byte line into L2 (no register, no scoreboard)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cluster_barrier_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 pack(double a, int b, int c) {
    uint4 r;
    r.x = (unsigned)__double2loint(a);
    r.y = (unsigned)__double2hiint(a);
    r.z = (unsigned)b;
    r.w = (unsigned)c;
    return r;
}
__device__ __forceinline__ double unpack_d(const uint4& v) { return __hiloint2double((int)v.y, (int)v.x); }
}  // namespace lc

// One exchange buffer: four 16-byte pieces per publishing warp + the state of the row being recomputed.
//   A = (smallest bound v1, its row i1, multiplicity c1)      B = (cur[i1], nbr[i1], -)
//   C = (second smallest bound v2 of the warp, -, -)          D = (partial value, partial row, -)
//   P0 = (lb, cur) P1 = (nbr, -, -, -) of the pending row (the survivor of the merge just swept)
template <int E>
struct __align__(16) ClusterXch {
    uint4 A[E], B[E], C[E], D[E], P[2];
};

// Requests that need the whole cluster (a merge sweep, or a refill of the candidate lists) end with an exchange;
// revalidations of stale candidates (clustering.cpp:329-337) do not: every CTA keeps a replica of the cluster
// sizes / ids, rescans the row redundantly from the L2-resident matrix and reaches the same result, so a
// revalidation costs one L2 round trip and one block barrier instead of a DSMEM round.  To keep deciding after
// a row of some warp has been revalidated, every warp also publishes the value of its second smallest bound:
// as long as the running minimum stays strictly below it, the rows that warp did not publish cannot matter.
// PRE (used from 512 threads per CTA on): the warps of a CTA first reduce their entries through shared memory (one
// block barrier) and warp 0 publishes ONE entry per CTA, so that the exchange and the redundant decide work on 8
// entries whatever the thread count is -- that is what lets large problems use 512 threads per CTA (fewer sequential
// load rounds per thread in the sweep and the rescans) without the decide growing with the number of warps.
// Register budget of the 128-thread form (the one every file of a batch runs): capped at 128 per thread = 16 K per CTA.
// An SM has 64 K registers and an STFT CTA takes 15.4 K (96 x 160): next to an uncapped merge-loop CTA (167 registers,
// 21 K) only two STFT CTAs fit, next to a capped one three -- and with 16 files in flight the merge loops sit on 128
// of the 148 SMs, so this decides how fast the bandwidth-bound STFTs of the other files run underneath them.
template <int T, bool GREPL, bool PRE>
__global__ void __cluster_dims__(kLcCtas, 1, 1) __launch_bounds__(T, T == 128 ? 4 : 1)
    linkage_cluster_kernel(const LinkWork* __restrict__ works, const int* ns, int* __restrict__ need_exact) {
    constexpr int NW = T / 32;
    constexpr int E = PRE ? kLcCtas : kLcCtas * NW;
    constexpr int EPL = (E + 31) / 32;  // exchange entries per lane
    constexpr int PF = 4;               // sweep / rescan groups in flight per warp
    constexpr int LOOSE = 4;            // rows revalidated since the last exchange that every warp tracks
    constexpr uint32_t kTxBytes = E * 64 + 32;
    const int prob = blockIdx.x / kLcCtas;
    int rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const LinkWork w = works[prob];
    const int n = ns[prob];
    extern __shared__ __align__(16) unsigned char lc_smem[];
    __shared__ ClusterXch<E> xch[2];
    __shared__ __align__(16) uint4 rx[2][E];  // revalidation exchange: (partial value, partial row) per warp
    __shared__ __align__(8) unsigned long long mbar[2], rbar[2];
    // PRE: per-warp entries of this CTA, combined by warp 0 before they leave the CTA
    constexpr int PW = PRE ? NW : 1;
    __shared__ double pre_v[PW], pre_v2[PW], pre_cur[PW], pre_pv[PW], prx_v[PW];
    __shared__ int pre_i[PW], pre_c[PW], pre_nbr[PW], pre_pi[PW], prx_i[PW];
    __shared__ __align__(16) uint4 pre_pend[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool scribe = rank == 0 && tid == 0;  // writes Z / flags / counters
    if (scribe) need_exact[prob] = 0;
    if (n < 2) return;

    const int NG = (n + 31) / 32;
    const int NGl = NG > rank ? (NG - rank + kLcCtas - 1) / kLcCtas : 0;  // groups of this CTA: g = lg * 8 + rank
    const int NGmax = (NG + kLcCtas - 1) / kLcCtas;
    // A matrix beyond the L2 capacity streams its rows from HBM on every merge: a warp then has only PF groups of
    // loads in flight per ~1 us round trip.  In that regime the rows' later groups are prefetched into L2 while the
    // first round is on its way (lanes 0/1: the two 128-byte lines of a 32-row group).
    const bool cold = (size_t)n * (size_t)w.ld * sizeof(double) > ((size_t)64 << 20);
    // own rows (private to the owning thread)
    double* lb = reinterpret_cast<double*>(lc_smem);
    double* cur = lb + (size_t)NGmax * 32;
    int* nbr = reinterpret_cast<int*>(cur + (size_t)NGmax * 32);
    int* alive = nbr + (size_t)NGmax * 32;
    // replicas, all rows: cluster size and id -- in shared memory, or (large N) this CTA's copy in global memory
    int* rsize = GREPL ? reinterpret_cast<int*>(w.fast_scratch) + (size_t)rank * 2 * NG * 32 : alive + (size_t)NGmax * 32;
    int* rcid = rsize + (size_t)NG * 32;

    constexpr uint32_t kRxBytes = E * 16;
    const uint32_t bar_local[2] = {lc::smem_u32(&mbar[0]), lc::smem_u32(&mbar[1])};
    const uint32_t rbar_local[2] = {lc::smem_u32(&rbar[0]), lc::smem_u32(&rbar[1])};
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            lc::mbar_init(bar_local[b], 1);
            lc::mbar_init(rbar_local[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            lc::mbar_expect_tx(bar_local[b], kTxBytes);
            lc::mbar_expect_tx(rbar_local[b], kRxBytes);
        }
    }
    // this lane's st.async destination: CTA (lane & 7), piece (lane >> 3)
    const uint32_t dst_rank = (uint32_t)(lane & 7);
    const int piece = lane >> 3;
    uint32_t dst_piece[2], dst_pend[2], dst_bar[2], dst_rx[2], dst_rbar[2];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const uint4* base = piece == 0 ? xch[b].A : piece == 1 ? xch[b].B : piece == 2 ? xch[b].C : xch[b].D;
        dst_piece[b] = lc::mapa(lc::smem_u32(base + (PRE ? rank : rank * NW + warp)), dst_rank);
        dst_pend[b] = lc::mapa(lc::smem_u32(&xch[b].P[piece & 1]), dst_rank);
        dst_bar[b] = lc::mapa(bar_local[b], dst_rank);
        dst_rx[b] = lc::mapa(lc::smem_u32(&rx[b][PRE ? rank : rank * NW + warp]), dst_rank);
        dst_rbar[b] = lc::mapa(rbar_local[b], dst_rank);
    }

    auto row_of = [&](int lg) { return ((lg * kLcCtas + rank) << 5) + lane; };
    auto owner_rank = [&](int z) { return (z >> 5) % kLcCtas; };
    auto owner_lg = [&](int z) { return (z >> 5) / kLcCtas; };
    auto entry_of = [&](int z) { return PRE ? owner_rank(z) : owner_rank(z) * NW + owner_lg(z) % NW; };
    auto mine = [&](int z) { return z >= 0 && owner_rank(z) == rank && owner_lg(z) % NW == warp && (z & 31) == lane; };

    for (int lg = warp; lg < NGl; lg += NW) {
        const int z = row_of(lg), s = lg * 32 + lane;
        nbr[s] = z < n - 1 ? w.nbr[z] : -1;
        lb[s] = z < n - 1 ? w.lb[z] : INFINITY;
        cur[s] = lb[s];
        alive[s] = z < n ? 1 : 0;
    }
    for (int i = tid; i < NG * 32; i += T) {
        rsize[i] = i < n ? 1 : 0;
        rcid[i] = i;
    }
    __syncthreads();
    lc::cluster_barrier_relaxed();  // every mbarrier of the cluster is initialised before the first st.async

    // the two smallest bounds among this warp's live heap rows (skip_a / skip_b left out): (v1, i1, c1) and v2
    auto warp_rows_top2 = [&](int skip_a, int skip_b, double& v2) -> Top {
        Top m;
        m.v = INFINITY;
        m.i = -1;
        m.c = 0;
        double second = INFINITY;  // smallest value of this lane not counted in m
        for (int lg = warp; lg < NGl; lg += NW) {
            const int z = row_of(lg), s = lg * 32 + lane;
            if (z < n - 1 && z != skip_a && z != skip_b && alive[s] != 0) {
                const double v = lb[s];
                if (m.c == 0 || v < m.v) {
                    if (m.c) second = m.v;  // the old best (and anything tied with it) is now second
                    m.v = v;
                    m.i = z;
                    m.c = 1;
                } else if (v == m.v) {
                    ++m.c;
                } else if (v < second)
                    second = v;
            }
        }
        const Top t = warp_top(m.c ? m.v : INFINITY, m.i, m.c);
        // second smallest of the warp: lanes that lost contribute their best, the winner (or lanes tied with it:
        // a tie aborts the fast path anyway) its own second
        const bool at_min = m.c > 0 && m.v == t.v;
        double cand = at_min ? second : (m.c ? m.v : INFINITY);
        const unsigned hi = (unsigned)__double2hiint(cand);
        const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? (unsigned)__double2loint(cand) : 0xffffffffu);
        v2 = __hiloint2double((int)mh, (int)ml);
        return t;
    };

    // publish this warp's entry and, from the warp that owns it, the state of the pending row, to every CTA of
    // the cluster: one 16-byte st.async per lane
    auto publish = [&](int b, const Top& t, double v2, double pv, int pi, int pend, bool dummy_pend) {
        double d_cur = 0.0;
        int d_nbr = -1;
        if (t.i >= 0) {
            const int src = t.i & 31, s = owner_lg(t.i) * 32 + src;
            const bool me = lane == src;
            d_cur = __shfl_sync(0xffffffffu, me ? cur[s] : 0.0, src);
            d_nbr = __shfl_sync(0xffffffffu, me ? nbr[s] : 0, src);
        }
        const bool own_pend = pend >= 0 ? (owner_rank(pend) == rank && owner_lg(pend) % NW == warp) : false;
        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
        if (own_pend) {
            const int src = pend & 31, s = owner_lg(pend) * 32 + src;
            const bool me = lane == src;
            const double q_lb = __shfl_sync(0xffffffffu, me ? lb[s] : 0.0, src);
            const double q_cur = __shfl_sync(0xffffffffu, me ? cur[s] : 0.0, src);
            q0.x = (unsigned)__double2loint(q_lb);
            q0.y = (unsigned)__double2hiint(q_lb);
            q0.z = (unsigned)__double2loint(q_cur);
            q0.w = (unsigned)__double2hiint(q_cur);
            q1.x = (unsigned)__shfl_sync(0xffffffffu, me ? nbr[s] : 0, src);
        }
        Top pt = t;
        bool send_pend = own_pend || (pend < 0 && dummy_pend);
        if (PRE) {
            // per-warp entries -> shared memory -> warp 0 combines them into the CTA's one entry
            if (lane == 0) {
                pre_v[warp] = t.v;
                pre_i[warp] = t.i;
                pre_c[warp] = t.c;
                pre_cur[warp] = d_cur;
                pre_nbr[warp] = d_nbr;
                pre_v2[warp] = v2;
                pre_pv[warp] = pv;
                pre_pi[warp] = pi;
                if (own_pend) {
                    pre_pend[0] = q0;
                    pre_pend[1] = q1;
                }
            }
            __syncthreads();
            if (warp != 0) return;
            const bool has = lane < NW;
            const double ev = has && pre_c[lane] > 0 ? pre_v[lane] : INFINITY;
            pt = warp_top(ev, has ? pre_i[lane] : -1, has ? pre_c[lane] : 0);
            // second smallest bound of the CTA: the winner's own second, everybody else's first
            const bool winner = has && pre_c[lane] > 0 && pre_i[lane] == pt.i;
            v2 = warp_top(has ? (winner ? pre_v2[lane] : ev) : INFINITY, 0, 1).v;
            const unsigned wb = __ballot_sync(0xffffffffu, winner);
            const int wsrc = wb ? __ffs(wb) - 1 : 0;
            d_cur = __shfl_sync(0xffffffffu, has ? pre_cur[lane] : 0.0, wsrc);
            d_nbr = __shfl_sync(0xffffffffu, has ? pre_nbr[lane] : -1, wsrc);
            if (!wb) {
                d_cur = 0.0;
                d_nbr = -1;
            }
            const int ppi = has ? pre_pi[lane] : -1;
            const Top pp = warp_top(ppi >= 0 ? pre_pv[lane] : INFINITY, ppi, ppi >= 0 ? 1 : 0);
            pv = pp.i >= 0 ? pp.v : INFINITY;
            pi = pp.i;
            if (pend >= 0) {
                send_pend = owner_rank(pend) == rank;
                q0 = pre_pend[0];
                q1 = pre_pend[1];
            } else
                send_pend = dummy_pend;
        }
        uint4 v;
        if (piece == 0)
            v = lc::pack(pt.v, pt.i, pt.c);
        else if (piece == 1)
            v = lc::pack(d_cur, d_nbr, 0);
        else if (piece == 2)
            v = lc::pack(v2, 0, 0);
        else
            v = lc::pack(pv, pi, 0);
        lc::st_async16(dst_piece[b], dst_bar[b], v);
        if (send_pend && piece < 2) lc::st_async16(dst_pend[b], dst_bar[b], piece == 0 ? q0 : q1);
    };

    int par = 0;
    unsigned phase[2] = {0u, 0u};
    int pending = -1;        // survivor of the merge whose sweep preceded the exchange in flight
    int dead = -1;           // the row that merge removed
    int pend_size = 0, pend_cid = 0;
    {
        double v2;
        const Top t = warp_rows_top2(-1, -1, v2);
        publish(par, t, v2, INFINITY, -1, -1, rank == 0 && warp == 0);
    }

    int k = 0, tries = 0, rpar = 0;
    unsigned rphase[2] = {0u, 0u};
    unsigned long long rescans = 0, refills = 0;
    long long c_dec = 0, c_work = 0, c_res = 0, c_bar = 0, t0, t1;
    t1 = clock64();
    for (;;) {
        lc::mbar_wait(bar_local[par], phase[par] & 1u);
        ++phase[par];
        if (tid == 0) {
            lc::mbar_expect_tx(bar_local[par], kTxBytes);  // re-arm for the round after next
            if (pending >= 0) {                            // replicas follow the merge (clustering.cpp:347-358)
                rsize[dead] = 0;
                rsize[pending] = pend_size;
                rcid[pending] = pend_cid;
            }
        }
        __syncthreads();
        t0 = clock64();
        c_bar += t0 - t1;
        const ClusterXch<E>& X = xch[par];
        // ---------------- every warp: load its share of the exchange ----------------
        double a_v[EPL], b_cur[EPL], c_v2[EPL];
        int a_i[EPL], a_c[EPL], b_nbr[EPL];
        bool used[EPL];
        double pv = INFINITY;
        int pi = -1;
#pragma unroll
        for (int u = 0; u < EPL; ++u) {
            const int e = lane + 32 * u;
            a_v[u] = INFINITY;
            a_i[u] = -1;
            a_c[u] = 0;
            b_cur[u] = 0.0;
            b_nbr[u] = -1;
            c_v2[u] = INFINITY;
            used[u] = false;
            if (e < E) {
                const uint4 a = X.A[e], bq = X.B[e], cq = X.C[e], dq = X.D[e];
                a_v[u] = lc::unpack_d(a);
                a_i[u] = (int)a.z;
                a_c[u] = (int)a.w;
                b_cur[u] = lc::unpack_d(bq);
                b_nbr[u] = (int)bq.z;
                c_v2[u] = lc::unpack_d(cq);
                const int qi = (int)dq.z;
                if (qi >= 0) {
                    const double qv = lc::unpack_d(dq);
                    if (pi < 0 || qv < pv || (qv == pv && qi < pi)) {
                        pv = qv;
                        pi = qi;
                    }
                }
            }
        }
        // rows whose bound changed since the exchange was published (same in every warp of the cluster)
        double l_v[LOOSE], l_cur[LOOSE];
        int l_i[LOOSE], l_nbr[LOOSE];
        int nl = 0;
        if (pending >= 0) {
            const Top pt = warp_top(pi >= 0 ? pv : INFINITY, pi, pi >= 0 ? 1 : 0);
            const uint4 p0 = X.P[0], p1 = X.P[1];
            double q_lb = __hiloint2double((int)p0.y, (int)p0.x), q_cur = __hiloint2double((int)p0.w, (int)p0.z);
            int q_nbr = (int)p1.x;
            if (pt.i >= 0) {  // new nearest neighbour of the survivor (clustering.cpp:395-404)
                q_lb = pt.v;
                q_cur = pt.v;
                q_nbr = pt.i;
            }
            if (mine(pending)) {
                const int s = owner_lg(pending) * 32 + lane;
                lb[s] = q_lb;
                cur[s] = q_cur;
                nbr[s] = q_nbr;
            }
            if (pending < n - 1) {
                l_v[0] = q_lb;
                l_cur[0] = q_cur;
                l_i[0] = pending;
                l_nbr[0] = q_nbr;
                nl = 1;
            }
            pending = -1;
        }
        // ---------------- pop until a valid candidate, a tie, or a list runs dry ----------------
        int op = 0;  // 0 merge, 1 refill, 2 abort (need the exact kernel)
        int x = -1, y = -1;
        double dist = 0.0;
        for (;;) {
            Top m;
            m.v = INFINITY;
            m.i = -1;
            m.c = 0;
#pragma unroll
            for (int u = 0; u < EPL; ++u) {
                // a list whose head was consumed stands in with the value of its second entry (row unknown: -2)
                const double v = used[u] ? c_v2[u] : a_v[u];
                const int i = used[u] ? -2 : a_i[u];
                const int c = used[u] ? (c_v2[u] < INFINITY ? 1 : 0) : a_c[u];
                if (c > 0) {
                    if (m.c == 0 || v < m.v) {
                        m.v = v;
                        m.i = i;
                        m.c = c;
                    } else if (v == m.v) {
                        m.c += c;
                        m.i = i < m.i ? i : m.i;
                    }
                }
            }
            Top top = warp_top(m.c ? m.v : INFINITY, m.i, m.c);
            int loose_src = -1;
#pragma unroll
            for (int j = 0; j < LOOSE; ++j)
                if (j < nl) {
                    if (top.c == 0 || l_v[j] < top.v) {
                        top.v = l_v[j];
                        top.i = l_i[j];
                        top.c = 1;
                        loose_src = j;
                    } else if (l_v[j] == top.v)
                        top.c += 1;
                }
            x = top.i;
            dist = top.v;
            if (top.c != 1 || x == -1 || tries >= n - k) {  // tied minimum: the heap order would matter
                op = 2;
                break;
            }
            if (x == -2) {  // the minimum may be a row nobody published: refill the lists
                op = 1;
                break;
            }
            double x_cur;
            int x_nbr;
            if (loose_src >= 0) {
                x_cur = l_cur[0];
                x_nbr = l_nbr[0];
#pragma unroll
                for (int j = 1; j < LOOSE; ++j)
                    if (j == loose_src) {
                        x_cur = l_cur[j];
                        x_nbr = l_nbr[j];
                    }
            } else {
                const int e = entry_of(x), src = e & 31, uu = e >> 5;
                double sc = b_cur[0];
                int sn = b_nbr[0];
#pragma unroll
                for (int u = 1; u < EPL; ++u)
                    if (u == uu) {
                        sc = b_cur[u];
                        sn = b_nbr[u];
                    }
                x_cur = __shfl_sync(0xffffffffu, sc, src);
                x_nbr = __shfl_sync(0xffffffffu, sn, src);
            }
            y = x_nbr;
            if (y >= 0 && dist == x_cur) {  // valid candidate (clustering.cpp:329)
                op = 0;
                break;
            }
            // ---- stale: find_min_dist(x) (clustering.cpp:259-276).  Every warp scans its own rows i > x, the 8*NW
            // partial minima are exchanged (one 16-byte st.async per destination CTA) and reduced by everyone ----
            long long tr0 = clock64();
            {
                const double* r = w.D + (size_t)x * w.ld;
                double bv = INFINITY;
                int bi = -1;
                if (cold && lane < 2)
                    for (int lg = warp + NW * PF; lg < NGl; lg += NW) {
                        const int i0 = (lg * kLcCtas + rank) << 5;
                        if (i0 + 31 > x) lc::prefetch_l2(r + i0 + 16 * lane);
                    }
                for (int lg0 = warp; lg0 < NGl; lg0 += NW * PF) {
                    double d[PF];
                    bool ok[PF];
#pragma unroll
                    for (int u = 0; u < PF; ++u) {
                        const int lg = lg0 + u * NW;
                        const int i = row_of(lg);
                        ok[u] = lg < NGl && i > x && i < n && alive[lg * 32 + lane] != 0;
                        d[u] = ok[u] ? __ldcg(r + i) : INFINITY;
                    }
#pragma unroll
                    for (int u = 0; u < PF; ++u)
                        if (ok[u] && d[u] < bv) {  // rows ascend within a lane: first minimum in index order
                            bv = d[u];
                            bi = row_of(lg0 + u * NW);
                        }
                }
                Top part = warp_top(bi >= 0 ? bv : INFINITY, bi, bi >= 0 ? 1 : 0);
                bool sender = true;
                if (PRE) {  // one partial per CTA
                    if (lane == 0) {
                        prx_v[warp] = part.v;
                        prx_i[warp] = part.i;
                    }
                    __syncthreads();
                    sender = warp == 0;
                    if (sender) {
                        const int qi_ = lane < NW ? prx_i[lane] : -1;
                        part = warp_top(qi_ >= 0 ? prx_v[lane] : INFINITY, qi_, qi_ >= 0 ? 1 : 0);
                    }
                }
                if (sender && lane < kLcCtas)
                    lc::st_async16(dst_rx[rpar], dst_rbar[rpar], lc::pack(part.i >= 0 ? part.v : INFINITY, part.i, 0));
                lc::mbar_wait(rbar_local[rpar], rphase[rpar] & 1u);
                ++rphase[rpar];
                if (tid == 0) lc::mbar_expect_tx(rbar_local[rpar], kRxBytes);  // re-arm for the revalidation after next
                double qv = INFINITY;
                int qi = -1;
#pragma unroll
                for (int u = 0; u < EPL; ++u) {
                    const int e = lane + 32 * u;
                    if (e < E) {
                        const uint4 q = rx[rpar][e];
                        const int ci = (int)q.z;
                        const double cv = lc::unpack_d(q);
                        if (ci >= 0 && (qi < 0 || cv < qv || (cv == qv && ci < qi))) {
                            qv = cv;
                            qi = ci;
                        }
                    }
                }
                const Top res = warp_top(qi >= 0 ? qv : INFINITY, qi, qi >= 0 ? 1 : 0);
                rpar ^= 1;
                const double nv = res.i >= 0 ? res.v : INFINITY;
                if (mine(x)) {
                    const int s = owner_lg(x) * 32 + lane;
                    lb[s] = nv;
                    cur[s] = nv;
                    nbr[s] = res.i;
                }
                ++tries;
                ++rescans;
                if (loose_src >= 0) {
#pragma unroll
                    for (int j = 0; j < LOOSE; ++j)
                        if (j == loose_src) {
                            l_v[j] = nv;
                            l_cur[j] = nv;
                            l_nbr[j] = res.i;
                        }
                } else {
                    const int e = entry_of(x);
                    if (lane == (e & 31)) {
#pragma unroll
                        for (int u = 0; u < EPL; ++u)
                            if (u == (e >> 5)) used[u] = true;
                    }
                    if (nl == LOOSE) {  // no room to track another row: republish everything
                        op = 1;
                        c_res += clock64() - tr0;
                        break;
                    }
#pragma unroll
                    for (int j = 0; j < LOOSE; ++j)
                        if (j == nl) {
                            l_v[j] = nv;
                            l_cur[j] = nv;
                            l_i[j] = x;
                            l_nbr[j] = res.i;
                        }
                    ++nl;
                }
            }
            c_res += clock64() - tr0;
        }
        if (op == 2) {
            if (scribe) need_exact[prob] = 1;
            break;
        }
        par ^= 1;
        t1 = clock64();
        c_dec += t1 - t0;
        if (op == 1) {  // refill: every warp republishes its two smallest bounds
            double v2;
            const Top t = warp_rows_top2(-1, -1, v2);
            publish(par, t, v2, INFINITY, -1, -1, rank == 0 && warp == 0);
            ++refills;
            t0 = clock64();
            c_work += t0 - t1;
            t1 = t0;
            continue;
        }
        // ---------------- merge x into y (clustering.cpp:347-404) ----------------
        const int nx = rsize[x], ny = rsize[y], ix = rcid[x], iy = rcid[y];
        if (scribe) {
            double* z = w.Z + 4 * (size_t)k;
            z[0] = ix < iy ? ix : iy;
            z[1] = ix < iy ? iy : ix;
            z[2] = dist;
            z[3] = nx + ny;
        }
        const double* rowx = w.D + (size_t)x * w.ld;
        double* rowy = w.D + (size_t)y * w.ld;
        double fx = 0.0, fy = 0.0, fs = 1.0, t3 = 0.0, rs = 1.0;
        bool have_terms = false;
        double ymv = INFINITY;
        int ymi = -1;
        if (cold && lane < 4)
            for (int lg = warp + NW * PF; lg < NGl; lg += NW) {
                const int z0 = (lg * kLcCtas + rank) << 5;
                lc::prefetch_l2((lane < 2 ? rowx : rowy) + z0 + 16 * (lane & 1));
            }
        for (int lg0 = warp; lg0 < NGl; lg0 += NW * PF) {
            double dx[PF], dy[PF];
            bool live[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int lg = lg0 + u * NW;
                const int z = row_of(lg);
                live[u] = lg < NGl && z < n && z != x && z != y && alive[lg * 32 + lane] != 0;
                if (lg < NGl && z == x) alive[lg * 32 + lane] = 0;
                dx[u] = live[u] ? __ldcg(rowx + z) : 0.0;
                dy[u] = live[u] ? __ldcg(rowy + z) : 0.0;
            }
            if (!have_terms) {
                // centroid update, clustering.cpp:250-256: the third term and the divisor do not depend on z;
                // computed under the shadow of the loads just issued
                fx = (double)nx;
                fy = (double)ny;
                fs = (double)(nx + ny);
                t3 = __ddiv_rn(__dmul_rn(__dmul_rn((double)(nx * ny), dist), dist), fs);
                rs = __drcp_rn(fs);
                have_terms = true;
            }
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int lg = lg0 + u * NW;
                if (lg >= NGl) break;
                if (!live[u]) continue;
                const int z = row_of(lg), s = lg * 32 + lane;
                const double t1q = __dmul_rn(__dmul_rn(fx, dx[u]), dx[u]);
                const double t2q = __dmul_rn(__dmul_rn(fy, dy[u]), dy[u]);
                const double nd = __dsqrt_rn(div_by(__dsub_rn(__dadd_rn(t1q, t2q), t3), fs, rs));
                __stcg(rowy + z, nd);
                __stcg(w.D + (size_t)z * w.ld + y, nd);
                if (z < y) {
                    int nb = nbr[s];
                    const double lbz = lb[s];
                    if (z < x && nb == x) nb = y;  // clustering.cpp:374-378
                    if (nd < lbz) {                // clustering.cpp:381-392
                        nb = y;
                        lb[s] = nd;
                    }
                    if (nb == y) cur[s] = nd;  // cur mirrors D[z][nbr[z]]
                    nbr[s] = nb;
                } else if (nd < ymv) {  // rows ascend within a lane: first strict minimum
                    ymv = nd;
                    ymi = z;
                }
            }
        }
        pending = y;
        dead = x;
        pend_size = nx + ny;
        pend_cid = n + k;
        ++k;
        tries = 0;
        if (k >= n - 1) break;  // that was the last merge: nothing left to exchange
        const Top part = warp_top(ymi >= 0 ? ymv : INFINITY, ymi, ymi >= 0 ? 1 : 0);
        double v2;
        const Top t = warp_rows_top2(x, y, v2);
#if defined(SD_LC_FENCE_CLUSTER)
        asm volatile("fence.acq_rel.cluster;" ::: "memory");  // experiment: every consumer is in this cluster
#else
        __threadfence();  // the sweep's stores are visible device-wide before any peer can see this publish
#endif
        publish(par, t, v2, part.i >= 0 ? part.v : INFINITY, part.i, y, false);
        t0 = clock64();
        c_work += t0 - t1;
        t1 = t0;
    }
    if (scribe && w.stats) {
        w.stats[0] += rescans;
        w.stats[1] += refills;
        w.stats[3] += c_dec - c_res;  // reduce the exchange + decide
        w.stats[4] += c_res;          // revalidation (local row rescans)
        w.stats[5] += c_work;         // sweep + fence + publish
        w.stats[6] += c_bar;          // mbarrier wait
        w.stats[7] += (unsigned long long)(n - 1);
    }
    lc::cluster_barrier_relaxed();  // no CTA may exit while a peer can still write into its shared memory
}

// ------------------------------------------------------------------------------------------------
// linkage_wide_kernel: the heap-free merge loop spread over the whole GPU (large N)
// ------------------------------------------------------------------------------------------------
// One 8-CTA cluster streams a matrix that no longer fits L2 through 8 SMs: 11 us per merge at N = 10 773, 25 us at
// N = 50 000.  Here G co-resident CTAs (one per SM, cooperative launch) share the rows, 32-row groups dealt round
// robin, and the per-round exchange goes through a global-memory mailbox instead of DSMEM:
//   * every round each CTA publishes one 64-byte box -- the smallest bound among its own live rows (value, row,
//     multiplicity), that row's cached candidate (cur, nbr), its share of the recomputed row's new nearest neighbour,
//     and, from the owner, the old state of the recomputed row -- then raises its flag (fence + store);
//   * G threads of every CTA poll one flag each (ld.acquire.gpu) and read one box each; two block reductions later
//     every CTA has reached the same decision, exactly as in linkage_cluster_kernel: nothing is broadcast;
//   * a round executes ONE request: revalidate a stale candidate (every thread scans its rows of that matrix row) or
//     merge (Lance-Williams sweep over the own rows, both D[y][z] and D[z][y] written).  The row whose bound the
//     request recomputes (x after a rescan, the survivor y after a merge) is "pending": its owner leaves it out of the
//     published minimum and everybody derives its new state from the partial results of the next exchange.
// Same algorithm, same uniqueness argument and the same hand-over to the exact heap kernel on a tied minimum as the
// other two heap-free kernels, so Z is the reference's bit for bit.  Cluster size / id replicas are private to each
// CTA (global memory, L1-cacheable since nobody else touches them).
// Measured (profiles/r02_cluster_sizes_wide*.log): a round costs ~7 us (flag + box round trips through L2, the
// per-thread fences, four block barriers, the skew of 148 CTAs) and a merge needs ~3 rounds (one per revalidation):
// 20.4 us per merge at N = 10 773, 25.5 us at N = 50 000 -- slower than the cluster kernel at the first size, level
// at the second.  Kept as an opt-in (SD_OPT_LINKAGE_WIDE), bit-exact (tests/test_gpu_linkage_wide.py); what it
// shows is that the round count, not the width, is what a faster large-N merge loop has to attack.
constexpr int LW_T = 256;
constexpr int LW_NW = LW_T / 32;
constexpr int LW_MAX_CTAS = 160;

struct __align__(16) WideBox {
    uint4 A;  // v1 (double), i1, c1
    uint4 B;  // cur[i1] (double), nbr[i1], -
    uint4 C;  // partial value (double), partial row, old nbr of the pending row (owner only)
    uint4 D;  // old lb, old cur of the pending row (owner only)
};

__host__ __device__ inline size_t linkwide_global_bytes(int n, int ctas) {
    const size_t N32 = (size_t)(n + 31) / 32 * 32;
    return (size_t)ctas * N32 * 8 + 2 * (size_t)ctas * sizeof(WideBox) + (size_t)LW_MAX_CTAS * LW_MAX_CTAS * 4 + 512;
}

namespace lw {
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_cg16(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 packd(double a, int b, int c) {
    uint4 r;
    r.x = (unsigned)__double2loint(a);
    r.y = (unsigned)__double2hiint(a);
    r.z = (unsigned)b;
    r.w = (unsigned)c;
    return r;
}
__device__ __forceinline__ double lod(const uint4& v) { return __hiloint2double((int)v.y, (int)v.x); }
__device__ __forceinline__ double hid(const uint4& v) { return __hiloint2double((int)v.w, (int)v.z); }
}  // namespace lw

__global__ void __launch_bounds__(LW_T)
    linkage_wide_kernel(const LinkWork* __restrict__ works, const int* ns, int* __restrict__ need_exact) {
    const LinkWork w = works[0];
    const int n = ns[0];
    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    extern __shared__ __align__(16) unsigned char lw_smem[];
    __shared__ double s_tv[LW_NW], s_pv[LW_NW];
    __shared__ int s_ti[LW_NW], s_tc[LW_NW], s_pi[LW_NW];
    __shared__ double s_xcur, s_plb, s_pcur;
    __shared__ int s_xnbr, s_pnbr;
    const bool scribe = cta == 0 && tid == 0;
    if (scribe) need_exact[0] = 0;
    if (n < 2) return;

    const int NG = (n + 31) / 32;
    const int NGl = NG > cta ? (NG - cta + G - 1) / G : 0;  // groups of this CTA: g = lg * G + cta
    const int NGmax = (NG + G - 1) / G;
    double* lb = reinterpret_cast<double*>(lw_smem);
    double* cur = lb + (size_t)NGmax * 32;
    int* nbr = reinterpret_cast<int*>(cur + (size_t)NGmax * 32);
    int* alive = nbr + (size_t)NGmax * 32;
    // global scratch: private size / id replicas, the two box arrays, the flags
    const size_t N32 = (size_t)NG * 32;
    int* rsize = reinterpret_cast<int*>(w.fast_scratch) + (size_t)cta * 2 * N32;
    int* rcid = rsize + N32;
    WideBox* boxes = reinterpret_cast<WideBox*>(reinterpret_cast<char*>(w.fast_scratch) + (size_t)G * N32 * 8);
    int* flags = reinterpret_cast<int*>(boxes + 2 * (size_t)G);

    auto row_of = [&](int lg) { return ((lg * G + cta) << 5) + lane; };
    auto owner_cta = [&](int z) { return (z >> 5) % G; };
    auto owner_lg = [&](int z) { return (z >> 5) / G; };
    auto mine = [&](int z) { return z >= 0 && owner_cta(z) == cta && owner_lg(z) % LW_NW == warp && (z & 31) == lane; };

    for (int lg = warp; lg < NGl; lg += LW_NW) {
        const int z = row_of(lg), s = lg * 32 + lane;
        nbr[s] = z < n - 1 ? w.nbr[z] : -1;
        lb[s] = z < n - 1 ? w.lb[z] : INFINITY;
        cur[s] = lb[s];
        alive[s] = z < n ? 1 : 0;
    }
    for (int i = tid; i < (int)N32; i += LW_T) {
        rsize[i] = i < n ? 1 : 0;
        rcid[i] = i;
    }
    __syncthreads();

    // smallest bound among this CTA's live heap rows, `skip` left out; result in every thread of warp 0 ... and,
    // through shared memory, published by thread 0
    auto local_top = [&](int skip) -> Top {
        Top m;
        m.v = INFINITY;
        m.i = -1;
        m.c = 0;
        for (int lg = warp; lg < NGl; lg += LW_NW) {
            const int z = row_of(lg), s = lg * 32 + lane;
            if (z < n - 1 && z != skip && alive[s] != 0) {
                const double v = lb[s];
                if (m.c == 0 || v < m.v) {
                    m.v = v;
                    m.i = z;
                    m.c = 1;
                } else if (v == m.v)
                    ++m.c;  // rows ascend within a lane: the stored row stays the lowest
            }
        }
        return warp_top(m.c ? m.v : INFINITY, m.i, m.c);
    };
    // combine the per-warp results left in shared memory (every thread, redundantly)
    auto combine_tops = [&]() -> Top {
        Top t;
        t.v = INFINITY;
        t.i = -1;
        t.c = 0;
#pragma unroll
        for (int q = 0; q < LW_NW; ++q) {
            const int c = s_tc[q];
            if (c == 0) continue;
            const double v = s_tv[q];
            const int i = s_ti[q];
            if (t.c == 0 || v < t.v) {
                t.v = v;
                t.i = i;
                t.c = c;
            } else if (v == t.v) {
                t.c += c;
                t.i = i < t.i ? i : t.i;
            }
        }
        return t;
    };
    auto combine_parts = [&](double& pv, int& pi) {
        pv = INFINITY;
        pi = -1;
#pragma unroll
        for (int q = 0; q < LW_NW; ++q) {
            const int i = s_pi[q];
            if (i < 0) continue;
            const double v = s_pv[q];
            if (pi < 0 || v < pv || (v == pv && i < pi)) {
                pv = v;
                pi = i;
            }
        }
    };
    // publish this CTA's box of round `round` (thread 0, after a block barrier that follows every thread's fence)
    auto publish = [&](int round, const Top& t, double pv, int pi, int pend) {
        WideBox* b = boxes + (size_t)(round & 1) * G + cta;
        double t_cur = 0.0;
        int t_nbr = -1;
        if (t.i >= 0) {
            const int s = owner_lg(t.i) * 32 + (t.i & 31);
            t_cur = cur[s];
            t_nbr = nbr[s];
        }
        double q_lb = 0.0, q_cur = 0.0;
        int q_nbr = -1;
        if (pend >= 0 && owner_cta(pend) == cta) {
            const int s = owner_lg(pend) * 32 + (pend & 31);
            q_lb = lb[s];
            q_cur = cur[s];
            q_nbr = nbr[s];
        }
        b->A = lw::packd(t.v, t.i, t.c);
        b->B = lw::packd(t_cur, t_nbr, 0);
        b->C = lw::packd(pv, pi, q_nbr);
        uint4 d;
        d.x = (unsigned)__double2loint(q_lb);
        d.y = (unsigned)__double2hiint(q_lb);
        d.z = (unsigned)__double2loint(q_cur);
        d.w = (unsigned)__double2hiint(q_cur);
        b->D = d;
        __threadfence();  // the box is visible before any of the flags that the other threads raise after the barrier
    };
    // every reader polls its own copy of the flags (flags[reader][writer]): one hot line per reader instead of five
    // lines polled by all G * G threads of the grid
    auto raise_flags = [&](int round) {
        if (tid < G) lw::st_release(flags + (size_t)tid * G + cta, round);
    };

    int round = 1;
    int pending = -1;       // row whose bound the previous request recomputed
    bool pend_merge = false;  // it was a merge (the survivor keeps its old state when no partial exists)
    {
        const Top wt = local_top(-1);
        if (lane == 0) {
            s_tv[warp] = wt.v;
            s_ti[warp] = wt.i;
            s_tc[warp] = wt.c;
        }
        __syncthreads();
        if (tid == 0) publish(round, combine_tops(), INFINITY, -1, -1);
        __syncthreads();
        raise_flags(round);
    }
    int k = 0, tries = 0;
    unsigned long long rescans = 0;
    for (;;) {
        // ---------------- exchange: wait for every CTA's box of this round, reduce ----------------
        Top bt;
        bt.v = INFINITY;
        bt.i = -1;
        bt.c = 0;
        double bpv = INFINITY;
        int bpi = -1;
        uint4 bB = make_uint4(0, 0, 0, 0), bC = bB, bD = bB;
        if (tid < G) {
            while (lw::ld_acquire(flags + (size_t)cta * G + tid) < round) {
            }
            const WideBox* b = boxes + (size_t)(round & 1) * G + tid;
            const uint4 a = lw::ld_cg16(&b->A);
            bB = lw::ld_cg16(&b->B);
            bC = lw::ld_cg16(&b->C);
            bD = lw::ld_cg16(&b->D);
            bt.v = lw::lod(a);
            bt.i = (int)a.z;
            bt.c = (int)a.w;
            if (bt.c == 0) {
                bt.v = INFINITY;
                bt.i = -1;
            }
            bpi = (int)bC.z;
            bpv = bpi >= 0 ? lw::lod(bC) : INFINITY;
        }
        {
            const Top wt = warp_top(bt.c ? bt.v : INFINITY, bt.i, bt.c);
            const Top wp = warp_top(bpi >= 0 ? bpv : INFINITY, bpi, bpi >= 0 ? 1 : 0);
            if (lane == 0) {
                s_tv[warp] = wt.v;
                s_ti[warp] = wt.i;
                s_tc[warp] = wt.c;
                s_pv[warp] = wp.v;
                s_pi[warp] = wp.i;
            }
        }
        if (pending >= 0 && tid == owner_cta(pending)) {  // the owner's box carries the pending row's old state
            s_plb = lw::lod(bD);
            s_pcur = lw::hid(bD);
            s_pnbr = (int)bC.w;
        }
        __syncthreads();  // S1
        Top top = combine_tops();
        double pv;
        int pi;
        combine_parts(pv, pi);
        // new state of the pending row (clustering.cpp:395-404 after a merge, 259-276 after a rescan)
        double p_lb = INFINITY, p_cur = INFINITY;
        int p_nbr = -1;
        bool p_cand = false;
        if (pending >= 0) {
            if (pi >= 0) {
                p_lb = pv;
                p_cur = pv;
                p_nbr = pi;
            } else if (pend_merge) {
                p_lb = s_plb;
                p_cur = s_pcur;
                p_nbr = s_pnbr;
            }
            if (mine(pending)) {
                const int s = owner_lg(pending) * 32 + lane;
                lb[s] = p_lb;
                cur[s] = p_cur;
                nbr[s] = p_nbr;
            }
            p_cand = pending < n - 1;
            if (p_cand) {  // it was left out of its owner's published minimum
                if (top.c == 0 || p_lb < top.v) {
                    top.v = p_lb;
                    top.i = pending;
                    top.c = 1;
                } else if (p_lb == top.v) {
                    top.c += 1;
                    top.i = pending < top.i ? pending : top.i;
                }
            }
        }
        if (k >= n - 1) break;
        const int x = top.i;
        const double dist = top.v;
        if (top.c != 1 || x < 0 || tries >= n - k) {  // tied minimum: the heap order would matter
            if (scribe) need_exact[0] = 1;
            break;
        }
        // cached candidate of row x: from the pending state or from the box that published x
        if (x != pending && tid < G && bt.c > 0 && bt.i == x) {
            s_xcur = lw::lod(bB);
            s_xnbr = (int)bB.z;
        }
        __syncthreads();  // S2 (also: every thread has consumed s_tv.. before the request reuses them)
        const double x_cur = x == pending ? p_cur : s_xcur;
        const int y = x == pending ? p_nbr : s_xnbr;
        const bool valid = y >= 0 && dist == x_cur;  // clustering.cpp:329
        ++round;
        double part_v = INFINITY;
        int part_i = -1;
        if (!valid) {
            // ---------------- revalidate: find_min_dist(x) over live i > x ----------------
            const double* r = w.D + (size_t)x * w.ld;
            for (int lg = warp; lg < NGl; lg += LW_NW) {
                const int i = row_of(lg);
                if (i > x && i < n && alive[lg * 32 + lane] != 0) {
                    const double d = __ldcg(r + i);
                    if (d < part_v) {  // rows ascend within a lane: first minimum in index order
                        part_v = d;
                        part_i = i;
                    }
                }
            }
            pending = x;
            pend_merge = false;
            ++tries;
            ++rescans;
        } else {
            // ---------------- merge x into y (clustering.cpp:347-404) ----------------
            const int nx = rsize[x], ny = rsize[y], ix = rcid[x], iy = rcid[y];
            if (scribe) {
                double* z = w.Z + 4 * (size_t)k;
                z[0] = ix < iy ? ix : iy;
                z[1] = ix < iy ? iy : ix;
                z[2] = dist;
                z[3] = nx + ny;
            }
            const double* rowx = w.D + (size_t)x * w.ld;
            double* rowy = w.D + (size_t)y * w.ld;
            const double fx = (double)nx, fy = (double)ny, fs = (double)(nx + ny);
            const double t3 = __ddiv_rn(__dmul_rn(__dmul_rn((double)(nx * ny), dist), dist), fs);
            const double rs = __drcp_rn(fs);
            for (int lg = warp; lg < NGl; lg += LW_NW) {
                const int z = row_of(lg), s = lg * 32 + lane;
                if (z == x) alive[s] = 0;
                if (z < n && z != x && z != y && alive[s] != 0) {
                    const double dx = __ldcg(rowx + z), dy = __ldcg(rowy + z);
                    const double t1q = __dmul_rn(__dmul_rn(fx, dx), dx);
                    const double t2q = __dmul_rn(__dmul_rn(fy, dy), dy);
                    const double nd = __dsqrt_rn(div_by(__dsub_rn(__dadd_rn(t1q, t2q), t3), fs, rs));
                    __stcg(rowy + z, nd);
                    __stcg(w.D + (size_t)z * w.ld + y, nd);
                    if (z < y) {
                        int nb = nbr[s];
                        const double lbz = lb[s];
                        if (z < x && nb == x) nb = y;  // clustering.cpp:374-378
                        if (nd < lbz) {                // clustering.cpp:381-392
                            nb = y;
                            lb[s] = nd;
                        }
                        if (nb == y) cur[s] = nd;  // cur mirrors D[z][nbr[z]]
                        nbr[s] = nb;
                    } else if (nd < part_v) {  // rows ascend within a lane: first strict minimum
                        part_v = nd;
                        part_i = z;
                    }
                }
            }
            if (tid == 0) {  // private replicas follow the merge
                rsize[x] = 0;
                rsize[y] = nx + ny;
                rcid[y] = n + k;
            }
            pending = y;
            pend_merge = true;
            ++k;
            tries = 0;
            __threadfence();  // this thread's matrix stores are visible device-wide before the CTA's flag goes up
        }
        // ---------------- publish: own minimum without the pending row, own share of its new neighbour ----------
        {
            const Top wt = local_top(pending);
            const Top wp = warp_top(part_i >= 0 ? part_v : INFINITY, part_i, part_i >= 0 ? 1 : 0);
            if (lane == 0) {
                s_tv[warp] = wt.v;
                s_ti[warp] = wt.i;
                s_tc[warp] = wt.c;
                s_pv[warp] = wp.v;
                s_pi[warp] = wp.i;
            }
        }
        __syncthreads();  // S3
        if (tid == 0) {
            double pv2;
            int pi2;
            combine_parts(pv2, pi2);
            publish(round, combine_tops(), pv2, pi2, pending);
        }
        __syncthreads();  // S4: the box is out and fenced; shared scratch is free again
        raise_flags(round);
    }
    if (scribe && w.stats) {
        w.stats[0] += rescans;
        w.stats[7] += (unsigned long long)(n - 1);
    }
}

// ------------------------------------------------------------------------------------------------
// fcluster (criterion "distance")
// ------------------------------------------------------------------------------------------------

// Sequential form (single thread), kept as the exact fallback for dendrograms that contain NaN merge distances
// (the reference's max-propagation ignores NaN children, which the scan formulation below does not model).
// T[n] receives labels 1..K in the reference's depth-first numbering; *num_out = K.
__global__ void fcluster_seq_kernel(const double* __restrict__ Z, int n, double cutoff, double* __restrict__ MD,
                                    int* __restrict__ stack, unsigned char* __restrict__ stage, int* __restrict__ T,
                                    int* __restrict__ num_out, const int* __restrict__ run_flag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (run_flag && !*run_flag) return;
    if (n < 2) {
        if (n == 1) T[0] = 1;
        *num_out = n;
        return;
    }
    // max merge distance inside each subtree (clustering.cpp:121-172); children precede parents in Z
    for (int k = 0; k < n - 1; ++k) {
        const int lc = (int)Z[4 * (size_t)k], rc = (int)Z[4 * (size_t)k + 1];
        double m = Z[4 * (size_t)k + 2];
        if (lc >= n) {
            const double v = MD[lc - n];
            if (v > m) m = v;
        }
        if (rc >= n) {
            const double v = MD[rc - n];
            if (v > m) m = v;
        }
        MD[k] = m;
        stage[k] = 0;
    }
    // depth-first flat-cluster numbering (clustering.cpp:174-232)
    int sp = 0, ncl = 0, leader = -1;
    stack[0] = n - 2;
    while (sp >= 0) {
        const int r = stack[sp];
        const int lc = (int)Z[4 * (size_t)r], rc = (int)Z[4 * (size_t)r + 1];
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
    *num_out = ncl;
}

// in-place inclusive scan of a[0..m) by one CTA (a in shared or global memory)
__device__ void block_inclusive_scan(int* a, int m, int* warp_tot) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int per = (m + nt - 1) / nt;
    const int lo = tid * per, hi = min(lo + per, m);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += a[i];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int v = lane < nw ? warp_tot[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        warp_tot[lane] = v;  // inclusive totals of warps
    }
    __syncthreads();
    int run = inc - sum + (warp > 0 ? warp_tot[warp - 1] : 0);
    for (int i = lo; i < hi; ++i) {
        run += a[i];
        a[i] = run;
    }
    __syncthreads();
}

// Parallel fcluster (criterion "distance"), one CTA.
//
// The reference assigns leaf labels in a fixed depth-first order (clustering.cpp:174-232): at a node, first
// the leaves of an internal left child, then those of an internal right child, then a leaf left child, then a
// leaf right child.  In that order every subtree occupies one contiguous interval of leaf positions whose
// length is the cluster size stored in Z, so
//   start[node] = sum of per-edge offsets on the path to the root          (pointer jumping, log depth rounds)
//   every internal node owns exactly one gap between two consecutive positions (where its two blocks meet)
//   MD[node] <= cutoff  <=>  no gap inside its interval belongs to a merge with distance > cutoff (prefix sums)
//   two consecutive leaves share a flat cluster  <=>  the owner of the gap between them satisfies that
// and the flat-cluster number of a leaf is the number of cluster starts at or before its position, which is the
// order in which the reference's traversal discovers clusters.  NaN distances set *nan_flag and the caller
// runs the sequential kernel instead.
struct FcWork {
    int* up;     // [2n] pointer-jumping parent
    int* acc;    // [2n] accumulated offset
    int* up2;    // [2n] second buffer
    int* acc2;   // [2n]
    int* gapbad; // [n]  1 when the merge owning gap g has distance > cutoff; then inclusive prefix sums
    int* owner;  // [n]  internal node owning gap g
    int* newc;   // [n]  cluster starts, then inclusive prefix sums = labels by position
};

__global__ void __launch_bounds__(1024)
    fcluster_par_kernel(const double* __restrict__ Z, int n, double cutoff, FcWork w, int use_smem,
                        int* __restrict__ T, int* __restrict__ num_out, int* __restrict__ nan_flag) {
    extern __shared__ __align__(16) unsigned char fc_smem[];
    __shared__ int warp_tot[32];
    __shared__ int s_nan, s_live;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (n < 2) {
        if (tid == 0) {
            if (n == 1) T[0] = 1;
            *num_out = n;
            *nan_flag = 0;
        }
        return;
    }
    const int nodes = 2 * n - 1;
    if (use_smem) {
        int* p = reinterpret_cast<int*>(fc_smem);
        w.up = p;
        p += 2 * n;
        w.acc = p;
        p += 2 * n;
        w.up2 = p;
        p += 2 * n;
        w.acc2 = p;
        p += 2 * n;
        w.gapbad = p;
        p += n;
        w.owner = p;
        p += n;
        w.newc = p;
    }
    if (tid == 0) s_nan = 0;
    for (int g = tid; g < n; g += nt) {
        w.gapbad[g] = 0;
        w.owner[g] = -1;
    }
    if (tid == 0) {
        w.up[nodes - 1] = -1;  // root
        w.acc[nodes - 1] = 0;
    }
    __syncthreads();
    // per-edge offsets inside the parent's block
    for (int k = tid; k < n - 1; k += nt) {
        const int lc = (int)Z[4 * (size_t)k], rc = (int)Z[4 * (size_t)k + 1];
        const double d = Z[4 * (size_t)k + 2];
        if (isnan(d)) s_nan = 1;
        const int szl = lc >= n ? (int)Z[4 * (size_t)(lc - n) + 3] : 1;
        const int szr = rc >= n ? (int)Z[4 * (size_t)(rc - n) + 3] : 1;
        int offl, offr;
        if (lc >= n) {  // internal left child goes first
            offl = 0;
            offr = szl;
        } else if (rc >= n) {  // leaf left child follows the internal right subtree
            offr = 0;
            offl = szr;
        } else {
            offl = 0;
            offr = 1;
        }
        w.up[lc] = n + k;
        w.acc[lc] = offl;
        w.up[rc] = n + k;
        w.acc[rc] = offr;
    }
    __syncthreads();
    if (s_nan) {
        if (tid == 0) *nan_flag = 1;
        return;
    }
    // start[node] = sum of offsets up to the root: pointer jumping between two buffers
    int *upA = w.up, *accA = w.acc, *upB = w.up2, *accB = w.acc2;
    for (;;) {
        __syncthreads();  // every thread has read the verdict of the previous round before it is reset (without this
                          // barrier a fast thread 0 could clear s_live while slower threads had not yet tested it:
                          // they left the loop one round early and the labels came out wrong -- seen only when other
                          // kernels shared the SM, i.e. with many contexts in flight)
        if (tid == 0) s_live = 0;
        __syncthreads();
        int live = 0;
        for (int v = tid; v < nodes; v += nt) {
            const int u = upA[v];
            int nu = -1, na = accA[v];
            if (u >= 0) {
                nu = upA[u];
                na += accA[u];
            }
            upB[v] = nu;
            accB[v] = na;
            live |= nu >= 0;
        }
        if (live) s_live = 1;
        __syncthreads();
        const int again = s_live;
        int* t = upA;
        upA = upB;
        upB = t;
        t = accA;
        accA = accB;
        accB = t;
        if (!again) break;
    }
    w.acc = accA;  // start[] of every node
    // gaps: node k splits [start, start + size) at start + size(first block)
    for (int k = tid; k < n - 1; k += nt) {
        const int lc = (int)Z[4 * (size_t)k], rc = (int)Z[4 * (size_t)k + 1];
        const double d = Z[4 * (size_t)k + 2];
        const int szl = lc >= n ? (int)Z[4 * (size_t)(lc - n) + 3] : 1;
        const int szr = rc >= n ? (int)Z[4 * (size_t)(rc - n) + 3] : 1;
        const int first = (lc >= n) ? szl : ((rc >= n) ? szr : 1);
        const int g = w.acc[n + k] + first;
        w.owner[g] = k;
        w.gapbad[g] = !(d <= cutoff);
    }
    __syncthreads();
    block_inclusive_scan(w.gapbad, n, warp_tot);
    // cluster starts by position
    for (int p = tid; p < n; p += nt) {
        int nc = 1;
        if (p > 0) {
            const int k = w.owner[p];
            const int st = w.acc[n + k];
            const int sz = (int)Z[4 * (size_t)k + 3];
            nc = (w.gapbad[st + sz - 1] - w.gapbad[st]) > 0;  // bad gaps strictly inside the interval
        }
        w.newc[p] = nc;
    }
    __syncthreads();
    block_inclusive_scan(w.newc, n, warp_tot);
    for (int leaf = tid; leaf < n; leaf += nt) T[leaf] = w.newc[w.acc[leaf]];
    if (tid == 0) {
        *num_out = w.newc[n - 1];
        *nan_flag = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Cluster::cluster post-processing and Cluster::assign_embeddings
// ------------------------------------------------------------------------------------------------

// Helper::cosineDistance (speakerDiarizer.cpp:476-498).  Returns false on zero magnitude.
__device__ __forceinline__ bool cosine_distance(const double* a, const double* b, int D, double* out) {
    double dot = 0.0, ma = 0.0, mb = 0.0;
    for (int k = 0; k < D; ++k) {
        const double x = a[k], y = b[k];
        dot = __dadd_rn(dot, __dmul_rn(x, y));
        ma = __dadd_rn(ma, __dmul_rn(x, x));
        mb = __dadd_rn(mb, __dmul_rn(y, y));
    }
    if (ma == 0.0 || mb == 0.0) return false;
    *out = __dsub_rn(1.0, __ddiv_rn(dot, __dmul_rn(__dsqrt_rn(ma), __dsqrt_rn(mb))));
    return true;
}

__global__ void __launch_bounds__(256)
    cosine_cdist_kernel(const double* __restrict__ a, int na, const double* __restrict__ b, int nb, int D,
                        double* __restrict__ out, int* __restrict__ status) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)na * nb) return;
    const int i = (int)(idx / nb), j = (int)(idx - (long)i * nb);
    double d;
    if (cosine_distance(a + (size_t)i * D, b + (size_t)j * D, D, &d))
        out[idx] = d;
    else
        atomicExch(status, SD_ERR_ZERO_MAGNITUDE);
}

struct PostWork {
    const double* x;   // [N][D] un-normalised filtered embeddings
    int* labels;       // [N] in: fcluster labels 1..K ; out: final labels 0..K'-1
    int* count;        // [K+1]
    int* map;          // [K+1]
    int* rank;         // [K+1]
    double* sums;      // [K][D]
    int* num_clusters; // out: max label + 1
    int* need_means;   // out of phase 1: 1 when small clusters must be re-assigned
    int* status;
};

// Cluster means in index order (calculateClusterMeans, speakerDiarizer.cpp:443-473; assign_embeddings
// 2147-2167).  One CTA per cluster, one thread per dimension: the member rows of the cluster are compacted
// in increasing row order, then added one by one into a register -- the same sequence of fp64 additions as the
// reference -- with the loads issued eight at a time so their latency is off the dependent add chain.
constexpr int CM_TILE = 1024;
__global__ void __launch_bounds__(512)
    cluster_means_kernel(const double* __restrict__ x, const int* __restrict__ labels, int N, int D, int K,
                         const int* __restrict__ enable, double* __restrict__ means, const int* __restrict__ d_k) {
    __shared__ int list[CM_TILE];
    __shared__ int wcount[16];
    __shared__ int s_total;
    if (enable && !*enable) return;
    const int k = blockIdx.x;
    if (d_k && k >= *d_k) return;  // asynchronous path: the grid covers the upper bound of the cluster count
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    double acc[2] = {0.0, 0.0};  // dimensions tid and tid + blockDim.x (D <= 1024)
    long members = 0;
    for (int i0 = 0; i0 < N; i0 += CM_TILE) {
        // ordered compaction of the rows of this tile that belong to cluster k
        int base = 0;
        for (int r0 = 0; r0 < CM_TILE; r0 += blockDim.x) {
            const int i = i0 + r0 + tid;
            const bool mine = (r0 + tid < CM_TILE) && i < N && labels[i] == k;
            const unsigned bal = __ballot_sync(0xffffffffu, mine);
            if (lane == 0) wcount[warp] = __popc(bal);
            __syncthreads();
            int off = base;
            for (int w2 = 0; w2 < warp; ++w2) off += wcount[w2];
            if (mine) list[off + __popc(bal & ((1u << lane) - 1u))] = i;
            int tot = 0;
            for (int w2 = 0; w2 < nwarp; ++w2) tot += wcount[w2];
            base += tot;
            __syncthreads();
        }
        const int cnt = base;
        members += cnt;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = tid + h * blockDim.x;
            if (j < D) {
                int m = 0;
                for (; m + 8 <= cnt; m += 8) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = x[(size_t)list[m + u] * D + j];
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc[h] = __dadd_rn(acc[h], v[u]);
                }
                for (; m < cnt; ++m) acc[h] = __dadd_rn(acc[h], x[(size_t)list[m] * D + j]);
            }
        }
        __syncthreads();
    }
    if (tid == 0) s_total = (int)members;
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = tid + h * blockDim.x;
        if (j < D) means[(size_t)k * D + j] = __ddiv_rn(acc[h], (double)s_total);  // 0/0 -> NaN like the reference
    }
}

// phase 1 of Cluster::cluster's post-processing: labels - 1, counts, large/small split
__global__ void __launch_bounds__(1024)
    cluster_post1_kernel(PostWork w, int N, int K, int min_cluster_size, const int* __restrict__ d_k, int k_cap) {
    __shared__ int s_nl, s_ns;
    const int tid = threadIdx.x;
    if (d_k) {  // asynchronous path: the cluster count of fcluster is only known on the device
        K = *d_k;
        __syncthreads();  // everybody has read *d_k before thread 0 may overwrite it (num_clusters aliases it)
        if (K < 1 || K > k_cap) {
            // labels run up to K-1 but count/map/rank/sums are sized for k_cap: latch the error and leave a safe
            // one-cluster result behind instead of indexing past the workspace (the later kernels see K = 1)
            for (int i = tid; i < N; i += blockDim.x) w.labels[i] = 0;
            if (tid == 0) {
                atomicExch(w.status, SD_ERR_CAPACITY);
                *w.num_clusters = 1;
                *w.need_means = 0;
            }
            return;
        }
    }
    for (int l = tid; l <= K; l += blockDim.x) {
        w.count[l] = 0;
        w.map[l] = l;
    }
    if (tid == 0) {
        s_nl = 0;
        s_ns = 0;
    }
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) {  // speakerDiarizer.cpp:2324-2341
        const int l = w.labels[i] - 1;
        w.labels[i] = l;
        atomicAdd(&w.count[l], 1);
    }
    __syncthreads();
    long mcs = (long)round(0.1 * (double)N);  // speakerDiarizer.cpp:2308-2309
    if (mcs < 1) mcs = 1;
    if (mcs > min_cluster_size) mcs = min_cluster_size;
    for (int l = tid; l < K; l += blockDim.x) {
        if (w.count[l] >= mcs)
            atomicAdd(&s_nl, 1);
        else if (w.count[l] > 0)
            atomicAdd(&s_ns, 1);
    }
    __syncthreads();
    const int nl = s_nl, ns = s_ns;
    if (nl == 0) {  // speakerDiarizer.cpp:2371-2375
        for (int i = tid; i < N; i += blockDim.x) w.labels[i] = 0;
        if (tid == 0) {
            *w.num_clusters = 1;
            *w.need_means = 0;
        }
        return;
    }
    if (tid == 0) {
        *w.num_clusters = K;  // speakerDiarizer.cpp:2377-2380 when there is no small cluster
        *w.need_means = ns > 0;
    }
}

// phase 2: each small cluster -> nearest large cluster by centroid cosine distance with a float running
// minimum (speakerDiarizer.cpp:2390-2415; large clusters visited in ascending label order), then the rank of
// each surviving label among the sorted unique labels (speakerDiarizer.cpp:519-548)
__global__ void __launch_bounds__(1024)
    cluster_post2_kernel(PostWork w, int N, int D, int K, int min_cluster_size, const int* __restrict__ d_k) {
    __shared__ int s_next;
    if (!*w.need_means) return;
    const int tid = threadIdx.x;
    if (d_k) {
        K = *d_k;
        __syncthreads();
    }
    long mcs = (long)round(0.1 * (double)N);
    if (mcs < 1) mcs = 1;
    if (mcs > min_cluster_size) mcs = min_cluster_size;
    for (int b = tid; b < K; b += blockDim.x) {
        if (w.count[b] == 0 || w.count[b] >= mcs) continue;
        float best = FLT_MAX;
        int arg = -1;
        for (int a = 0; a < K; ++a) {
            if (w.count[a] < mcs) continue;
            double d;
            if (!cosine_distance(w.sums + (size_t)a * D, w.sums + (size_t)b * D, D, &d)) {
                atomicExch(w.status, SD_ERR_ZERO_MAGNITUDE);
                continue;
            }
            if (d < best) {
                best = (float)d;
                arg = a;
            }
        }
        if (arg >= 0) w.map[b] = arg;
    }
    __syncthreads();
    if (tid == 0) {
        int next = 0;
        for (int l = 0; l < K; ++l) w.rank[l] = (w.count[l] >= mcs) ? next++ : -1;
        s_next = next;
    }
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) w.labels[i] = w.rank[w.map[w.labels[i]]];
    if (tid == 0) *w.num_clusters = s_next;
}

struct AssignWork {
    const double* emb;   // [R][D] all embeddings (NaN rows included)
    const double* x;     // [N][D] filtered, un-normalised
    const int* labels;   // [N]
    double* cent;        // [K][D]
    double* soft;        // optional [R][soft_k_cap]
    double* dist;        // optional [R][soft_k_cap]: the cosine distances themselves (cpp_dist stage dump)
    double* dist_k;      // scratch [R][K] behind `dist`
    int* hard;           // [R]
    const double* binarized;  // optional [C][F][S]
    int* status;
};

// one thread per (embedding row, centroid): cosine distance in the reference's sequential order
// (speakerDiarizer.cpp:2180-2203); soft = 2 - d
__global__ void __launch_bounds__(256) assign_dist_kernel(AssignWork w, int R, int D, int K, double* __restrict__ soft,
                                                         int ld_soft, const int* __restrict__ d_k) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)R * K) return;  // K is the grid's (upper-bound) cluster count
    const int r = (int)(idx / K), k = (int)(idx - (long)r * K);
    if (d_k && k >= *d_k) return;
    double d;
    if (!cosine_distance(w.emb + (size_t)r * D, w.cent + (size_t)k * D, D, &d)) {
        atomicExch(w.status, SD_ERR_ZERO_MAGNITUDE);
        d = NAN;
    }
    soft[(size_t)r * ld_soft + k] = __dsub_rn(2.0, d);
    if (w.dist_k) w.dist_k[(size_t)r * ld_soft + k] = d;
}

// one warp per embedding row: first strict maximum over the centroids (Helper::argmax, speakerDiarizer.cpp:
// 293-316; an all-NaN row gives 0), then the inactive-speaker mask (3166-3191): the reference sums the 0/1
// activity of the speaker over the chunk's frames in float, which is exact in any order
__global__ void __launch_bounds__(256) assign_argmax_kernel(AssignWork w, int R, int S, int K, int F,
                                                           const double* __restrict__ soft, int ld_soft,
                                                           int user_cap, const int* __restrict__ d_k) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    if (d_k) K = *d_k;
    int arg = 0;
    if (lane == 0) {
        double best = -DBL_MAX;
        for (int k = 0; k < K; ++k) {
            const double v = soft[(size_t)r * ld_soft + k];
            if (w.soft && k < user_cap) w.soft[(size_t)r * user_cap + k] = v;
            if (w.dist && k < user_cap) w.dist[(size_t)r * user_cap + k] = w.dist_k[(size_t)r * ld_soft + k];
            if (v > best) {
                best = v;
                arg = k;
            }
        }
    }
    if (w.binarized) {
        const int c = r / S, s = r - c * S;
        float acc = 0.0f;
        for (int f = lane; f < F; f += 32) acc += (float)w.binarized[((size_t)c * F + f) * S + s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (fabsf(acc) < DBL_EPSILON) arg = -2;
    }
    if (lane == 0) w.hard[r] = arg;
}

__global__ void fill_int_kernel(int* p, long n, int v) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// inactive-speaker mask alone (used when clustering is skipped because fewer than two embeddings exist)
__global__ void inactive_mask_kernel(const double* __restrict__ binarized, int C, int F, int S, int* hard) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= C * S) return;
    const int c = r / S, s = r - c * S;
    float acc = 0.0f;
    for (int f = 0; f < F; ++f) acc += (float)binarized[((size_t)c * F + f) * S + s];
    if (fabsf(acc) < DBL_EPSILON) hard[r] = -2;
}

// ------------------------------------------------------------------------------------------------
// host-side orchestration (device pointers in, enqueue only unless noted)
// ------------------------------------------------------------------------------------------------

int upload_small(sd_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int means_threads(int D) {
    int t = (D + 31) / 32 * 32;
    if (t > 512) t = 512;
    if (t < 64) t = 64;
    return t;
}

struct LinkLayout {
    long ld;
    size_t off_D, off_size, off_cid, off_nbr, off_lb, off_cur, off_pos, off_key, off_hval, off_bitmap, off_work, off_n,
        off_fast, total;
};

static LinkLayout link_layout(int N) {
    LinkLayout L;
    L.ld = (long)align_up((size_t)N, 16);
    size_t o = 0;
    L.off_D = o;
    o += align_up((size_t)N * L.ld * sizeof(double), 256);
    L.off_size = o;
    o += align_up((size_t)N * sizeof(int), 256);
    L.off_cid = o;
    o += align_up((size_t)N * sizeof(int), 256);
    L.off_nbr = o;
    o += align_up((size_t)N * sizeof(int), 256);
    L.off_lb = o;
    o += align_up((size_t)N * sizeof(double), 256);
    L.off_cur = o;
    o += align_up((size_t)N * sizeof(double), 256);
    L.off_pos = o;
    o += align_up((size_t)N * sizeof(int), 256);
    L.off_key = o;
    o += align_up((size_t)N * sizeof(int), 256);
    L.off_hval = o;
    o += align_up((size_t)N * sizeof(double), 256);
    L.off_bitmap = o;
    o += align_up(((size_t)N / 32 + 2) * sizeof(unsigned), 256);
    L.off_work = o;
    o += align_up(sizeof(LinkWork), 256);
    L.off_n = o;
    o += 256;
    L.off_fast = o;
    o += align_up(std::max(std::max(linkfast_smem_bytes(N), linkcluster_global_bytes(N)),
                           linkwide_global_bytes(N, LW_MAX_CTAS)), 256);
    L.total = o;
    return L;
}

int pdist_tc_launch(sd_ctx* ctx, const double* d_xn, int N, int D, double* Dm, long ld, double refine_below);

// squared distance below which the tensor-core mode recomputes a pair exactly (d < 0.3)
constexpr double kTcRefineBelow = 0.09;

// pdist into the square matrix of the linkage workspace; returns the workspace base
static int pdist_square(sd_ctx* ctx, const double* d_xn, int N, int D, int mode, char** base_out, LinkLayout* L_out) {
    LinkLayout L = link_layout(N);
    char* base = (char*)ctx->scratch(BUF_CL_DIST, L.total);
    if (!base) return SD_ERR_NOMEM;
    if (mode == SD_PDIST_GEMM_TF32X3) {
        int rc = pdist_tc_launch(ctx, d_xn, N, D, reinterpret_cast<double*>(base + L.off_D), L.ld, kTcRefineBelow);
        if (rc) return rc;
    } else if (mode == SD_PDIST_EXACT_F64) {
        dim3 grid((N + PD_TILE - 1) / PD_TILE, (N + PD_TILE - 1) / PD_TILE);
        pdist_f64_kernel<<<grid, 256, 0, ctx->stream>>>(d_xn, N, D, reinterpret_cast<double*>(base + L.off_D), L.ld,
                                                        nullptr);
        SD_LAUNCH_CHECK(ctx);
    } else
        return ctx->fail(SD_ERR_UNSUPPORTED, "unknown pdist mode %d", mode);
    *base_out = base;
    *L_out = L;
    return SD_OK;
}

template <int MODE>
static int linkage_launch_mode(sd_ctx* ctx, const LinkWork* d_works, const int* d_ns, int problems, int max_n,
                               const int* d_run_flags) {
    const size_t smem = link_smem_bytes(MODE, max_n);
    if (kernel_setup(ctx, linkage_kernel<MODE>, (int)(MODE == LK_GLOBAL ? smem : (size_t)227 * 1024 - 1024)) < 0)
        return SD_ERR_CUDA;
    linkage_kernel<MODE><<<problems, LK_THREADS, smem, ctx->stream>>>(d_works, d_ns, d_run_flags);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

// picks where the merge state lives from the largest problem in the launch
static int linkage_dispatch(sd_ctx* ctx, const LinkWork* d_works, const int* d_ns, int problems, int max_n,
                            const int* d_run_flags) {
    const size_t budget = (size_t)227 * 1024 - 2048;  // static shared memory of the kernel comes on top
    if (link_smem_bytes(LK_SMEM_ALL, max_n) <= budget)
        return linkage_launch_mode<LK_SMEM_ALL>(ctx, d_works, d_ns, problems, max_n, d_run_flags);
    if (link_smem_bytes(LK_SMEM_HEAP, max_n) <= budget)
        return linkage_launch_mode<LK_SMEM_HEAP>(ctx, d_works, d_ns, problems, max_n, d_run_flags);
    return linkage_launch_mode<LK_GLOBAL>(ctx, d_works, d_ns, problems, max_n, d_run_flags);
}

template <int MODE, int T>
static int linkage_fast_launch_mode(sd_ctx* ctx, const LinkWork* d_works, const int* d_ns, int problems, int max_n,
                                    int* d_need_exact) {
    const size_t smem = MODE == LF_SMEM ? linkfast_smem_bytes(max_n) : 64;
    if (kernel_setup(ctx, linkage_fast_kernel<MODE, T>, 220 * 1024) < 0) return SD_ERR_CUDA;
    linkage_fast_kernel<MODE, T><<<problems, T, smem, ctx->stream>>>(d_works, d_ns, d_need_exact);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

template <int T, bool GREPL, bool PRE>
static int linkage_cluster_launch(sd_ctx* ctx, const LinkWork* d_works, const int* d_ns, int problems, int max_n,
                                  int* d_need_exact) {
    const size_t smem = linkcluster_smem_bytes(max_n, GREPL);
    if (kernel_setup(ctx, linkage_cluster_kernel<T, GREPL, PRE>, 200 * 1024) < 0) return SD_ERR_CUDA;
    linkage_cluster_kernel<T, GREPL, PRE><<<problems * kLcCtas, T, smem, ctx->stream>>>(d_works, d_ns, d_need_exact);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

// The whole-GPU merge loop: G = one CTA per SM, cooperative launch (all CTAs must be co-resident: they wait for each
// other's flags).  `fast_scratch` = base of the problem's scratch area (replicas, boxes, flags).
static int linkage_wide_launch(sd_ctx* ctx, const LinkWork* d_works, const int* d_ns, int max_n, int* d_need_exact,
                               char* fast_scratch) {
    const int G = std::min(ctx->num_sms, LW_MAX_CTAS);
    const size_t N32 = (size_t)(max_n + 31) / 32 * 32;
    const size_t ngmax = (N32 / 32 + G - 1) / G;
    const size_t smem = ngmax * 32 * (8 + 8 + 4 + 4) + 64;
    if (kernel_setup(ctx, linkage_wide_kernel, (int)std::max(smem, (size_t)48 * 1024)) < 0) return SD_ERR_CUDA;
    int* d_flags = reinterpret_cast<int*>(fast_scratch + (size_t)G * N32 * 8 + 2 * (size_t)G * sizeof(WideBox));
    SD_CUDA(ctx, cudaMemsetAsync(d_flags, 0, (size_t)G * G * sizeof(int), ctx->stream));
    void* args[] = {(void*)&d_works, (void*)&d_ns, (void*)&d_need_exact};
    SD_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)linkage_wide_kernel, dim3((unsigned)G), dim3(LW_T), args, smem,
                                             ctx->stream));
    ctx->launches++;
    return SD_OK;
}

static int linkage_fast_dispatch(sd_ctx* ctx, const LinkWork* d_works, const int* d_ns, int problems, int max_n,
                                 int* d_need_exact, char* fast_scratch) {
    // Opt-in (SD_OPT_LINKAGE_WIDE): measured 20.4 us per merge at N = 10 773 and 25.5 us at N = 50 000 against 11.4 /
    // 26.5 us for the cluster kernel -- a round through the global-memory mailbox costs ~7 us and a merge needs ~3 of
    // them, so it only breaks even at the largest size
    if (problems == 1 && fast_scratch && (ctx->linkage_wide == 2 || (ctx->linkage_wide == 1 && max_n >= 32768)))
        return linkage_wide_launch(ctx, d_works, d_ns, max_n, d_need_exact, fast_scratch);
    constexpr size_t kClusterSmem = (size_t)180 * 1024;
    if (ctx->linkage_cluster && linkcluster_smem_bytes(max_n, false) <= kClusterSmem) {
        // SD_OPT_LINKAGE_THREADS: 128 / 256 = one exchange entry per warp, 512 = pre-reduced to one entry per CTA,
        // 1024 = 512 threads with one entry per warp (the round-1 form, kept for comparison)
        const int t = ctx->linkage_threads ? ctx->linkage_threads : (max_n <= 4096 ? 128 : 512);
        if (t <= 128) return linkage_cluster_launch<128, false, false>(ctx, d_works, d_ns, problems, max_n, d_need_exact);
        if (t <= 256) return linkage_cluster_launch<256, false, false>(ctx, d_works, d_ns, problems, max_n, d_need_exact);
        if (t <= 512) return linkage_cluster_launch<512, false, true>(ctx, d_works, d_ns, problems, max_n, d_need_exact);
        return linkage_cluster_launch<512, false, false>(ctx, d_works, d_ns, problems, max_n, d_need_exact);
    }
    if (ctx->linkage_cluster && linkcluster_smem_bytes(max_n, true) <= kClusterSmem)  // replicas in global memory
        return ctx->linkage_threads == 1024
                   ? linkage_cluster_launch<512, true, false>(ctx, d_works, d_ns, problems, max_n, d_need_exact)
                   : linkage_cluster_launch<512, true, true>(ctx, d_works, d_ns, problems, max_n, d_need_exact);
    const bool fits = linkfast_smem_bytes(max_n) <= (size_t)220 * 1024;
    const int threads = ctx->linkage_threads ? ctx->linkage_threads : (max_n <= 4096 ? 512 : 1024);
    if (fits && threads == 512) return linkage_fast_launch_mode<LF_SMEM, 512>(ctx, d_works, d_ns, problems, max_n, d_need_exact);
    if (fits) return linkage_fast_launch_mode<LF_SMEM, 1024>(ctx, d_works, d_ns, problems, max_n, d_need_exact);
    return linkage_fast_launch_mode<LF_GLOBAL, 1024>(ctx, d_works, d_ns, problems, max_n, d_need_exact);
}

// d_x: the rows the distance matrix was built from (needed to rebuild it if the exact kernel must run)
static int linkage_on_square(sd_ctx* ctx, char* base, const LinkLayout& L, const double* d_x, int N, int D,
                             double* d_Z) {
    LinkWork w;
    w.D = reinterpret_cast<double*>(base + L.off_D);
    w.ld = L.ld;
    w.size = reinterpret_cast<int*>(base + L.off_size);
    w.cid = reinterpret_cast<int*>(base + L.off_cid);
    w.nbr = reinterpret_cast<int*>(base + L.off_nbr);
    w.lb = reinterpret_cast<double*>(base + L.off_lb);
    w.cur = reinterpret_cast<double*>(base + L.off_cur);
    w.pos_of = reinterpret_cast<int*>(base + L.off_pos);
    w.key_at = reinterpret_cast<int*>(base + L.off_key);
    w.hval = reinterpret_cast<double*>(base + L.off_hval);
    w.bitmap = reinterpret_cast<unsigned*>(base + L.off_bitmap);
    w.Z = d_Z;
    w.status = ctx->d_status;
    w.stats = ctx->d_stats;
    w.fast_scratch = base + L.off_fast;
    int rc = upload_small(ctx, base + L.off_work, &w, sizeof(w));
    if (rc) return rc;
    rc = upload_small(ctx, base + L.off_n, &N, sizeof(int));
    if (rc) return rc;
    const LinkWork* d_work = reinterpret_cast<const LinkWork*>(base + L.off_work);
    const int* d_n = reinterpret_cast<const int*>(base + L.off_n);
    int* d_need_exact = reinterpret_cast<int*>(base + L.off_n) + 8;
    const unsigned rm_grid = (unsigned)(((long)N * 32 + 255) / 256);
    rowmin_init_kernel<<<rm_grid, 256, 0, ctx->stream>>>(w, N, nullptr);
    SD_LAUNCH_CHECK(ctx);
    if (!ctx->force_exact_linkage) {
        rc = linkage_fast_dispatch(ctx, d_work, d_n, 1, N, d_need_exact, base + L.off_fast);
        if (rc) return rc;
        // exact chain, skipped on the device unless the fast kernel met a tied minimum
        dim3 grid((N + PD_TILE - 1) / PD_TILE, (N + PD_TILE - 1) / PD_TILE);
        pdist_f64_kernel<<<grid, 256, 0, ctx->stream>>>(d_x, N, D, w.D, L.ld, d_need_exact);
        SD_LAUNCH_CHECK(ctx);
        rowmin_init_kernel<<<rm_grid, 256, 0, ctx->stream>>>(w, N, d_need_exact);
        SD_LAUNCH_CHECK(ctx);
        return linkage_dispatch(ctx, d_work, d_n, 1, N, d_need_exact);
    }
    return linkage_dispatch(ctx, d_work, d_n, 1, N, nullptr);
}

int normalize_launch(sd_ctx* ctx, const double* d_x, int N, int D, double* d_xn) {
    gather_normalize_kernel<<<(unsigned)(((long)N * 32 + 255) / 256), 256, 0, ctx->stream>>>(d_x, nullptr, N, D, nullptr, d_xn);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

int pdist_condensed_launch(sd_ctx* ctx, const double* d_x, int N, int D, int mode, double* d_cond) {
    char* base;
    LinkLayout L;
    int rc = pdist_square(ctx, d_x, N, D, mode, &base, &L);
    if (rc) return rc;
    dim3 grid((N + 255) / 256, N);
    condense_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<double*>(base + L.off_D), L.ld, N, d_cond);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

// Clustering::linkage on device rows d_x[N][D] (already normalised by the caller, as in the reference)
int linkage_launch(sd_ctx* ctx, const double* d_x, int N, int D, double* d_Z, int mode) {
    if (N < 2) return SD_OK;
    char* base;
    LinkLayout L;
    if (!ctx->ev_lk[0])
        for (auto& e : ctx->ev_lk) SD_CUDA(ctx, cudaEventCreate(&e));
    SD_CUDA(ctx, cudaEventRecord(ctx->ev_lk[0], ctx->stream));
    int rc = pdist_square(ctx, d_x, N, D, mode, &base, &L);
    if (rc) return rc;
    SD_CUDA(ctx, cudaEventRecord(ctx->ev_lk[1], ctx->stream));
    trace_stamp(ctx, 3);
    rc = linkage_on_square(ctx, base, L, d_x, N, D, d_Z);
    if (rc) return rc;
    SD_CUDA(ctx, cudaEventRecord(ctx->ev_lk[2], ctx->stream));
    trace_stamp(ctx, 4);
    return SD_OK;
}

// Clustering::fcluster; d_T[N] labels 1..K, d_num receives K
int fcluster_launch(sd_ctx* ctx, const double* d_Z, int N, double cutoff, int* d_T, int* d_num) {
    // workspace: [flag | par: up, acc (2N each), gapbad, owner, newc (N each) | seq: MD, stack, stage]
    const size_t o_par = 256;
    const size_t par_bytes = align_up(sizeof(int) * (size_t)(11 * (size_t)N + 8), 256);
    const size_t o_md = o_par + par_bytes;
    const size_t o_stack = o_md + align_up((size_t)N * sizeof(double), 256);
    const size_t o_stage = o_stack + align_up((size_t)N * sizeof(int), 256);
    const size_t bytes = o_stage + align_up((size_t)N, 256);
    char* base = (char*)ctx->scratch(BUF_CL_WORK, bytes);
    if (!base) return SD_ERR_NOMEM;
    int* d_flag = reinterpret_cast<int*>(base);
    int* ip = reinterpret_cast<int*>(base + o_par);
    FcWork w;
    w.up = ip;
    w.acc = ip + 2 * (size_t)N;
    w.up2 = ip + 4 * (size_t)N;
    w.acc2 = ip + 6 * (size_t)N;
    w.gapbad = ip + 8 * (size_t)N;
    w.owner = ip + 9 * (size_t)N;
    w.newc = ip + 10 * (size_t)N;
    const size_t smem = sizeof(int) * (11 * (size_t)N + 8);
    const int use_smem = smem <= (size_t)200 * 1024;
    if (kernel_setup(ctx, fcluster_par_kernel, 200 * 1024) < 0) return SD_ERR_CUDA;
    fcluster_par_kernel<<<1, 1024, use_smem ? smem : 0, ctx->stream>>>(d_Z, N, cutoff, w, use_smem, d_T, d_num, d_flag);
    SD_LAUNCH_CHECK(ctx);
    fcluster_seq_kernel<<<1, 32, 0, ctx->stream>>>(d_Z, N, cutoff, reinterpret_cast<double*>(base + o_md),
                                                   reinterpret_cast<int*>(base + o_stack),
                                                   reinterpret_cast<unsigned char*>(base + o_stage), d_T, d_num, d_flag);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

int cosine_cdist_launch(sd_ctx* ctx, const double* d_a, int na, const double* d_b, int nb, int D, double* d_out) {
    const long total = (long)na * nb;
    cosine_cdist_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_a, na, d_b, nb, D, d_out,
                                                                                 ctx->d_status);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

// Cluster::cluster on filtered device rows.  d_x un-normalised [N][D]; d_labels[N] out; d_num out (device int).
// Synchronises once (the number of flat clusters sizes the post-processing workspace).
// k_cap == 0: the cluster count is read back after fcluster (one stream synchronisation) and sizes everything;
// k_cap > 0 : nothing is read back -- buffers and grids are sized for k_cap clusters, the kernels take the count
//             from *d_num and flag SD_ERR_CAPACITY in the status word if it does not fit.
int cluster_labels_launch(sd_ctx* ctx, const double* d_x, int N, int D, const sd_cluster_params* p, int* d_labels,
                          int* d_num, int k_cap) {
    if (p->num_clusters != -1)
        return ctx->fail(SD_ERR_UNSUPPORTED, "num_clusters != -1 is not implemented by the reference (SD:2368-2369)");
    double* d_xn = (double*)ctx->scratch(BUF_CL_XN, sizeof(double) * (size_t)N * D);
    if (!d_xn) return SD_ERR_NOMEM;
    int rc = normalize_launch(ctx, d_x, N, D, d_xn);
    if (rc) return rc;
    double* d_Z = (double*)ctx->scratch(BUF_CL_Z, sizeof(double) * 4 * (size_t)N);
    if (!d_Z) return SD_ERR_NOMEM;
    rc = linkage_launch(ctx, d_xn, N, D, d_Z, p->pdist_mode);
    if (rc) return rc;
    rc = fcluster_launch(ctx, d_Z, N, (double)p->threshold, d_labels, d_num);  // float threshold, SD:2049/2323
    if (rc) return rc;
    int K = k_cap;
    const int* d_k = k_cap > 0 ? d_num : nullptr;
    if (k_cap <= 0) {
        SD_CUDA(ctx, cudaMemcpyAsync(&K, d_num, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (K < 1 || K > N) return ctx->fail(SD_ERR_CUDA, "fcluster produced %d clusters for %d points", K, N);
    }
    const size_t o_count = 0, o_map = align_up(sizeof(int) * (size_t)(K + 1), 256),
                 o_rank = o_map + align_up(sizeof(int) * (size_t)(K + 1), 256),
                 o_flag = o_rank + align_up(sizeof(int) * (size_t)(K + 1), 256), o_sums = o_flag + 256,
                 total = o_sums + sizeof(double) * (size_t)K * D;
    char* base = (char*)ctx->scratch(BUF_CL_CENT, total);
    if (!base) return SD_ERR_NOMEM;
    PostWork w;
    w.x = d_x;
    w.labels = d_labels;
    w.count = reinterpret_cast<int*>(base + o_count);
    w.map = reinterpret_cast<int*>(base + o_map);
    w.rank = reinterpret_cast<int*>(base + o_rank);
    w.sums = reinterpret_cast<double*>(base + o_sums);
    w.num_clusters = d_num;
    w.status = ctx->d_status;
    w.need_means = reinterpret_cast<int*>(base + o_flag);
    cluster_post1_kernel<<<1, 1024, 0, ctx->stream>>>(w, N, K, p->min_cluster_size, d_k, k_cap);
    SD_LAUNCH_CHECK(ctx);
    if (D > 1024) return ctx->fail(SD_ERR_UNSUPPORTED, "embedding dimension %d > 1024", D);
    // after phase 1 *d_num is still fcluster's count whenever phase 2 / the means are needed at all
    cluster_means_kernel<<<K, means_threads(D), 0, ctx->stream>>>(d_x, d_labels, N, D, K, w.need_means, w.sums, d_k);
    SD_LAUNCH_CHECK(ctx);
    cluster_post2_kernel<<<1, 1024, 0, ctx->stream>>>(w, N, D, K, p->min_cluster_size, d_k);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

// Cluster::clustering.  d_emb[C*S][D]; h_keep = rows with a non-NaN first element (chunk-major).
// d_num_out == nullptr: synchronous (cluster counts are read back, *num_clusters_out is set);
// d_num_out != nullptr: nothing is read back, the final cluster count is copied to *d_num_out on the device.
int clustering_launch(sd_ctx* ctx, const double* d_emb, int C, int S, int D, const int* h_keep, int n_keep,
                      const sd_cluster_params* p, const double* d_binarized, int F, int* d_hard, double* d_soft,
                      int soft_k_cap, int* num_clusters_out, int* d_num_out, double* d_dist) {
    const int R = C * S;
    const int N = n_keep;
    const bool async = d_num_out != nullptr;
    const int k_cap = async ? std::min(N, 1024) : 0;
    // set_num_clusters (speakerDiarizer.cpp:2261-2296) with the defaults: min 1, max N
    int min_c = p->min_clusters == -1 ? 1 : p->min_clusters;
    int max_c = p->max_clusters == -1 ? N : p->max_clusters;
    min_c = std::max(1, std::min(N, min_c));
    max_c = std::max(1, std::min(N, max_c));
    if (max_c < 2) {  // speakerDiarizer.cpp:2081-2088: everything in cluster 0
        fill_int_kernel<<<(R + 255) / 256, 256, 0, ctx->stream>>>(d_hard, R, 0);
        SD_LAUNCH_CHECK(ctx);
        if (d_binarized) {
            inactive_mask_kernel<<<(R + 127) / 128, 128, 0, ctx->stream>>>(d_binarized, C, F, S, d_hard);
            SD_LAUNCH_CHECK(ctx);
        }
        if (num_clusters_out) *num_clusters_out = 1;
        if (async) {
            fill_int_kernel<<<1, 32, 0, ctx->stream>>>(d_num_out, 1, 1);
            SD_LAUNCH_CHECK(ctx);
        }
        return SD_OK;
    }
    int* d_keep = (int*)ctx->scratch(BUF_CL_MISC, sizeof(int) * (size_t)N + 256);
    double* d_x = (double*)ctx->scratch(BUF_CL_X, sizeof(double) * (size_t)N * D);
    double* d_xn = (double*)ctx->scratch(BUF_CL_XN, sizeof(double) * (size_t)N * D);
    int* d_labels = (int*)ctx->scratch(BUF_CL_LABELS, sizeof(int) * (size_t)N + 256);
    if (!d_keep || !d_x || !d_xn || !d_labels) return SD_ERR_NOMEM;
    int* d_num = d_labels + N;  // spare slot after the labels (scratch is over-allocated by 256 B)
    int rc = upload_small(ctx, d_keep, h_keep, sizeof(int) * (size_t)N);
    if (rc) return rc;
    // filter_embeddings (speakerDiarizer.cpp:2214-2259); x is also what cluster_labels normalises again
    gather_normalize_kernel<<<(unsigned)(((long)N * 32 + 255) / 256), 256, 0, ctx->stream>>>(d_emb, d_keep, N, D, d_x, nullptr);
    SD_LAUNCH_CHECK(ctx);
    rc = cluster_labels_launch(ctx, d_x, N, D, p, d_labels, d_num, k_cap);
    if (rc) return rc;
    int K = k_cap;
    const int* d_k = async ? d_num : nullptr;
    if (!async) {
        SD_CUDA(ctx, cudaMemcpyAsync(&K, d_num, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        // on the context stream, not cudaMemcpy: a copy on the legacy default stream would wait for every other
        // (blocking) stream of the process, i.e. for the other files of a batch
        if (ctx->h_status)
            SD_CUDA(ctx, cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->h_status && *ctx->h_status) return status_message(ctx, *ctx->h_status);
        if (K < 1 || K > N) return ctx->fail(SD_ERR_CUDA, "cluster post-processing produced %d clusters", K);
    }
    char* base = (char*)ctx->scratch(BUF_CL_OUT, sizeof(double) * (size_t)K * D);
    if (!base) return SD_ERR_NOMEM;
    AssignWork w;
    w.emb = d_emb;
    w.x = d_x;
    w.labels = d_labels;
    w.cent = reinterpret_cast<double*>(base);
    w.soft = d_soft;
    w.dist = d_dist;
    w.dist_k = nullptr;
    if (d_dist) {
        w.dist_k = (double*)ctx->scratch(BUF_GENERIC_A, sizeof(double) * (size_t)R * K);
        if (!w.dist_k) return SD_ERR_NOMEM;
    }
    w.hard = d_hard;
    w.binarized = d_binarized;
    w.status = ctx->d_status;
    if (D > 1024) return ctx->fail(SD_ERR_UNSUPPORTED, "embedding dimension %d > 1024", D);
    cluster_means_kernel<<<K, means_threads(D), 0, ctx->stream>>>(d_x, d_labels, N, D, K, nullptr, w.cent, d_k);
    SD_LAUNCH_CHECK(ctx);
    double* d_softk = (double*)ctx->scratch(BUF_CL_WORK, sizeof(double) * (size_t)R * K);
    if (!d_softk) return SD_ERR_NOMEM;
    assign_dist_kernel<<<(unsigned)(((long)R * K + 255) / 256), 256, 0, ctx->stream>>>(w, R, D, K, d_softk, K, d_k);
    SD_LAUNCH_CHECK(ctx);
    assign_argmax_kernel<<<(unsigned)(((long)R * 32 + 255) / 256), 256, 0, ctx->stream>>>(w, R, S, K, F, d_softk, K,
                                                                                      soft_k_cap, d_k);
    SD_LAUNCH_CHECK(ctx);
    if (async)
        SD_CUDA(ctx, cudaMemcpyAsync(d_num_out, d_num, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    else if (num_clusters_out)
        *num_clusters_out = K;
    return SD_OK;
}

int row_valid_launch(sd_ctx* ctx, const double* d_emb, int R, int D, unsigned char* d_valid) {
    row_valid_kernel<<<(R + 255) / 256, 256, 0, ctx->stream>>>(d_emb, R, D, d_valid);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

}  // namespace sdb
