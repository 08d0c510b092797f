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

#include <cfloat>
#include <cmath>

namespace sdb {

// ------------------------------------------------------------------------------------------------
// gather + normalise
// ------------------------------------------------------------------------------------------------

// x[i][:] = emb[keep[i]][:];  xn[i][:] = x[i][:] / (double)(float)sqrt(sum_k x^2)  (sum in k order, fp64)
__global__ void __launch_bounds__(128)
    gather_normalize_kernel(const double* __restrict__ emb, const int* __restrict__ keep, int N, int D,
                            double* __restrict__ x, double* __restrict__ xn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double* src = emb + (size_t)(keep ? keep[i] : i) * D;
    double ss = 0.0;
    for (int k = 0; k < D; ++k) {
        const double v = src[k];
        if (x) x[(size_t)i * D + k] = v;
        ss = __dadd_rn(ss, __dmul_rn(v, v));
    }
    const double norm = (double)(float)__dsqrt_rn(ss);  // L2Norm returns float, speakerDiarizer.cpp:332
    for (int k = 0; k < D; ++k) {
        const double v = src[k];
        xn[(size_t)i * D + k] = norm != 0.0 ? __ddiv_rn(v, norm) : v;
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

// Dm[i][j] = sqrt(sum_k (x_ik - x_jk)^2), k ascending, mul then add, no FMA.  Full square (both triangles
// are produced by identical arithmetic since (a-b)^2 == (b-a)^2 bit for bit).
__global__ void __launch_bounds__(256)
    pdist_f64_kernel(const double* __restrict__ x, int N, int D, double* __restrict__ Dm, long ld) {
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
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int gi = bi + ty * 4 + p;
        if (gi >= N) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int gj = bj + tx * 4 + q;
            if (gj < N) Dm[(size_t)gi * ld + gj] = __dsqrt_rn(acc[p][q]);
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
    int* pos_of;    // heap: key -> slot
    int* key_at;    // heap: slot -> key
    double* hval;   // heap: slot -> value
    unsigned* bitmap;  // [ceil(N/32)] rows whose bound dropped in this merge
    double* Z;      // [N-1][4]
    int* status;    // device status word
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
__global__ void __launch_bounds__(256) rowmin_init_kernel(LinkWork w, int n) {
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
struct Heap {
    int* pos_of;
    int* key_at;
    double* val;
    int n;
    __device__ __forceinline__ void swp(int a, int b) {
        const double tv = val[a];
        val[a] = val[b];
        val[b] = tv;
        const int ka = key_at[a], kb = key_at[b];
        key_at[a] = kb;
        key_at[b] = ka;
        pos_of[ka] = b;
        pos_of[kb] = a;
    }
    __device__ __forceinline__ void down(int i) {
        for (int c = 2 * i + 1; c < n; c = 2 * i + 1) {
            if (c + 1 < n && val[c + 1] < val[c]) ++c;
            if (!(val[i] > val[c])) break;
            swp(i, c);
            i = c;
        }
    }
    __device__ __forceinline__ void up(int i) {
        while (i > 0) {
            const int p = (i - 1) >> 1;
            if (!(val[p] > val[i])) break;
            swp(i, p);
            i = p;
        }
    }
    __device__ __forceinline__ void set(int key, double v) {
        const int i = pos_of[key];
        const double old = val[i];
        val[i] = v;
        if (v < old)
            up(i);
        else
            down(i);
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

// One persistent CTA performs all N-1 merges of one problem (grid.x = number of independent problems).
__global__ void __launch_bounds__(LK_THREADS) linkage_kernel(const LinkWork* __restrict__ works, const int* ns) {
    const LinkWork w = works[blockIdx.x];
    const int n = ns[blockIdx.x];
    __shared__ MinIdx red[33];
    __shared__ int s_x, s_y, s_nx, s_ny, s_stale;
    __shared__ double s_dist;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n < 2) return;

    Heap h;
    h.pos_of = w.pos_of;
    h.key_at = w.key_at;
    h.val = w.hval;
    h.n = n - 1;

    // heapify: the reference sifts down i = size/2 .. 0 (clustering.cpp:94-96).  Sift-downs of nodes on one
    // level touch disjoint subtrees, and every deeper level is finished before a shallower one in the
    // sequential order too, so processing level by level (deepest first) yields the same heap.
    {
        const int last = h.n / 2;
        int top = 0;
        while (((2 << top) - 1) <= last) ++top;  // deepest level that contains an index <= last
        for (int lev = top; lev >= 0; --lev) {
            const int first = (1 << lev) - 1;
            int end = (2 << lev) - 2;
            if (end > last) end = last;
            for (int i = first + tid; i <= end; i += LK_THREADS)
                if (i < h.n) h.down(i);
            __syncthreads();
        }
    }

    const int nwords = (n + 31) >> 5;
    for (int k = 0; k < n - 1; ++k) {
        // ---- pop the closest pair, revalidating stale candidates (clustering.cpp:323-339) ----
        int tries = 0;
        for (;;) {
            if (tid == 0) {
                const int x = h.key_at[0];
                const double dist = h.val[0];
                const int y = w.nbr[x];
                s_x = x;
                s_y = y;
                s_dist = dist;
                s_stale = (y < 0) || !(dist == w.D[(size_t)x * w.ld + y]);
            }
            __syncthreads();
            if (!s_stale) break;
            const int x = s_x;
            MinIdx m;
            m.v = INFINITY;
            m.i = -1;
            const double* r = w.D + (size_t)x * w.ld;
            for (int i = x + 1 + tid; i < n; i += LK_THREADS) {
                if (w.size[i] == 0) continue;
                const double d = r[i];
                if (d < m.v) {
                    m.v = d;
                    m.i = i;
                }
            }
            m = block_min(m, red);
            if (tid == 0) {
                const double v = m.i >= 0 ? m.v : INFINITY;
                w.nbr[x] = m.i;
                w.lb[x] = v;
                h.set(x, v);
                s_y = m.i;
                s_dist = v;
            }
            ++tries;
            if (tries >= n - k) {  // loop bound of the reference: fall through with the recomputed pair
                __syncthreads();
                break;
            }
            __syncthreads();
        }
        const int x = s_x, y = s_y;
        const double dist = s_dist;
        if (y < 0) {  // no live partner (only reachable through NaN distances): the reference indexes out of range
            if (tid == 0) atomicExch(w.status, SD_ERR_INVALID);
            return;
        }
        if (tid == 0) {
            h.swp(0, h.n - 1);  // remove_min, clustering.cpp:103-107
            h.n -= 1;
            h.down(0);
            int ix = w.cid[x], iy = w.cid[y];
            const int nx = w.size[x], ny = w.size[y];
            if (ix > iy) {
                const int t = ix;
                ix = iy;
                iy = t;
            }
            double* z = w.Z + 4 * (size_t)k;
            z[0] = ix;
            z[1] = iy;
            z[2] = dist;
            z[3] = nx + ny;
            s_nx = nx;
            s_ny = ny;
            w.size[x] = 0;
            w.size[y] = nx + ny;
            w.cid[y] = n + k;
        }
        __syncthreads();
        h.n = (n - 1) - (k + 1);  // every thread tracks the heap size
        const int nx = s_nx, ny = s_ny;

        // ---- Lance-Williams update of row/column y, neighbour fix-ups, and y's own nearest neighbour
        //      (clustering.cpp:361-404), fused into one sweep over z ----
        const double* rowx = w.D + (size_t)x * w.ld;
        double* rowy = w.D + (size_t)y * w.ld;
        MinIdx ym;
        ym.v = INFINITY;
        ym.i = -1;
        for (int z0 = 0; z0 < n; z0 += LK_THREADS) {
            const int z = z0 + tid;
            bool changed = false;
            if (z < n && z != y && w.size[z] != 0) {
                const double nd = centroid_update(rowx[z], rowy[z], dist, nx, ny);
                rowy[z] = nd;
                w.D[(size_t)z * w.ld + y] = nd;
                if (z < x && w.nbr[z] == x) w.nbr[z] = y;
                if (z < y) {
                    if (nd < w.lb[z]) {
                        w.nbr[z] = y;
                        w.lb[z] = nd;
                        changed = true;
                    }
                } else if (nd < ym.v) {  // z > y, ascending within a thread: first strict minimum
                    ym.v = nd;
                    ym.i = z;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, changed);
            if (lane == 0 && (z0 >> 5) + warp < nwords) w.bitmap[(z0 >> 5) + warp] = bal;
        }
        ym = block_min(ym, red);  // contains __syncthreads: bitmap, lb, nbr are visible afterwards

        // ---- replay the heap updates in increasing z, exactly as the sequential loop would ----
        if (warp == 0) {
            for (int wb = 0; wb < nwords; wb += 32) {
                const unsigned word = (wb + lane < nwords) ? w.bitmap[wb + lane] : 0u;
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
                            h.set(z, w.lb[z]);
                        }
                    }
                }
            }
            if (lane == 0 && y < n - 1 && ym.i != -1) {  // clustering.cpp:395-404
                w.nbr[y] = ym.i;
                w.lb[y] = ym.v;
                h.set(y, ym.v);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// fcluster (criterion "distance")
// ------------------------------------------------------------------------------------------------

// Single thread: both passes are O(n) chains of dependent steps.  T[n] receives labels 1..K in the
// reference's depth-first numbering; *num_out = K.
__global__ void fcluster_kernel(const double* __restrict__ Z, int n, double cutoff, double* __restrict__ MD,
                                int* __restrict__ stack, unsigned char* __restrict__ stage, int* __restrict__ T,
                                int* __restrict__ num_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
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
    int* status;
};

// Per-cluster sums in index order (calculateClusterMeans, speakerDiarizer.cpp:443-473; assign_embeddings
// 2147-2167): thread j owns dimension j and walks the rows once, so each cluster's sum is formed in
// increasing row order exactly like the reference, then divided by the count.
__device__ void cluster_means(const double* x, const int* labels, int N, int D, int K, const int* count,
                              double* sums) {
    for (long e = threadIdx.x; e < (long)K * D; e += blockDim.x) sums[e] = 0.0;
    __syncthreads();
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        for (int i = 0; i < N; ++i) {
            const int l = labels[i];
            double* s = sums + (size_t)l * D + j;
            *s = __dadd_rn(*s, x[(size_t)i * D + j]);
        }
    }
    __syncthreads();
    for (long e = threadIdx.x; e < (long)K * D; e += blockDim.x) sums[e] = __ddiv_rn(sums[e], (double)count[e / D]);
    __syncthreads();
}

__global__ void __launch_bounds__(1024) cluster_post_kernel(PostWork w, int N, int D, int K, int min_cluster_size) {
    __shared__ int s_nl, s_ns, s_next;
    const int tid = threadIdx.x;
    // labels - 1, counts (speakerDiarizer.cpp:2324-2341)
    for (int l = tid; l <= K; l += blockDim.x) {
        w.count[l] = 0;
        w.map[l] = l;
    }
    if (tid == 0) {
        s_nl = 0;
        s_ns = 0;
    }
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) {
        const int l = w.labels[i] - 1;
        w.labels[i] = l;
        atomicAdd(&w.count[l], 1);
    }
    __syncthreads();
    // min_cluster_size heuristic (speakerDiarizer.cpp:2308-2309)
    long mcs = (long)round(0.1 * (double)N);
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
        if (tid == 0) *w.num_clusters = 1;
        return;
    }
    if (ns == 0) {  // speakerDiarizer.cpp:2377-2380: labels as they are (all K clusters are large and present)
        if (tid == 0) *w.num_clusters = K;
        return;
    }
    cluster_means(w.x, w.labels, N, D, K, w.count, w.sums);
    // each small cluster -> nearest large cluster by centroid cosine distance, float running minimum
    // (speakerDiarizer.cpp:2390-2415); large clusters are visited in ascending label order
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
    // rank of each surviving label among the sorted unique labels (speakerDiarizer.cpp:519-548)
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
    int* count;          // [K]
    double* cent;        // [K][D]
    double* soft;        // optional [R][soft_k_cap]
    int* hard;           // [R]
    const double* binarized;  // optional [C][F][S]
    int* status;
};

// centroids (one CTA; see cluster_means)
__global__ void __launch_bounds__(1024) centroid_kernel(AssignWork w, int N, int D, int K) {
    for (int l = threadIdx.x; l < K; l += blockDim.x) w.count[l] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(&w.count[w.labels[i]], 1);
    __syncthreads();
    cluster_means(w.x, w.labels, N, D, K, w.count, w.cent);
}

// one thread per embedding row: distances to every centroid in order, soft = 2 - d, first strict maximum
// (speakerDiarizer.cpp:2180-2211), then the inactive-speaker mask (3166-3191)
__global__ void __launch_bounds__(128)
    assign_kernel(AssignWork w, int R, int S, int D, int K, int F, int soft_k_cap) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const double* e = w.emb + (size_t)r * D;
    int arg = 0;
    double best = -DBL_MAX;
    for (int k = 0; k < K; ++k) {
        double d;
        if (!cosine_distance(e, w.cent + (size_t)k * D, D, &d)) {
            atomicExch(w.status, SD_ERR_ZERO_MAGNITUDE);
            d = NAN;
        }
        const double soft = __dsub_rn(2.0, d);
        if (w.soft && k < soft_k_cap) w.soft[(size_t)r * soft_k_cap + k] = soft;
        if (soft > best) {
            best = soft;
            arg = k;
        }
    }
    if (w.binarized) {
        const int c = r / S, s = r - c * S;
        float acc = 0.0f;  // the reference accumulates 0/1 values in float (exact)
        for (int f = 0; f < F; ++f) acc += (float)w.binarized[((size_t)c * F + f) * S + s];
        if (fabsf(acc) < DBL_EPSILON) arg = -2;
    }
    w.hard[r] = arg;
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

struct LinkLayout {
    long ld;
    size_t off_D, off_size, off_cid, off_nbr, off_lb, off_pos, off_key, off_hval, off_bitmap, off_work, off_n, total;
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
    L.total = o;
    return L;
}

// pdist (exact) into the square matrix of the linkage workspace; returns the workspace base
static int pdist_square(sd_ctx* ctx, const double* d_xn, int N, int D, char** base_out, LinkLayout* L_out) {
    LinkLayout L = link_layout(N);
    char* base = (char*)ctx->scratch(BUF_CL_DIST, L.total);
    if (!base) return SD_ERR_NOMEM;
    dim3 grid((N + PD_TILE - 1) / PD_TILE, (N + PD_TILE - 1) / PD_TILE);
    pdist_f64_kernel<<<grid, 256, 0, ctx->stream>>>(d_xn, N, D, reinterpret_cast<double*>(base + L.off_D), L.ld);
    SD_LAUNCH_CHECK(ctx);
    *base_out = base;
    *L_out = L;
    return SD_OK;
}

static int linkage_on_square(sd_ctx* ctx, char* base, const LinkLayout& L, int N, double* d_Z) {
    LinkWork w;
    w.D = reinterpret_cast<double*>(base + L.off_D);
    w.ld = L.ld;
    w.size = reinterpret_cast<int*>(base + L.off_size);
    w.cid = reinterpret_cast<int*>(base + L.off_cid);
    w.nbr = reinterpret_cast<int*>(base + L.off_nbr);
    w.lb = reinterpret_cast<double*>(base + L.off_lb);
    w.pos_of = reinterpret_cast<int*>(base + L.off_pos);
    w.key_at = reinterpret_cast<int*>(base + L.off_key);
    w.hval = reinterpret_cast<double*>(base + L.off_hval);
    w.bitmap = reinterpret_cast<unsigned*>(base + L.off_bitmap);
    w.Z = d_Z;
    w.status = ctx->d_status;
    int rc = upload_small(ctx, base + L.off_work, &w, sizeof(w));
    if (rc) return rc;
    rc = upload_small(ctx, base + L.off_n, &N, sizeof(int));
    if (rc) return rc;
    rowmin_init_kernel<<<(unsigned)(((long)N * 32 + 255) / 256), 256, 0, ctx->stream>>>(w, N);
    SD_LAUNCH_CHECK(ctx);
    linkage_kernel<<<1, LK_THREADS, 0, ctx->stream>>>(reinterpret_cast<const LinkWork*>(base + L.off_work),
                                                      reinterpret_cast<const int*>(base + L.off_n));
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

int normalize_launch(sd_ctx* ctx, const double* d_x, int N, int D, double* d_xn) {
    gather_normalize_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(d_x, nullptr, N, D, nullptr, d_xn);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

int pdist_condensed_launch(sd_ctx* ctx, const double* d_x, int N, int D, double* d_cond) {
    char* base;
    LinkLayout L;
    int rc = pdist_square(ctx, d_x, N, D, &base, &L);
    if (rc) return rc;
    dim3 grid((N + 255) / 256, N);
    condense_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<double*>(base + L.off_D), L.ld, N, d_cond);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

// Clustering::linkage on device rows d_x[N][D] (already normalised by the caller, as in the reference)
int linkage_launch(sd_ctx* ctx, const double* d_x, int N, int D, double* d_Z) {
    if (N < 2) return SD_OK;
    char* base;
    LinkLayout L;
    int rc = pdist_square(ctx, d_x, N, D, &base, &L);
    if (rc) return rc;
    return linkage_on_square(ctx, base, L, N, d_Z);
}

// Clustering::fcluster; d_T[N] labels 1..K, d_num receives K
int fcluster_launch(sd_ctx* ctx, const double* d_Z, int N, double cutoff, int* d_T, int* d_num) {
    size_t bytes = align_up((size_t)N * sizeof(double), 256) + align_up((size_t)N * sizeof(int), 256) +
                   align_up((size_t)N, 256);
    char* base = (char*)ctx->scratch(BUF_CL_WORK, bytes);
    if (!base) return SD_ERR_NOMEM;
    double* MD = reinterpret_cast<double*>(base);
    int* stack = reinterpret_cast<int*>(base + align_up((size_t)N * sizeof(double), 256));
    unsigned char* stage = reinterpret_cast<unsigned char*>(base + align_up((size_t)N * sizeof(double), 256) +
                                                            align_up((size_t)N * sizeof(int), 256));
    fcluster_kernel<<<1, 32, 0, ctx->stream>>>(d_Z, N, cutoff, MD, stack, stage, d_T, d_num);
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
int cluster_labels_launch(sd_ctx* ctx, const double* d_x, int N, int D, const sd_cluster_params* p, int* d_labels,
                          int* d_num) {
    if (p->num_clusters != -1)
        return ctx->fail(SD_ERR_UNSUPPORTED, "num_clusters != -1 is not implemented by the reference (SD:2368-2369)");
    double* d_xn = (double*)ctx->scratch(BUF_CL_XN, sizeof(double) * (size_t)N * D);
    if (!d_xn) return SD_ERR_NOMEM;
    int rc = normalize_launch(ctx, d_x, N, D, d_xn);
    if (rc) return rc;
    double* d_Z = (double*)ctx->scratch(BUF_CL_Z, sizeof(double) * 4 * (size_t)N);
    if (!d_Z) return SD_ERR_NOMEM;
    if (p->pdist_mode != SD_PDIST_EXACT_F64)
        return ctx->fail(SD_ERR_UNSUPPORTED, "pdist_mode %d is not available in this build", p->pdist_mode);
    rc = linkage_launch(ctx, d_xn, N, D, d_Z);
    if (rc) return rc;
    rc = fcluster_launch(ctx, d_Z, N, (double)p->threshold, d_labels, d_num);  // float threshold, SD:2049/2323
    if (rc) return rc;
    int K = 0;
    SD_CUDA(ctx, cudaMemcpyAsync(&K, d_num, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (K < 1 || K > N) return ctx->fail(SD_ERR_CUDA, "fcluster produced %d clusters for %d points", K, N);
    const size_t o_count = 0, o_map = align_up(sizeof(int) * (size_t)(K + 1), 256),
                 o_rank = o_map + align_up(sizeof(int) * (size_t)(K + 1), 256),
                 o_sums = o_rank + align_up(sizeof(int) * (size_t)(K + 1), 256),
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
    cluster_post_kernel<<<1, 1024, 0, ctx->stream>>>(w, N, D, K, p->min_cluster_size);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

// Cluster::clustering.  d_emb[C*S][D]; h_keep = rows with a non-NaN first element (chunk-major).
int clustering_launch(sd_ctx* ctx, const double* d_emb, int C, int S, int D, const std::vector<int>& h_keep,
                      const sd_cluster_params* p, const double* d_binarized, int F, int* d_hard, double* d_soft,
                      int soft_k_cap, int* num_clusters_out) {
    const int R = C * S;
    const int N = (int)h_keep.size();
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
        return SD_OK;
    }
    int* d_keep = (int*)ctx->scratch(BUF_CL_MISC, sizeof(int) * (size_t)N + 256);
    double* d_x = (double*)ctx->scratch(BUF_CL_X, sizeof(double) * (size_t)N * D);
    double* d_xn = (double*)ctx->scratch(BUF_CL_XN, sizeof(double) * (size_t)N * D);
    int* d_labels = (int*)ctx->scratch(BUF_CL_LABELS, sizeof(int) * (size_t)N + 256);
    if (!d_keep || !d_x || !d_xn || !d_labels) return SD_ERR_NOMEM;
    int* d_num = d_labels + N;  // spare slot after the labels (scratch is over-allocated by 256 B)
    int rc = upload_small(ctx, d_keep, h_keep.data(), sizeof(int) * (size_t)N);
    if (rc) return rc;
    // filter_embeddings (speakerDiarizer.cpp:2214-2259); x is also what cluster_labels normalises again
    gather_normalize_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(d_emb, d_keep, N, D, d_x, d_xn);
    SD_LAUNCH_CHECK(ctx);
    rc = cluster_labels_launch(ctx, d_x, N, D, p, d_labels, d_num);
    if (rc) return rc;
    int K = 0;
    SD_CUDA(ctx, cudaMemcpyAsync(&K, d_num, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_status) {
        SD_CUDA(ctx, cudaMemcpy(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost));
        if (*ctx->h_status) return ctx->fail(*ctx->h_status, "Vectors have zero magnitude.");
    }
    if (K < 1 || K > N) return ctx->fail(SD_ERR_CUDA, "cluster post-processing produced %d clusters", K);
    const size_t o_cent = align_up(sizeof(int) * (size_t)K, 256);
    char* base = (char*)ctx->scratch(BUF_CL_OUT, o_cent + sizeof(double) * (size_t)K * D);
    if (!base) return SD_ERR_NOMEM;
    AssignWork w;
    w.emb = d_emb;
    w.x = d_x;
    w.labels = d_labels;
    w.count = reinterpret_cast<int*>(base);
    w.cent = reinterpret_cast<double*>(base + o_cent);
    w.soft = d_soft;
    w.hard = d_hard;
    w.binarized = d_binarized;
    w.status = ctx->d_status;
    centroid_kernel<<<1, 1024, 0, ctx->stream>>>(w, N, D, K);
    SD_LAUNCH_CHECK(ctx);
    assign_kernel<<<(R + 127) / 128, 128, 0, ctx->stream>>>(w, R, S, D, K, F, soft_k_cap);
    SD_LAUNCH_CHECK(ctx);
    if (num_clusters_out) *num_clusters_out = K;
    return SD_OK;
}

int row_valid_launch(sd_ctx* ctx, const double* d_emb, int R, int D, unsigned char* d_valid) {
    row_valid_kernel<<<(R + 255) / 256, 256, 0, ctx->stream>>>(d_emb, R, D, d_valid);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

}  // namespace sdb
