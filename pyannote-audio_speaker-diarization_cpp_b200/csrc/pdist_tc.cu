// pdist, tensor-core mode (SD_PDIST_GEMM_TF32X3): the N x N Euclidean distance matrix of the normalised
// embeddings as a Gram GEMM on the 5th-generation tensor cores.
//
//   d_ij = sqrt(max(0, n_i + n_j - 2 g_ij)),   g = X X^T
//
// * operands are fed by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 3-stage shared-memory ring,
// * tcgen05.mma (cta_group::1, kind::tf32, M = N = 128, K = 8 per instruction) accumulates in TMEM,
// * fp32 accuracy is recovered with the 3xTF32 split: x = hi + lo with hi = tf32(x), lo = fp32(x - hi);
//   g = hi.hi + hi.lo + lo.hi (three MMAs per K step into the same accumulator),
// * the epilogue reads the accumulator with tcgen05.ld, forms d in fp64 and writes both triangles of the square
//   matrix from the same value (the linkage needs an exactly symmetric matrix); pairs closer than `refine_below`
//   are recomputed with the exact fp64 difference form (where sqrt(2 - 2g) loses its digits).
//
// This is the *approximate* mode: the linkage compares distances for exact equality, so bit-exact
// merge order needs SD_PDIST_EXACT_F64; the measured error of this mode is reported by tests/bench.
#include <cuda.h>

#include "common.cuh"

#include <cmath>

namespace sdb {

namespace tc {

constexpr int BM = 128;      // rows of X per A tile (TMEM lanes)
constexpr int BN = 128;      // rows of X per B tile (accumulator columns)
constexpr int BK = 32;       // fp32 elements per K block = one 128-byte swizzle row
constexpr int UMMA_K = 8;    // K per tcgen05.mma for tf32
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;            // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;        // A_hi, A_lo, B_hi, B_lo
constexpr int THREADS = 320;                       // warp 0: TMA, warp 1: MMA + TMEM, warps 2..9: epilogue
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "LAB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity));
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
            "r"(smem_u32(smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle: rows of 128 B, 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_k_sw128(unsigned smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);        // start address
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(kIdesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}

struct Params {
    const double* xn;     // [N][D] fp64 normalised rows (exact refinement)
    const double* norms;  // [N] sum of squares, fp64
    double* Dm;           // [N][ld] output
    long ld;
    int N, D, tiles;      // tiles per dimension
    double refine_below;  // squared-distance threshold below which the pair is recomputed exactly
};

__global__ void __launch_bounds__(THREADS, 1)
    pdist_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, Params p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + STAGES;       // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // upper-triangle tile (bi <= bj) from the linear block index
    int bi = 0, rem = blockIdx.x;
    while (rem >= p.tiles - bi) {
        rem -= p.tiles - bi;
        ++bi;
    }
    const int bj = bi + rem;
    const int num_kb = (p.D + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_hi));
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_lo));
    }
    if (warp == 1) {  // TMEM: 128 lanes x 128 fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "n"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const unsigned ph = (unsigned)(kb / STAGES) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);  // first pass over the ring falls through
            unsigned char* st = smem + (size_t)s * STAGE_BYTES;
            mbar_expect_tx(&full[s], STAGE_BYTES);
            tma_load_2d(st + 0 * TILE_BYTES, &map_hi, kb * BK, bi * BM, &full[s]);
            tma_load_2d(st + 1 * TILE_BYTES, &map_lo, kb * BK, bi * BM, &full[s]);
            tma_load_2d(st + 2 * TILE_BYTES, &map_hi, kb * BK, bj * BN, &full[s]);
            tma_load_2d(st + 3 * TILE_BYTES, &map_lo, kb * BK, bj * BN, &full[s]);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer (one thread) =====
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const unsigned ph = (unsigned)(kb / STAGES) & 1u;
            mbar_wait(&full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            const unsigned base = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                const unsigned koff = ks * UMMA_K * 4;  // bytes inside the 128-byte swizzle row
                const uint64_t a_hi = umma_desc_k_sw128(base + 0 * TILE_BYTES + koff);
                const uint64_t a_lo = umma_desc_k_sw128(base + 1 * TILE_BYTES + koff);
                const uint64_t b_hi = umma_desc_k_sw128(base + 2 * TILE_BYTES + koff);
                const uint64_t b_lo = umma_desc_k_sw128(base + 3 * TILE_BYTES + koff);
                umma_tf32(tmem_base, a_hi, b_hi, (kb | ks) != 0);  // hi . hi
                umma_tf32(tmem_base, a_hi, b_lo, 1);               // hi . lo
                umma_tf32(tmem_base, a_lo, b_hi, 1);               // lo . hi
            }
            umma_commit(&empty[s]);  // frees the stage when these MMAs retire
        }
        umma_commit(tmem_full);
    } else if (warp >= 2) {
        // ===== epilogue: TMEM -> registers -> distances -> both triangles =====
        // eight warps: warp w reads TMEM lanes 32*(w%4).. (hardware restriction) and one half of the columns
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        const int lane_base = 32 * (warp & 3);
        const int col_half = (warp - 2) >> 2;  // 0: columns 0..63, 1: columns 64..127
        const int il = lane_base + lane;
        const int i = bi * BM + il;
        const double ni = i < p.N ? p.norms[i] : 0.0;
        for (int c0 = col_half * 64; c0 < col_half * 64 + 64; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            const int j0 = bj * BN + c0;
            if (bi == bj && c0 + 31 < lane_base) {
                // chunk entirely left of this warp's rows (j < i): produced by the mirrored stores of other rows
            } else if (i < p.N) {
                double dv[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int j = j0 + c;
                    const double g = (double)__uint_as_float(r[c]);
                    double d2 = ni + (j < p.N ? p.norms[j] : 0.0) - 2.0 * g;
                    if (d2 < p.refine_below && j < p.N && j > i) {  // cancellation zone: exact difference form
                        const double* a = p.xn + (size_t)i * p.D;
                        const double* b2 = p.xn + (size_t)j * p.D;
                        double s2 = 0.0;
                        for (int k = 0; k < p.D; ++k) {
                            const double df = __dsub_rn(a[k], b2[k]);
                            s2 = __dadd_rn(s2, __dmul_rn(df, df));
                        }
                        dv[c] = sqrt(s2);
                    } else
                        dv[c] = d2 > 0.0 ? (double)sqrtf((float)d2) : 0.0;  // fp32 sqrt: far below the GEMM's own error
                    if (j == i) dv[c] = 0.0;
                }
                // strictly-upper values go to row i (16-byte stores) and, mirrored, to column i (coalesced over lanes)
                double* row = p.Dm + (size_t)i * p.ld + j0;
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const int j = j0 + c;
                    if (j + 1 < p.N && j >= i) {
                        if (j > i)
                            *reinterpret_cast<double2*>(row + c) = make_double2(dv[c], dv[c + 1]);
                        else {
                            row[c] = 0.0;  // j == i
                            row[c + 1] = dv[c + 1];
                        }
                    } else {
                        if (j < p.N && j >= i) row[c] = dv[c];
                        if (j + 1 < p.N && j + 1 >= i) row[c + 1] = dv[c + 1];
                    }
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int j = j0 + c;
                    if (j < p.N && j > i) p.Dm[(size_t)j * p.ld + i] = dv[c];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(BN));
    }
}

// x (fp64) -> hi = tf32-truncated fp32, lo = fp32(x - hi); rows padded to ldx floats (zeros), norms in fp64
__global__ void __launch_bounds__(256)
    split_tf32_kernel(const double* __restrict__ x, int N, int D, int rows_pad, int ldx, float* __restrict__ hi,
                      float* __restrict__ lo, double* __restrict__ norms) {
    const int i = blockIdx.x;
    if (i >= rows_pad) return;
    double ss = 0.0;
    for (int k = threadIdx.x; k < ldx; k += blockDim.x) {
        float h = 0.f, l = 0.f;
        if (i < N && k < D) {
            const double v = x[(size_t)i * D + k];
            h = __uint_as_float(__float_as_uint((float)v) & 0xffffe000u);
            l = (float)(v - (double)h);
            ss += v * v;
        }
        hi[(size_t)i * ldx + k] = h;
        lo[(size_t)i * ldx + k] = l;
    }
    // block reduction of the squared norm
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0 && i < N) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        norms[i] = t;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace tc

// Dm[N][ld] <- pairwise distances of d_xn[N][D] (fp64) through the tensor cores.  Enqueue only.
int pdist_tc_launch(sd_ctx* ctx, const double* d_xn, int N, int D, double* Dm, long ld, double refine_below) {
    using namespace tc;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return ctx->fail(SD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const int tiles = (N + BM - 1) / BM;
    const int rows_pad = tiles * BM;
    const int ldx = (D + BK - 1) / BK * BK;
    const size_t mat = sizeof(float) * (size_t)rows_pad * ldx;
    char* base = (char*)ctx->scratch(BUF_TC_OPERANDS, 2 * mat + sizeof(double) * (size_t)rows_pad + 256);
    if (!base) return SD_ERR_NOMEM;
    float* d_hi = reinterpret_cast<float*>(base);
    float* d_lo = reinterpret_cast<float*>(base + mat);
    double* d_norm = reinterpret_cast<double*>(base + 2 * mat);
    split_tf32_kernel<<<rows_pad, 256, 0, ctx->stream>>>(d_xn, N, D, rows_pad, ldx, d_hi, d_lo, d_norm);
    SD_LAUNCH_CHECK(ctx);

    CUtensorMap map_hi, map_lo;
    const cuuint64_t dims[2] = {(cuuint64_t)ldx, (cuuint64_t)rows_pad};
    const cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r1 = enc(&map_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_hi, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&map_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_lo, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS)
        return ctx->fail(SD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);

    if (kernel_setup(ctx, pdist_tc_kernel, (int)SMEM_BYTES) < 0) return SD_ERR_CUDA;
    Params p;
    p.xn = d_xn;
    p.norms = d_norm;
    p.Dm = Dm;
    p.ld = ld;
    p.N = N;
    p.D = D;
    p.tiles = tiles;
    p.refine_below = refine_below;
    const unsigned grid = (unsigned)((long)tiles * (tiles + 1) / 2);
    pdist_tc_kernel<<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(map_hi, map_lo, p);
    SD_LAUNCH_CHECK(ctx);
    return SD_OK;
}

}  // namespace sdb
