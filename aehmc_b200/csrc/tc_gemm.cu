// tcgen05 / TMA tensor-core contraction for the batched logistic-regression gradient (config 3 / 5).
//
//   D[M x N] (fp32) = sum_{p < pieces} A_p[M x K] . B[N x K]^T        A_p, B: bf16, K-major (row-major, K contiguous)
//
// The gradient of all chains is two such products:
//   S[C x n] = sum_p Beta_p[C x D] . X[n x D]^T          (Beta split into 3 bf16 pieces: exact to 24 bits)
//   G[C x D] = sum_p R_p[C x n]    . Xt[D x n]^T         (R = sigmoid(S) - y split into 3 bf16 pieces)
// X is bf16-representable by construction, so every product is exact and accumulation is fp32 in TMEM: the result
// has fp32-class accuracy, which is what the float32 parity bar (1e-4) needs; the FMA path stays the exactness
// reference.
//
// Kernel anatomy (one 128 x 128 output tile per CTA, optional split-K over blockIdx.z):
//   warp 0 : TMA producer  -- cp.async.bulk.tensor 2D loads of 128 x 64 bf16 boxes (SWIZZLE_128B) into a 6-stage ring
//   warp 1 : MMA issuer    -- one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M128 N128 K16) from smem
//                             descriptors, tcgen05.commit releases the smem stage / signals the epilogue
//   warp 2 : TMEM allocator (128 columns)
//   warps 4-7 : epilogue   -- tcgen05.ld 32x32b.x32 (TMEM lane quadrant = warp % 4), fp32 -> global
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "launch.h"

namespace b2h {

namespace tc {

constexpr int BM = 128, BN = 128, BK = 64;          // BK bf16 = one 128-byte swizzle row
constexpr int STAGES = 6;
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 128;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// K-major, SWIZZLE_128B shared-memory operand descriptor (sm_100 "version 1"): rows of 128 bytes, 8-row groups
// 1024 bytes apart (SBO), leading offset unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)1 << 16;                              // leading byte offset (ignored)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}

// instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__global__ void __launch_bounds__(THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out,
               int M, int N, int K, int pieces, int piece_rows, int ldo, int k_blocks_per_split, long long split_stride) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(tiles + (size_t)STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kb_total = (K + BK - 1) / BK;
    const int kb_begin = blockIdx.z * k_blocks_per_split;
    const int kb_end = min(kb_total, kb_begin + k_blocks_per_split);
    const int n_kb = max(kb_end - kb_begin, 0);
    const int iters = n_kb * pieces;                       // (piece, k-block) pairs accumulated into one tile

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES, round = it / STAGES;
                mbar_wait(&empty_bar[s], (round & 1) ^ 1);
                const int p = it / n_kb, kb = kb_begin + it % n_kb;
                uint8_t* a_dst = tiles + (size_t)s * STAGE_BYTES;
                uint8_t* b_dst = a_dst + A_BYTES;
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                tma_load_2d(a_dst, &map_a, &full_bar[s], kb * BK, p * piece_rows + m0);
                tma_load_2d(b_dst, &map_b, &full_bar[s], kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES, round = it / STAGES;
                mbar_wait(&full_bar[s], round & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t ad = make_desc(a_addr + k * UMMA_K * 2);
                    const uint64_t bd = make_desc(b_addr + k * UMMA_K * 2);
                    umma_bf16(tmem_base, ad, bd, IDESC, (it > 0 || k > 0) ? 1u : 0u);
                }
                tcgen05_commit(&empty_bar[s]);             // frees the smem stage when these MMAs retire
            }
            tcgen05_commit(tmem_full);                     // accumulator complete
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp & 3;                            // TMEM lane quadrant this warp may access
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + q * 32 + lane;
        float* orow = out + (long long)blockIdx.z * split_stride + (long long)row * ldo + n0;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < M) {
                if (iters == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int col = n0 + c0 + j;
                    if (col + 3 < N) {
                        *reinterpret_cast<float4*>(orow + c0 + j) =
                            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                        __uint_as_float(r[j + 3]));
                    } else {
                        for (int e = 0; e < 4; ++e)
                            if (col + e < N) orow[c0 + j + e] = __uint_as_float(r[j + e]);
                    }
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2D bf16 row-major [rows][cols] (cols contiguous, row pitch ld elements), box = 64 x 128, 128-byte swizzle
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return B2H_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return B2H_ERR_CUDA; }
    return 0;
}

}  // namespace tc

// out[z][M x N] (fp32, row pitch ldo, plane stride split_stride) = sum_p A[p*piece_rows + m][k] * B[n][k]
// over the k-blocks of split z.  A: [pieces*piece_rows x K] bf16 (pitch lda), B: [N x K] bf16 (pitch ldb).
int tc_gemm(cudaStream_t st, const void* A, long long lda, const void* B, long long ldb, float* out, int M, int N, int K,
            int pieces, int piece_rows, int ldo, int nsplit, long long split_stride) {
    using namespace tc;
    if ((lda * 2) % 16 || (ldb * 2) % 16 || ((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)out & 15) || ldo % 4) {
        set_error("tc_gemm: operands must be 16-byte aligned with 16-byte row pitches");
        return B2H_ERR_ARG;
    }
    CUtensorMap ma, mb;
    int rc = make_map(&ma, A, (long long)pieces * piece_rows, K, lda);
    if (rc) return rc;
    rc = make_map(&mb, B, N, K, ldb);
    if (rc) return rc;
    const int kb_total = (K + BK - 1) / BK;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > kb_total) nsplit = kb_total;
    const int kb_per = (kb_total + nsplit - 1) / nsplit;
    nsplit = (kb_total + kb_per - 1) / kb_per;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, nsplit);
    B2H_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    tc_gemm_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ma, mb, out, M, N, K, pieces, piece_rows, ldo, kb_per, split_stride);
    B2H_LAUNCH_CHECK();
    return nsplit;
}

}  // namespace b2h

extern "C" int b2h_tc_gemm_bf16(b2h_ctx* ctx, const void* A, int64_t lda, const void* B, int64_t ldb, float* out, int64_t M,
                                int64_t N, int64_t K, int32_t pieces, int64_t piece_rows, int64_t ldo, int32_t nsplit,
                                int64_t split_stride) {
    if (!ctx) { b2h::set_error("null context"); return B2H_ERR_ARG; }
    int rc = b2h::tc_gemm(ctx->stream, A, lda, B, ldb, out, (int)M, (int)N, (int)K, pieces, (int)piece_rows, (int)ldo,
                          nsplit, split_stride);
    return rc < 0 ? rc : 0;
}
