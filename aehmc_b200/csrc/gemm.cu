// Dense apply for all chains at once:  out[M x N] = A[M x K] . B[K x N]
//
// This is the mass-matrix apply (metrics.py:71,95-96: velocity = imm @ p) and the
// correlated-Gaussian gradient (Lambda (q - mu)) of every chain in one launch.
// B is shared by all chains, so the per-chain mat-vecs of the reference become
// one [C x d].[d x d] product: FP64/FP32 FMA-pipe bound, not HBM bound
// (SURVEY.md section 8d).  FP64 has no tcgen05 path on sm_100a, so this is a
// register-tiled SIMT kernel: 128x128x8 CTA tile, 256 threads, 8x8 outputs per
// thread laid out as 4x4 chunks of 2 so that every shared-memory read is a
// conflict-free 128-bit load, global->shared software-pipelined through
// registers.  Rows may be limited by a DEVICE-side count (compacted momentum
// rows), so no host synchronisation is needed to size the problem.
#include "common.cuh"
#include "launch.h"

namespace b2h {

constexpr int BM = 128, BN = 128, BK = 8, GT = 256;

template <typename T>
__global__ void __launch_bounds__(GT, 1)
dense_apply_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ out, int M, int N, int K,
                   i64 lda, i64 ldb, i64 ldo, const int* __restrict__ m_dev, const T* __restrict__ sub,
                   int tiles_n, int k_chunk, i64 split_stride) {
    if (m_dev) M = min(M, *m_dev);
    const int tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x % tiles_n;
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    if (m0 >= M) return;
    // split-K: slice blockIdx.y handles k in [k_begin, k_end) and writes its own partial plane
    const int k_begin = blockIdx.y * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    out += (i64)blockIdx.y * split_stride;

    __shared__ __align__(16) T As[2][BK][BM];   // transposed: [k][m]
    __shared__ __align__(16) T Bs[2][BK][BN];   // [k][n]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    // global->smem assignment: A tile 128 rows x 8 k : thread loads 4 consecutive k of one row
    const int a_row = tid >> 1, a_k = (tid & 1) * 4;
    // B tile 8 k x 128 n : thread loads 4 consecutive n of one k row
    const int b_k = tid >> 5, b_n = (tid & 31) * 4;

    T acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = (T)0;

    T ra[4], rb[4];
    auto load_tile = [&](int k0) {
        const int gr = m0 + a_row;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int gk = k0 + a_k + i;
            T x = (T)0;
            if (gr < M && gk < k_end) {
                x = A[(i64)gr * lda + gk];
                if (sub) x -= sub[gk];
            }
            ra[i] = x;
        }
        const int gk = k0 + b_k;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int gn = n0 + b_n + i;
            rb[i] = (gk < k_end && gn < N) ? B[(i64)gk * ldb + gn] : (T)0;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) As[buf][a_k + i][a_row] = ra[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) Bs[buf][b_k][b_n + i] = rb[i];
    };

    const int nk = (k_end - k_begin + BK - 1) / BK;
    load_tile(k_begin);
    store_tile(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tile(k_begin + (kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[8], b[8];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                // rows ty*2 + 32*c + {0,1}, cols tx*2 + 32*c + {0,1}
                a[2 * c] = As[buf][kk][ty * 2 + 32 * c];
                a[2 * c + 1] = As[buf][kk][ty * 2 + 32 * c + 1];
                b[2 * c] = Bs[buf][kk][tx * 2 + 32 * c];
                b[2 * c + 1] = Bs[buf][kk][tx * 2 + 32 * c + 1];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tile(buf ^ 1);
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gr = m0 + ty * 2 + 32 * (i >> 1) + (i & 1);
        if (gr >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int gn = n0 + tx * 2 + 32 * (j >> 1) + (j & 1);
            if (gn < N) out[(i64)gr * ldo + gn] = acc[i][j];
        }
    }
}

// out = (A - sub) . B ; sub (optional, [K]) is subtracted from every row of A on load (q - mu).
template <typename T>
void launch_dense_apply(cudaStream_t st, const T* A, const T* B, T* out, int M, int N, int K, const int* m_dev,
                        const T* sub) {
    launch_gemm<T>(st, A, (i64)K, B, (i64)N, out, (i64)N, M, N, K, m_dev, sub, 1, 0);
}

// General strided form with split-K: slice s of nsplit writes out + s*split_stride (a partial plane).
template <typename T>
void launch_gemm(cudaStream_t st, const T* A, i64 lda, const T* B, i64 ldb, T* out, i64 ldo, int M, int N, int K,
                 const int* m_dev, const T* sub, int nsplit, i64 split_stride) {
    if (M <= 0 || N <= 0) return;
    int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
    if (nsplit < 1) nsplit = 1;
    int k_chunk = (K + nsplit - 1) / nsplit;
    k_chunk = ((k_chunk + BK - 1) / BK) * BK;
    dim3 grid(tiles_m * tiles_n, nsplit);
    dense_apply_kernel<T><<<grid, GT, 0, st>>>(A, B, out, M, N, K, lda, ldb, ldo, m_dev, sub, tiles_n, k_chunk,
                                               split_stride);
}

template void launch_gemm<float>(cudaStream_t, const float*, i64, const float*, i64, float*, i64, int, int, int,
                                 const int*, const float*, int, i64);
template void launch_gemm<double>(cudaStream_t, const double*, i64, const double*, i64, double*, i64, int, int, int,
                                  const int*, const double*, int, i64);

template void launch_dense_apply<float>(cudaStream_t, const float*, const float*, float*, int, int, int, const int*,
                                        const float*);
template void launch_dense_apply<double>(cudaStream_t, const double*, const double*, double*, int, int, int,
                                         const int*, const double*);

}  // namespace b2h
