// Pooled (cross-chain) Welford statistics for window adaptation.
//
// The reference adapts one chain at a time: welford_covariance.update folds one position into (mean, m2, n)
// (algorithms.py:166-197) and mass_matrix.covariance_adaptation.final turns (m2, n) into the inverse mass matrix
// (mass_matrix.py:81-118).  With many chains of the SAME target the statistic can be pooled: the positions of all
// chains in a slow window are one sample.  b2h_welford_pooled_update folds a whole block of draws [T][C][d] into a
// running float64 (n, mean[d], m2[d] or m2[d x d]) with Chan's pairwise update -- the group form of the reference's
// recurrence (same sums, batched) -- and b2h_welford_merge combines the states of two groups (the per-rank states
// after the all-gather over NVLink, SURVEY 8e "pooled adaptation").
//
// Everything accumulates in float64 and in a fixed order (per-block partials summed by one block), so the result
// does not depend on the launch schedule.
#include <algorithm>

#include "common.cuh"
#include "launch.h"

namespace b2h {

constexpr int kPoolBlocks = 592;     // 4 row slabs per SM
constexpr int kPoolSplit = 8;        // split-K planes of the d x d second-moment contraction

// partial[b][j] = sum over the rows of slab b of (x[r][j] - shift[j])^POW      (POW = 1: sums, 2: squares).
// The 256 threads of a block cover W columns x 256/W interleaved row groups (W = 32 .. 256, so narrow targets keep all
// lanes busy); the row groups are summed in a fixed order through shared memory.
template <typename T, int POW>
__global__ void __launch_bounds__(256) pool_colsum_kernel(const T* __restrict__ x, i64 n, int d, int W,
                                                          const double* __restrict__ shift, double* __restrict__ partial) {
    __shared__ double red[256];
    const i64 rows_per = (n + gridDim.x - 1) / gridDim.x;
    const i64 r0 = (i64)blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
    const int R = 256 / W, jj = threadIdx.x % W, rr = threadIdx.x / W;
    for (int jb = 0; jb < d; jb += W) {
        const int j = jb + jj;
        double a0 = 0.0, a1 = 0.0;
        if (j < d) {
            const double s = shift ? shift[j] : 0.0;
            i64 r = r0 + rr;
            for (; r + R < r1; r += 2 * R) {
                double v0 = (double)x[r * d + j] - s, v1 = (double)x[(r + R) * d + j] - s;
                if (POW == 2) { v0 *= v0; v1 *= v1; }
                a0 += v0; a1 += v1;
            }
            if (r < r1) {
                double v = (double)x[r * d + j] - s;
                a0 += (POW == 2) ? v * v : v;
            }
        }
        red[threadIdx.x] = a0 + a1;
        __syncthreads();
        if (rr == 0 && j < d) {
            double t = 0.0;
            for (int k = 0; k < R; ++k) t += red[k * W + jj];
            partial[(i64)blockIdx.x * d + j] = t;
        }
        __syncthreads();
    }
}

// out[j] = scale * sum_b partial[b][j]   (fixed order)
__global__ void pool_reduce_kernel(const double* __restrict__ partial, int nb, i64 len, double scale, double* out) {
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[(i64)b * len + j];
    out[j] = scale * s;
}

// xc[r][j] = x[r][j] - mean[j] (float64) and its transpose xct[j][r], 32 x 32 tiles through shared memory
template <typename T>
__global__ void __launch_bounds__(256) pool_center_kernel(const T* __restrict__ x, i64 n, int d,
                                                          const double* __restrict__ mean, double* __restrict__ xc,
                                                          double* __restrict__ xct, i64 ldt) {
    __shared__ double tile[32][33];
    const i64 r0 = (i64)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const i64 r = r0 + k;
        const int j = j0 + tx;
        double v = 0.0;
        if (r < n && j < d) {
            v = (double)x[r * d + j] - mean[j];
            xc[r * d + j] = v;
        }
        tile[k][tx] = v;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int j = j0 + k;
        const i64 r = r0 + tx;
        if (j < d && r < ldt) xct[(i64)j * ldt + r] = (r < n) ? tile[tx][k] : 0.0;
    }
}

// Chan's update of the running state (a) with a group (b): delta = mean_b - mean_a,
//   mean <- mean_a + delta n_b / n,  m2 <- m2_a + m2_b + outer(delta, delta) n_a n_b / n      (n = n_a + n_b)
__global__ void pool_merge_kernel(int d, int full, double na, double nb, double* mean_a, double* m2_a,
                                  const double* __restrict__ mean_b, const double* __restrict__ m2_b, double* mean_out) {
    const i64 len = full ? (i64)d * d : d;
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= len) return;
    const double n = na + nb;
    const int i = full ? (int)(k / d) : (int)k, j = full ? (int)(k % d) : (int)k;
    const double di = mean_b[i] - mean_a[i], dj = mean_b[j] - mean_a[j];
    m2_a[k] = m2_a[k] + m2_b[k] + di * dj * (na * nb / n);
    if (!full || i == j) mean_out[i] = mean_a[i] + di * (nb / n);   // written to a side buffer: mean_a is still being read
}

__global__ void pool_copy_kernel(const double* src, double* dst, i64 len) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < len) dst[k] = src[k];
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct PoolWs {
    double *partial, *mean_b, *m2_b, *mean_new, *xc, *xct, *planes;
    i64 ldt;
    size_t bytes;
};

static PoolWs pool_carve(char* base, i64 n, i64 d, int full) {
    PoolWs w;
    size_t off = 0;
    auto take = [&](size_t count) {
        double* p = base ? (double*)(base + off) : nullptr;
        off = align256(off + count * sizeof(double));
        return p;
    };
    w.partial = take((size_t)kPoolBlocks * d);
    w.mean_b = take(d);
    w.mean_new = take(d);
    w.m2_b = take(full ? (size_t)d * d : d);
    w.ldt = (n + 1) & ~(i64)1;                 // even pitch: 16-byte rows for the cp.async contraction path
    w.xc = w.xct = w.planes = nullptr;
    if (full) {
        w.xc = take((size_t)w.ldt * d);
        w.xct = take((size_t)w.ldt * d);
        w.planes = take((size_t)kPoolSplit * d * d);
    }
    w.bytes = off + 256;
    return w;
}

template <typename T>
static int pooled_update_typed(b2h_ctx* ctx, const T* draws, i64 n, int d, int full, i64 n_a, double* mean, double* m2,
                               void* ws, i64 ws_bytes) {
    PoolWs need = pool_carve(nullptr, n, d, full);
    if (!ws || (size_t)ws_bytes < need.bytes) {
        set_error("pooled welford: workspace too small: need " + std::to_string(need.bytes) + " bytes");
        return B2H_ERR_WORKSPACE;
    }
    PoolWs w = pool_carve((char*)ws, n, d, full);
    cudaStream_t st = ctx->stream;
    const int nb = (int)std::min<i64>(kPoolBlocks, n);
    const i64 len = full ? (i64)d * d : d;
    int W = 32;
    while (W < 256 && W < d) W *= 2;
    // group mean
    pool_colsum_kernel<T, 1><<<nb, 256, 0, st>>>(draws, n, d, W, nullptr, w.partial);
    pool_reduce_kernel<<<(d + 255) / 256, 256, 0, st>>>(w.partial, nb, d, 1.0 / (double)n, w.mean_b);
    // group m2 about the group mean
    if (!full) {
        pool_colsum_kernel<T, 2><<<nb, 256, 0, st>>>(draws, n, d, W, w.mean_b, w.partial);
        pool_reduce_kernel<<<(d + 255) / 256, 256, 0, st>>>(w.partial, nb, d, 1.0, w.m2_b);
    } else {
        dim3 grid((unsigned)((w.ldt + 31) / 32), (unsigned)((d + 31) / 32));
        pool_center_kernel<T><<<grid, 256, 0, st>>>(draws, n, d, w.mean_b, w.xc, w.xct, w.ldt);
        if (w.ldt > n) B2H_CUDA(cudaMemsetAsync(w.xc + (size_t)n * d, 0, (size_t)(w.ldt - n) * d * sizeof(double), st));
        // m2_b[d x d] = xct[d x n] . xc[n x d] on the FP64 tensor path, split over the rows
        launch_gemm<double>(st, w.xct, w.ldt, w.xc, (i64)d, w.planes, (i64)d, d, d, (int)w.ldt, nullptr, nullptr, kPoolSplit,
                            (i64)d * d);
        // (a slice whose row range is empty still writes its plane: zeros)
        pool_reduce_kernel<<<(int)((len + 255) / 256), 256, 0, st>>>(w.planes, kPoolSplit, len, 1.0, w.m2_b);
    }
    if (n_a == 0) {
        pool_copy_kernel<<<(d + 255) / 256, 256, 0, st>>>(w.mean_b, mean, d);
        pool_copy_kernel<<<(int)((len + 255) / 256), 256, 0, st>>>(w.m2_b, m2, len);
    } else {
        pool_merge_kernel<<<(int)((len + 255) / 256), 256, 0, st>>>(d, full, (double)n_a, (double)n, mean, m2, w.mean_b, w.m2_b,
                                                                    w.mean_new);
        pool_copy_kernel<<<(d + 255) / 256, 256, 0, st>>>(w.mean_new, mean, d);
    }
    B2H_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2h

using namespace b2h;

extern "C" {

int64_t b2h_welford_pooled_workspace_bytes(int64_t T_, int64_t C, int64_t d, int32_t full) {
    if (T_ <= 0 || C <= 0 || d <= 0) return -1;
    return (int64_t)pool_carve(nullptr, T_ * C, d, full).bytes;
}

int b2h_welford_pooled_update(b2h_ctx* ctx, int dtype, const void* draws, int64_t T_, int64_t C, int64_t d, int32_t full,
                              int64_t n, double* mean, double* m2, void* workspace, int64_t workspace_bytes) {
    if (!ctx || !draws || !mean || !m2) { set_error("null argument"); return B2H_ERR_ARG; }
    if (T_ <= 0 || C <= 0 || d <= 0 || n < 0 || d > (1 << 20)) { set_error("bad shape"); return B2H_ERR_ARG; }
    if (dtype == B2H_F64)
        return pooled_update_typed<double>(ctx, (const double*)draws, T_ * C, (int)d, full, n, mean, m2, workspace, workspace_bytes);
    if (dtype == B2H_F32)
        return pooled_update_typed<float>(ctx, (const float*)draws, T_ * C, (int)d, full, n, mean, m2, workspace, workspace_bytes);
    set_error("bad dtype");
    return B2H_ERR_ARG;
}

int b2h_welford_merge(b2h_ctx* ctx, int64_t d, int32_t full, int64_t n_a, double* mean_a, double* m2_a, int64_t n_b,
                      const double* mean_b, const double* m2_b, double* scratch_mean) {
    if (!ctx || !mean_a || !m2_a || !mean_b || !m2_b || !scratch_mean) { set_error("null argument"); return B2H_ERR_ARG; }
    if (d <= 0 || n_a < 0 || n_b < 0) { set_error("bad shape"); return B2H_ERR_ARG; }
    if (n_b == 0) return 0;
    const i64 len = full ? d * d : d;
    cudaStream_t st = ctx->stream;
    if (n_a == 0) {
        pool_copy_kernel<<<(int)((d + 255) / 256), 256, 0, st>>>(mean_b, mean_a, d);
        pool_copy_kernel<<<(int)((len + 255) / 256), 256, 0, st>>>(m2_b, m2_a, len);
    } else {
        pool_merge_kernel<<<(int)((len + 255) / 256), 256, 0, st>>>((int)d, full, (double)n_a, (double)n_b, mean_a, m2_a, mean_b,
                                                                    m2_b, scratch_mean);
        pool_copy_kernel<<<(int)((d + 255) / 256), 256, 0, st>>>(scratch_mean, mean_a, d);
    }
    B2H_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
