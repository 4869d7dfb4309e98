// Bayesian logistic regression, all chains at once (exactness-reference FMA path).
//
//   s = X b_c ; U_c = sum_n softplus(s_n) - y_n s_n + 1/2 |b_c|^2 / sigma_b^2
//   dU/db_c = X^T (sigmoid(s) - y) + b_c / sigma_b^2            (oracle/models.py:LogisticRegression)
//
// Batched over chains this is S[C x N] = B[C x D] . X^T[D x N] followed by
// G[C x D] = R[C x N] . X[N x D], the only dense contraction of the engine.
// This file is the FP32/FP64 FMA-pipe version built on the tiled GEMM of gemm.cu,
// chunked over the data dimension so that S never exceeds a bounded scratch,
// with a deterministic split-K reduction for the tall-skinny second product.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "launch.h"

namespace b2h {

static const i64 kChunkBytes = 256ll << 20;     // S-chunk scratch budget

struct LogregPlan {
    int n_chunk;       // data rows per chunk
    int nsplit;        // split-K slices of the second product
    size_t off_S, off_part, off_U, total;
};

static LogregPlan plan_logreg(const b2h_model* m, int dtype, i64 C) {
    LogregPlan p;
    const size_t ts = dtype_size(dtype);
    i64 nc = kChunkBytes / (i64)(C * ts);
    nc = std::max<i64>(128, (nc / 128) * 128);
    if (nc > m->n_data) nc = ((m->n_data + 127) / 128) * 128;
    p.n_chunk = (int)nc;
    // enough CTAs to fill 148 SMs: tiles_m * tiles_n * nsplit >= ~2 waves
    i64 tiles = ((C + 127) / 128) * ((m->dim + 127) / 128);
    i64 ns = (296 + tiles - 1) / tiles;
    ns = std::max<i64>(1, std::min<i64>(ns, (p.n_chunk + 255) / 256));
    p.nsplit = (int)ns;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    p.off_S = take((size_t)C * p.n_chunk * ts);
    p.off_part = take((size_t)p.nsplit * C * m->dim * ts);
    p.off_U = take((size_t)C * sizeof(double));
    p.total = off + 256;
    return p;
}



// R = sigmoid(S) - y in place; U_acc[c] += sum_n softplus(s) - y s   (one CTA per chain: deterministic)
template <typename T>
__global__ void __launch_bounds__(256)
logistic_resid_kernel(T* S, const T* y, double* U_acc, int n_valid, i64 ld, int first) {
    const i64 c = blockIdx.x;
    T* row = S + c * ld;
    double acc = 0.0;
    for (int n = threadIdx.x; n < n_valid; n += 256) {
        T s = row[n], yy = y[n];
        T sp = fmax(s, (T)0) + log1p(exp(-fabs(s)));
        acc += (double)(sp - yy * s);
        row[n] = (T)1 / ((T)1 + exp(-s)) - yy;
    }
    __shared__ double red[8];
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        U_acc[c] = first ? s : U_acc[c] + s;
    }
}

// g = (first ? 0 : g) + sum_s partial[s]
template <typename T>
__global__ void split_reduce_kernel(const T* part, T* g, i64 n, int nsplit, int first) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T s = first ? (T)0 : g[i];
    for (int k = 0; k < nsplit; ++k) s += part[(i64)k * n + i];
    g[i] = s;
}

// prior: g += b / sigma^2 ; U = U_acc + 1/2 |b|^2 / sigma^2
template <typename T>
__global__ void __launch_bounds__(128)
logistic_prior_kernel(const T* q, T* g, T* U, const double* U_acc, T inv_prior_var, i64 C, int d) {
    const i64 c = (i64)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    T acc = 0;
    for (int j = lane; j < d; j += 32) {
        T b = q[c * d + j];
        g[c * d + j] += inv_prior_var * b;
        acc += b * b;
    }
    double s = Group<32>::sum1((double)acc, nullptr);
    if (lane == 0) U[c] = (T)U_acc[c] + (T)0.5 * inv_prior_var * (T)s;
}

// =============================================================================================================
// tensor-core path (tcgen05 / TMA, tc_gemm.cu): bf16 operands, fp32 accumulation in TMEM
// =============================================================================================================
struct LogregTcPlan {
    int n_chunk, nsplit;
    size_t off_bp, off_y, off_R, off_part, off_U, off_upart, total;
};

static LogregTcPlan plan_logreg_tc(const b2h_model* m, i64 C) {
    LogregTcPlan p;
    i64 nc = (2 * kChunkBytes) / (i64)(C * 6);              // residual pieces: 3 x bf16 per (chain, data row)
    nc = std::max<i64>(128, (nc / 128) * 128);
    if (nc > m->n_data) nc = ((m->n_data + 127) / 128) * 128;
    p.n_chunk = (int)nc;
    i64 tiles = ((C + 127) / 128) * ((m->dim + 127) / 128);
    i64 ns = (296 + tiles - 1) / tiles;
    ns = std::max<i64>(1, std::min<i64>(ns, (p.n_chunk + 511) / 512));
    p.nsplit = (int)ns;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) & ~(size_t)1023; return o; };
    p.off_bp = take((size_t)3 * C * m->dim * 2);
    p.off_y = take((size_t)m->n_data * 4);
    p.off_R = take((size_t)3 * ((C + 127) / 128) * 128 * p.n_chunk * 2);     // tile-blocked, rows padded to 128
    p.off_part = take((size_t)p.nsplit * C * m->dim * 4);
    p.off_U = take((size_t)C * sizeof(double));
    p.off_upart = take((size_t)4 * std::max<i64>(p.n_chunk / 128, 160) * C * sizeof(double));
    p.total = off + 1024;
    return p;
}

__device__ __forceinline__ void split3(double x, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
    a = __double2bfloat16(x);
    double r = x - (double)__bfloat162float(a);
    b = __double2bfloat16(r);
    r -= (double)__bfloat162float(b);
    c = __double2bfloat16(r);
}

// beta[C x d] (T) -> three stacked bf16 pieces [3*C x d] whose sum carries 24 significant bits
template <typename T>
__global__ void beta_split_kernel(const T* q, __nv_bfloat16* bp, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __nv_bfloat16 a, b, c;
    split3((double)q[i], a, b, c);
    bp[i] = a; bp[n + i] = b; bp[2 * n + i] = c;
}

template <typename T>
__global__ void to_float_kernel(const T* x, float* y, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = (float)x[i];
}

// U_acc[c] (+)= sum over the partials written by the contraction's epilogue (fixed order: deterministic).
// per_cta == 0: planes [column tile][4]; per_cta > 0: planes [CTA][4] where CTA b covered tiles [b*per, (b+1)*per)
// of the (row tile, column tile) grid, column tile fastest -- only the CTAs that touched chain c's row tile count.
__global__ void upart_reduce_kernel(const double* upart, double* U_acc, i64 C, int tiles_n, int per_cta, int first) {
    i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    int p_begin = 0, p_end = 4 * tiles_n;
    if (per_cta > 0) {
        const i64 mt = c / 128;
        p_begin = (int)((mt * tiles_n) / per_cta) * 4;
        p_end = (int)(((mt + 1) * tiles_n - 1) / per_cta + 1) * 4;
    }
    double s = first ? 0.0 : U_acc[c];
    for (int t = p_begin; t < p_end; ++t) s += upart[(i64)t * C + c];
    U_acc[c] = s;
}

// g (T) = (first ? 0 : g) + sum_s partial[s] (fp32 planes)
template <typename T>
__global__ void split_reduce_f32_kernel(const float* part, T* g, i64 n, int nsplit, int first) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = first ? 0.0 : (double)g[i];
    for (int k = 0; k < nsplit; ++k) s += (double)part[(i64)k * n + i];
    g[i] = (T)s;
}


// ---- fully fused tensor-core gradient (dim <= 128): one contraction kernel, no residual traffic ----------------
struct LogregFusedPlan {
    int planes;
    size_t off_bp, off_y, off_gpart, off_U, off_upart, total;
};

static LogregFusedPlan plan_logreg_fused(const b2h_model* m, i64 C) {
    LogregFusedPlan p;
    p.planes = logistic_fused_planes((int)C, (int)m->n_data);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) & ~(size_t)1023; return o; };
    p.off_bp = take((size_t)3 * C * m->dim * 2);
    p.off_y = take((size_t)m->n_data * 4);
    p.off_gpart = take((size_t)p.planes * C * m->dim * 4);
    p.off_U = take((size_t)C * sizeof(double));
    p.off_upart = take((size_t)4 * 160 * C * sizeof(double));
    p.total = off + 1024;
    return p;
}

static bool logistic_use_fused16(const b2h_model* m) { return (int)m->s1 == 4 && m->dim <= 128 && m->x_f16 && m->u_lin; }
static bool logistic_use_fused(const b2h_model* m) {
    return ((int)m->s1 == 2 || (int)m->s1 == 4) && m->dim <= 128;
}

// g (T) = sum over the planes written by the CTAs that touched the chain tile (fixed order) + b / sigma^2;
// U = sum of the CTAs' potential partials + 1/2 |b|^2 / sigma^2.  One warp per chain.
// VEC = 4 (float, dim % 4 == 0): every lane takes four consecutive coordinates, so the partial planes, q and g move as
// 128-bit pieces (the scalar version ran at a quarter of the HBM bandwidth: 114 us per launch at 131072 chains x 128).
template <typename T, int VEC>
__global__ void __launch_bounds__(128)
logistic_fused_finish_kernel(const float* __restrict__ gpart, i64 plane_stride, const double* __restrict__ upart,
                             const T* __restrict__ q, T* __restrict__ g, T* __restrict__ U, T inv_prior_var, i64 C, int d,
                             int tiles_n, int per_cta, double g_scale, const double* __restrict__ u_lin, double beta_limit) {
    const i64 c = (i64)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    const i64 mt = c / 128;
    const int b_first = (int)((mt * tiles_n) / per_cta);
    const int b_last = (int)(((mt + 1) * tiles_n - 1) / per_cta);
    T acc = 0;
    double lin = 0.0;
    const int np = b_last - b_first + 1;
    if constexpr (VEC == 4) {
        for (int j = lane * 4; j < d; j += 128) {
            double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 2
            for (int b = 0; b < np; ++b) {
                const float4 t = *reinterpret_cast<const float4*>(gpart + (i64)b * plane_stride + c * d + j);
                s[0] += (double)t.x; s[1] += (double)t.y; s[2] += (double)t.z; s[3] += (double)t.w;
            }
            const float4 b4 = *reinterpret_cast<const float4*>(q + c * d + j);
            const float bj[4] = {b4.x, b4.y, b4.z, b4.w};
            float out[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                out[x] = (float)(s[x] * g_scale) + (float)inv_prior_var * bj[x];
                acc += (T)(bj[x] * bj[x]);
                if (u_lin) lin += (double)bj[x] * u_lin[j + x];
                if (fabs((double)bj[x]) > beta_limit) lin = INFINITY;   // outside the fp16 pieces' range: reject the state
            }
            *reinterpret_cast<float4*>(g + c * d + j) = make_float4(out[0], out[1], out[2], out[3]);
        }
    } else {
#pragma unroll 4
        for (int j = lane; j < d; j += 32) {
            double s = 0.0;
#pragma unroll 2
            for (int b = 0; b < np; ++b) s += (double)gpart[(i64)b * plane_stride + c * d + j];
            const T bj = q[c * d + j];
            g[c * d + j] = (T)(s * g_scale) + inv_prior_var * bj;
            acc += bj * bj;
            if (u_lin) lin += (double)bj * u_lin[j];
            if (fabs((double)bj) > beta_limit) lin = INFINITY;       // outside the fp16 pieces' range: reject the state
        }
    }
    // potential partials of the CTAs that touched the chain tile: the lanes share the loads (fixed order: deterministic)
    double u = 0.0;
    for (int t = b_first * 4 + lane; t < (b_last + 1) * 4; t += 32) u += upart[(i64)t * C + c];
    double red[3] = {(double)acc, lin, u};
    Group<32>::sum<3>(red, nullptr);
    if (lane == 0) U[c] = (T)(red[1] + red[2]) + (T)0.5 * inv_prior_var * (T)red[0];
}

template <typename T>
static void launch_logistic_finish(cudaStream_t st, const float* gpart, i64 plane_stride, const double* upart, const T* q, T* g,
                                   T* U, T inv_prior_var, i64 C, int d, int tiles_n, int per_cta, double g_scale,
                                   const double* u_lin, double beta_limit) {
    const int grid = (int)((C + 3) / 4);
    if constexpr (sizeof(T) == sizeof(float)) {
        if (d % 4 == 0 && ((uintptr_t)q % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)gpart % 16 == 0) && plane_stride % 4 == 0) {
            logistic_fused_finish_kernel<T, 4><<<grid, 128, 0, st>>>(gpart, plane_stride, upart, q, g, U, inv_prior_var, C, d,
                                                                     tiles_n, per_cta, g_scale, u_lin, beta_limit);
            return;
        }
    }
    logistic_fused_finish_kernel<T, 1><<<grid, 128, 0, st>>>(gpart, plane_stride, upart, q, g, U, inv_prior_var, C, d, tiles_n,
                                                             per_cta, g_scale, u_lin, beta_limit);
}

template <typename T>
static int logistic_tc_fused(b2h_ctx* ctx, const b2h_model* m, const T* q, T* U, T* g, i64 C, void* ws, i64 ws_bytes) {
    if (!m->x_bf16) { set_error("logistic tensor-core path needs x_bf16"); return B2H_ERR_ARG; }
    if (m->dim % 8 || m->n_data % 8) { set_error("logistic tensor-core path needs dim and n_data multiples of 8"); return B2H_ERR_ARG; }
    LogregFusedPlan p = plan_logreg_fused(m, C);
    if (!ws || (size_t)ws_bytes < p.total) {
        set_error("logistic workspace too small: need " + std::to_string(p.total) + " bytes");
        return B2H_ERR_WORKSPACE;
    }
    cudaStream_t st = ctx->stream;
    const int d = m->dim;
    const i64 N = m->n_data;
    char* base = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* bp = (__nv_bfloat16*)(base + p.off_bp);
    float* yf = (float*)(base + p.off_y);
    float* gpart = (float*)(base + p.off_gpart);
    double* upart = (double*)(base + p.off_upart);
    const i64 ng = C * d;
    beta_split_kernel<T><<<(int)((ng + 255) / 256), 256, 0, st>>>(q, bp, ng);
    to_float_kernel<T><<<(int)((N + 255) / 256), 256, 0, st>>>((const T*)m->b, yf, N);
    int per_cta = 0, planes = 0;
    int rc = tc_logistic_fused(st, bp, (int)C, m->x_bf16, (int)C, (int)N, d, yf, gpart, upart, &per_cta, &planes);
    if (rc < 0) return rc;
    launch_logistic_finish<T>(st, gpart, ng, upart, q, g, U, (T)m->s0, C, d, (int)((N + 63) / 64), per_cta, 1.0, nullptr, INFINITY);
    B2H_LAUNCH_CHECK();
    return 0;
}


// beta[C x d] (T) -> two stacked fp16 pieces of beta * scale (22 significant bits; clamped to the fp16 range: the
// finish kernel turns a chain that left the range into a divergence)
template <typename T>
__global__ void beta_split16_kernel(const T* q, __half* bp, i64 n, double scale) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = (double)q[i] * scale;
    x = fmin(fmax(x, -65000.0), 65000.0);
    const __half a = __double2half(x);
    bp[i] = a;
    bp[n + i] = __double2half(x - (double)__half2float(a));
}

// float positions, four per thread: beta * 2^k is exact in float, the high piece is its fp16 rounding and the remainder
// float(x) - float(hi) is exact too, so the pieces are the ones the double-precision kernel above produces
__global__ void beta_split16_f32x4_kernel(const float* __restrict__ q, __half* __restrict__ bp, i64 n, float scale) {
    const i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const float4 v = *reinterpret_cast<const float4*>(q + i);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __half hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float t = fminf(fmaxf(x[k] * scale, -65000.f), 65000.f);
        hi[k] = __float2half_rn(t);
        lo[k] = __float2half_rn(t - __half2float(hi[k]));
    }
    *reinterpret_cast<uint2*>(bp + i) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(bp + n + i) = *reinterpret_cast<const uint2*>(lo);
}

template <typename T>
static void launch_beta_split16(cudaStream_t st, const T* q, __half* bp, i64 n, double scale) {
    if constexpr (sizeof(T) == sizeof(float)) {
        if (n % 4 == 0 && (uintptr_t)q % 16 == 0 && (uintptr_t)bp % 8 == 0) {
            beta_split16_f32x4_kernel<<<(int)((n / 4 + 255) / 256), 256, 0, st>>>((const float*)q, bp, n, (float)scale);
            return;
        }
    }
    beta_split16_kernel<T><<<(int)((n + 255) / 256), 256, 0, st>>>(q, bp, n, scale);
}

template <typename T>
static int logistic_tc_fused16(b2h_ctx* ctx, const b2h_model* m, const T* q, T* U, T* g, i64 C, void* ws, i64 ws_bytes) {
    if (m->dim % 8 || m->n_data % 8) { set_error("logistic tensor-core path needs dim and n_data multiples of 8"); return B2H_ERR_ARG; }
    LogregFusedPlan p = plan_logreg_fused(m, C);
    if (!ws || (size_t)ws_bytes < p.total) {
        set_error("logistic workspace too small: need " + std::to_string(p.total) + " bytes");
        return B2H_ERR_WORKSPACE;
    }
    cudaStream_t st = ctx->stream;
    const int d = m->dim;
    const i64 N = m->n_data;
    char* base = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    __half* bp = (__half*)(base + p.off_bp);                 // the plan reserves three 16-bit pieces; two are used
    float* yf = (float*)(base + p.off_y);
    float* gpart = (float*)(base + p.off_gpart);
    double* upart = (double*)(base + p.off_upart);
    const i64 ng = C * d;
    // Scales are powers of two tied together: X16 = X * 2^shift (largest entry just below 2^15), beta pieces hold
    // beta * 2^(20 - shift), so the accumulator is always s * 2^20 and the representable |beta| grows as X shrinks
    // (|beta| < 255 for data of unit scale).
    const int beta_exp = 20 - m->x_f16_shift;
    launch_beta_split16<T>(st, q, bp, ng, ldexp(1.0, beta_exp));
    if (sizeof(T) == sizeof(float)) yf = (float*)m->b;           // the responses are float already
    else to_float_kernel<T><<<(int)((N + 255) / 256), 256, 0, st>>>((const T*)m->b, yf, N);
    int per_cta = 0;
    int rc = tc_logistic_fused16(st, bp, m->x_f16, 20, (int)C, (int)N, d, yf, gpart, upart, &per_cta);
    if (rc < 0) return rc;
    launch_logistic_finish<T>(st, gpart, ng, upart, q, g, U, (T)m->s0, C, d, (int)((N + 63) / 64), per_cta,
                              ldexp(1.0, -m->x_f16_shift), m->u_lin, 65000.0 * ldexp(1.0, -beta_exp));
    B2H_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int logistic_tc(b2h_ctx* ctx, const b2h_model* m, const T* q, T* U, T* g, i64 C, void* ws, i64 ws_bytes) {
    if (!m->x_bf16 || !m->xt_bf16) { set_error("logistic tensor-core path needs x_bf16 and xt_bf16"); return B2H_ERR_ARG; }
    if (m->dim % 8 || m->n_data % 8) { set_error("logistic tensor-core path needs dim and n_data multiples of 8"); return B2H_ERR_ARG; }
    LogregTcPlan p = plan_logreg_tc(m, C);
    if (!ws || (size_t)ws_bytes < p.total) {
        set_error("logistic workspace too small: need " + std::to_string(p.total) + " bytes");
        return B2H_ERR_WORKSPACE;
    }
    cudaStream_t st = ctx->stream;
    const int d = m->dim;
    const i64 N = m->n_data;
    char* base = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* bp = (__nv_bfloat16*)(base + p.off_bp);
    float* yf = (float*)(base + p.off_y);
    double* upart = (double*)(base + p.off_upart);
    __nv_bfloat16* R = (__nv_bfloat16*)(base + p.off_R);
    float* part = (float*)(base + p.off_part);
    double* U_acc = (double*)(base + p.off_U);
    const __nv_bfloat16* Xb = (const __nv_bfloat16*)m->x_bf16;      // [N x d]
    const __nv_bfloat16* Xtb = (const __nv_bfloat16*)m->xt_bf16;    // [d x N]
    const T* y = (const T*)m->b;
    const i64 ng = C * d;
    beta_split_kernel<T><<<(int)((ng + 255) / 256), 256, 0, st>>>(q, bp, ng);
    to_float_kernel<T><<<(int)((N + 255) / 256), 256, 0, st>>>(y, yf, N);
    for (i64 n0 = 0, it = 0; n0 < N; n0 += p.n_chunk, ++it) {
        const int nv = (int)std::min<i64>(p.n_chunk, N - n0);
        // S[C x nv] = sum_p Beta_p[C x d] . X[n0.., d]^T stays in TMEM; the epilogue emits the residual pieces
        const i64 r_piece = (i64)((C + 127) / 128) * 128 * p.n_chunk;
        int per_cta = 0;
        int rc = tc_gemm_logistic(st, bp, d, Xb + n0 * d, d, (int)C, nv, d, 3, (int)C, yf + n0, R, r_piece, upart, &per_cta);
        if (rc < 0) return rc;
        upart_reduce_kernel<<<(int)((C + 63) / 64), 64, 0, st>>>(upart, U_acc, C, (nv + 127) / 128, per_cta, it == 0);
        // G[C x d] += sum_p R_p[C x nv] . Xt[d, n0..]^T   (split-K over the data rows)
        rc = tc_gemm_blocked_a(st, R, r_piece, Xtb + n0, N, part, (int)C, d, nv, 3, d, p.nsplit, ng);
        if (rc < 0) return rc;
        split_reduce_f32_kernel<T><<<(int)((ng + 255) / 256), 256, 0, st>>>(part, g, ng, rc, it == 0);
    }
    logistic_prior_kernel<T><<<(int)((C + 3) / 4), 128, 0, st>>>(q, g, U, U_acc, (T)m->s0, C, d);
    B2H_LAUNCH_CHECK();
    return 0;
}

i64 logistic_workspace_bytes(const b2h_model* m, int dtype, i64 C) {
    if (logistic_use_fused(m)) return (i64)plan_logreg_fused(m, C).total + 1024;
    if ((int)m->s1 >= 2) return (i64)plan_logreg_tc(m, C).total + 1024;
    return (i64)plan_logreg(m, dtype, C).total;
}

template <typename T>
int logistic_potential_and_grad(b2h_ctx* ctx, const b2h_model* m, const T* q, T* U, T* g, i64 C, void* ws,
                                i64 ws_bytes, int path) {
    // model flag s1: 0 = FMA / DMMA exactness reference, 2 = tensor cores (fully fused when dim <= 128), 3 = tensor
    // cores with the two-kernel formulation (residual pieces through memory)
    if (path == 0 && logistic_use_fused16(m)) return logistic_tc_fused16<T>(ctx, m, q, U, g, C, ws, ws_bytes);
    if (path == 0 && logistic_use_fused(m)) return logistic_tc_fused<T>(ctx, m, q, U, g, C, ws, ws_bytes);
    if (path == 2 || (path == 0 && (int)m->s1 >= 2)) return logistic_tc<T>(ctx, m, q, U, g, C, ws, ws_bytes);
    if (!m->a || !m->b || !m->c) { set_error("logistic model needs X (a), y (b) and X^T (c)"); return B2H_ERR_ARG; }
    LogregPlan p = plan_logreg(m, Num<T>::dtype, C);
    if (!ws || (size_t)ws_bytes < p.total) {
        set_error("logistic workspace too small: need " + std::to_string(p.total) + " bytes");
        return B2H_ERR_WORKSPACE;
    }
    cudaStream_t st = ctx->stream;
    const int d = m->dim;
    const i64 N = m->n_data;
    const T* X = (const T*)m->a;      // [N x d]
    const T* y = (const T*)m->b;      // [N]
    const T* Xt = (const T*)m->c;     // [d x N]
    T* S = (T*)((char*)ws + p.off_S);
    T* part = (T*)((char*)ws + p.off_part);
    double* U_acc = (double*)((char*)ws + p.off_U);
    const i64 ng = C * d;
    for (i64 n0 = 0, it = 0; n0 < N; n0 += p.n_chunk, ++it) {
        const int nv = (int)std::min<i64>(p.n_chunk, N - n0);
        // S[C x nv] = B[C x d] . X^T[d x nv]
        launch_gemm<T>(st, q, (i64)d, Xt + n0, N, S, (i64)p.n_chunk, (int)C, nv, d, nullptr, nullptr, 1, 0);
        logistic_resid_kernel<T><<<(int)C, 256, 0, st>>>(S, y + n0, U_acc, nv, (i64)p.n_chunk, it == 0);
        // G[C x d] += R[C x nv] . X[nv x d]   (split-K over the data rows, deterministic reduce)
        launch_gemm<T>(st, S, (i64)p.n_chunk, X + n0 * d, (i64)d, part, (i64)d, (int)C, d, nv, nullptr, nullptr,
                       p.nsplit, ng);
        split_reduce_kernel<T><<<(int)((ng + 255) / 256), 256, 0, st>>>(part, g, ng, p.nsplit, it == 0);
    }
    logistic_prior_kernel<T><<<(int)((C + 3) / 4), 128, 0, st>>>(q, g, U, U_acc, (T)m->s0, C, d);
    B2H_LAUNCH_CHECK();
    return 0;
}

template int logistic_potential_and_grad<float>(b2h_ctx*, const b2h_model*, const float*, float*, float*, i64, void*,
                                                i64, int);
template int logistic_potential_and_grad<double>(b2h_ctx*, const b2h_model*, const double*, double*, double*, i64,
                                                 void*, i64, int);

}  // namespace b2h
