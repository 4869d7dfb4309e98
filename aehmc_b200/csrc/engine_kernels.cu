// Kernels and launch logic of the tick engine (see engine.cuh).
//
//  fused mode : model gradient is a device function (iid Gaussian, funnel, eight
//               schools) and the metric is scalar/diagonal -> ONE persistent kernel
//               runs every chain through all its ticks; nothing returns to the host.
//  split mode : gradient and/or metric need an all-chain contraction (dense metric,
//               correlated Gaussian, logistic regression) -> each tick is
//               pre (half kick + drift) -> gradient -> [dense metric: w' = imm.g', one contraction] -> post,
//               all chains in lock-step, chains restarting transitions independently.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "engine.cuh"
#include "launch.h"
#include "models.cuh"

namespace b2h {

// ---------------------------------------------------------------------------
// workspace carving
// ---------------------------------------------------------------------------
struct Carver {
    char* base;
    size_t off;
    template <typename U>
    U* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        U* p = base ? (U*)(base + off) : nullptr;
        off += n * sizeof(U);
        return p;
    }
};

constexpr int kRiderSplit = 5;      // split-K slices of the momentum contractions (few rows, full K)

struct EnginePlan {
    int G;
    bool dense, split, hmc, per_chain_imm, scalar_imm;
    size_t model_ws_off, model_ws_bytes;
};

static int auto_group(int d, long long C) {
    // tiny targets: thread-per-chain only when there are enough chains to fill the machine with threads
    // (measured on config 4, 65536 chains: 8 lanes per chain is 20-35 % faster than 1)
    if (d <= 16) return C >= 262144 ? 1 : 8;
    if (d <= 64) return 8;
    if (d <= 512) return 32;
    return 256;
}

static bool model_is_fused(int kind) {
    return kind == B2H_MODEL_IID_GAUSSIAN || kind == B2H_MODEL_FUNNEL || kind == B2H_MODEL_EIGHT_SCHOOLS;
}

static int make_plan(const b2h_model* model, const b2h_metric* metric, const b2h_cfg* cfg, EnginePlan& pl,
                     long long C = 0) {
    pl.dense = metric->kind == B2H_IMM_DENSE;
    pl.per_chain_imm = metric->kind == B2H_IMM_DIAG_PER_CHAIN;
    pl.scalar_imm = metric->kind == B2H_IMM_SCALAR;
    pl.split = pl.dense || !model_is_fused(model->kind);
    int G = cfg->group > 0 ? cfg->group : auto_group(model->dim, C);
    if (pl.split && G < 8) G = 8;          // split scratch is row-major: needs the row-major layout
    if (G != 1 && G != 8 && G != 32 && G != 256) {
        set_error("group must be one of 0 (auto), 1, 8, 32, 256");
        return B2H_ERR_ARG;
    }
    if ((model->kind == B2H_MODEL_FUNNEL || model->kind == B2H_MODEL_EIGHT_SCHOOLS) && !pl.split && G > 8) G = 8;
    pl.G = G;
    return 0;
}

template <typename T>
static size_t carve(EngineView<T>& v, char* base, const EnginePlan& pl, const b2h_model* model, int C, int d,
                    int maxd, bool adapt, size_t model_ws_bytes, size_t* model_ws_off) {
    Carver cv{base, 0};
    const size_t n = (size_t)C * d;
    v.ql = cv.take<T>(n); v.pl = cv.take<T>(n); v.gl = cv.take<T>(n);
    v.qr = cv.take<T>(n); v.pr = cv.take<T>(n); v.gr = cv.take<T>(n);
    v.qs = cv.take<T>(n); v.ps = cv.take<T>(n); v.gs = cv.take<T>(n);
    v.qp = cv.take<T>(n); v.pp = cv.take<T>(n); v.gp = cv.take<T>(n);
    v.msum = cv.take<T>(n); v.sms = cv.take<T>(n);
    v.mck = cv.take<T>(n * maxd); v.sckp = cv.take<T>(n * maxd);
    v.vl = v.vr = v.vck = nullptr;
    v.wl = v.wr = v.ws = v.wp = nullptr;
    if (pl.dense) {
        v.vl = cv.take<T>(n); v.vr = cv.take<T>(n); v.vck = cv.take<T>(n * maxd);
        v.wl = cv.take<T>(n); v.wr = cv.take<T>(n); v.ws = cv.take<T>(n); v.wp = cv.take<T>(n);
    }
    v.rec = cv.take<ChainRec>(C);
    T* imm_own = nullptr;
    if (pl.per_chain_imm) imm_own = cv.take<T>(n);
    else if (pl.scalar_imm) imm_own = cv.take<T>(1);
    v.imm = imm_own;
    v.adapt.wc_mean = v.adapt.wc_m2 = nullptr;
    if (adapt) { v.adapt.wc_mean = cv.take<T>(n); v.adapt.wc_m2 = cv.take<T>(n); }
    v.xa = v.xb = v.xc = v.Unew = nullptr;
    v.mom_p = v.mom_v = v.mom_z = nullptr; v.mom_count = nullptr; v.mom_list = nullptr;
    if (pl.split) {
        v.xa = cv.take<T>(n); v.xb = cv.take<T>(n); v.Unew = cv.take<T>(C);
    }
    if (pl.dense) {
        v.xc = cv.take<T>(n);
        v.mom_p = cv.take<T>(n); v.mom_v = cv.take<T>(n); v.mom_z = cv.take<T>(2 * n);
        v.mom_count = cv.take<int>(4); v.mom_list = cv.take<int>(2 * (size_t)C);
        v.mom_part = cv.take<T>(2 * (size_t)kRiderSplit * n);
    }
    v.scratch = cv.take<int>(64);
    if (model_ws_bytes) {
        cv.off = (cv.off + 255) & ~(size_t)255;
        if (model_ws_off) *model_ws_off = cv.off;
        cv.off += model_ws_bytes;
    }
    return cv.off + 256;
}

// ---------------------------------------------------------------------------
// entry / exit kernels (user row-major arrays <-> engine layout)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void copy_in_kernel(EngineView<T> v, const T* q, const T* g, const T* U, const double* eps,
                               const T* imm_user, T imm_scalar, int write_scalar, double init_step_size) {
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0 && write_scalar) v.imm[0] = imm_scalar;
    if (idx >= (i64)v.C * v.d) return;
    int c = (int)(idx / v.d), j = (int)(idx % v.d);
    i64 a = (i64)c * v.sc + (i64)j * v.sj;
    v.qp[a] = q[idx];
    v.gp[a] = g[idx];
    if (imm_user) v.imm[a] = imm_user[idx];
    if (v.adapt.enabled) { ((T*)v.adapt.wc_mean)[a] = 0; ((T*)v.adapt.wc_m2)[a] = 0; }
    if (j == 0) {
        ChainRec r;
        memset(&r, 0, sizeof(r));
        r.phase = PH_START;
        r.U_prop = (double)U[c];
        r.eps = eps[c];
        if (v.adapt.enabled) {
            // window_adaptation.init (window_adaptation.py:132-144): mu = initial step size, step size = exp(0)
            v.adapt.da_step[c] = 1; v.adapt.da_x[c] = 0.0; v.adapt.da_x_avg[c] = 0.0; v.adapt.da_g_avg[c] = 0.0;
            v.adapt.da_mu[c] = init_step_size;
            v.adapt.wc_n[c] = 0;
            r.eps = 1.0;
        }
        v.rec[c] = r;
    }
}

template <typename T>
__global__ void copy_out_kernel(EngineView<T> v, T* q, T* p, T* g, T* U, double* eps, T* imm_user) {
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (i64)v.C * v.d) return;
    int c = (int)(idx / v.d), j = (int)(idx % v.d);
    i64 a = (i64)c * v.sc + (i64)j * v.sj;
    q[idx] = v.qp[a];
    if (p) p[idx] = v.pp[a];
    g[idx] = v.gp[a];
    if (imm_user) imm_user[idx] = v.imm[a];
    if (j == 0) {
        const ChainRec& r = v.rec[c];
        U[c] = (T)r.U_prop;
        eps[c] = r.eps;
        if (v.out.acceptance_probability) v.out.acceptance_probability[c] = r.accept_prob;
        if (v.out.num_doublings) v.out.num_doublings[c] = r.last_nd;
        if (v.out.is_turning) v.out.is_turning[c] = (r.last_flags & 1) ? 1 : 0;
        if (v.out.is_diverging) v.out.is_diverging[c] = (r.last_flags & 2) ? 1 : 0;
        if (v.out.n_leapfrog) v.out.n_leapfrog[c] = r.last_nleap;
        if (v.counters) {
            atomicAdd((unsigned long long*)&v.counters[0], (unsigned long long)r.total_leap);
            atomicAdd((unsigned long long*)&v.counters[1], (unsigned long long)(r.t - r.t_base));
        }
    }
}

// resume: refresh only what the caller may have changed between calls (nothing) and clear per-call counters
template <typename T>
__global__ void resume_kernel(EngineView<T> v) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= v.C) return;
    v.rec[c].total_leap = 0;
    v.rec[c].t_base = v.rec[c].t;
    if (v.rec[c].phase == PH_DONE) v.rec[c].phase = PH_START;
}

// ---------------------------------------------------------------------------
// fused persistent kernel
// ---------------------------------------------------------------------------
template <int G>
struct Geo {
    static constexpr int kThreads = G > 32 ? G : 128;
    static constexpr int kChainsPerBlock = G > 32 ? 1 : 128 / G;
    // the per-tick (split) kernels are latency-bound streams: cap registers at 64 for 50% occupancy
    static constexpr int kMinBlocksSplit = G > 32 ? 4 : 8;
    // the persistent fused kernel is latency / instruction-fetch bound: favour resident warps over registers
    static constexpr int kMinBlocksFused = G > 32 ? 3 : 5;
    __device__ static int chain() {
        return G > 32 ? (int)blockIdx.x : (int)(blockIdx.x * kChainsPerBlock + threadIdx.x / G);
    }
    static int grid(int C) { return (C + kChainsPerBlock - 1) / kChainsPerBlock; }
};

// E > 0: the integration front stays in registers (E elements per lane) between sub-tree boundaries.
template <typename T, int E> struct FrontOf { typedef RegFront<T, E> type; };
template <typename T> struct FrontOf<T, 0> { typedef MemFront<T> type; };

template <typename T, int G, int MODEL, bool HMC, int E>
__global__ void __launch_bounds__(Geo<G>::kThreads, Geo<G>::kMinBlocksFused)
fused_run_kernel(EngineView<T> v, ModelDev m, i64 max_ticks) {
    typedef typename FrontOf<T, E>::type Front;
    __shared__ double red_s[128];
    const int c = Geo<G>::chain();
    if (c >= v.C) return;
    Chain<T, G> ch(v, c, red_s);
    ch.load();
    Front f;
    bool bound = false;
    i64 tick = 0;
    while (max_ticks <= 0 || tick < max_ticks) {
        if (ch.r.phase == PH_DONE) break;
        if (ch.r.phase == PH_START) {
            if (HMC) hmc_begin<T, G, false>(ch);
            else begin_transition<T, G, false>(ch);
            Group<G>::sync();
            bound = false;
        }
        if (!bound) { f.bind(ch); bound = Front::kRegs; }
        half_kick_drift<T, G, false, false>(ch, f);
        T U;
        if constexpr (Front::kRegs) {
            U = model_grad_front<T, G, MODEL>(m, f, ch.lane, ch.red);
        } else {
            Group<G>::sync();
            U = model_grad<T, G, MODEL>(m, f.Q + ch.base, f.Gd + ch.base, v.sj, ch.lane, ch.red);
            Group<G>::sync();
        }
        bool ended;
        if (HMC) ended = hmc_post<T, G, false, false>(ch, U, f);
        else ended = post_gradient<T, G, false, false>(ch, U, f);
        if (ended) bound = false;
        Group<G>::sync();
        ++tick;
    }
    if (bound) f.flush(ch);                // max_ticks ran out in the middle of a sub-tree
    ch.store();
    if (v.counters && ch.lane == 0) atomicAdd((unsigned long long*)&v.counters[3], (unsigned long long)tick);
}

// ---------------------------------------------------------------------------
// split-mode kernels
// ---------------------------------------------------------------------------
template <typename T, int G, bool DENSE, bool HMC>
__global__ void __launch_bounds__(Geo<G>::kThreads, Geo<G>::kMinBlocksSplit) split_pre_kernel(EngineView<T> v) {
    __shared__ double red_s[128];
    const int c = Geo<G>::chain();
    if (c >= v.C) return;
    Chain<T, G> ch(v, c, red_s);
    ch.load();
    if (ch.r.phase == PH_DONE) return;
    if (ch.r.phase == PH_START) {
        if (HMC) hmc_begin<T, G, DENSE>(ch);
        else begin_transition<T, G, DENSE>(ch);
    }
    MemFront<T> f;
    f.bind(ch);
    half_kick_drift<T, G, DENSE, true>(ch, f);
    ch.store();
}

template <typename T, int G, bool DENSE, bool HMC>
__global__ void __launch_bounds__(Geo<G>::kThreads, Geo<G>::kMinBlocksSplit) split_post_kernel(EngineView<T> v, int* not_done) {
    __shared__ double red_s[128];
    const int c = Geo<G>::chain();
    if (c >= v.C) return;
    Chain<T, G> ch(v, c, red_s);
    ch.load();
    if (ch.r.phase != PH_RUN) return;
    const T U = v.Unew[c];
    MemFront<T> f;
    f.bind(ch);
    if (HMC) hmc_post<T, G, DENSE, true>(ch, U, f);
    else post_gradient<T, G, DENSE, true>(ch, U, f);
    ch.store();
    if (ch.lane == 0) {
        if (ch.r.phase != PH_DONE && not_done) atomicAdd(not_done, 1);
        if (v.counters) atomicAdd((unsigned long long*)&v.counters[3], 1ull);
    }
}

// sum the split-K planes of a rider contraction and scatter the rows to their chains:
// out[list[r]][:] = sum_s part[s][r][:] for r < *count   (two riders per launch: blockIdx.y)
template <typename T>
__global__ void rider_reduce_kernel(const T* part0, const int* count0, const int* list0, T* out0, const T* part1,
                                    const int* count1, const int* list1, T* out1, int nsplit, i64 plane, int d) {
    const T* part = blockIdx.y ? part1 : part0;
    const int* count = blockIdx.y ? count1 : count0;
    const int* list = blockIdx.y ? list1 : list0;
    T* out = blockIdx.y ? out1 : out0;
    const int r = blockIdx.x;
    if (r >= *count) return;
    const i64 dst = (i64)list[r] * d;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        T s = 0;
        for (int k = 0; k < nsplit; ++k) s += part[(i64)k * plane + (i64)r * d + j];
        out[dst + j] = s;
    }
}

// dense metric momentum at the start of a run: normals of every chain's first transition (row c of mom_z)
template <typename T>
__global__ void mom_init_kernel(EngineView<T> v) {
    const int c = blockIdx.x;
    const int t = v.rec[c].t;
    for (int j = threadIdx.x; j < v.d; j += blockDim.x) v.mom_z[(i64)c * v.d + j] = (T)draw_z(v.rng, c, t, j, v.d);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static ModelDev to_dev(const b2h_model* m) {
    ModelDev d;
    d.kind = m->kind; d.dim = m->dim; d.n_data = m->n_data;
    d.a = m->a; d.b = m->b; d.c = m->c; d.s0 = m->s0; d.s1 = m->s1;
    return d;
}

template <typename T, int G, bool HMC, int MODEL, int E>
static void launch_fused_e(cudaStream_t st, const EngineView<T>& v, const ModelDev& m, i64 max_ticks) {
    fused_run_kernel<T, G, MODEL, HMC, E><<<Geo<G>::grid(v.C), Geo<G>::kThreads, 0, st>>>(v, m, max_ticks);
}

// Register front when the chain's row fits 2, 4 or 8 elements per lane (B2H_REG_FRONT=0 disables it).
// Measured on B200 (benchmarks/workloads.py): HMC keeps its whole trajectory in registers (c1: 0.78 -> 1.67 G
// evals/s) and the funnel gains 33 %; NUTS on wide elementwise targets is instruction/latency bound, not memory
// bound, and loses 10 % to the extra register pressure -- it keeps the memory front.
template <typename T, int G, bool HMC, int MODEL>
static void launch_fused_model(cudaStream_t st, const EngineView<T>& v, const ModelDev& m, i64 max_ticks) {
    static int use_regs = -1;
    if (use_regs < 0) { const char* e = getenv("B2H_REG_FRONT"); use_regs = e ? atoi(e) : 1; }
    const int epl = (v.d + G - 1) / G;
    if constexpr (HMC || MODEL != MODEL_IID) {
        if (use_regs && epl <= 2) return launch_fused_e<T, G, HMC, MODEL, 2>(st, v, m, max_ticks);
        if (use_regs && epl <= 4) return launch_fused_e<T, G, HMC, MODEL, 4>(st, v, m, max_ticks);
        if (use_regs && epl <= 8) return launch_fused_e<T, G, HMC, MODEL, 8>(st, v, m, max_ticks);
    }
    (void)epl;
    launch_fused_e<T, G, HMC, MODEL, 0>(st, v, m, max_ticks);
}

template <typename T, int G, bool HMC>
static int launch_fused(cudaStream_t st, const EngineView<T>& v, const b2h_model* model, i64 max_ticks) {
    ModelDev m = to_dev(model);
    constexpr int GS = G > 8 ? 8 : G;          // funnel / eight schools are instantiated for 1 and 8 lanes only
    switch (model->kind) {
        case B2H_MODEL_IID_GAUSSIAN:
            launch_fused_model<T, G, HMC, MODEL_IID>(st, v, m, max_ticks);
            break;
        case B2H_MODEL_FUNNEL:
            if (G > 8) { set_error("funnel: group must be 1 or 8"); return B2H_ERR_ARG; }
            launch_fused_model<T, GS, HMC, MODEL_FUNNEL>(st, v, m, max_ticks);
            break;
        case B2H_MODEL_EIGHT_SCHOOLS:
            if (G > 8) { set_error("eight schools: group must be 1 or 8"); return B2H_ERR_ARG; }
            launch_fused_model<T, GS, HMC, MODEL_SCHOOLS>(st, v, m, max_ticks);
            break;
        default:
            set_error("model has no fused gradient");
            return B2H_ERR_UNSUPPORTED;
    }
    B2H_LAUNCH_CHECK();
    return 0;
}

template <typename T, int G, bool HMC>
static int run_split(b2h_ctx* ctx, EngineView<T>& v, const EnginePlan& pl, const b2h_model* model,
                     const b2h_metric* metric, const b2h_cfg* cfg, i64 max_ticks, int n_transitions, void* model_ws,
                     i64 model_ws_bytes, int* not_done_dev, int resume) {
    cudaStream_t st = ctx->stream;
    const int grid = Geo<G>::grid(v.C), thr = Geo<G>::kThreads;
    const int C = v.C, d = v.d;
    const T* imm_dense = (const T*)metric->imm;
    const T* sqrt_t = (const T*)metric->sqrt_t;
    i64 bound = max_ticks > 0 ? max_ticks
                              : (i64)n_transitions * (HMC ? (i64)cfg->num_integration_steps
                                                          : (((i64)1 << v.maxd) - 1 + v.maxd)) + 1;
    int* host_flag = ctx->host_flag;
    int rc = 0;
    GemmGroup<T> none{nullptr, 0, nullptr, 0, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr};
    int last_parity = 0;
    if (pl.dense) {
        B2H_CUDA(cudaMemsetAsync(v.mom_count, 0, 4 * sizeof(int), st));
        if (!resume) {
            // p0 = z . S^T (metrics.py:56-59,67), v0 = p0 . imm (metrics.py:71) for every chain's first transition,
            // and w = imm . g of the starting positions
            mom_init_kernel<T><<<C, 128, 0, st>>>(v);
            launch_dense_apply<T>(st, v.mom_z, sqrt_t, v.mom_p, C, d, d, nullptr, nullptr);
            launch_dense_apply<T>(st, v.mom_p, imm_dense, v.mom_v, C, d, d, nullptr, nullptr);
            launch_dense_apply<T>(st, v.gp, imm_dense, v.wp, C, d, d, nullptr, nullptr);
        }
    }
    bool side_pending[2] = {false, false};
    static int use_side = -1;
    if (use_side < 0) { const char* e = getenv("B2H_SIDE_STREAM"); use_side = e ? atoi(e) : 1; }
    cudaStream_t rider_stream = use_side ? ctx->side : st;
    for (i64 tick = 0; tick < bound; ++tick) {
        const int b = (int)(tick & 1);
        if (pl.dense) {
            last_parity = b;
            v.mom_parity = b;
            // all momentum contractions launched so far must have landed: a chain that started a transition two
            // ticks ago may start the next one now (its v0 came from the previous tick's side launch), and this
            // parity's request list is about to be reused
            for (int k = 0; k < 2; ++k)
                if (side_pending[k]) { B2H_CUDA(cudaStreamWaitEvent(st, ctx->ev_side[k], 0)); side_pending[k] = false; }
            B2H_CUDA(cudaMemsetAsync(v.mom_count + b, 0, sizeof(int), st));
            split_pre_kernel<T, G, true, HMC><<<grid, thr, 0, st>>>(v);      // half kick + drift by the v/w recurrence
            // side stream: p0 = z . S^T of the transitions queued by this pre kernel, and v0 = imm . p0 of the
            // transitions queued one tick ago (their p0 was produced by the previous side launch)
            if (use_side) {
                B2H_CUDA(cudaEventRecord(ctx->ev_pre[b], st));
                B2H_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_pre[b], 0));
            }
            // few rows, full reduction length: split K so that the tiles spread over all SMs, then reduce + scatter
            const i64 plane = (i64)C * d;
            T* part1 = v.mom_part;
            T* part2 = v.mom_part + (size_t)kRiderSplit * plane;
            GemmGroup<T> g1{v.mom_z + (size_t)b * C * d, (i64)d, sqrt_t, (i64)d, part1, (i64)d, C, v.mom_count + b,
                            nullptr, nullptr, nullptr};
            GemmGroup<T> g2{v.mom_p, (i64)d, imm_dense, (i64)d, part2, (i64)d, C, v.mom_count + (b ^ 1), nullptr,
                            v.mom_list + (size_t)(b ^ 1) * C, nullptr};
            launch_gemm_grouped<T>(rider_stream, g1, g2, none, d, d, kRiderSplit, plane, 0);
            rider_reduce_kernel<T><<<dim3(C, 2), 128, 0, rider_stream>>>(
                part1, v.mom_count + b, v.mom_list + (size_t)b * C, v.mom_p, part2, v.mom_count + (b ^ 1),
                v.mom_list + (size_t)(b ^ 1) * C, v.mom_v, kRiderSplit, plane, d);
            if (use_side) {
                B2H_CUDA(cudaEventRecord(ctx->ev_side[b], ctx->side));
                side_pending[b] = true;
            }
        } else {
            split_pre_kernel<T, G, false, HMC><<<grid, thr, 0, st>>>(v);
        }
        rc = potential_and_grad_impl<T>(ctx, model, v.xa, v.Unew, v.xb, C, model_ws, model_ws_bytes);
        if (rc) break;
        const bool check = (max_ticks <= 0) && ((tick & 3) == 3 || tick + 1 == bound);
        if (check) B2H_CUDA(cudaMemsetAsync(not_done_dev, 0, sizeof(int), st));
        if (pl.dense) {
            // the tick's only metric contraction on the main stream: w' = imm . g'
            launch_dense_apply<T>(st, v.xb, imm_dense, v.xc, C, d, d, nullptr, nullptr);
            split_post_kernel<T, G, true, HMC><<<grid, thr, 0, st>>>(v, check ? not_done_dev : nullptr);
        } else {
            split_post_kernel<T, G, false, HMC><<<grid, thr, 0, st>>>(v, check ? not_done_dev : nullptr);
        }
        if (check) {
            cudaError_t e = cudaMemcpyAsync(host_flag, not_done_dev, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { rc = cuda_fail(e, "tick poll"); break; }
            if (*host_flag == 0) break;
        }
    }
    if (pl.dense && rc == 0) {
        // join the side stream, then flush: v0 of the transitions queued in the last tick, so that a resumed run
        // starts with no request pending
        for (int b = 0; b < 2; ++b)
            if (side_pending[b]) B2H_CUDA(cudaStreamWaitEvent(st, ctx->ev_side[b], 0));
        GemmGroup<T> g0{v.mom_p, (i64)d, imm_dense, (i64)d, v.mom_v, (i64)d, C, v.mom_count + last_parity, nullptr,
                        v.mom_list + (size_t)last_parity * C, v.mom_list + (size_t)last_parity * C};
        launch_gemm_grouped<T>(st, g0, none, none, d, d, 1, 0, 0);
    }
    if (rc) return rc;
    B2H_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int run_typed(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng,
                     const b2h_cfg* cfg, const b2h_adapt* adapt, void* q, void* p, void* U, void* g,
                     double* step_size, i64 C64, int n_transitions, i64 max_ticks, int resume, b2h_diag* diag,
                     void* draws, double* draw_stats, int n_store, int64_t* counters, void* ws, i64 ws_bytes,
                     bool hmc) {
    EnginePlan pl;
    int rc = make_plan(model, metric, cfg, pl, C64);
    if (rc) return rc;
    const int C = (int)C64, d = model->dim;
    const int maxd = hmc ? 1 : cfg->max_num_expansions;
    if (!hmc && (maxd < 1 || maxd > 24)) { set_error("max_num_expansions must be in [1, 24]"); return B2H_ERR_ARG; }
    const bool adapting = adapt && adapt->enabled;
    if (adapting && !pl.per_chain_imm) {
        set_error("window adaptation needs a DIAG_PER_CHAIN inverse mass matrix (adapted in place)");
        return B2H_ERR_ARG;
    }
    if (pl.dense && (!metric->imm || !metric->sqrt_t)) { set_error("dense metric needs imm and sqrt_t"); return B2H_ERR_ARG; }
    if (rng->mode == B2H_RNG_INJECTED) {
        if (!rng->z || rng->n_injected < n_transitions || n_transitions <= 0) {
            set_error("injected draws need z and n_injected >= n_transitions > 0");
            return B2H_ERR_ARG;
        }
        if (!hmc && (!rng->u_dir || !rng->u_biased || !rng->u_uniform)) { set_error("NUTS needs u_dir/u_biased/u_uniform"); return B2H_ERR_ARG; }
        if (hmc && !rng->u_accept) { set_error("HMC needs u_accept"); return B2H_ERR_ARG; }
    }

    EngineView<T> v;
    memset(&v, 0, sizeof(v));
    size_t model_ws_bytes = pl.split ? (size_t)potential_workspace_bytes_impl(model, Num<T>::dtype, C) : 0;
    size_t model_ws_off = 0;
    size_t need = carve<T>(v, nullptr, pl, model, C, d, maxd, adapting, model_ws_bytes, &model_ws_off);
    if (!ws || (size_t)ws_bytes < need) {
        set_error("workspace too small: need " + std::to_string(need) + " bytes");
        return B2H_ERR_WORKSPACE;
    }
    carve<T>(v, (char*)ws, pl, model, C, d, maxd, adapting, model_ws_bytes, &model_ws_off);
    void* model_ws = model_ws_bytes ? (char*)ws + model_ws_off : nullptr;
    v.C = C; v.d = d; v.maxd = maxd;
    if (pl.G == 1) { v.sc = 1; v.sj = C; v.sck = 1; }
    else { v.sc = d; v.sj = 1; v.sck = (i64)maxd * d; }
    v.imm_kind = metric->kind;
    if (pl.scalar_imm) { v.imm_sc = 0; v.imm_sj = 0; }
    else if (metric->kind == B2H_IMM_DIAG) { v.imm = (T*)metric->imm; v.imm_sc = 0; v.imm_sj = 1; }
    else if (pl.per_chain_imm) { v.imm_sc = v.sc; v.imm_sj = v.sj; }
    v.rng.mode = rng->mode;
    v.rng.key.k0 = (uint32_t)rng->seed; v.rng.key.k1 = (uint32_t)(rng->seed >> 32);
    v.rng.chain_offset = rng->chain_offset; v.rng.transition_offset = rng->transition_offset;
    v.rng.n_injected = rng->n_injected;
    v.rng.z = rng->z; v.rng.u_dir = rng->u_dir; v.rng.u_biased = rng->u_biased; v.rng.u_uniform = rng->u_uniform;
    v.rng.u_accept = rng->u_accept;
    v.adapt.enabled = adapting ? 1 : 0;
    if (adapting) {
        v.adapt.num_steps = adapt->num_steps; v.adapt.stage = adapt->stage; v.adapt.window_end = adapt->window_end;
        v.adapt.target = adapt->target_acceptance_rate; v.adapt.gamma = adapt->gamma; v.adapt.t0 = adapt->t0;
        v.adapt.kappa = adapt->kappa;
        v.adapt.da_step = (i64*)adapt->da_step; v.adapt.da_x = adapt->da_x; v.adapt.da_x_avg = adapt->da_x_avg;
        v.adapt.da_g_avg = adapt->da_g_avg; v.adapt.da_mu = adapt->da_mu; v.adapt.wc_n = (i64*)adapt->wc_n;
    }
    v.out.draws = draws; v.out.draw_stats = draw_stats; v.out.n_store = n_store;
    if (diag) {
        v.out.acceptance_probability = diag->acceptance_probability; v.out.num_doublings = diag->num_doublings;
        v.out.is_turning = diag->is_turning; v.out.is_diverging = diag->is_diverging;
        v.out.n_leapfrog = diag->n_leapfrog;
    }
    v.div_thr = cfg->divergence_threshold;
    v.n_transitions = max_ticks > 0 ? 0 : n_transitions;
    v.hmc_L = cfg->num_integration_steps;
    v.counters = (i64*)counters;
    if (hmc && v.hmc_L < 1) { set_error("num_integration_steps must be >= 1"); return B2H_ERR_ARG; }
    if (max_ticks <= 0 && n_transitions <= 0) { set_error("need n_transitions > 0 or max_ticks > 0"); return B2H_ERR_ARG; }

    cudaStream_t st = ctx->stream;
    const i64 n = (i64)C * d;
    const int eb = 256, eg = (int)((n + eb - 1) / eb);
    if (!resume) {
        copy_in_kernel<T><<<eg, eb, 0, st>>>(v, (const T*)q, (const T*)g, (const T*)U, step_size,
                                             pl.per_chain_imm ? (const T*)metric->imm : nullptr,
                                             (T)metric->scalar, pl.scalar_imm ? 1 : 0,
                                             adapting ? adapt->initial_step_size : 0.0);
    } else {
        resume_kernel<T><<<(C + 255) / 256, 256, 0, st>>>(v);
    }
    B2H_LAUNCH_CHECK();
    int* not_done_dev = v.scratch;

#define B2H_DISPATCH_G(FN, ...)                                          \
    switch (pl.G) {                                                      \
        case 1: rc = FN<T, 1, false> __VA_ARGS__; break;                 \
        case 8: rc = FN<T, 8, false> __VA_ARGS__; break;                 \
        case 32: rc = FN<T, 32, false> __VA_ARGS__; break;               \
        default: rc = FN<T, 256, false> __VA_ARGS__; break;              \
    }
#define B2H_DISPATCH_G_HMC(FN, ...)                                      \
    switch (pl.G) {                                                      \
        case 1: rc = FN<T, 1, true> __VA_ARGS__; break;                  \
        case 8: rc = FN<T, 8, true> __VA_ARGS__; break;                  \
        case 32: rc = FN<T, 32, true> __VA_ARGS__; break;                \
        default: rc = FN<T, 256, true> __VA_ARGS__; break;               \
    }

    if (!pl.split) {
        if (hmc) { B2H_DISPATCH_G_HMC(launch_fused, (st, v, model, max_ticks)) }
        else { B2H_DISPATCH_G(launch_fused, (st, v, model, max_ticks)) }
    } else {
        if (pl.G == 1) { set_error("split mode needs group >= 8"); return B2H_ERR_ARG; }
        if (hmc) {
            switch (pl.G) {
                case 8: rc = run_split<T, 8, true>(ctx, v, pl, model, metric, cfg, max_ticks, n_transitions, model_ws, model_ws_bytes, not_done_dev, resume); break;
                case 32: rc = run_split<T, 32, true>(ctx, v, pl, model, metric, cfg, max_ticks, n_transitions, model_ws, model_ws_bytes, not_done_dev, resume); break;
                default: rc = run_split<T, 256, true>(ctx, v, pl, model, metric, cfg, max_ticks, n_transitions, model_ws, model_ws_bytes, not_done_dev, resume); break;
            }
        } else {
            switch (pl.G) {
                case 8: rc = run_split<T, 8, false>(ctx, v, pl, model, metric, cfg, max_ticks, n_transitions, model_ws, model_ws_bytes, not_done_dev, resume); break;
                case 32: rc = run_split<T, 32, false>(ctx, v, pl, model, metric, cfg, max_ticks, n_transitions, model_ws, model_ws_bytes, not_done_dev, resume); break;
                default: rc = run_split<T, 256, false>(ctx, v, pl, model, metric, cfg, max_ticks, n_transitions, model_ws, model_ws_bytes, not_done_dev, resume); break;
            }
        }
    }
    if (rc) return rc;
    copy_out_kernel<T><<<eg, eb, 0, st>>>(v, (T*)q, (T*)p, (T*)g, (T*)U, step_size,
                                          pl.per_chain_imm ? (T*)metric->imm : nullptr);
    B2H_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// stand-alone trajectory builders (caller-supplied tree state): the same state machine, entered after
// begin_transition (expand) or for exactly one sub-tree (integrate)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void tree_in_kernel(EngineView<T> v, b2h_tree t, const double* eps) {
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0 && v.imm_kind == B2H_IMM_SCALAR) {}
    if (idx >= (i64)v.C * v.d) return;
    int c = (int)(idx / v.d), j = (int)(idx % v.d);
    i64 a = (i64)c * v.sc + (i64)j * v.sj;
    v.qp[a] = ((const T*)t.proposal.q)[idx]; v.pp[a] = ((const T*)t.proposal.p)[idx]; v.gp[a] = ((const T*)t.proposal.g)[idx];
    v.ql[a] = ((const T*)t.left.q)[idx]; v.pl[a] = ((const T*)t.left.p)[idx]; v.gl[a] = ((const T*)t.left.g)[idx];
    v.qr[a] = ((const T*)t.right.q)[idx]; v.pr[a] = ((const T*)t.right.p)[idx]; v.gr[a] = ((const T*)t.right.g)[idx];
    v.msum[a] = ((const T*)t.momentum_sum)[idx];
    for (int l = 0; l < v.maxd; ++l) {
        i64 b = (i64)c * v.sck + ((i64)l * v.d + j) * v.sj;
        v.mck[b] = ((const T*)t.momentum_ckpts)[((i64)c * v.maxd + l) * v.d + j];
        v.sckp[b] = ((const T*)t.momentum_sum_ckpts)[((i64)c * v.maxd + l) * v.d + j];
    }
    if (j == 0) {
        ChainRec r;
        memset(&r, 0, sizeof(r));
        r.phase = PH_RUN;
        r.eps = eps[c];
        r.E0 = (double)((const T*)t.initial_energy)[c];
        r.U_prop = (double)((const T*)t.proposal.U)[c];
        r.E_prop = (double)((const T*)t.proposal_energy)[c];
        r.w_prop = t.proposal_weight[c];
        r.slpa_prop = t.proposal_slpa[c];
        r.U_left = (double)((const T*)t.left.U)[c];
        r.U_right = (double)((const T*)t.right.U)[c];
        r.imin = (int)t.idx_min[c]; r.imax = (int)t.idx_max[c];
        r.k = 0; r.s = 0;
        double u = draw_u(v.rng, DRAW_DIR, c, 0, 0, v.maxd);          // trajectory.py:516 for expansion 0
        r.go_right = bern(u, 0.5) ? 1 : 0;
        v.rec[c] = r;
    }
}

template <typename T>
__global__ void tree_out_kernel(EngineView<T> v, b2h_tree t) {
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (i64)v.C * v.d) return;
    int c = (int)(idx / v.d), j = (int)(idx % v.d);
    i64 a = (i64)c * v.sc + (i64)j * v.sj;
    ((T*)t.proposal.q)[idx] = v.qp[a]; ((T*)t.proposal.p)[idx] = v.pp[a]; ((T*)t.proposal.g)[idx] = v.gp[a];
    ((T*)t.left.q)[idx] = v.ql[a]; ((T*)t.left.p)[idx] = v.pl[a]; ((T*)t.left.g)[idx] = v.gl[a];
    ((T*)t.right.q)[idx] = v.qr[a]; ((T*)t.right.p)[idx] = v.pr[a]; ((T*)t.right.g)[idx] = v.gr[a];
    ((T*)t.momentum_sum)[idx] = v.msum[a];
    for (int l = 0; l < v.maxd; ++l) {
        i64 b = (i64)c * v.sck + ((i64)l * v.d + j) * v.sj;
        ((T*)t.momentum_ckpts)[((i64)c * v.maxd + l) * v.d + j] = v.mck[b];
        ((T*)t.momentum_sum_ckpts)[((i64)c * v.maxd + l) * v.d + j] = v.sckp[b];
    }
    if (j == 0) {
        const ChainRec& r = v.rec[c];
        ((T*)t.proposal.U)[c] = (T)r.U_prop;
        ((T*)t.proposal_energy)[c] = (T)r.E_prop;
        t.proposal_weight[c] = r.w_prop;
        t.proposal_slpa[c] = r.slpa_prop;
        ((T*)t.left.U)[c] = (T)r.U_left;
        ((T*)t.right.U)[c] = (T)r.U_right;
        t.idx_min[c] = r.imin; t.idx_max[c] = r.imax;
        if (v.out.acceptance_probability) v.out.acceptance_probability[c] = r.accept_prob;
        if (v.out.num_doublings) v.out.num_doublings[c] = r.last_nd;
        if (v.out.is_turning) v.out.is_turning[c] = (r.last_flags & 1) ? 1 : 0;
        if (v.out.is_diverging) v.out.is_diverging[c] = (r.last_flags & 2) ? 1 : 0;
        if (v.out.n_leapfrog) v.out.n_leapfrog[c] = r.last_nleap;
    }
}

template <typename T>
__global__ void subtree_in_kernel(EngineView<T> v, b2h_subtree t, const double* eps) {
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (i64)v.C * v.d) return;
    int c = (int)(idx / v.d), j = (int)(idx % v.d);
    i64 a = (i64)c * v.sc + (i64)j * v.sj;
    T q = ((const T*)t.state.q)[idx], p = ((const T*)t.state.p)[idx], g = ((const T*)t.state.g)[idx];
    v.ql[a] = q; v.pl[a] = p; v.gl[a] = g;
    v.qr[a] = q; v.pr[a] = p; v.gr[a] = g;
    for (int l = 0; l < v.maxd; ++l) {
        i64 b = (i64)c * v.sck + ((i64)l * v.d + j) * v.sj;
        v.mck[b] = ((const T*)t.momentum_ckpts)[((i64)c * v.maxd + l) * v.d + j];
        v.sckp[b] = ((const T*)t.momentum_sum_ckpts)[((i64)c * v.maxd + l) * v.d + j];
    }
    if (j == 0) {
        ChainRec r;
        memset(&r, 0, sizeof(r));
        r.phase = PH_RUN;
        r.eps = eps[c];
        r.E0 = (double)((const T*)t.initial_energy)[c];
        r.imin = (int)t.idx_min[c]; r.imax = (int)t.idx_max[c];
        r.k = t.expansion; r.s = 0;
        r.go_right = t.direction[c] > 0 ? 1 : 0;
        v.rec[c] = r;
    }
}

template <typename T>
__global__ void subtree_out_kernel(EngineView<T> v, b2h_subtree t) {
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (i64)v.C * v.d) return;
    int c = (int)(idx / v.d), j = (int)(idx % v.d);
    i64 a = (i64)c * v.sc + (i64)j * v.sj;
    const ChainRec& r = v.rec[c];
    const T* Q = r.go_right ? v.qr : v.ql; const T* P = r.go_right ? v.pr : v.pl; const T* Gd = r.go_right ? v.gr : v.gl;
    ((T*)t.state.q)[idx] = Q[a]; ((T*)t.state.p)[idx] = P[a]; ((T*)t.state.g)[idx] = Gd[a];
    ((T*)t.proposal.q)[idx] = v.qs[a]; ((T*)t.proposal.p)[idx] = v.ps[a]; ((T*)t.proposal.g)[idx] = v.gs[a];
    ((T*)t.momentum_sum)[idx] = v.sms[a];
    for (int l = 0; l < v.maxd; ++l) {
        i64 b = (i64)c * v.sck + ((i64)l * v.d + j) * v.sj;
        ((T*)t.momentum_ckpts)[((i64)c * v.maxd + l) * v.d + j] = v.mck[b];
        ((T*)t.momentum_sum_ckpts)[((i64)c * v.maxd + l) * v.d + j] = v.sckp[b];
    }
    if (j == 0) {
        ((T*)t.state.U)[c] = (T)r.U_front;
        ((T*)t.proposal.U)[c] = (T)r.U_sub;
        ((T*)t.proposal_energy)[c] = (T)r.E_sub;
        t.proposal_weight[c] = r.w_sub;
        t.proposal_slpa[c] = r.slpa_sub;
        t.idx_min[c] = r.imin; t.idx_max[c] = r.imax;
        t.trajectory_length[c] = r.sub_len;
        t.is_diverging[c] = (r.last_flags & 2) ? 1 : 0;
        t.has_terminated[c] = (r.last_flags & 4) ? 1 : 0;
    }
}

// shared set-up of the two entry points: fused models, diagonal-family metrics
template <typename T>
static int tree_setup(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng, const b2h_cfg* cfg,
                      i64 C64, void* ws, i64 ws_bytes, EngineView<T>& v, EnginePlan& pl) {
    int rc = make_plan(model, metric, cfg, pl, C64);
    if (rc) return rc;
    if (pl.split) { set_error("stand-alone trajectory builders support the fused models with scalar/diagonal metrics"); return B2H_ERR_UNSUPPORTED; }
    const int C = (int)C64, d = model->dim, maxd = cfg->max_num_expansions;
    if (maxd < 1 || maxd > 24) { set_error("max_num_expansions must be in [1, 24]"); return B2H_ERR_ARG; }
    memset(&v, 0, sizeof(v));
    size_t need = carve<T>(v, nullptr, pl, model, C, d, maxd, false, 0, nullptr);
    if (!ws || (size_t)ws_bytes < need) { set_error("workspace too small: need " + std::to_string(need) + " bytes"); return B2H_ERR_WORKSPACE; }
    carve<T>(v, (char*)ws, pl, model, C, d, maxd, false, 0, nullptr);
    v.C = C; v.d = d; v.maxd = maxd;
    if (pl.G == 1) { v.sc = 1; v.sj = C; v.sck = 1; }
    else { v.sc = d; v.sj = 1; v.sck = (i64)maxd * d; }
    v.imm_kind = metric->kind;
    if (pl.scalar_imm) {
        v.imm_sc = 0; v.imm_sj = 0;
        T val = (T)metric->scalar;
        B2H_CUDA(cudaMemcpyAsync(v.imm, &val, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        B2H_CUDA(cudaStreamSynchronize(ctx->stream));
    } else if (metric->kind == B2H_IMM_DIAG) { v.imm = (T*)metric->imm; v.imm_sc = 0; v.imm_sj = 1; }
    else { set_error("per-chain metrics are not supported here"); return B2H_ERR_UNSUPPORTED; }
    v.rng.mode = rng->mode;
    v.rng.key.k0 = (uint32_t)rng->seed; v.rng.key.k1 = (uint32_t)(rng->seed >> 32);
    v.rng.chain_offset = rng->chain_offset; v.rng.transition_offset = rng->transition_offset;
    v.rng.n_injected = rng->n_injected;
    v.rng.z = rng->z; v.rng.u_dir = rng->u_dir; v.rng.u_biased = rng->u_biased; v.rng.u_uniform = rng->u_uniform;
    v.rng.u_accept = rng->u_accept;
    v.div_thr = cfg->divergence_threshold;
    v.n_transitions = 1;
    return 0;
}

template <typename T>
static int expand_typed(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng,
                        const b2h_cfg* cfg, b2h_tree* tree, const double* eps, i64 C, b2h_diag* diag, void* ws, i64 ws_bytes) {
    EngineView<T> v;
    EnginePlan pl;
    int rc = tree_setup<T>(ctx, model, metric, rng, cfg, C, ws, ws_bytes, v, pl);
    if (rc) return rc;
    if (diag) {
        v.out.acceptance_probability = diag->acceptance_probability; v.out.num_doublings = diag->num_doublings;
        v.out.is_turning = diag->is_turning; v.out.is_diverging = diag->is_diverging; v.out.n_leapfrog = diag->n_leapfrog;
    }
    cudaStream_t st = ctx->stream;
    const i64 n = C * model->dim;
    const int eb = 256, eg = (int)((n + eb - 1) / eb);
    tree_in_kernel<T><<<eg, eb, 0, st>>>(v, *tree, eps);
    switch (pl.G) {
        case 1: rc = launch_fused<T, 1, false>(st, v, model, 0); break;
        case 8: rc = launch_fused<T, 8, false>(st, v, model, 0); break;
        case 32: rc = launch_fused<T, 32, false>(st, v, model, 0); break;
        default: rc = launch_fused<T, 256, false>(st, v, model, 0); break;
    }
    if (rc) return rc;
    tree_out_kernel<T><<<eg, eb, 0, st>>>(v, *tree);
    B2H_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int subtree_typed(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng,
                         const b2h_cfg* cfg, b2h_subtree* sub, const double* eps, i64 C, void* ws, i64 ws_bytes) {
    EngineView<T> v;
    EnginePlan pl;
    int rc = tree_setup<T>(ctx, model, metric, rng, cfg, C, ws, ws_bytes, v, pl);
    if (rc) return rc;
    if (sub->max_num_steps < 1 || sub->expansion < 0 ||
        ((i64)1 << sub->expansion) - 1 + sub->max_num_steps > ((i64)1 << cfg->max_num_expansions) - 1) {
        set_error("sub-tree: the uniform-draw slots 2**expansion - 1 + step must fit 2**max_num_expansions - 1");
        return B2H_ERR_ARG;
    }
    v.sub_max_steps = sub->max_num_steps;
    v.stop_at_subtree_end = 1;
    cudaStream_t st = ctx->stream;
    const i64 n = C * model->dim;
    const int eb = 256, eg = (int)((n + eb - 1) / eb);
    subtree_in_kernel<T><<<eg, eb, 0, st>>>(v, *sub, eps);
    switch (pl.G) {
        case 1: rc = launch_fused<T, 1, false>(st, v, model, 0); break;
        case 8: rc = launch_fused<T, 8, false>(st, v, model, 0); break;
        case 32: rc = launch_fused<T, 32, false>(st, v, model, 0); break;
        default: rc = launch_fused<T, 256, false>(st, v, model, 0); break;
    }
    if (rc) return rc;
    subtree_out_kernel<T><<<eg, eb, 0, st>>>(v, *sub);
    B2H_LAUNCH_CHECK();
    return 0;
}

int nuts_expand_impl(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng,
                     const b2h_cfg* cfg, b2h_tree* tree, const double* eps, i64 C, b2h_diag* diag, void* ws, i64 ws_bytes) {
    if (!ctx || !model || !metric || !rng || !cfg || !tree || !eps) { set_error("null argument"); return B2H_ERR_ARG; }
    if (cfg->dtype == B2H_F64) return expand_typed<double>(ctx, model, metric, rng, cfg, tree, eps, C, diag, ws, ws_bytes);
    if (cfg->dtype == B2H_F32) return expand_typed<float>(ctx, model, metric, rng, cfg, tree, eps, C, diag, ws, ws_bytes);
    set_error("bad dtype");
    return B2H_ERR_ARG;
}

int nuts_subtree_impl(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng,
                      const b2h_cfg* cfg, b2h_subtree* sub, const double* eps, i64 C, void* ws, i64 ws_bytes) {
    if (!ctx || !model || !metric || !rng || !cfg || !sub || !eps) { set_error("null argument"); return B2H_ERR_ARG; }
    if (cfg->dtype == B2H_F64) return subtree_typed<double>(ctx, model, metric, rng, cfg, sub, eps, C, ws, ws_bytes);
    if (cfg->dtype == B2H_F32) return subtree_typed<float>(ctx, model, metric, rng, cfg, sub, eps, C, ws, ws_bytes);
    set_error("bad dtype");
    return B2H_ERR_ARG;
}

int nuts_run_impl(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng,
                  const b2h_cfg* cfg, const b2h_adapt* adapt, void* q, void* p, void* U, void* g, double* step_size,
                  i64 C, int n_transitions, i64 max_ticks, int resume, b2h_diag* diag, void* draws,
                  double* draw_stats, int n_store, int64_t* counters, void* ws, i64 ws_bytes, bool hmc) {
    if (!ctx || !model || !metric || !rng || !cfg || !q || !U || !g || !step_size) {
        set_error("null argument");
        return B2H_ERR_ARG;
    }
    if (C <= 0 || C > (1ll << 30) || model->dim <= 0) { set_error("bad C or dim"); return B2H_ERR_ARG; }
    if (cfg->dtype == B2H_F64)
        return run_typed<double>(ctx, model, metric, rng, cfg, adapt, q, p, U, g, step_size, C, n_transitions,
                                 max_ticks, resume, diag, draws, draw_stats, n_store, counters, ws, ws_bytes, hmc);
    if (cfg->dtype == B2H_F32)
        return run_typed<float>(ctx, model, metric, rng, cfg, adapt, q, p, U, g, step_size, C, n_transitions,
                                max_ticks, resume, diag, draws, draw_stats, n_store, counters, ws, ws_bytes, hmc);
    set_error("dtype must be B2H_F32 or B2H_F64");
    return B2H_ERR_ARG;
}

i64 engine_workspace_bytes(const b2h_model* model, const b2h_metric* metric, const b2h_cfg* cfg, i64 C) {
    EnginePlan pl;
    if (make_plan(model, metric, cfg, pl, C)) return -1;
    const int maxd = cfg->max_num_expansions > 0 ? cfg->max_num_expansions : 1;
    size_t mws = pl.split ? (size_t)potential_workspace_bytes_impl(model, cfg->dtype, C) : 0;
    if (cfg->dtype == B2H_F64) {
        EngineView<double> v;
        return (i64)carve<double>(v, nullptr, pl, model, (int)C, model->dim, maxd, true, mws, nullptr);
    }
    EngineView<float> v;
    return (i64)carve<float>(v, nullptr, pl, model, (int)C, model->dim, maxd, true, mws, nullptr);
}

}  // namespace b2h
