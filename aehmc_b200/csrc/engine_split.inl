// The per-tick (split) kernels and their driver loop; included by engine_split_f32.cu / engine_split_f64.cu.
#include <stdlib.h>

#include "engine_host.cuh"
#include "engine_tile.inl"

namespace b2h {

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

// post of tick t and pre of tick t+1 in one pass with the front in registers (TickFront): p (V) of the edge are
// read once and written once per tick, q goes through xa only, g (W) are not written at all inside a sub-tree.
// PRE = false is the last tick of a call: the post part only, then the whole front goes back to the edge arrays so
// that the state in memory is complete (the next call starts with split_pre_kernel).
template <typename T, int G, bool DENSE, bool HMC, int E, bool PRE>
__global__ void __launch_bounds__(Geo<G>::kThreads, Geo<G>::kMinBlocksTick)
split_postpre_kernel(EngineView<T> v, int* not_done) {
    __shared__ double red_s[128];
    const int c = Geo<G>::chain();
    if (c >= v.C) return;
    Chain<T, G> ch(v, c, red_s);
    ch.load();
    if (ch.r.phase == PH_DONE) return;
    TickFront<T, E, DENSE> f;
    bool rebind = true;
    if (ch.r.phase == PH_RUN) {
        f.bind_post(ch);
        const T U = v.Unew[c];
        bool ended;
        if (HMC) ended = hmc_post<T, G, DENSE, true>(ch, U, f);
        else ended = post_gradient<T, G, DENSE, true>(ch, U, f);
        rebind = ended;                        // the front was written back (or abandoned) at a sub-tree end
        if (ch.lane == 0) {
            if (ch.r.phase != PH_DONE && not_done) atomicAdd(not_done, 1);
            if (v.counters) atomicAdd((unsigned long long*)&v.counters[3], 1ull);
        }
        if (!PRE) {
            if (!ended) f.flush(ch);
            ch.store();
            return;
        }
        if (ch.r.phase == PH_DONE) { ch.store(); return; }
    }
    if (!PRE) return;
    if (ch.r.phase == PH_START) {
        if (HMC) hmc_begin<T, G, DENSE>(ch);
        else begin_transition<T, G, DENSE>(ch);
        rebind = true;
    }
    if (rebind) f.bind(ch);
    half_kick_drift<T, G, DENSE, true>(ch, f);
    f.store(ch, false);                        // p (V) only: q is in xa, g (W) stay in the edge / come from xb (xc)
    ch.store();
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
    const int n = *count;
    for (int r = blockIdx.x; r < n; r += gridDim.x) {             // a few hundred of the C rows restart on a tick
        const i64 dst = (i64)list[r] * d;
        for (int j = threadIdx.x; j < d; j += blockDim.x) {
            T s = 0;
            for (int k = 0; k < nsplit; ++k) s += part[(i64)k * plane + (i64)r * d + j];
            out[dst + j] = s;
        }
    }
}

// dense metric momentum at the start of a run: normals of every chain's first transition (row c of mom_z)
template <typename T>
__global__ void mom_init_kernel(EngineView<T> v) {
    const int c = blockIdx.x;
    const int t = v.rec[c].t;
    for (int j = threadIdx.x; j < v.d; j += blockDim.x) v.mom_z[(i64)c * v.d + j] = (T)draw_z(v.rng, c, t, j, v.d);
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
    const T* chol_t = (const T*)metric->chol_t;
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
            // p0 = z . S^T (metrics.py:56-59,67), v0 = imm . p0 = z . L^T (metrics.py:71 with imm = L L^T, S = L^-T)
            // for every chain's first transition, and w = imm . g of the starting positions
            mom_init_kernel<T><<<C, 128, 0, st>>>(v);
            launch_dense_apply<T>(st, v.mom_z, sqrt_t, v.mom_p, C, d, d, nullptr, nullptr);
            launch_dense_apply<T>(st, v.mom_z, chol_t, v.mom_v, C, d, d, nullptr, nullptr);
            launch_dense_apply<T>(st, v.gp, imm_dense, v.wp, C, d, d, nullptr, nullptr);
        }
    }
    bool side_pending[2] = {false, false};
    static int use_side = -1, use_fuse = -1;
    if (use_side < 0) { const char* e = getenv("B2H_SIDE_STREAM"); use_side = e ? atoi(e) : 1; }
    if (use_fuse < 0) { const char* e = getenv("B2H_FUSE_TICK"); use_fuse = e ? atoi(e) : 1; }
    int use_tile = 1;                                 // read at every call: the parity tests compare both tick kernels
    { const char* e = getenv("B2H_TILE_TICK"); if (e) use_tile = atoi(e); }
    cudaStream_t rider_stream = use_side ? ctx->side : st;
    const int epl = (d + G - 1) / G;                 // front elements per lane
    // NUTS: the tile kernel (engine_tile.inl) is the tick for every row length; HMC (and B2H_TILE_TICK=0) keep the
    // register-front kernel, which needs the chain's row to fit its group's registers
    const bool tile_tick = use_fuse && use_tile && !HMC;
    const bool fuse = use_fuse && (tile_tick || epl <= 4);

    // Dense metric, before the kernel that holds the pre part of tick t: all momentum contractions launched so far
    // must have landed (a chain that started a transition one tick ago -- HMC with L = 1, a first-step divergence --
    // may start the next one now: p0 and v0 both came from the previous tick's side launch), and this parity's
    // request list is about to be reused.
    auto pre_prologue = [&](i64 t) -> int {
        const int b = (int)(t & 1);
        last_parity = b;
        v.mom_parity = b;
        for (int k = 0; k < 2; ++k)
            if (side_pending[k]) { B2H_CUDA(cudaStreamWaitEvent(st, ctx->ev_side[k], 0)); side_pending[k] = false; }
        B2H_CUDA(cudaMemsetAsync(v.mom_count + b, 0, sizeof(int), st));
        return 0;
    };
    // ... and after it, on the side stream: p0 = z . S^T and v0 = imm . p0 = z . L^T of the transitions queued by
    // this pre part, both from the same normals in one grouped launch.
    auto pre_epilogue = [&](i64 t) -> int {
        const int b = (int)(t & 1);
        if (use_side) {
            B2H_CUDA(cudaEventRecord(ctx->ev_pre[b], st));
            B2H_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_pre[b], 0));
        }
        // few rows, full reduction length: split K so that the tiles spread over all SMs, then reduce + scatter
        const i64 plane = (i64)C * d;
        T* part1 = v.mom_part;
        T* part2 = v.mom_part + (size_t)kRiderSplit * plane;
        // metric->reserved bit 0: the factors are the reference's triangular ones (sqrt_t = L^-1 lower, chol_t = L^T
        // upper: metrics.py:56-58), so half of each contraction is skipped
        const int tri = (metric->reserved & 1) ? 1 : 0;
        GemmGroup<T> g1{v.mom_z + (size_t)b * C * d, (i64)d, sqrt_t, (i64)d, part1, (i64)d, C, v.mom_count + b,
                        nullptr, nullptr, nullptr, tri ? 2 : 0};
        GemmGroup<T> g2{v.mom_z + (size_t)b * C * d, (i64)d, chol_t, (i64)d, part2, (i64)d, C, v.mom_count + b,
                        nullptr, nullptr, nullptr, tri ? 1 : 0};
        static int rsplit = 0;
        if (rsplit == 0) { const char* e = getenv("B2H_RIDER_SPLIT"); rsplit = e ? atoi(e) : kRiderSplit; if (rsplit < 1 || rsplit > kRiderSplit) rsplit = kRiderSplit; }
        launch_gemm_grouped<T>(rider_stream, g1, g2, none, d, d, rsplit, plane, 0);
        rider_reduce_kernel<T><<<dim3(C < 592 ? C : 592, 2), 128, 0, rider_stream>>>(
            part1, v.mom_count + b, v.mom_list + (size_t)b * C, v.mom_p, part2, v.mom_count + b,
            v.mom_list + (size_t)b * C, v.mom_v, rsplit, plane, d);
        if (use_side) {
            B2H_CUDA(cudaEventRecord(ctx->ev_side[b], ctx->side));
            side_pending[b] = true;
        }
        return 0;
    };
    auto launch_pre = [&](i64 t) -> int {
        if (pl.dense) {
            if (int e = pre_prologue(t)) return e;
            split_pre_kernel<T, G, true, HMC><<<grid, thr, 0, st>>>(v);      // half kick + drift by the v/w recurrence
            return pre_epilogue(t);
        }
        split_pre_kernel<T, G, false, HMC><<<grid, thr, 0, st>>>(v);
        return 0;
    };
#define B2H_POSTPRE(D_, E_, P_) split_postpre_kernel<T, G, D_, HMC, E_, P_><<<grid, thr, 0, st>>>(v, nd)
    // b2h_tick_timer: an event before and after the tick kernel on the engine's stream
    auto mark_event = [&]() {
        if (ctx->tick_events_used == ctx->tick_events.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            ctx->tick_events.push_back(e);
        }
        cudaEventRecord(ctx->tick_events[ctx->tick_events_used++], st);
    };
    auto grad_mark = [&]() { if (ctx->tick_timer == 2) mark_event(); };
    auto tick_mark = [&]() { if (ctx->tick_timer == 1) mark_event(); };
    auto launch_postpre = [&](i64 t, int* nd) -> int {                       // post of tick t - 1, pre of tick t
        if (tile_tick) {
            if (pl.dense) {
                if (int e = pre_prologue(t)) return e;
                tick_mark();
                launch_tile_tick<T, true>(st, v, nd, true, ctx->sm_count);
                tick_mark();
                return pre_epilogue(t);
            }
            tick_mark();
            launch_tile_tick<T, false>(st, v, nd, true, ctx->sm_count);
            tick_mark();
            return 0;
        }
        if (pl.dense) {
            if (int e = pre_prologue(t)) return e;
            tick_mark();
            if (epl <= 1) B2H_POSTPRE(true, 1, true);
            else if (epl <= 2) B2H_POSTPRE(true, 2, true);
            else B2H_POSTPRE(true, 4, true);
            tick_mark();
            return pre_epilogue(t);
        }
        tick_mark();
        if (epl <= 1) B2H_POSTPRE(false, 1, true);
        else if (epl <= 2) B2H_POSTPRE(false, 2, true);
        else B2H_POSTPRE(false, 4, true);
        tick_mark();
        return 0;
    };
    auto launch_post_last = [&](int* nd) {                                   // post of the call's last tick
        if (tile_tick) {
            if (pl.dense) launch_tile_tick<T, true>(st, v, nd, false, ctx->sm_count);
            else launch_tile_tick<T, false>(st, v, nd, false, ctx->sm_count);
            return;
        }
        if (pl.dense) {
            if (epl <= 1) B2H_POSTPRE(true, 1, false);
            else if (epl <= 2) B2H_POSTPRE(true, 2, false);
            else B2H_POSTPRE(true, 4, false);
        } else {
            if (epl <= 1) B2H_POSTPRE(false, 1, false);
            else if (epl <= 2) B2H_POSTPRE(false, 2, false);
            else B2H_POSTPRE(false, 4, false);
        }
    };
#undef B2H_POSTPRE

    // Correlated Gaussian target under the tile tick kernel's one-chain layouts: U = 0.5 (q' - mu) . g' is formed by the
    // tick kernel's pass A (it holds q' and g' already); the gradient call is the contraction alone.
    // B2H_TICK_POTENTIAL=0: the separate potential kernel (A/B measurements, parity tests)
    bool grad_only = false;
    v.u_center = nullptr;
    if (tile_tick && model->kind == B2H_MODEL_CORR_GAUSSIAN && ((uintptr_t)model->a % 16) == 0) {
        const char* e = getenv("B2H_TICK_POTENTIAL");
        const bool layout = pl.dense ? tile_tick_unit_chain_layout<T, true>(v, ctx->sm_count)
                                     : tile_tick_unit_chain_layout<T, false>(v, ctx->sm_count);
        if (layout && !(e && atoi(e) == 0)) { grad_only = true; v.u_center = (const T*)model->a; }
    }
    rc = launch_pre(0);
    for (i64 tick = 0; tick < bound && rc == 0; ++tick) {
        grad_mark();
        rc = potential_and_grad_impl<T>(ctx, model, v.xa, v.Unew, v.xb, C, model_ws, model_ws_bytes, grad_only);
        grad_mark();
        if (rc) break;
        const bool last = tick + 1 == bound;
        const bool check = (max_ticks <= 0) && ((tick & 3) == 3 || last);
        if (check) B2H_CUDA(cudaMemsetAsync(not_done_dev, 0, sizeof(int), st));
        // the tick's only metric contraction on the main stream: w' = imm . g'
        if (pl.dense) launch_dense_apply<T>(st, v.xb, imm_dense, v.xc, C, d, d, nullptr, nullptr);
        int* nd = check ? not_done_dev : nullptr;
        if (fuse && !last) {
            rc = launch_postpre(tick + 1, nd);
            if (rc) break;
        } else if (fuse) {
            launch_post_last(nd);
        } else {
            if (pl.dense) split_post_kernel<T, G, true, HMC><<<grid, thr, 0, st>>>(v, nd);
            else split_post_kernel<T, G, false, HMC><<<grid, thr, 0, st>>>(v, nd);
        }
        if (check) {
            cudaError_t e = cudaMemcpyAsync(host_flag, not_done_dev, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { rc = cuda_fail(e, "tick poll"); break; }
            if (*host_flag == 0) break;
        }
        if (!fuse && !last) rc = launch_pre(tick + 1);
    }
    if (pl.dense && rc == 0) {
        // join the side stream so that a resumed run (or the caller) starts with no momentum request pending
        for (int b = 0; b < 2; ++b)
            if (side_pending[b]) B2H_CUDA(cudaStreamWaitEvent(st, ctx->ev_side[b], 0));
    }
    (void)last_parity;
    if (rc) return rc;
    B2H_LAUNCH_CHECK();
    return 0;
}

template <typename T>
int run_split_g(b2h_ctx* ctx, EngineView<T>& v, const EnginePlan& pl, const b2h_model* model, const b2h_metric* metric,
                const b2h_cfg* cfg, i64 max_ticks, int n_transitions, void* model_ws, i64 model_ws_bytes,
                int* not_done_dev, int resume, bool hmc) {
#define B2H_RUN_SPLIT(G, H) \
    run_split<T, G, H>(ctx, v, pl, model, metric, cfg, max_ticks, n_transitions, model_ws, model_ws_bytes, not_done_dev, resume)
    // Dense metric: the whole call runs on the context's high-priority stream, ordered after the caller's stream at
    // entry and before it at exit (launch.h: the momentum tiles of the side stream then only fill idle SMs).
    struct Hop {
        b2h_ctx* c;
        cudaStream_t user;
        bool on;
        ~Hop() {
            if (!on) return;
            cudaEventRecord(c->ev_hop, c->stream);
            cudaStreamWaitEvent(user, c->ev_hop, 0);
            c->stream = user;
        }
    } hop{ctx, ctx->stream, false};
    if (pl.dense && ctx->hi) {
        B2H_CUDA(cudaEventRecord(ctx->ev_hop, ctx->stream));
        B2H_CUDA(cudaStreamWaitEvent(ctx->hi, ctx->ev_hop, 0));
        ctx->stream = ctx->hi;
        hop.on = true;
    }
    if (hmc) {
        switch (pl.G) {
            case 8: return B2H_RUN_SPLIT(8, true);
            case 32: return B2H_RUN_SPLIT(32, true);
            default: return B2H_RUN_SPLIT(256, true);
        }
    }
    switch (pl.G) {
        case 8: return B2H_RUN_SPLIT(8, false);
        case 32: return B2H_RUN_SPLIT(32, false);
        default: return B2H_RUN_SPLIT(256, false);
    }
#undef B2H_RUN_SPLIT
}

}  // namespace b2h
