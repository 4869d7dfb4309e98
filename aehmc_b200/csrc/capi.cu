// Context management, error reporting and the sampler entry points of the C-ABI.
#include <stdlib.h>
#include <string.h>

#include "launch.h"

namespace b2h {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what) {
    g_last_error = std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e);
    return B2H_ERR_CUDA;
}

}  // namespace b2h

using namespace b2h;

extern "C" {

const char* b2h_last_error(void) { return g_last_error.c_str(); }

int b2h_version(void) { return 100; }

int b2h_ctx_create(int device, void* stream, b2h_ctx** out) {
    if (!out) { set_error("null out pointer"); return B2H_ERR_ARG; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error(std::string("no CUDA device available: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                  " (libb200hmc has no CPU fallback)");
        return B2H_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_error("bad device ordinal"); return B2H_ERR_ARG; }
    B2H_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2H_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("libb200hmc is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
        return B2H_ERR_UNSUPPORTED;
    }
    b2h_ctx* ctx = new b2h_ctx;
    ctx->device = device;
    ctx->stream = (cudaStream_t)stream;
    ctx->sm_count = prop.multiProcessorCount;
    // the momentum tiles of restarting chains run at the LOWEST priority beside the full-size launches of a dense-metric
    // run, which hop onto the high-priority stream (engine_split.inl); B2H_HI_STREAM=0: the former arrangement (side
    // stream at the highest priority, everything else on the caller's stream)
    int prio_least = 0, prio_greatest = 0;
    B2H_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    const char* hop = getenv("B2H_HI_STREAM");
    const bool use_hi = !(hop && atoi(hop) == 0);
    B2H_CUDA(cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, use_hi ? prio_least : prio_greatest));
    ctx->hi = nullptr;
    if (use_hi) B2H_CUDA(cudaStreamCreateWithPriority(&ctx->hi, cudaStreamNonBlocking, prio_greatest));
    B2H_CUDA(cudaEventCreateWithFlags(&ctx->ev_hop, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        B2H_CUDA(cudaEventCreateWithFlags(&ctx->ev_pre[i], cudaEventDisableTiming));
        B2H_CUDA(cudaEventCreateWithFlags(&ctx->ev_side[i], cudaEventDisableTiming));
    }
    B2H_CUDA(cudaMallocHost(&ctx->host_flag, sizeof(int)));
    *out = ctx;
    return 0;
}

int b2h_ctx_destroy(b2h_ctx* ctx) {
    if (ctx) {
        cudaStreamDestroy(ctx->side);
        if (ctx->hi) cudaStreamDestroy(ctx->hi);
        cudaEventDestroy(ctx->ev_hop);
        cudaFreeHost(ctx->host_flag);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(ctx->ev_pre[i]); cudaEventDestroy(ctx->ev_side[i]); }
        for (cudaEvent_t e : ctx->tick_events) cudaEventDestroy(e);
    }
    delete ctx;
    return 0;
}

int b2h_ctx_sync(b2h_ctx* ctx) {
    if (!ctx) { set_error("null context"); return B2H_ERR_ARG; }
    B2H_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// fold the recorded event pairs into the totals (synchronises the stream)
static int tick_timer_collect(b2h_ctx* ctx) {
    B2H_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i + 1 < ctx->tick_events_used; i += 2) {
        float ms = 0.f;
        B2H_CUDA(cudaEventElapsedTime(&ms, ctx->tick_events[i], ctx->tick_events[i + 1]));
        ctx->tick_ms += ms;
        ctx->tick_launches += 1;
    }
    ctx->tick_events_used = 0;
    return 0;
}

int b2h_tick_timer(b2h_ctx* ctx, int32_t enable) {
    if (!ctx) { set_error("null context"); return B2H_ERR_ARG; }
    if (int rc = tick_timer_collect(ctx)) return rc;
    ctx->tick_timer = (enable == 1 || enable == 2) ? enable : 0;
    ctx->tick_ms = 0.0;
    ctx->tick_launches = 0;
    return 0;
}

int b2h_tick_timer_read(b2h_ctx* ctx, double* total_ms, int64_t* launches) {
    if (!ctx || !total_ms || !launches) { set_error("null argument"); return B2H_ERR_ARG; }
    if (int rc = tick_timer_collect(ctx)) return rc;
    *total_ms = ctx->tick_ms;
    *launches = ctx->tick_launches;
    return 0;
}

int b2h_nuts_run(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng,
                 const b2h_cfg* cfg, const b2h_adapt* adapt, void* q, void* p, void* U, void* g, double* step_size,
                 int64_t C, int32_t n_transitions, int64_t max_ticks, int32_t resume, b2h_diag* diag, void* draws,
                 double* draw_stats, int32_t n_store, int64_t* counters, void* workspace, int64_t workspace_bytes) {
    return nuts_run_impl(ctx, model, metric, rng, cfg, adapt, q, p, U, g, step_size, C, n_transitions, max_ticks,
                         resume, diag, draws, draw_stats, n_store, counters, workspace, workspace_bytes, false);
}

int b2h_hmc_run(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng,
                const b2h_cfg* cfg, const b2h_adapt* adapt, void* q, void* p, void* U, void* g, double* step_size,
                int64_t C, int32_t n_transitions, b2h_diag* diag, void* draws, double* draw_stats, int32_t n_store,
                int64_t* counters, void* workspace, int64_t workspace_bytes) {
    return nuts_run_impl(ctx, model, metric, rng, cfg, adapt, q, p, U, g, step_size, C, n_transitions, 0, 0, diag,
                         draws, draw_stats, n_store, counters, workspace, workspace_bytes, true);
}

int b2h_nuts_expand(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng,
                    const b2h_cfg* cfg, b2h_tree* tree, const double* step_size, int64_t C, b2h_diag* diag,
                    void* workspace, int64_t workspace_bytes) {
    return nuts_expand_impl(ctx, model, metric, rng, cfg, tree, step_size, C, diag, workspace, workspace_bytes);
}

int b2h_nuts_subtree(b2h_ctx* ctx, const b2h_model* model, const b2h_metric* metric, const b2h_rng* rng,
                     const b2h_cfg* cfg, b2h_subtree* sub, const double* step_size, int64_t C, void* workspace,
                     int64_t workspace_bytes) {
    return nuts_subtree_impl(ctx, model, metric, rng, cfg, sub, step_size, C, workspace, workspace_bytes);
}

int64_t b2h_nuts_workspace_bytes(const b2h_model* model, const b2h_metric* metric, const b2h_cfg* cfg, int64_t C) {
    if (!model || !metric || !cfg) return -1;
    return engine_workspace_bytes(model, metric, cfg, C);
}

int32_t b2h_nuts_plan_group(const b2h_model* model, const b2h_metric* metric, const b2h_cfg* cfg, int64_t C,
                            int32_t free_running) {
    if (!model || !metric || !cfg) return -1;
    return engine_plan_group(model, metric, cfg, C, free_running != 0);
}

int64_t b2h_hmc_workspace_bytes(const b2h_model* model, const b2h_metric* metric, const b2h_cfg* cfg, int64_t C) {
    if (!model || !metric || !cfg) return -1;
    return engine_workspace_bytes(model, metric, cfg, C);
}

}  // extern "C"
