// The persistent fused kernel and its launch policy; included by engine_fused_f32.cu / engine_fused_f64.cu.
#include <stdlib.h>

#include "engine_host.cuh"
#include "models.cuh"

namespace b2h {

#ifndef B2H_DEFER_PATHS
#define B2H_DEFER_PATHS 1
#endif
// Lanes of a warp that must wait for a deferred path before it runs (it also runs when no lane of the warp can step).
// Measured on config 4 (65536 chains, thread per chain, 200 transitions / free-running, ms; benchmarks/defer_ab.sh):
//   eight schools  no deferral 128 / 94   8 lanes 107 / 85   16 lanes 90 / 68   24 lanes + 3-tick limit 95 / 73   32 + 4: 106 / 95
//   funnel         no deferral 1982 / 740 8 lanes 1520 / 741 16 lanes 1444 / 1040
// A limit on the ticks a lane may wait did not pay (16 lanes + 2 ticks: 107 / 77 and 1977 / 828).
#ifndef B2H_DEFER_LANES
#define B2H_DEFER_LANES 16
#endif
#ifndef B2H_DEFER_LANES_FUNNEL
#define B2H_DEFER_LANES_FUNNEL 8     // deep, uneven trees: the path is rare, waiting for half a warp idles the parked lanes
#endif

// E > 8 only exists for thread-per-chain (G = 1): the chain's whole front is the thread's registers, so the launch
// bound leaves the thread its 255 registers instead of trading them for resident warps.
template <typename T, int G, int MODEL, bool HMC, int E>
__global__ void __launch_bounds__(Geo<G>::kThreads, (E > 8 ? 2 : Geo<G>::kMinBlocksFused))
fused_run_kernel(EngineView<T> v, ModelDev m, i64 max_ticks, int stage_ck) {
    typedef typename FrontOf<T, E>::type Front;
    __shared__ double red_s[128];
    extern __shared__ __align__(16) unsigned char ck_smem[];    // stage_ck: [chains of the CTA][2][maxd][d] checkpoints
    // Thread per chain: drawing the d normals of a new momentum inside begin_transition runs with the few lanes of the
    // warp that start a transition on that tick (ncu: 45 % of the kernel's warp-instructions at 5 of 32 lanes).  The
    // normals of the NEXT transition are drawn ahead instead, one Box-Muller pair per tick at the point of the loop
    // where all lanes are converged, into this per-thread column of shared memory.
    constexpr bool kAhead = (G == 1 && E > 0);
    constexpr int ZE = kAhead ? ((E + 1) & ~1) : 1;
    __shared__ double zsm[ZE][kAhead ? Geo<G>::kThreads : 1];
    // deferred paths (below) vote with the whole warp: a lane without a chain stays, as a finished one
    constexpr bool kDefer = (G == 1 && E > 0 && !HMC) && B2H_DEFER_PATHS;
    const bool has_chain = Geo<G>::chain() < v.C;
    if (!has_chain && !kDefer) return;
    const int c = has_chain ? Geo<G>::chain() : 0;
    Chain<T, G> ch(v, c, red_s);
    if (has_chain) ch.load();
    else ch.r.phase = PH_DONE;
    // U-turn checkpoints staged in shared memory for the whole launch when the CTA's chains fit (termination.py:63-131:
    // written on even steps, read on odd ones, 2 x max_num_expansions x d values per chain)
    if (stage_ck)
        ch.stage_checkpoints(reinterpret_cast<T*>(ck_smem) + (size_t)(G > 32 ? 0 : threadIdx.x / G) * 2 * v.maxd * v.d);
    Front f;
    bool bound = false;
    i64 tick = 0;
    int zfill = 0, ztrans = -1;                 // normals drawn ahead, and the transition they belong to
    const bool ahead = kAhead && v.rng.mode == 0;
    const double* zs = kAhead ? &zsm[0][threadIdx.x] : nullptr;
    // Thread per chain, NUTS: the paths only SOME lanes of a warp need on a given tick -- the end of a sub-tree
    // (expand_once: a third of the ticks of a chain with short trees) and the start of a transition -- used to run
    // with those few lanes while the others waited (ncu: 11.5 of 32 lanes active).  Here a lane that reaches such a path
    // parks until kDeferLanes lanes of its warp wait for the same path (or no lane of the warp can step), and the path
    // then runs once for all of them.  A chain's result does not depend on when its own steps run.
    if constexpr (kDefer) {
        constexpr unsigned kAll = 0xffffffffu;
        int pend = 0;                               // kSubtreeEnd | flags: this lane's sub-tree end waits to be run
        constexpr int kLanes = MODEL == MODEL_FUNNEL ? B2H_DEFER_LANES_FUNNEL : B2H_DEFER_LANES;
        while (true) {
            const bool live = ch.r.phase != PH_DONE && (max_ticks <= 0 || tick < max_ticks);
            if (!__any_sync(kAll, live)) break;
            const bool can_step = live && ch.r.phase == PH_RUN && pend == 0;
            const bool none_steps = !__any_sync(kAll, can_step);
            // (1) sub-tree ends
            const bool want_end = live && pend != 0;
            if (__popc(__ballot_sync(kAll, want_end)) >= kLanes || none_steps) {
                if (want_end) {
                    subtree_end<T, G, false, false, Front>(ch, f, (pend & kEndDiv) != 0, (pend & kEndTerm) != 0);
                    pend = 0;
                    bound = false;
                }
            }
            // (2) transition starts (a sub-tree end above may have produced some)
            const bool want_begin = live && pend == 0 && ch.r.phase == PH_START;
            const bool none_steps2 = !__any_sync(kAll, live && ch.r.phase == PH_RUN && pend == 0);
            if (__popc(__ballot_sync(kAll, want_begin)) >= kLanes || none_steps2) {
                if (want_begin) {
                    const int zready = (ahead && ztrans == ch.r.t) ? zfill : 0;
                    begin_transition<T, G, false>(ch, zs, zready, Geo<G>::kThreads);
                    bound = false;
                    zfill = 0;
                    ztrans = ch.r.t + 1;
                }
            }
            // (3) one leapfrog of every lane that can step
            if (live && ch.r.phase == PH_RUN && pend == 0) {
                if (!bound) { f.bind(ch); bound = true; }
                half_kick_drift<T, G, false, false>(ch, f);
                const T U = model_grad_front<T, G, MODEL>(m, f, ch.lane, ch.red);
                if (ahead && zfill < v.d) {
                    double z0, z1;
                    philox_normal_pair(v.rng.key, v.rng.chain_offset + (uint64_t)c,
                                       (uint32_t)(v.rng.transition_offset + (uint64_t)ztrans), (uint32_t)(zfill >> 1), &z0, &z1);
                    zsm[zfill][threadIdx.x] = z0;
                    zsm[zfill + 1][threadIdx.x] = z1;
                    zfill += 2;
                }
                pend = post_gradient<T, G, false, false, Front, true>(ch, U, f);
                if (!(pend & kSubtreeEnd)) pend = 0;
                ++tick;
            }
        }
        // a sub-tree end still parked when the tick budget ran out belongs to the last tick
        if (pend != 0) { subtree_end<T, G, false, false, Front>(ch, f, (pend & kEndDiv) != 0, (pend & kEndTerm) != 0); bound = false; }
    } else
    while (max_ticks <= 0 || tick < max_ticks) {
        if (ch.r.phase == PH_DONE) break;
        if (ch.r.phase == PH_START) {
            const int zready = (ahead && ztrans == ch.r.t) ? zfill : 0;
            if (HMC) hmc_begin<T, G, false>(ch, zs, zready, Geo<G>::kThreads);
            else begin_transition<T, G, false>(ch, zs, zready, Geo<G>::kThreads);
            Group<G>::sync();
            bound = false;
            zfill = 0;
            ztrans = ch.r.t + 1;
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
        if constexpr (kAhead) {
            if (ahead && zfill < v.d) {
                double z0, z1;
                philox_normal_pair(v.rng.key, v.rng.chain_offset + (uint64_t)c,
                                   (uint32_t)(v.rng.transition_offset + (uint64_t)ztrans), (uint32_t)(zfill >> 1), &z0, &z1);
                zsm[zfill][threadIdx.x] = z0;
                zsm[zfill + 1][threadIdx.x] = z1;
                zfill += 2;
            }
        }
        bool ended;
        if (HMC) ended = hmc_post<T, G, false, false>(ch, U, f);
        else ended = post_gradient<T, G, false, false>(ch, U, f);
        if (ended) bound = false;
        Group<G>::sync();
        ++tick;
    }
    if (!has_chain) return;
    if (bound) f.flush(ch);                // max_ticks ran out in the middle of a sub-tree
    if (stage_ck) ch.unstage_checkpoints();
    ch.store();
    if (v.counters && ch.lane == 0) atomicAdd((unsigned long long*)&v.counters[3], (unsigned long long)tick);
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

// shared-memory staging of the checkpoints: when the chains of one CTA need at most this many bytes (all resident CTAs
// of an SM then still fit its 227 KB)
constexpr size_t kStageCkMaxBytes = 40 * 1024;

template <typename T, int G, bool HMC, int MODEL, int E>
static void launch_fused_e(cudaStream_t st, const EngineView<T>& v, const ModelDev& m, i64 max_ticks) {
    // Measured (config 4, 8 lanes per chain, benchmarks/c4_tail_probe.py): free-running, where every lane is busy and
    // the checkpoint traffic competes for L2, staging gives the funnel 0.74 -> 0.89 G evals/s; a run of a fixed number
    // of transitions ends with a few lone chains whose step latency is what counts, and there the staged kernel's
    // generic-address loads make a leapfrog 9 % slower (1338 -> 1473 ms for 200 transitions).  B2H_STAGE_CKPT=0 / 1
    // forces it off / on.
    static int use_stage = -2;
    if (use_stage == -2) { const char* e = getenv("B2H_STAGE_CKPT"); use_stage = e ? atoi(e) : -1; }
    const size_t bytes = (size_t)Geo<G>::kChainsPerBlock * 2 * v.maxd * v.d * sizeof(T);
    const bool want = use_stage < 0 ? max_ticks > 0 : use_stage != 0;
    const bool stage = want && !HMC && v.sj == 1 && bytes <= kStageCkMaxBytes;
    fused_run_kernel<T, G, MODEL, HMC, E><<<Geo<G>::grid(v.C), Geo<G>::kThreads, stage ? bytes : 0, st>>>(v, m, max_ticks,
                                                                                                          stage ? 1 : 0);
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
        if constexpr (G == 1) {
            // thread per chain, tiny targets (funnel, eight schools: d = 10): the whole front in the thread's registers
            if (use_regs && epl <= 10) return launch_fused_e<T, G, HMC, MODEL, 10>(st, v, m, max_ticks);
            if (use_regs && epl <= 16) return launch_fused_e<T, G, HMC, MODEL, 16>(st, v, m, max_ticks);
        }
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

template <typename T>
int launch_fused_g(cudaStream_t st, const EngineView<T>& v, const b2h_model* model, i64 max_ticks, int G, bool hmc) {
    if (hmc) {
        switch (G) {
            case 1: return launch_fused<T, 1, true>(st, v, model, max_ticks);
            case 8: return launch_fused<T, 8, true>(st, v, model, max_ticks);
            case 32: return launch_fused<T, 32, true>(st, v, model, max_ticks);
            default: return launch_fused<T, 256, true>(st, v, model, max_ticks);
        }
    }
    switch (G) {
        case 1: return launch_fused<T, 1, false>(st, v, model, max_ticks);
        case 8: return launch_fused<T, 8, false>(st, v, model, max_ticks);
        case 32: return launch_fused<T, 32, false>(st, v, model, max_ticks);
        default: return launch_fused<T, 256, false>(st, v, model, max_ticks);
    }
}

}  // namespace b2h
