// TEST TOOLING ONLY (see cuda_shim.h): runs engine.cuh's fused tick loop on the host,
// one "thread" per chain, dim-major layout, for the fused models.
#include "cuda_shim.h"

#include <vector>

#include "../../aehmc_b200/csrc/engine.cuh"
#include "../../aehmc_b200/csrc/models.cuh"

using namespace b2h;

static int g_reg_front = 0;      // 0: memory front; 16: register front with 16 elements per (single) lane

template <int MODEL, bool HMC, class Front>
static void run_chain_f(EngineView<double>& v, const ModelDev& m, int c, long long max_ticks) {
    Chain<double, 1> ch(v, c, nullptr);
    ch.load();
    Front f;
    bool bound = false;
    long long tick = 0;
    while (max_ticks <= 0 || tick < max_ticks) {
        if (ch.r.phase == PH_DONE) break;
        if (ch.r.phase == PH_START) {
            if (HMC) hmc_begin<double, 1, false>(ch);
            else begin_transition<double, 1, false>(ch);
            bound = false;
        }
        if (!bound) { f.bind(ch); bound = Front::kRegs; }
        half_kick_drift<double, 1, false, false>(ch, f);
        double U;
        if constexpr (Front::kRegs) {
            U = model_grad_front<double, 1, MODEL>(m, f, 0, nullptr);
        } else {
            U = model_grad<double, 1, MODEL>(m, f.Q + ch.base, f.Gd + ch.base, v.sj, 0, nullptr);
        }
        bool ended;
        if (HMC) ended = hmc_post<double, 1, false, false>(ch, U, f);
        else ended = post_gradient<double, 1, false, false>(ch, U, f);
        if (ended) bound = false;
        ++tick;
    }
    if (bound) f.flush(ch);
    ch.store();
}

template <int MODEL, bool HMC>
static void run_chain(EngineView<double>& v, const ModelDev& m, int c, long long max_ticks) {
    if (g_reg_front) run_chain_f<MODEL, HMC, RegFront<double, 16>>(v, m, c, max_ticks);
    else run_chain_f<MODEL, HMC, MemFront<double>>(v, m, c, max_ticks);
}

extern "C" void sim_set_reg_front(int on) { g_reg_front = on; }
static int g_exact_doubling = 0;
extern "C" void sim_set_exact_doubling(int on) { g_exact_doubling = on; }

extern "C" int sim_run(int model_kind, int hmc, int C, int d, int maxd, const double* a, const double* b, double s0,
                       int imm_kind, const double* imm, double imm_scalar, double* q, double* p, double* U, double* g,
                       double* eps, long long n_inj, const double* z, const double* u_dir, const double* u_biased,
                       const double* u_uniform, const double* u_accept, int n_transitions, int hmc_L, double div_thr,
                       double* acc, int* nd, unsigned char* turning, unsigned char* diverging, int* nleap,
                       double* draws, int n_store,
                       int adapt_steps, const unsigned char* stage, const unsigned char* wend, double target,
                       double init_step_size, double* imm_out) {
    EngineView<double> v;
    memset(&v, 0, sizeof(v));
    const size_t n = (size_t)C * d;
    std::vector<std::vector<double>> store;
    auto mk = [&](size_t k) { store.emplace_back(k, 0.0); return store.back().data(); };
    store.reserve(64);
    v.C = C; v.d = d; v.maxd = hmc ? 1 : maxd;
    v.exact_doubling = g_exact_doubling;
    v.sc = 1; v.sj = C; v.sck = 1;
    v.ql = mk(n); v.pl = mk(n); v.gl = mk(n); v.qr = mk(n); v.pr = mk(n); v.gr = mk(n);
    v.qs = mk(n); v.ps = mk(n); v.gs = mk(n); v.qp = mk(n); v.pp = mk(n); v.gp = mk(n);
    v.msum = mk(n); v.sms = mk(n); v.mck = mk(n * v.maxd); v.sckp = mk(n * v.maxd);
    std::vector<ChainRec> rec(C);
    v.rec = rec.data();
    std::vector<double> imm_own;
    v.imm_kind = imm_kind;
    if (imm_kind == 0) { imm_own.assign(1, imm_scalar); v.imm = imm_own.data(); v.imm_sc = 0; v.imm_sj = 0; }
    else if (imm_kind == 1) { imm_own.assign(imm, imm + d); v.imm = imm_own.data(); v.imm_sc = 0; v.imm_sj = 1; }
    else {
        imm_own.resize(n);
        for (int c = 0; c < C; ++c) for (int j = 0; j < d; ++j) imm_own[(size_t)j * C + c] = imm[(size_t)c * d + j];
        v.imm = imm_own.data(); v.imm_sc = 1; v.imm_sj = C;
    }
    v.rng.mode = 1; v.rng.n_injected = n_inj; v.rng.z = z; v.rng.u_dir = u_dir; v.rng.u_biased = u_biased;
    v.rng.u_uniform = u_uniform; v.rng.u_accept = u_accept;
    v.div_thr = div_thr; v.n_transitions = n_transitions; v.hmc_L = hmc_L;
    v.out.draws = draws; v.out.n_store = n_store; v.out.thin = 1;
    std::vector<long long> da_step(C, 1), wc_n(C, 0);
    std::vector<double> da_x(C, 0.0), da_xa(C, 0.0), da_g(C, 0.0), da_mu(C, init_step_size), wc_mean(n, 0.0), wc_m2(n, 0.0);
    if (adapt_steps > 0) {
        v.adapt.enabled = 1; v.adapt.num_steps = adapt_steps; v.adapt.stage = stage; v.adapt.window_end = wend;
        v.adapt.target = target; v.adapt.gamma = 0.05; v.adapt.t0 = 10; v.adapt.kappa = 0.75;
        v.adapt.da_step = da_step.data(); v.adapt.da_x = da_x.data(); v.adapt.da_x_avg = da_xa.data();
        v.adapt.da_g_avg = da_g.data(); v.adapt.da_mu = da_mu.data(); v.adapt.wc_mean = wc_mean.data();
        v.adapt.wc_m2 = wc_m2.data(); v.adapt.wc_n = wc_n.data();
    }
    ModelDev m;
    m.kind = model_kind; m.dim = d; m.n_data = 0; m.a = a; m.b = b; m.c = nullptr; m.s0 = s0; m.s1 = 0;
    for (int c = 0; c < C; ++c) {
        for (int j = 0; j < d; ++j) v.qp[(size_t)j * C + c] = q[(size_t)c * d + j];
        double u;
        if (model_kind == MODEL_IID) u = model_grad<double, 1, MODEL_IID>(m, v.qp + c, v.gp + c, C, 0, nullptr);
        else if (model_kind == MODEL_FUNNEL) u = model_grad<double, 1, MODEL_FUNNEL>(m, v.qp + c, v.gp + c, C, 0, nullptr);
        else u = model_grad<double, 1, MODEL_SCHOOLS>(m, v.qp + c, v.gp + c, C, 0, nullptr);
        memset(&rec[c], 0, sizeof(ChainRec));
        rec[c].phase = PH_START; rec[c].U_prop = u; rec[c].eps = adapt_steps > 0 ? 1.0 : eps[c];
    }
    for (int c = 0; c < C; ++c) {
        if (hmc) {
            if (model_kind == MODEL_IID) run_chain<MODEL_IID, true>(v, m, c, 0);
            else if (model_kind == MODEL_FUNNEL) run_chain<MODEL_FUNNEL, true>(v, m, c, 0);
            else run_chain<MODEL_SCHOOLS, true>(v, m, c, 0);
        } else {
            if (model_kind == MODEL_IID) run_chain<MODEL_IID, false>(v, m, c, 0);
            else if (model_kind == MODEL_FUNNEL) run_chain<MODEL_FUNNEL, false>(v, m, c, 0);
            else run_chain<MODEL_SCHOOLS, false>(v, m, c, 0);
        }
    }
    for (int c = 0; c < C; ++c) {
        for (int j = 0; j < d; ++j) {
            size_t s = (size_t)j * C + c, r = (size_t)c * d + j;
            q[r] = v.qp[s]; p[r] = v.pp[s]; g[r] = v.gp[s];
            if (imm_out && imm_kind == 2) imm_out[r] = v.imm[s];
        }
        U[c] = rec[c].U_prop; eps[c] = rec[c].eps;
        acc[c] = rec[c].accept_prob; nd[c] = rec[c].last_nd; turning[c] = rec[c].last_flags & 1;
        diverging[c] = (rec[c].last_flags >> 1) & 1; nleap[c] = rec[c].last_nleap;
    }
    return 0;
}
