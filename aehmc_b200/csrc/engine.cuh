// Per-chain HMC/NUTS state machine, one leapfrog per "tick".
//
// A chain is advanced by a group of G threads (G=1: thread per chain over a
// dim-major [d][C] layout; G<=32: sub-warp/warp per chain; G>32: CTA per chain,
// both over row-major [C][d]).  Every tick is exactly one velocity-Verlet step
// (reference integrators.py:58-73) followed by the bookkeeping the reference
// performs in trajectory.dynamic_integration.add_one_state (trajectory.py:195-273)
// and, at sub-tree ends, multiplicative_expansion.expand_once (trajectory.py:463-608).
// Chains that finish a transition start the next one on the following tick, so
// heterogeneous tree depths never idle the machine.
//
// The integration front lives IN PLACE in the left/right edge arrays: going
// right advances (qr,pr,gr), going left advances (ql,pl,gl); no "current state"
// copy exists.  The chain's position between transitions is the proposal
// (qp,gp,U_prop).
//
// The numbered quirks Q1..Q13 are the ones listed in SURVEY.md section 3.5.
#pragma once

#include "common.cuh"

namespace b2h {

enum Phase : int { PH_START = 0, PH_RUN = 1, PH_DONE = 2 };

// Per-chain scalar record (one 256-byte line per chain).
struct __align__(16) ChainRec {
    double w_sub, slpa_sub, w_prop, slpa_prop;  // Q13: always float64
    double E0, E_sub, E_prop, U_sub, U_prop;
    double eps;
    double accept_prob;                          // of the last completed transition (Q9)
    double hmc_E0;
    int phase, k, s, go_right, imin, imax, sub_len, mom_slot;
    int t;                                       // completed transitions in this run
    int nleap;                                   // integrator steps of the running transition
    int last_nd, last_flags, last_nleap, hmc_step;
    i64 total_leap;
    int t_base, sub_term;                        // t at the start of the current call (resume); sub-tree U-turn flag
    double U_left, U_right, U_front;             // potential energy at the edges / at the front (explicit-state API)
    double pad[8];
};
static_assert(sizeof(ChainRec) == 256, "ChainRec must be one 256-byte record");

struct RngView {
    int mode;                // 0 philox, 1 injected
    PhiloxKey key;
    uint64_t chain_offset, transition_offset;
    i64 n_injected;
    const double *z, *u_dir, *u_biased, *u_uniform, *u_accept;
};

struct AdaptView {
    int enabled, num_steps;
    const uint8_t *stage, *window_end;
    double target, gamma, t0, kappa;
    i64* da_step;
    double *da_x, *da_x_avg, *da_g_avg, *da_mu;
    void *wc_mean, *wc_m2;   // engine layout, dtype T
    i64* wc_n;
};

struct OutView {
    void* draws;             // [n_store][C][d] row-major, dtype T
    double* draw_stats;      // [n_store][C][4]
    int n_store;
    double* acceptance_probability;
    int32_t* num_doublings;
    uint8_t *is_turning, *is_diverging;
    int32_t* n_leapfrog;
};

// All engine arrays.  elem(c, j) = c*sc + j*sj ; ckpt(c, level, j) = c*sck + (level*d + j)*sj.
template <typename T>
struct EngineView {
    int C, d, maxd;
    i64 sc, sj, sck;
    T *ql, *pl, *gl, *qr, *pr, *gr;      // trajectory edges (integration fronts)
    T *qs, *ps, *gs;                     // sub-tree proposal
    T *qp, *pp, *gp;                     // transition proposal == chain position between transitions
    T *msum, *sms;                       // momentum sums: whole trajectory / current sub-tree
    T *mck, *sckp;                       // U-turn checkpoints [C][maxd][d]
    T *vl, *vr, *vck;                    // dense metric only: velocities imm.p of pl, pr, checkpoints
    T *wl, *wr, *ws, *wp;                // dense metric only: imm.g of the edges / sub-tree proposal / proposal, so that
                                         // the half-step velocity is a recurrence: imm.(p - h g) = v - h w (no contraction)
    ChainRec* rec;
    // metric (diag family): imm(c, j) = imm[c*imm_sc + j*imm_sj]
    int imm_kind;
    T* imm;
    i64 imm_sc, imm_sj;
    // split-mode scratch (row-major [C][d]) and dense-momentum compaction
    T *xa, *xb, *xc, *Unew;              // xa = q' (gradient input), xb = g' (gradient output), xc = imm.g' (dense)
    // dense-metric momentum, one transition of lookahead: mom_p/mom_v [C][d] hold p0 = sqrt z and v0 = imm p0 of
    // each chain's NEXT transition; a chain that starts a transition consumes them and queues a request
    // (list/count of parity mom_parity, normals in mom_z) that rides along the following dense applies.
    T *mom_p, *mom_v, *mom_z;            // mom_z: [2][C][d] compact request rows
    T *mom_part;                         // split-K partial planes of the two momentum contractions
    int* mom_count;                      // [2]
    int* mom_list;                       // [2][C]
    int mom_parity;
    int* scratch;                        // [0] chains not yet done (split-mode poll)
    RngView rng;
    AdaptView adapt;
    OutView out;
    double div_thr;
    int n_transitions;                   // per chain; <=0 : free running
    int hmc_L;
    int sub_max_steps;                   // > 0: stand-alone dynamic_integration: scan length of the one sub-tree
    int stop_at_subtree_end;             // stand-alone dynamic_integration: stop instead of running expand_once
    i64* counters;                       // [0] leapfrogs [1] transitions [2] ticks [3] active chain-ticks
};

// ---------------------------------------------------------------------------
// draws
// ---------------------------------------------------------------------------
B2H_DEVINL double draw_u(const RngView& r, int kind, int c, int t, int idx, int maxd) {
    if (r.mode == 1) {
        i64 row = (i64)c * r.n_injected + t;
        if (kind == DRAW_DIR) return r.u_dir[row * maxd + idx];
        if (kind == DRAW_BIASED) return r.u_biased[row * maxd + idx];
        if (kind == DRAW_UNIFORM) return r.u_uniform[row * (((i64)1 << maxd) - 1) + idx];
        return r.u_accept[row];
    }
    return philox_uniform(r.key, r.chain_offset + (uint64_t)c, (uint32_t)(r.transition_offset + (uint64_t)t),
                          (uint32_t)kind, (uint32_t)idx);
}

B2H_DEVINL double draw_z(const RngView& r, int c, int t, int j, int d) {
    if (r.mode == 1) return r.z[((i64)c * r.n_injected + t) * d + j];
    return philox_normal(r.key, r.chain_offset + (uint64_t)c, (uint32_t)(r.transition_offset + (uint64_t)t),
                         (uint32_t)j);
}

// ---------------------------------------------------------------------------
// per-chain context
// ---------------------------------------------------------------------------
template <typename T, int G>
struct Chain {
    const EngineView<T>& v;
    int c, lane;
    i64 base;        // c*sc
    i64 ckbase;      // c*sck
    double* red;     // smem scratch for block groups
    ChainRec r;      // register copy

    B2H_DEVINL Chain(const EngineView<T>& v_, int c_, double* red_)
        : v(v_), c(c_), lane(Group<G>::lane()), base((i64)c_ * v_.sc), ckbase((i64)c_ * v_.sck), red(red_) {}

    B2H_DEVINL i64 at(int j) const { return base + (i64)j * v.sj; }
    B2H_DEVINL i64 ck(int level, int j) const { return ckbase + ((i64)level * v.d + j) * v.sj; }
    B2H_DEVINL T imm(int j) const { return v.imm[(i64)c * v.imm_sc + (i64)j * v.imm_sj]; }
    B2H_DEVINL void load() { r = v.rec[c]; }
    B2H_DEVINL void store() {
        if (lane == 0) v.rec[c] = r;
    }
};

// ---------------------------------------------------------------------------
// The integration front (q, p, dU/dq of the edge being extended).  MemFront reads and writes the edge arrays
// every tick (split engine, dense metric, large rows).  RegFront keeps the front in registers of the chain's
// group for the whole sub-tree and writes it back only at sub-tree ends, which removes 8 of the ~15 row
// accesses per leapfrog of the persistent fused kernel.  Element e of lane l is coordinate j = l + e*G.
// ---------------------------------------------------------------------------
template <typename T>
struct MemFront {
    static constexpr int kE = 1 << 30;
    static constexpr bool kRegs = false;
    T *Q, *P, *Gd;
    template <int G> B2H_DEVINL void bind(const Chain<T, G>& ch) {
        Q = ch.r.go_right ? ch.v.qr : ch.v.ql;
        P = ch.r.go_right ? ch.v.pr : ch.v.pl;
        Gd = ch.r.go_right ? ch.v.gr : ch.v.gl;
    }
    template <int G> B2H_DEVINL void flush(const Chain<T, G>&) {}
    B2H_DEVINL T q(int, i64 a) const { return Q[a]; }
    B2H_DEVINL T p(int, i64 a) const { return P[a]; }
    B2H_DEVINL T g(int, i64 a) const { return Gd[a]; }
    B2H_DEVINL void set_q(int, i64 a, T x) { Q[a] = x; }
    B2H_DEVINL void set_p(int, i64 a, T x) { P[a] = x; }
    B2H_DEVINL void set_g(int, i64 a, T x) { Gd[a] = x; }
};

template <typename T, int E>
struct RegFront {
    static constexpr int kE = E;
    static constexpr bool kRegs = true;
    T fq[E], fp[E], fg[E];
    template <int G> B2H_DEVINL void bind(const Chain<T, G>& ch) {
        const T* Q = ch.r.go_right ? ch.v.qr : ch.v.ql;
        const T* P = ch.r.go_right ? ch.v.pr : ch.v.pl;
        const T* Gd = ch.r.go_right ? ch.v.gr : ch.v.gl;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int j = ch.lane + e * G;
            if (j < ch.v.d) { i64 a = ch.at(j); fq[e] = Q[a]; fp[e] = P[a]; fg[e] = Gd[a]; }
            else { fq[e] = 0; fp[e] = 0; fg[e] = 0; }
        }
    }
    template <int G> B2H_DEVINL void flush(const Chain<T, G>& ch) {
        T* Q = ch.r.go_right ? ch.v.qr : ch.v.ql;
        T* P = ch.r.go_right ? ch.v.pr : ch.v.pl;
        T* Gd = ch.r.go_right ? ch.v.gr : ch.v.gl;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int j = ch.lane + e * G;
            if (j < ch.v.d) { i64 a = ch.at(j); Q[a] = fq[e]; P[a] = fp[e]; Gd[a] = fg[e]; }
        }
    }
    B2H_DEVINL T q(int e, i64) const { return fq[e]; }
    B2H_DEVINL T p(int e, i64) const { return fp[e]; }
    B2H_DEVINL T g(int e, i64) const { return fg[e]; }
    B2H_DEVINL void set_q(int e, i64, T x) { fq[e] = x; }
    B2H_DEVINL void set_p(int e, i64, T x) { fp[e] = x; }
    B2H_DEVINL void set_g(int e, i64, T x) { fg[e] = x; }
};

// loop over this lane's coordinates: fully unrolled for a register front, a plain strided loop otherwise
#define B2H_ELEMS(Front, e, j, lane, d, G) \
    _Pragma("unroll") for (int e = 0, j = (lane); e < Front::kE && j < (d); ++e, j += (G))

// ---------------------------------------------------------------------------
// sub-tree start: direction draw (trajectory.py:516-518)
// ---------------------------------------------------------------------------
template <typename T, int G>
B2H_DEVINL void begin_subtree(Chain<T, G>& ch) {
    double u = draw_u(ch.v.rng, DRAW_DIR, ch.c, ch.r.t, ch.r.k, ch.v.maxd);
    ch.r.go_right = bern(u, 0.5) ? 1 : 0;
    ch.r.s = 0;
}

// ---------------------------------------------------------------------------
// transition start (nuts.py:113-135): momentum, initial energy, edges = start
// state, proposal = start state with weight 0 / sum_log_p_accept -inf.
// Diagonal metric family: p = sqrt(1/imm) z (metrics.py:46,50,67).
// Dense: (p0, v0) were produced by the momentum GEMMs into compact row mom_slot.
// ---------------------------------------------------------------------------
template <typename T, int G, bool DENSE, bool NUTS = true>
B2H_DEVINL void begin_transition(Chain<T, G>& ch) {
    const EngineView<T>& v = ch.v;
    T kacc = 0;
    for (int j = ch.lane; j < v.d; j += G) {
        i64 a = ch.at(j);
        T p0, vel;
        if (DENSE) {
            i64 m = (i64)ch.c * v.d + j;
            p0 = v.mom_p[m];
            vel = v.mom_v[m];
            v.vl[a] = vel;
            v.vr[a] = vel;
            T w0 = v.wp[a];
            v.wl[a] = w0;
            v.wr[a] = w0;
        } else {
            T im = ch.imm(j);
            T z = (T)draw_z(v.rng, ch.c, ch.r.t, j, v.d);
            p0 = sqrt((T)1 / im) * z;
            vel = im * p0;
        }
        T q0 = v.qp[a], g0 = v.gp[a];
        v.ql[a] = q0; v.qr[a] = q0;
        v.pl[a] = p0; v.pr[a] = p0;
        v.gl[a] = g0; v.gr[a] = g0;
        v.pp[a] = p0;
        v.msum[a] = p0;
        kacc += vel * p0;
    }
    T K0 = (T)0.5 * (T)Group<G>::sum1((double)kacc, ch.red);
    T E0 = (T)ch.r.U_prop + K0;                       // nuts.py:117-119
    ch.r.E0 = (double)E0;
    ch.r.U_left = ch.r.U_prop; ch.r.U_right = ch.r.U_prop;
    ch.r.E_prop = (double)E0;
    ch.r.w_prop = 0.0;                                // nuts.py:123
    ch.r.slpa_prop = -INFINITY;                       // nuts.py:124
    ch.r.imin = 0; ch.r.imax = 0;                     // termination.py:63-83
    ch.r.k = 0;
    ch.r.nleap = 0;
    ch.r.phase = PH_RUN;
    if (DENSE) {
        // queue the momentum of the NEXT transition (consumed at the earliest two ticks from now)
        int slot = 0;
        if (ch.lane == 0) {
            slot = atomicAdd(v.mom_count + v.mom_parity, 1);
            v.mom_list[(i64)v.mom_parity * v.C + slot] = ch.c;
        }
        slot = Group<G>::bcast(slot, ch.red);
        const int tn = ch.r.t + 1;
        const bool have = (v.rng.mode == 0) || (tn < v.rng.n_injected);
        T* zrow = v.mom_z + ((i64)v.mom_parity * v.C + slot) * v.d;
        for (int j = ch.lane; j < v.d; j += G) zrow[j] = have ? (T)draw_z(v.rng, ch.c, tn, j, v.d) : (T)0;
    }
    if (NUTS) begin_subtree(ch);
}

// ---------------------------------------------------------------------------
// first half of velocity Verlet (integrators.py:59-62) in place on the edge:
//   p_half = p - (0.5*e) g ;  q' = q + e * (imm p_half)
// Dense metric: imm p_half = v - (0.5*e) w with v = imm p and w = imm g carried along with the state.
// ---------------------------------------------------------------------------
template <typename T, int G, bool DENSE, bool SPLIT, class Front>
B2H_DEVINL void half_kick_drift(Chain<T, G>& ch, Front& f) {
    const EngineView<T>& v = ch.v;
    T* V = ch.r.go_right ? v.vr : v.vl;     // dense only
    T* W = ch.r.go_right ? v.wr : v.wl;     // dense only
    T e = (T)(ch.r.go_right ? ch.r.eps : -ch.r.eps);
    T he = (T)0.5 * e;
    B2H_ELEMS(Front, ee, j, ch.lane, v.d, G) {
        i64 a = ch.at(j);
        T ph = f.p(ee, a) - he * f.g(ee, a);
        f.set_p(ee, a, ph);
        T vh;
        if (DENSE) {
            vh = V[a] - he * W[a];           // imm.(p - h g) by linearity
            V[a] = vh;
        } else {
            vh = ch.imm(j) * ph;
        }
        T qn = f.q(ee, a) + e * vh;
        f.set_q(ee, a, qn);
        if (SPLIT) v.xa[(i64)ch.c * v.d + j] = qn;
    }
}

// ---------------------------------------------------------------------------
// warm-up adaptation at the end of transition `step` (window_adaptation.py:194-215),
// per chain: dual averaging (algorithms.py:104-115, step_size.py:97), Welford
// (algorithms.py:187-197), slow-window end (window_adaptation.py:165-190,
// mass_matrix.py:103-116) and the final averaged step size (:185-190).
// ---------------------------------------------------------------------------
template <typename T, int G>
B2H_DEVINL void adapt_update(Chain<T, G>& ch, int step, double p_accept) {
    const EngineView<T>& v = ch.v;
    const AdaptView& ad = v.adapt;
    const int c = ch.c;
    i64 dstep = ad.da_step[c];
    double x_old = ad.da_x[c], xavg = ad.da_x_avg[c], gavg = ad.da_g_avg[c], mu = ad.da_mu[c];
    double grad = ad.target - p_accept;
    double eta = 1.0 / ((double)dstep + ad.t0);
    double new_gavg = (1.0 - eta) * gavg + eta * grad;
    double new_x = mu - (sqrt((double)dstep) / ad.gamma) * new_gavg;
    double x_eta = pow((double)dstep, -ad.kappa);
    double new_xavg = x_eta * x_old + (1.0 - x_eta) * xavg;     // Q16: OLD iterate
    dstep += 1;
    double eps = exp(new_x);

    const bool slow = ad.stage[step] != 0;
    const bool wend = ad.window_end[step] != 0;
    T* mean = (T*)ad.wc_mean;
    T* m2 = (T*)ad.wc_m2;
    i64 n = ad.wc_n[c];
    if (slow) {
        n += 1;
        for (int j = ch.lane; j < v.d; j += G) {
            i64 a = ch.at(j);
            T val = v.qp[a];
            T delta = val - mean[a];
            T mn = mean[a] + delta / (T)n;
            T ud = val - mn;
            mean[a] = mn;
            m2[a] = m2[a] + ud * delta;
        }
    }
    if (wend) {
        // imm = (n/(n+5)) * m2/(n-1) + 1e-3 * (5/(n+5));  Welford re-init; da re-init with mu = step size
        T scale = (T)((double)n / ((double)n + 5.0));
        T shrink = (T)(1e-3 * (5.0 / ((double)n + 5.0)));
        for (int j = ch.lane; j < v.d; j += G) {
            i64 a = ch.at(j);
            T cov = m2[a] / (T)(n - 1);
            v.imm[(i64)c * v.imm_sc + (i64)j * v.imm_sj] = scale * cov + shrink;
            mean[a] = 0;
            m2[a] = 0;
        }
        n = 0;
        mu = eps;
        dstep = 1; new_x = 0.0; new_xavg = 0.0; new_gavg = 0.0;
    }
    if (step == ad.num_steps - 1) eps = exp(new_xavg);
    Group<G>::sync();
    if (ch.lane == 0) {
        ad.da_step[c] = dstep; ad.da_x[c] = new_x; ad.da_x_avg[c] = new_xavg; ad.da_g_avg[c] = new_gavg;
        ad.da_mu[c] = mu;
        ad.wc_n[c] = n;
    }
    ch.r.eps = eps;
}

// ---------------------------------------------------------------------------
// end of a transition: publish Diagnostics (nuts.py:138-151), store the draw,
// adapt, and arm the next transition.
// ---------------------------------------------------------------------------
template <typename T, int G>
B2H_DEVINL void end_transition(Chain<T, G>& ch, int num_doublings, bool is_turning, bool is_diverging) {
    const EngineView<T>& v = ch.v;
    const int t = ch.r.t;
    const int tl = t - ch.r.t_base;              // index within this call
    ch.r.last_nd = num_doublings;
    ch.r.last_flags = (is_turning ? 1 : 0) | (is_diverging ? 2 : 0) | (ch.r.sub_term ? 4 : 0);
    ch.r.last_nleap = ch.r.nleap;
    if (tl < v.out.n_store) {
        if (v.out.draws) {
            T* dr = (T*)v.out.draws + ((i64)tl * v.C + ch.c) * v.d;
            for (int j = ch.lane; j < v.d; j += G) dr[j] = v.qp[ch.at(j)];
        }
        if (v.out.draw_stats && ch.lane == 0) {
            double* ds = v.out.draw_stats + ((i64)tl * v.C + ch.c) * 4;
            ds[0] = ch.r.accept_prob; ds[1] = (double)num_doublings; ds[2] = (double)ch.r.nleap;
            ds[3] = (double)ch.r.last_flags;
        }
    }
    if (v.adapt.enabled && t < v.adapt.num_steps) adapt_update(ch, t, ch.r.accept_prob);
    ch.r.t = t + 1;
    ch.r.phase = (v.n_transitions > 0 && (ch.r.t - ch.r.t_base) >= v.n_transitions) ? PH_DONE : PH_START;
}

// ---------------------------------------------------------------------------
// second half of the leapfrog + everything the reference does per integration
// step.  Inputs: the edge holds (q', p_half), the new gradient is in the edge's g
// (fused) or in xb (split), the new potential is U_new.
// DENSE: xb holds g' and xc = imm g' (the one metric contraction of the tick); the edge holds p_half and
// V = imm p_half.
// ---------------------------------------------------------------------------
// Returns true when the tick ended a sub-tree (the front was written back and must be re-bound).
template <typename T, int G, bool DENSE, bool SPLIT, class Front>
B2H_DEVINL bool post_gradient(Chain<T, G>& ch, T U_new, Front& f) {
    const EngineView<T>& v = ch.v;
    ChainRec& r = ch.r;
    const int d = v.d;
    T* V = r.go_right ? v.vr : v.vl;     // dense only
    T* W = r.go_right ? v.wr : v.wl;     // dense only
    const T e = (T)(r.go_right ? r.eps : -r.eps);
    const T he = (T)0.5 * e;

    // ---- one pass over the chain's vectors, one group reduction --------------------------------
    //  p' = p_half - (0.5 e) g' (integrators.py:66), K(p') (metrics.py:70-73), sub-tree momentum sum
    //  (trajectory.py:243,278), checkpoint write on even steps at the step's storage index
    //  (termination.py:109-124; stale index at step 0, Q2/Q3) and the U-turn dot products of the first LV
    //  checkpoint levels (termination.py:164-187 with metrics.py:95-102).  Which levels are read depends
    //  only on the step number, so nothing here waits for the energy.
    const int s = r.s, k = r.k;
    int imin, imax;
    if (s == 0) { imin = r.imin; imax = r.imax; }
    else storage_indices(s, imin, imax);
    r.imin = imin; r.imax = imax;
    const bool even = (s & 1) == 0;
    const int nlev = (s >= 1 && imax >= imin) ? (imax - imin + 1) : 0;
    constexpr int LV = 4;
    T kacc = 0, dl[LV], dr[LV];
#pragma unroll
    for (int l = 0; l < LV; ++l) { dl[l] = 0; dr[l] = 0; }
    B2H_ELEMS(Front, ee, j, ch.lane, d, G) {
        i64 a = ch.at(j);
        T p, vel, im = 0;
        if (DENSE) {
            T g = v.xb[(i64)ch.c * d + j], wv = v.xc[(i64)ch.c * d + j];
            f.set_g(ee, a, g);
            W[a] = wv;
            p = f.p(ee, a) - he * g;
            f.set_p(ee, a, p);
            vel = V[a] - he * wv;            // imm p' = imm p_half - (0.5 e) imm g'
            V[a] = vel;
        } else {
            T g;
            if (SPLIT) { g = v.xb[(i64)ch.c * d + j]; f.set_g(ee, a, g); }
            else g = f.g(ee, a);
            p = f.p(ee, a) - he * g;
            f.set_p(ee, a, p);
            im = ch.imm(j);
            vel = im * p;
        }
        kacc += vel * p;
        T sm = (s == 0) ? p : v.sms[a] + p;
        v.sms[a] = sm;
        if (even) {
            i64 b = ch.ck(imax, j);
            v.mck[b] = p;
            v.sckp[b] = sm;
            if (DENSE) v.vck[b] = vel;
        }
#pragma unroll
        for (int l = 0; l < LV; ++l) {
            if (l < nlev) {
                i64 b = ch.ck(imax - l, j);
                T m = v.mck[b], sc = v.sckp[b];
                T subsum = sm - sc + m;
                T rho = subsum - (p + m) / (T)2;
                T vleft = DENSE ? v.vck[b] : im * m;
                dl[l] += vleft * rho;
                dr[l] += vel * rho;
            }
        }
    }
    double red[1 + 2 * LV];
    red[0] = (double)kacc;
#pragma unroll
    for (int l = 0; l < LV; ++l) { red[1 + 2 * l] = (double)dl[l]; red[2 + 2 * l] = (double)dr[l]; }
    Group<G>::template sum<1 + 2 * LV>(red, ch.red);
    const T K = (T)0.5 * (T)red[0];
    bool term = false;
#pragma unroll
    for (int l = 0; l < LV; ++l)
        if (l < nlev && ((T)red[1 + 2 * l] <= (T)0 || (T)red[2 + 2 * l] <= (T)0)) term = true;
    if (!term && nlev > LV) {                       // deeper levels (a step with >= 5 trailing one-bits): rare
        for (int i = imax - LV; i >= imin; --i) {
            T xl = 0, xr = 0;
            B2H_ELEMS(Front, ee, j, ch.lane, d, G) {
                i64 a = ch.at(j), b = ch.ck(i, j);
                T m = v.mck[b], sc = v.sckp[b], p = f.p(ee, a), sm = v.sms[a];
                T subsum = sm - sc + m;
                T rho = subsum - (p + m) / (T)2;
                T vleft, vright;
                if (DENSE) { vleft = v.vck[b]; vright = V[a]; }
                else { T im = ch.imm(j); vleft = im * m; vright = im * p; }
                xl += vleft * rho;
                xr += vright * rho;
            }
            double r2[2] = {(double)xl, (double)xr};
            Group<G>::template sum<2>(r2, ch.red);
            if ((T)r2[0] <= (T)0 || (T)r2[1] <= (T)0) { term = true; break; }
        }
    }

    // ---- proposal for the new state (proposals.py:41-52, Q10)
    const T E = U_new + K;
    double delta = (double)((T)r.E0 - E);
    if (isnan(delta)) delta = -INFINITY;
    const bool div = fabs(delta) > v.div_thr;
    const double w_new = delta;
    const double lpa = delta > 0 ? 0.0 : delta;

    // ---- progressive uniform sampling inside the sub-tree (proposals.py:96-100, Q7)
    bool take;
    if (s == 0) {
        take = true;                                   // trajectory.py:276-277: proposal = first state
        r.w_sub = w_new; r.slpa_sub = lpa;
    } else {
        double pa = expit(w_new - r.w_sub);
        if (isnan(pa)) pa = 0.0;
        double u = draw_u(v.rng, DRAW_UNIFORM, ch.c, r.t, uniform_slot(k, s), v.maxd);
        take = bern(u, pa);
        r.w_sub = lae(r.w_sub, w_new);                 // proposals.py:141-144
        r.slpa_sub = lae(r.slpa_sub, lpa);
    }
    if (take) {
        r.E_sub = (double)E; r.U_sub = (double)U_new;
        B2H_ELEMS(Front, ee, j, ch.lane, d, G) {
            i64 a = ch.at(j);
            v.qs[a] = f.q(ee, a); v.ps[a] = f.p(ee, a); v.gs[a] = f.g(ee, a);
            if (DENSE) v.ws[a] = W[a];
        }
    }
    r.sub_len = (s == 0) ? 1 : r.sub_len + 1;
    r.nleap += 1;
    r.total_leap += 1;
    r.U_front = (double)U_new;

    const int sub_limit = v.sub_max_steps > 0 ? v.sub_max_steps : (1 << k);
    const bool end_sub = div || term || (s == sub_limit);      // Q1: 2**k more steps after step 0
    if (!end_sub) { r.s = s + 1; return false; }
    f.flush(ch);                                                // the edge arrays must be current from here on
    Group<G>::sync();
    if (r.go_right) r.U_right = r.U_front; else r.U_left = r.U_front;
    r.sub_term = term ? 1 : 0;
    if (v.stop_at_subtree_end) {                                // trajectory.dynamic_integration.integrate on its own
        r.last_flags = (div ? 2 : 0) | (term ? 4 : 0);
        r.last_nleap = r.sub_len;
        r.phase = PH_DONE;
        return true;
    }

    // ================= end of the sub-tree: expand_once (trajectory.py:537-608) =================
    // edges are already in place; msum += sub-tree sum; top-level U-turn on (left, right, msum)
    T tl = 0, tr = 0;
    for (int j = ch.lane; j < d; j += G) {
        i64 a = ch.at(j);
        T ms = v.msum[a] + v.sms[a];
        v.msum[a] = ms;
        T plv = v.pl[a], prv = v.pr[a];
        T rho = ms - (prv + plv) / (T)2;
        T vleft, vright;
        if (DENSE) { vleft = v.vl[a]; vright = v.vr[a]; }
        else { T im = ch.imm(j); vleft = im * plv; vright = im * prv; }
        tl += vleft * rho;
        tr += vright * rho;
    }
    double red2[2] = {(double)tl, (double)tr};
    Group<G>::template sum<2>(red2, ch.red);
    const bool top_turn = ((T)red2[0] <= (T)0) || ((T)red2[1] <= (T)0);

    r.accept_prob = exp(r.slpa_sub) / (double)r.sub_len;        // Q9 (trajectory.py:551-553)

    // biased progressive sampling is always drawn (Q8, proposals.py:130-131)
    double diff = r.w_sub - r.w_prop;
    double pb = fmin(fmax(exp(diff), 0.0), 1.0);
    double ub = draw_u(v.rng, DRAW_BIASED, ch.c, r.t, k, v.maxd);
    bool accb = bern(ub, pb);
    if (div || term) {
        r.slpa_prop = lae(r.slpa_sub, r.slpa_prop);             // trajectory.py:560-564
    } else {
        if (accb) {
            for (int j = ch.lane; j < d; j += G) {
                i64 a = ch.at(j);
                v.qp[a] = v.qs[a]; v.pp[a] = v.ps[a]; v.gp[a] = v.gs[a];
                if (DENSE) v.wp[a] = v.ws[a];
            }
            r.E_prop = r.E_sub; r.U_prop = r.U_sub;
        }
        r.w_prop = lae(r.w_prop, r.w_sub);
        r.slpa_prop = lae(r.slpa_prop, r.slpa_sub);
    }
    const int nd = k + 1;
    if (div || top_turn || term || nd >= v.maxd) {              // trajectory.py:577 / scan length
        end_transition(ch, nd, top_turn, div);
    } else {
        r.k = nd;
        begin_subtree(ch);
    }
    return true;
}

// ---------------------------------------------------------------------------
// HMC (hmc.py:110-123,157-204): momentum, L leapfrogs in place on the "right"
// edge, flip, Metropolis accept (Q15: divergent transitions are not force-rejected).
// ---------------------------------------------------------------------------
template <typename T, int G, bool DENSE>
B2H_DEVINL void hmc_begin(Chain<T, G>& ch) {
    begin_transition<T, G, DENSE, false>(ch);   // edges = state + fresh momentum, E0
    ch.r.go_right = 1;
    ch.r.hmc_step = 0;
}

template <typename T, int G, bool DENSE, bool SPLIT, class Front>
B2H_DEVINL bool hmc_post(Chain<T, G>& ch, T U_new, Front& f) {
    const EngineView<T>& v = ch.v;
    ChainRec& r = ch.r;
    const int d = v.d;
    const T he = (T)0.5 * (T)r.eps;
    T kacc = 0;
    B2H_ELEMS(Front, ee, j, ch.lane, d, G) {
        i64 a = ch.at(j);
        if (DENSE) {
            T g = v.xb[(i64)ch.c * d + j], wv = v.xc[(i64)ch.c * d + j];
            f.set_g(ee, a, g);
            v.wr[a] = wv;
            T p = f.p(ee, a) - he * g;
            f.set_p(ee, a, p);
            T vel = v.vr[a] - he * wv;
            v.vr[a] = vel;
            kacc += vel * p;
        } else {
            T g;
            if (SPLIT) { g = v.xb[(i64)ch.c * d + j]; f.set_g(ee, a, g); }
            else g = f.g(ee, a);
            T p = f.p(ee, a) - he * g;
            f.set_p(ee, a, p);
            kacc += (ch.imm(j) * p) * p;
        }
    }
    r.hmc_step += 1;
    r.nleap += 1;
    r.total_leap += 1;
    if (r.hmc_step < v.hmc_L) return false;     // uniform across the group: no reduction skipped unevenly
    const T K = (T)0.5 * (T)Group<G>::sum1((double)kacc, ch.red);   // K(-p) == K(p)
    const T E = U_new + K;
    double delta = (double)((T)r.E0 - E);
    if (isnan(delta)) delta = -INFINITY;
    const bool div = fabs(delta) > v.div_thr;
    double p_accept = fmin(fmax(exp(delta), 0.0), 1.0);
    double u = draw_u(v.rng, DRAW_ACCEPT, ch.c, r.t, 0, v.maxd);
    bool acc = bern(u, p_accept);
    B2H_ELEMS(Front, ee, j, ch.lane, d, G) {
        i64 a = ch.at(j);
        if (acc) {
            v.qp[a] = f.q(ee, a); v.pp[a] = -f.p(ee, a); v.gp[a] = f.g(ee, a);
            if (DENSE) v.wp[a] = v.wr[a];
        }
        // on reject the state keeps (q, fresh momentum, g): pp already holds p0 (hmc.py:122,195)
    }
    if (acc) r.U_prop = (double)U_new;
    r.accept_prob = p_accept;
    end_transition(ch, 0, false, div);
    return true;
}

}  // namespace b2h
