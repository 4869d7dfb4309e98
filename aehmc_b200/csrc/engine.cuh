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

// elements per lane whose loads are issued together before the first store of the batch
constexpr int kBatch = 4;

// Per-chain scalar record (one 256-byte line per chain).  ChainRecLive is the part that carries state (192 bytes:
// what the tile kernel loads and stores); the rest of the line is padding.
struct __align__(16) ChainRecLive {
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
};
static_assert(sizeof(ChainRecLive) == 192, "ChainRecLive is the first 192 bytes of the record");
struct __align__(16) ChainRec : ChainRecLive {
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
    int pooled;              // 1: dual averaging only; the caller re-estimates the shared metric from pooled statistics
    int step_offset;         // schedule index of this call's first transition
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
    int thin;                // keep every thin-th transition of the call (<= 1: all)
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
    // Gaussian targets: U = 0.5 (q' - u_center) . g', which the tile tick kernel's pass A forms from the rows it reads
    // anyway instead of a separate kernel over q' and g' (nullptr: U comes from Unew)
    const T* u_center;
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
    int exact_doubling;                  // 1: sub-trees of 2**k leapfrogs instead of the reference's 2**k + 1 (Q1)
    int stop_at_subtree_end;             // stand-alone dynamic_integration: stop instead of running expand_once
    i64* counters;                       // [0] leapfrogs [1] transitions [2] ticks [3] active chain-ticks
};

// ---------------------------------------------------------------------------
// draws
// ---------------------------------------------------------------------------
template <int G = 32>
B2H_DEVINL double draw_u(const RngView& r, int kind, int c, int t, int idx, int maxd) {
    if (r.mode == 1) {
        i64 row = (i64)c * r.n_injected + t + (i64)r.transition_offset;
        if (kind == DRAW_DIR) return r.u_dir[row * maxd + idx];
        if (kind == DRAW_BIASED) return r.u_biased[row * maxd + idx];
        if (kind == DRAW_UNIFORM) return r.u_uniform[row * (((i64)1 << maxd) - 1) + idx];
        return r.u_accept[row];
    }
    if (Helpers<G>::kCall)
        return philox_uniform(r.key, r.chain_offset + (uint64_t)c, (uint32_t)(r.transition_offset + (uint64_t)t),
                              (uint32_t)kind, (uint32_t)idx);
    return philox_uniform_inl(r.key, r.chain_offset + (uint64_t)c, (uint32_t)(r.transition_offset + (uint64_t)t),
                              (uint32_t)kind, (uint32_t)idx);
}

template <int G = 32>
B2H_DEVINL double draw_z(const RngView& r, int c, int t, int j, int d) {
    if (r.mode == 1) return r.z[((i64)c * r.n_injected + t + (i64)r.transition_offset) * d + j];
    if (Helpers<G>::kCall)
        return philox_normal(r.key, r.chain_offset + (uint64_t)c, (uint32_t)(r.transition_offset + (uint64_t)t),
                             (uint32_t)j);
    return philox_normal_inl(r.key, r.chain_offset + (uint64_t)c, (uint32_t)(r.transition_offset + (uint64_t)t),
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
    // U-turn checkpoints of this chain (termination.py:63-131): rows in the workspace, or -- persistent kernel, when they
    // fit -- a shared-memory copy staged for the whole launch (stage_checkpoints); element (level, j) is at ck(level, j)
    T *mck, *sckp, *vck;
    i64 ck_ls, ck_sj;
    ChainRec r;      // register copy

    B2H_DEVINL Chain(const EngineView<T>& v_, int c_, double* red_)
        : v(v_), c(c_), lane(Group<G>::lane()), base((i64)c_ * v_.sc), ckbase((i64)c_ * v_.sck), red(red_),
          mck(v_.mck + (i64)c_ * v_.sck), sckp(v_.sckp + (i64)c_ * v_.sck),
          vck(v_.vck ? v_.vck + (i64)c_ * v_.sck : nullptr), ck_ls((i64)v_.d * v_.sj), ck_sj(v_.sj) {}

    // groups of several threads always work on the row-major layout (element stride 1): no stride multiplies there
    B2H_DEVINL i64 at(int j) const {
        if constexpr (G > 1) return base + j;
        else return base + (i64)j * v.sj;
    }
    B2H_DEVINL i64 ck(int level, int j) const {
        if constexpr (G > 1) return (i64)(level * v.d + j);
        else return (i64)level * ck_ls + (i64)j * ck_sj;
    }
    // checkpoints [2][maxd][d] of this chain into / out of shared memory (diagonal-family metrics: no vck)
    B2H_DEVINL void stage_checkpoints(T* smem) {
        const int n = v.maxd * v.d;
        for (int k = lane; k < n; k += G) {
            const i64 src = (i64)(k / v.d) * ck_ls + (i64)(k % v.d) * ck_sj;
            smem[k] = mck[src];
            smem[n + k] = sckp[src];
        }
        mck = smem; sckp = smem + n; ck_ls = v.d; ck_sj = 1;
        Group<G>::sync();
    }
    B2H_DEVINL void unstage_checkpoints() {
        Group<G>::sync();
        const int n = v.maxd * v.d;
        T* gm = v.mck + ckbase;
        T* gs = v.sckp + ckbase;
        for (int k = lane; k < n; k += G) {
            const i64 dst = ((i64)(k / v.d) * v.d + (k % v.d)) * v.sj;
            gm[dst] = mck[k];
            gs[dst] = sckp[k];
        }
    }
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
    static constexpr bool kSumRegs = false;     // the sub-tree momentum sum goes through v.sms every tick
    static constexpr bool kImmRegs = false;     // the diagonal metric is read from memory at every use
    B2H_DEVINL T sum(int) const { return 0; }
    B2H_DEVINL void set_sum(int, T) {}
    B2H_DEVINL T im(int) const { return 0; }
    T *Q, *P, *Gd, *V, *W;              // V = imm p, W = imm g: dense metric only
    template <int G> B2H_DEVINL void bind(const Chain<T, G>& ch) {
        Q = ch.r.go_right ? ch.v.qr : ch.v.ql;
        P = ch.r.go_right ? ch.v.pr : ch.v.pl;
        Gd = ch.r.go_right ? ch.v.gr : ch.v.gl;
        V = ch.r.go_right ? ch.v.vr : ch.v.vl;
        W = ch.r.go_right ? ch.v.wr : ch.v.wl;
    }
    template <int G> B2H_DEVINL void flush(const Chain<T, G>&) {}
    B2H_DEVINL T q(int, i64 a) const { return Q[a]; }
    B2H_DEVINL T p(int, i64 a) const { return P[a]; }
    B2H_DEVINL T g(int, i64 a) const { return Gd[a]; }
    B2H_DEVINL T vel(int, i64 a) const { return V[a]; }
    B2H_DEVINL T w(int, i64 a) const { return W[a]; }
    B2H_DEVINL void set_q(int, i64 a, T x) { Q[a] = x; }
    B2H_DEVINL void set_p(int, i64 a, T x) { P[a] = x; }
    B2H_DEVINL void set_g(int, i64 a, T x) { Gd[a] = x; }
    B2H_DEVINL void set_vel(int, i64 a, T x) { V[a] = x; }
    B2H_DEVINL void set_w(int, i64 a, T x) { W[a] = x; }
};

template <typename T, int E>
struct RegFront {
    static constexpr int kE = E;
    static constexpr bool kRegs = true;
    // the sub-tree momentum sum stays in registers too (v.sms is written by flush) unless the front already fills the
    // register file (thread per chain, E >= 10: measured 6 % slower with the extra spills)
    static constexpr bool kSumRegs = (E <= 8);
    // ... and so does this chain's diagonal inverse mass matrix (constant within a transition; re-read by bind)
    static constexpr bool kImmRegs = (E <= 8);
    T fq[E], fp[E], fg[E], fs[kSumRegs ? E : 1], fim[kImmRegs ? E : 1];
    template <int G> B2H_DEVINL void bind(const Chain<T, G>& ch) {
        const T* Q = ch.r.go_right ? ch.v.qr : ch.v.ql;
        const T* P = ch.r.go_right ? ch.v.pr : ch.v.pl;
        const T* Gd = ch.r.go_right ? ch.v.gr : ch.v.gl;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int j = ch.lane + e * G;
            if (j < ch.v.d) { i64 a = ch.at(j); fq[e] = Q[a]; fp[e] = P[a]; fg[e] = Gd[a]; if (kSumRegs) fs[e] = ch.v.sms[a]; if (kImmRegs) fim[e] = ch.imm(j); }
            else { fq[e] = 0; fp[e] = 0; fg[e] = 0; if (kSumRegs) fs[e] = 0; if (kImmRegs) fim[e] = 0; }
        }
    }
    template <int G> B2H_DEVINL void flush(const Chain<T, G>& ch) {
        T* Q = ch.r.go_right ? ch.v.qr : ch.v.ql;
        T* P = ch.r.go_right ? ch.v.pr : ch.v.pl;
        T* Gd = ch.r.go_right ? ch.v.gr : ch.v.gl;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int j = ch.lane + e * G;
            if (j < ch.v.d) { i64 a = ch.at(j); Q[a] = fq[e]; P[a] = fp[e]; Gd[a] = fg[e]; if (kSumRegs) ch.v.sms[a] = fs[e]; }
        }
    }
    B2H_DEVINL T im(int e) const { return fim[kImmRegs ? e : 0]; }
    B2H_DEVINL T sum(int e) const { return fs[kSumRegs ? e : 0]; }
    B2H_DEVINL void set_sum(int e, T x) { fs[kSumRegs ? e : 0] = x; }
    B2H_DEVINL T q(int e, i64) const { return fq[e]; }
    B2H_DEVINL T p(int e, i64) const { return fp[e]; }
    B2H_DEVINL T g(int e, i64) const { return fg[e]; }
    B2H_DEVINL T vel(int, i64) const { return 0; }      // diagonal-family metrics carry no V / W
    B2H_DEVINL T w(int, i64) const { return 0; }
    B2H_DEVINL void set_q(int e, i64, T x) { fq[e] = x; }
    B2H_DEVINL void set_p(int e, i64, T x) { fp[e] = x; }
    B2H_DEVINL void set_g(int e, i64, T x) { fg[e] = x; }
    B2H_DEVINL void set_vel(int, i64, T) {}
    B2H_DEVINL void set_w(int, i64, T) {}
};

// Split engine (one launch per tick part): the front of the tick lives in registers between the second half
// kick of one leapfrog (post) and the first half kick + drift of the next (pre), so each of p (and V for a dense
// metric) is read once and written once per tick.  Inside a sub-tree the edge arrays are NOT kept current:
//   q   the drifted position is written once, to xa (the gradient's input), and read back from there;
//   g,W arrive from the gradient / metric contractions (xb, xc) every tick and are consumed by the next half
//       kick in registers; nothing reads the edge's copy before the sub-tree ends.
// flush() (sub-tree end, or the last tick of a call) writes the whole front back to the edge arrays.
// bind_post loads what the post part needs; bind loads the whole front from the edge after a sub-tree or
// transition switch.  DENSE selects whether V / W exist.
template <typename T, int E, bool DENSE>
struct TickFront {
    static constexpr int kE = E;
    static constexpr bool kRegs = true;
    static constexpr bool kSumRegs = false;
    static constexpr bool kImmRegs = false;
    B2H_DEVINL T sum(int) const { return 0; }
    B2H_DEVINL void set_sum(int, T) {}
    B2H_DEVINL T im(int) const { return 0; }
    T fq[E], fp[E], fg[E], fv[DENSE ? E : 1], fw[DENSE ? E : 1];
    template <int G> B2H_DEVINL void load(const Chain<T, G>& ch, bool all) {
        const bool rt = ch.r.go_right != 0;
        const T* Q = rt ? ch.v.qr : ch.v.ql;
        const T* P = rt ? ch.v.pr : ch.v.pl;
        const T* Gd = rt ? ch.v.gr : ch.v.gl;
        const T* V = rt ? ch.v.vr : ch.v.vl;
        const T* W = rt ? ch.v.wr : ch.v.wl;
        const T* XA = ch.v.xa + (i64)ch.c * ch.v.d;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int j = ch.lane + e * G;
            const bool in = j < ch.v.d;
            const i64 a = in ? ch.at(j) : ch.at(0);
            fq[e] = in ? (all ? Q[a] : XA[in ? j : 0]) : (T)0;
            fp[e] = in ? P[a] : (T)0;
            if (DENSE) fv[e] = in ? V[a] : (T)0;
            fg[e] = (all && in) ? Gd[a] : (T)0;
            if (DENSE) fw[e] = (all && in) ? W[a] : (T)0;
        }
    }
    template <int G> B2H_DEVINL void bind(const Chain<T, G>& ch) { load(ch, true); }
    template <int G> B2H_DEVINL void bind_post(const Chain<T, G>& ch) { load(ch, false); }
    template <int G> B2H_DEVINL void store(const Chain<T, G>& ch, bool all) {
        const bool rt = ch.r.go_right != 0;
        T* Q = rt ? ch.v.qr : ch.v.ql;
        T* P = rt ? ch.v.pr : ch.v.pl;
        T* Gd = rt ? ch.v.gr : ch.v.gl;
        T* V = rt ? ch.v.vr : ch.v.vl;
        T* W = rt ? ch.v.wr : ch.v.wl;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int j = ch.lane + e * G;
            if (j < ch.v.d) {
                const i64 a = ch.at(j);
                P[a] = fp[e];
                if (DENSE) V[a] = fv[e];
                if (all) { Q[a] = fq[e]; Gd[a] = fg[e]; if (DENSE) W[a] = fw[e]; }
            }
        }
    }
    template <int G> B2H_DEVINL void flush(const Chain<T, G>& ch) { store(ch, true); }
    B2H_DEVINL T q(int e, i64) const { return fq[e]; }
    B2H_DEVINL T p(int e, i64) const { return fp[e]; }
    B2H_DEVINL T g(int e, i64) const { return fg[e]; }
    B2H_DEVINL T vel(int e, i64) const { return fv[DENSE ? e : 0]; }
    B2H_DEVINL T w(int e, i64) const { return fw[DENSE ? e : 0]; }
    B2H_DEVINL void set_q(int e, i64, T x) { fq[e] = x; }
    B2H_DEVINL void set_p(int e, i64, T x) { fp[e] = x; }
    B2H_DEVINL void set_g(int e, i64, T x) { fg[e] = x; }
    B2H_DEVINL void set_vel(int e, i64, T x) { fv[DENSE ? e : 0] = x; }
    B2H_DEVINL void set_w(int e, i64, T x) { fw[DENSE ? e : 0] = x; }
};

// loop over this lane's coordinates: fully unrolled for a register front, a plain strided loop otherwise
#define B2H_ELEMS(Front, e, j, lane, d, G) \
    _Pragma("unroll") for (int e = 0, j = (lane); e < Front::kE && j < (d); ++e, j += (G))

// ---------------------------------------------------------------------------
// sub-tree start: direction draw (trajectory.py:516-518)
// ---------------------------------------------------------------------------
template <typename T, int G>
B2H_DEVINL void begin_subtree(Chain<T, G>& ch) {
    double u = draw_u<G>(ch.v.rng, DRAW_DIR, ch.c, ch.r.t, ch.r.k, ch.v.maxd);
    ch.r.go_right = bern(u, 0.5) ? 1 : 0;
    ch.r.s = 0;
}

// ---------------------------------------------------------------------------
// transition start (nuts.py:113-135): momentum, initial energy, edges = start
// state, proposal = start state with weight 0 / sum_log_p_accept -inf.
// Diagonal metric family: p = sqrt(1/imm) z (metrics.py:46,50,67).
// Dense: (p0, v0) were produced by the momentum GEMMs into compact row mom_slot.
// ---------------------------------------------------------------------------
// zs / zready (thread-per-chain persistent kernel): standard normals of THIS transition that were drawn ahead of time,
// one Box-Muller pair per tick, while all lanes of the warp were converged (see fused_run_kernel): element j < zready
// is zs[j * zstride], the others are drawn here.
template <typename T, int G, bool DENSE, bool NUTS = true>
B2H_DEVINL void begin_transition(Chain<T, G>& ch, const double* zs = nullptr, int zready = 0, int zstride = 0) {
    const EngineView<T>& v = ch.v;
    constexpr int CH = DENSE ? kBatch : 1;             // dense metric == split engine
    T kacc = 0;
    for (int jb = ch.lane; jb < v.d; jb += CH * G) {
        T p0[CH], vel[CH], w0[CH], q0[CH], g0[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {                 // loads (and draws) of the batch first
            const int j = jb + i * G;
            p0[i] = 0; vel[i] = 0; w0[i] = 0; q0[i] = 0; g0[i] = 0;
            if (j < v.d) {
                const i64 a = ch.at(j);
                if (DENSE) {
                    const i64 m = (i64)ch.c * v.d + j;
                    p0[i] = v.mom_p[m];
                    vel[i] = v.mom_v[m];
                    w0[i] = v.wp[a];
                } else {
                    const T im = ch.imm(j);
                    const T z = (j < zready) ? (T)zs[j * zstride] : (T)draw_z<G>(v.rng, ch.c, ch.r.t, j, v.d);
                    p0[i] = sqrt((T)1 / im) * z;
                    vel[i] = im * p0[i];
                }
                q0[i] = v.qp[a]; g0[i] = v.gp[a];
            }
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int j = jb + i * G;
            if (j < v.d) {
                const i64 a = ch.at(j);
                if (DENSE) {
                    v.vl[a] = vel[i]; v.vr[a] = vel[i];
                    v.wl[a] = w0[i]; v.wr[a] = w0[i];
                }
                v.ql[a] = q0[i]; v.qr[a] = q0[i];
                v.pl[a] = p0[i]; v.pr[a] = p0[i];
                v.gl[a] = g0[i]; v.gr[a] = g0[i];
                v.pp[a] = p0[i];
                v.msum[a] = p0[i];
                kacc += vel[i] * p0[i];
            }
        }
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
        // queue the momentum of the NEXT transition: (p0, v0) land before the next tick's pre part, so a
        // transition that lasts a single tick (HMC with L = 1, a first-step divergence) finds them ready
        int slot = 0;
        if (ch.lane == 0) {
            slot = atomicAdd(v.mom_count + v.mom_parity, 1);
            v.mom_list[(i64)v.mom_parity * v.C + slot] = ch.c;
        }
        slot = Group<G>::bcast(slot, ch.red);
        const int tn = ch.r.t + 1;
        const bool have = (v.rng.mode == 0) || (tn + (i64)v.rng.transition_offset < v.rng.n_injected);
        T* zrow = v.mom_z + ((i64)v.mom_parity * v.C + slot) * v.d;
        for (int j = ch.lane; j < v.d; j += G) zrow[j] = have ? (T)draw_z<G>(v.rng, ch.c, tn, j, v.d) : (T)0;
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
    T e = (T)(ch.r.go_right ? ch.r.eps : -ch.r.eps);
    T he = (T)0.5 * e;
    constexpr int CH = SPLIT ? kBatch : 1;
    constexpr int NCH = Front::kRegs ? (Front::kE + CH - 1) / CH : (1 << 28);
#pragma unroll
    for (int cix = 0, jb = ch.lane; cix < NCH && jb < v.d; ++cix, jb += CH * G) {
        T pv[CH], gv[CH], vv[CH], wv[CH], qv[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {                     // loads of the batch first (see post_gradient)
            const int ee = cix * CH + i, j = jb + i * G;
            if (ee < Front::kE && j < v.d) {
                const i64 a = ch.at(j);
                pv[i] = f.p(ee, a); gv[i] = f.g(ee, a); qv[i] = f.q(ee, a);
                if (DENSE) { vv[i] = f.vel(ee, a); wv[i] = f.w(ee, a); }
                else vv[i] = Front::kImmRegs ? f.im(ee) : ch.imm(j);
            }
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int ee = cix * CH + i, j = jb + i * G;
            if (ee < Front::kE && j < v.d) {
                const i64 a = ch.at(j);
                const T ph = pv[i] - he * gv[i];
                f.set_p(ee, a, ph);
                T vh;
                if (DENSE) {
                    vh = vv[i] - he * wv[i];                 // imm.(p - h g) by linearity
                    f.set_vel(ee, a, vh);
                } else {
                    vh = vv[i] * ph;
                }
                const T qn = qv[i] + e * vh;
                f.set_q(ee, a, qn);
                if (SPLIT) v.xa[(i64)ch.c * v.d + j] = qn;
            }
        }
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

    const bool slow = ad.stage[step] != 0 && !ad.pooled;
    const bool wend = ad.window_end[step] != 0;
    T* mean = (T*)ad.wc_mean;
    T* m2 = (T*)ad.wc_m2;
    i64 n = ad.pooled ? 0 : ad.wc_n[c];
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
        // (pooled: the caller rebuilds the shared metric between two calls; only the dual averaging restarts here)
        if (!ad.pooled) {
            T scale = (T)((double)n / ((double)n + 5.0));
            T shrink = (T)(1e-3 * (5.0 / ((double)n + 5.0)));
            for (int j = ch.lane; j < v.d; j += G) {
                i64 a = ch.at(j);
                T cov = m2[a] / (T)(n - 1);
                v.imm[(i64)c * v.imm_sc + (i64)j * v.imm_sj] = scale * cov + shrink;
                mean[a] = 0;
                m2[a] = 0;
            }
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
        if (!ad.pooled) ad.wc_n[c] = n;
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
    const int thin = v.out.thin > 1 ? v.out.thin : 1;
    const int slot = tl / thin;
    if (slot < v.out.n_store && slot * thin == tl) {
        if (v.out.draws) {
            T* dr = (T*)v.out.draws + ((i64)slot * v.C + ch.c) * v.d;
            for (int j = ch.lane; j < v.d; j += G) dr[j] = v.qp[ch.at(j)];
        }
        if (v.out.draw_stats && ch.lane == 0) {
            double* ds = v.out.draw_stats + ((i64)slot * v.C + ch.c) * 4;
            ds[0] = ch.r.accept_prob; ds[1] = (double)num_doublings; ds[2] = (double)ch.r.nleap;
            ds[3] = (double)ch.r.last_flags;
        }
    }
    if (v.adapt.enabled && t + v.adapt.step_offset < v.adapt.num_steps)
        adapt_update(ch, t + v.adapt.step_offset, ch.r.accept_prob);
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
// DEFER (thread-per-chain persistent kernel): a step that ends a sub-tree returns kSubtreeEnd | flags WITHOUT running the
// sub-tree end; the caller runs subtree_end() later, when enough lanes of its warp wait for that path (the outcome does
// not depend on when it runs: nothing else touches the chain in between).
enum : int { kSubtreeEnd = 4, kEndDiv = 1, kEndTerm = 2 };
template <typename T, int G, bool DENSE, bool SPLIT, class Front>
B2H_DEVINL bool subtree_end(Chain<T, G>& ch, Front& f, bool div, bool term);

template <typename T, int G, bool DENSE, bool SPLIT, class Front, bool DEFER = false>
B2H_DEVINL int post_gradient(Chain<T, G>& ch, T U_new, Front& f) {
    const EngineView<T>& v = ch.v;
    ChainRec& r = ch.r;
    const int d = v.d;
    const T e = (T)(r.go_right ? r.eps : -r.eps);
    const T he = (T)0.5 * e;

    // ---- one pass over the chain's vectors, one group reduction --------------------------------
    //  p' = p_half - (0.5 e) g' (integrators.py:66), K(p') (metrics.py:70-73), sub-tree momentum sum
    //  (trajectory.py:243,278), checkpoint write on even steps at the step's storage index
    //  (termination.py:109-124; stale index at step 0, Q2/Q3) and the U-turn dot products of the first LV
    //  checkpoint levels (termination.py:164-187 with metrics.py:95-102).  Which levels are read depends
    //  only on the step number, so nothing here waits for the energy.
    const int s = r.s, k = r.k;
    // the step's progressive-sampling uniform does not depend on the energies: drawn first, so that its Philox rounds
    // overlap the row loads below instead of extending the scalar chain after the reduction
    const double u_step = (s != 0) ? draw_u<G>(v.rng, DRAW_UNIFORM, ch.c, r.t, uniform_slot(k, s), v.maxd) : 0.0;
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
    // Loads are issued CH elements at a time BEFORE any store of the batch: a store between two loads keeps the
    // compiler from hoisting the second one, which would serialise one memory round trip per element.  The
    // arithmetic per element and the accumulation order are unchanged.
    constexpr int CH = SPLIT ? kBatch : 1;     // the persistent kernel is register-bound: one element at a time
    constexpr int NCH = Front::kRegs ? (Front::kE + CH - 1) / CH : (1 << 28);
    [[maybe_unused]] T smreg[Front::kRegs ? Front::kE : 1];      // sub-tree momentum sum of this lane's elements
    const bool lev0 = nlev > 0;
#pragma unroll
    for (int cix = 0, jb = ch.lane; cix < NCH && jb < d; ++cix, jb += CH * G) {
        T gx[CH], wx[CH], so[CH], pv[CH], vv[CH], imv[CH], cm[CH], cs[CH], cv[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int ee = cix * CH + i, j = jb + i * G;
            if (ee < Front::kE && j < d) {
                const i64 a = ch.at(j), m = (i64)ch.c * d + j;
                pv[i] = f.p(ee, a);
                if (DENSE) { gx[i] = v.xb[m]; wx[i] = v.xc[m]; vv[i] = f.vel(ee, a); }
                else { gx[i] = SPLIT ? v.xb[m] : f.g(ee, a); imv[i] = Front::kImmRegs ? f.im(ee) : ch.imm(j); }
                if (s != 0) so[i] = Front::kSumRegs ? f.sum(ee) : v.sms[a];
                if (lev0) {
                    const i64 b = ch.ck(imax, j);
                    cm[i] = ch.mck[b]; cs[i] = ch.sckp[b];
                    if (DENSE) cv[i] = ch.vck[b];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int ee = cix * CH + i, j = jb + i * G;
            if (ee < Front::kE && j < d) {
                const i64 a = ch.at(j);
                T p, vel;
                if (DENSE) {
                    f.set_g(ee, a, gx[i]);
                    f.set_w(ee, a, wx[i]);
                    p = pv[i] - he * gx[i];
                    f.set_p(ee, a, p);
                    vel = vv[i] - he * wx[i];    // imm p' = imm p_half - (0.5 e) imm g'
                    f.set_vel(ee, a, vel);
                } else {
                    if (SPLIT) f.set_g(ee, a, gx[i]);
                    p = pv[i] - he * gx[i];
                    f.set_p(ee, a, p);
                    vel = imv[i] * p;
                }
                kacc += vel * p;
                const T sm = (s == 0) ? p : so[i] + p;
                if (Front::kSumRegs) f.set_sum(ee, sm); else v.sms[a] = sm;
                if (Front::kRegs) smreg[Front::kRegs ? ee : 0] = sm;
                if (even) {
                    const i64 b = ch.ck(imax, j);
                    ch.mck[b] = p;
                    ch.sckp[b] = sm;
                    if (DENSE) ch.vck[b] = vel;
                }
                if (lev0) {
                    const T subsum = sm - cs[i] + cm[i];
                    const T rho = subsum - (p + cm[i]) / (T)2;
                    const T vleft = DENSE ? cv[i] : imv[i] * cm[i];
                    dl[0] += vleft * rho;
                    dr[0] += vel * rho;
                }
            }
        }
    }
    // levels 1 .. LV-1 (steps with two or more trailing one-bits: a quarter of the ticks)
#pragma unroll
    for (int l = 1; l < LV; ++l) {
        if (l < nlev) {
#pragma unroll
            for (int cix = 0, jb = ch.lane; cix < NCH && jb < d; ++cix, jb += CH * G) {
                T pv[CH], vv[CH], sv[CH], cm[CH], cs[CH], cv[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int ee = cix * CH + i, j = jb + i * G;
                    if (ee < Front::kE && j < d) {
                        const i64 a = ch.at(j), b = ch.ck(imax - l, j);
                        pv[i] = f.p(ee, a);
                        sv[i] = Front::kRegs ? smreg[Front::kRegs ? ee : 0] : v.sms[a];
                        cm[i] = ch.mck[b]; cs[i] = ch.sckp[b];
                        if (DENSE) { cv[i] = ch.vck[b]; vv[i] = f.vel(ee, a); }
                        else { const T im = Front::kImmRegs ? f.im(ee) : ch.imm(j); cv[i] = im * cm[i]; vv[i] = im * pv[i]; }
                    }
                }
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int ee = cix * CH + i, j = jb + i * G;
                    if (ee < Front::kE && j < d) {
                        const T subsum = sv[i] - cs[i] + cm[i];
                        const T rho = subsum - (pv[i] + cm[i]) / (T)2;
                        dl[l] += cv[i] * rho;
                        dr[l] += vv[i] * rho;
                    }
                }
            }
        }
    }
    double red[1 + 2 * LV];
    red[0] = (double)kacc;
#pragma unroll
    for (int l = 0; l < LV; ++l) { red[1 + 2 * l] = (double)dl[l]; red[2 + 2 * l] = (double)dr[l]; }
    // 7 of 8 steps check at most one checkpoint level (the count depends on the step number only, so the whole group
    // agrees): reduce the kinetic energy and that level's two dot products, not all 1 + 2 LV values
    if (nlev <= 1) {
        double r3[3] = {red[0], red[1], red[2]};
        Group<G>::template sum<3>(r3, ch.red);
        red[0] = r3[0]; red[1] = r3[1]; red[2] = r3[2];
    } else {
        Group<G>::template sum<1 + 2 * LV>(red, ch.red);
    }
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
                T m = ch.mck[b], sc = ch.sckp[b], p = f.p(ee, a), sm = Front::kSumRegs ? f.sum(ee) : v.sms[a];
                T subsum = sm - sc + m;
                T rho = subsum - (p + m) / (T)2;
                T vleft, vright;
                if (DENSE) { vleft = ch.vck[b]; vright = f.vel(ee, a); }
                else { T im = Front::kImmRegs ? f.im(ee) : ch.imm(j); vleft = im * m; vright = im * p; }
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
        double pa = expit_g<G>(w_new - r.w_sub);
        if (isnan(pa)) pa = 0.0;
        take = bern(u_step, pa);
        r.w_sub = lae_g<G>(r.w_sub, w_new);                 // proposals.py:141-144
        r.slpa_sub = lae_g<G>(r.slpa_sub, lpa);
    }
    if (take) {
        r.E_sub = (double)E; r.U_sub = (double)U_new;
        B2H_ELEMS(Front, ee, j, ch.lane, d, G) {
            i64 a = ch.at(j);
            v.qs[a] = f.q(ee, a); v.ps[a] = f.p(ee, a); v.gs[a] = f.g(ee, a);
            if (DENSE) v.ws[a] = f.w(ee, a);
        }
    }
    r.sub_len = (s == 0) ? 1 : r.sub_len + 1;
    r.nleap += 1;
    r.total_leap += 1;
    r.U_front = (double)U_new;

    const int sub_limit = v.sub_max_steps > 0 ? v.sub_max_steps : (1 << k) - (v.exact_doubling ? 1 : 0);
    const bool end_sub = div || term || (s == sub_limit);      // Q1: 2**k more steps after step 0
    if (!end_sub) { r.s = s + 1; return 0; }
    if (DEFER) return kSubtreeEnd | (div ? kEndDiv : 0) | (term ? kEndTerm : 0);
    return subtree_end<T, G, DENSE, SPLIT, Front>(ch, f, div, term) ? 1 : 0;
}

// End of a sub-tree: the front goes back to the edge arrays, then expand_once (trajectory.py:537-608) or, for the
// stand-alone dynamic_integration, the stop.  Always returns true (the front must be re-bound).
template <typename T, int G, bool DENSE, bool SPLIT, class Front>
B2H_DEVINL bool subtree_end(Chain<T, G>& ch, Front& f, bool div, bool term) {
    const EngineView<T>& v = ch.v;
    ChainRec& r = ch.r;
    const int d = v.d;
    const int k = r.k;
    constexpr int CH = SPLIT ? kBatch : 1;
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
    for (int jb = ch.lane; jb < d; jb += CH * G) {
        T m0[CH], m1[CH], plv[CH], prv[CH], xl[CH], xr[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int j = jb + i * G;
            m0[i] = 0; m1[i] = 0; plv[i] = 0; prv[i] = 0; xl[i] = 0; xr[i] = 0;
            if (j < d) {
                const i64 a = ch.at(j);
                m0[i] = v.msum[a]; m1[i] = v.sms[a]; plv[i] = v.pl[a]; prv[i] = v.pr[a];
                if (DENSE) { xl[i] = v.vl[a]; xr[i] = v.vr[a]; }
                else { const T im = ch.imm(j); xl[i] = im * plv[i]; xr[i] = im * prv[i]; }
            }
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int j = jb + i * G;
            if (j < d) {
                const T ms = m0[i] + m1[i];
                v.msum[ch.at(j)] = ms;
                const T rho = ms - (prv[i] + plv[i]) / (T)2;
                tl += xl[i] * rho;
                tr += xr[i] * rho;
            }
        }
    }
    double red2[2] = {(double)tl, (double)tr};
    Group<G>::template sum<2>(red2, ch.red);
    const bool top_turn = ((T)red2[0] <= (T)0) || ((T)red2[1] <= (T)0);

    r.accept_prob = exp(r.slpa_sub) / (double)r.sub_len;        // Q9 (trajectory.py:551-553)

    // biased progressive sampling is always drawn (Q8, proposals.py:130-131)
    double diff = r.w_sub - r.w_prop;
    double pb = fmin(fmax(exp(diff), 0.0), 1.0);
    double ub = draw_u<G>(v.rng, DRAW_BIASED, ch.c, r.t, k, v.maxd);
    bool accb = bern(ub, pb);
    if (div || term) {
        r.slpa_prop = lae_g<G>(r.slpa_sub, r.slpa_prop);             // trajectory.py:560-564
    } else {
        if (accb) {
            for (int jb = ch.lane; jb < d; jb += CH * G) {
                T x0[CH], x1[CH], x2[CH], x3[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int j = jb + i * G;
                    x0[i] = 0; x1[i] = 0; x2[i] = 0; x3[i] = 0;
                    if (j < d) {
                        const i64 a = ch.at(j);
                        x0[i] = v.qs[a]; x1[i] = v.ps[a]; x2[i] = v.gs[a];
                        if (DENSE) x3[i] = v.ws[a];
                    }
                }
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int j = jb + i * G;
                    if (j < d) {
                        const i64 a = ch.at(j);
                        v.qp[a] = x0[i]; v.pp[a] = x1[i]; v.gp[a] = x2[i];
                        if (DENSE) v.wp[a] = x3[i];
                    }
                }
            }
            r.E_prop = r.E_sub; r.U_prop = r.U_sub;
        }
        r.w_prop = lae_g<G>(r.w_prop, r.w_sub);
        r.slpa_prop = lae_g<G>(r.slpa_prop, r.slpa_sub);
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
B2H_DEVINL void hmc_begin(Chain<T, G>& ch, const double* zs = nullptr, int zready = 0, int zstride = 0) {
    begin_transition<T, G, DENSE, false>(ch, zs, zready, zstride);   // edges = state + fresh momentum, E0
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
            f.set_w(ee, a, wv);
            T p = f.p(ee, a) - he * g;
            f.set_p(ee, a, p);
            T vel = f.vel(ee, a) - he * wv;
            f.set_vel(ee, a, vel);
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
    double u = draw_u<G>(v.rng, DRAW_ACCEPT, ch.c, r.t, 0, v.maxd);
    bool acc = bern(u, p_accept);
    B2H_ELEMS(Front, ee, j, ch.lane, d, G) {
        i64 a = ch.at(j);
        if (acc) {
            v.qp[a] = f.q(ee, a); v.pp[a] = -f.p(ee, a); v.gp[a] = f.g(ee, a);
            if (DENSE) v.wp[a] = f.w(ee, a);
        }
        // on reject the state keeps (q, fresh momentum, g): pp already holds p0 (hmc.py:122,195)
    }
    if (acc) r.U_prop = (double)U_new;
    r.accept_prob = p_accept;
    end_transition(ch, 0, false, div);
    return true;
}

}  // namespace b2h
