// The tile tick kernel of the split engine (NUTS): post of tick t fused with pre of tick t + 1, like
// split_postpre_kernel, but organised so that the work per chain-tick is a few hundred instructions instead of ~2000:
//
//  * a warp owns a TILE of TC consecutive chains (TC = 32 / 8 / 1, chosen from the number of chains);
//  * the SCALAR part of the state machine (energies, log-sum-exp weights, progressive / biased sampling, Philox
//    uniforms, U-turn and divergence decisions, dual averaging: trajectory.py:195-273,537-608, proposals.py:41-144,
//    algorithms.py:104-115) runs ONE LANE PER CHAIN: the double-precision exp / log chains and the Philox rounds are
//    issued once per 32 chains instead of once per chain, and divergence between chains costs predicated scalar
//    code only;
//  * the VECTOR part (kick, kinetic energy, momentum sums, checkpoint rows, U-turn dot products, proposal copies,
//    half kick + drift: integrators.py:58-73, termination.py:109-187, metrics.py:70-102) runs WARP PER CHAIN in a loop
//    over the tile with warp-uniform control flow (the chain's flags are broadcast from its lane), 128-bit row
//    accesses and one butterfly reduction per chain.  No register front: the passes stream the rows.
//  * Tiles of several chains (short rows) are bound by the latency of their row loads: the rows of the next three chains
//    are copied asynchronously (cp.async) into a shared-memory ring while the current chain is processed.  Layouts with
//    one chain per warp / per CTA of 4 or 8 warps (long rows; thread 0 owns the scalars, values travel through shared
//    memory) keep what pass D needs again in shared memory between the passes.
//
// Passes:  S0 (scalars of the step)  ->  A (kick, K, sums, checkpoints, U-turn dots)  ->  S1 (energy, sampling, end
// of sub-tree?)  ->  B (sub-tree ends: front back to the edge, proposal, trajectory sum, top-level U-turn)  ->  S2
// (expand_once decisions, end of transition, adaptation scalars)  ->  C (proposal copy, draw, Welford, next
// transition's momentum and edges)  ->  S3 (initial energy, direction)  ->  D (half kick + drift of the next tick,
// fused with the sub-tree proposal of chains that continue).
//
// The arithmetic per element and the decisions are those of engine.cuh's post_gradient / begin_transition /
// half_kick_drift / end_transition / adapt_update (same expressions, same rounding: this file is compiled with
// -fmad=false like the rest of the engine); only the order of the partial sums inside a reduction differs, as it
// already does between the thread-, warp- and CTA-per-chain layouts.
#pragma once

#include "engine_host.cuh"

namespace b2h {
namespace tile {

constexpr unsigned kFull = 0xffffffffu;

template <typename T, int VEC>
B2H_DEVINL void ldv(T (&x)[VEC], const T* __restrict__ p) {
    if constexpr (VEC == 1) {
        x[0] = *p;
    } else if constexpr (sizeof(T) == 4) {
        static_assert(VEC == 4, "float rows move as float4");
        const float4 t = *reinterpret_cast<const float4*>(p);
        x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else {
        static_assert(VEC == 2, "double rows move as double2");
        const double2 t = *reinterpret_cast<const double2*>(p);
        x[0] = t.x; x[1] = t.y;
    }
}

template <typename T, int VEC>
B2H_DEVINL void stv(T* __restrict__ p, const T (&x)[VEC]) {
    if constexpr (VEC == 1) {
        *p = x[0];
    } else if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
    } else {
        *reinterpret_cast<double2*>(p) = make_double2(x[0], x[1]);
    }
}

// diagonal-family inverse mass matrix of one chain: row (stride 1) or a single scalar (stride 0)
template <typename T, int VEC>
B2H_DEVINL void ld_imm(T (&x)[VEC], const T* __restrict__ row, int j, i64 sj) {
    if (sj == 0) {
        const T s = row[0];
#pragma unroll
        for (int i = 0; i < VEC; ++i) x[i] = s;
    } else {
        ldv<T, VEC>(x, row + j);
    }
}

// standard normals j .. j + VEC - 1 of transition t: the values draw_z returns, one Philox block per Box-Muller pair
template <typename T, int VEC>
B2H_DEVINL void draw_z_vec(const RngView& rg, int c, int t, int j, int d, T (&z)[VEC]) {
    if (rg.mode == 1 || VEC == 1) {
#pragma unroll
        for (int x = 0; x < VEC; ++x) z[x] = (T)draw_z<8>(rg, c, t, j + x, d);
    } else {
#pragma unroll
        for (int x = 0; x + 1 < VEC; x += 2) {
            double z0, z1;
            philox_normal_pair(rg.key, rg.chain_offset + (uint64_t)c, (uint32_t)(rg.transition_offset + (uint64_t)t),
                               (uint32_t)((j + x) >> 1), &z0, &z1);
            z[x] = (T)z0;
            z[x + 1] = (T)z1;
        }
    }
}


// Asynchronous 16-byte copy global -> shared (LDGSTS) and its group bookkeeping.  Tiles of several chains (short
// rows: one 16-byte piece per lane and row) are bound by the latency of their row loads, not by bandwidth: every lane
// copies ITS pieces of the rows of the next kRing - 1 chains into its own slots of a shared-memory ring while the
// current chain is processed, and later reads back exactly the bytes it copied (no cross-lane synchronisation).
B2H_DEVINL void cp16(unsigned sdst /* shared-space address */, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc) : "memory");
}
B2H_DEVINL void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
B2H_DEVINL void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef B2H_TILE_RING
#define B2H_TILE_RING 4
#endif
constexpr int kRing = B2H_TILE_RING;                         // stages of the ring (chains in flight per warp: kRing - 1)
constexpr int kRowBytes = 512;                   // one row of a stage: 32 lanes x 16 bytes
template <bool DENSE> struct RingRows { static constexpr int value = DENSE ? 9 : 6; };

// Who works on a chain.  WPC == 1: a warp owns a tile of TC chains, chain slot i is owned by lane i, values travel
// by shuffle.  WPC > 1 (TC == 1, long rows): the CTA's WPC warps share ONE chain, thread 0 owns it, values travel
// through shared memory (each broadcast has its own slot: nothing is reused within a launch).
template <int TC, int WPC>
struct Coop {
    static_assert(WPC == 1 || TC == 1, "several warps per chain: one chain per CTA");
    static constexpr int kThreads = WPC > 1 ? 32 * WPC : 128;
    B2H_DEVINL static int tid() { return WPC > 1 ? (int)threadIdx.x : (int)(threadIdx.x & 31); }
    B2H_DEVINL static bool owner(int i) { return tid() == i; }
    template <typename V>
    B2H_DEVINL static V get(V x, int i, double* slots, int slot) {
        if constexpr (WPC == 1) {
            return __shfl_sync(kFull, x, i);
        } else {
            V* s = reinterpret_cast<V*>(slots + slot);
            if (threadIdx.x == 0) *s = x;
            __syncthreads();
            return *s;
        }
    }
    // bit i set: chain slot i has the flag
    B2H_DEVINL static unsigned mask(bool flag, double* slots, int slot) {
        if constexpr (WPC == 1) return __ballot_sync(kFull, flag);
        else return get<int>(flag ? 1 : 0, 0, slots, slot) ? 1u : 0u;
    }
    template <int N>
    B2H_DEVINL static void sum(double (&v)[N], double* red) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(kFull, v[i], off);
        }
        if constexpr (WPC > 1) {
            const int w = threadIdx.x >> 5;
            __syncthreads();                       // the previous reduction's partials have been read
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) red[i * WPC + w] = v[i];
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < WPC; ++k) s += red[i * WPC + k];
                v[i] = s;
            }
        }
    }
};

#ifndef B2H_TILE_WPC_WARPS
#define B2H_TILE_WPC_WARPS 16        // resident warps per SM the several-warps-per-chain kernels are compiled for
#endif
#ifndef B2H_TILE_MINB
#define B2H_TILE_MINB 4
#endif

template <typename T, int TC, int WPC, int VEC, bool DENSE>
__global__ void __launch_bounds__((Coop<TC, WPC>::kThreads), (WPC > 1 ? B2H_TILE_WPC_WARPS / WPC : B2H_TILE_MINB)) tile_tick_kernel(EngineView<T> v, int* not_done, int flags) {
    typedef Coop<TC, WPC> Co;
    const bool pre = (flags & 1) != 0;                   // the half kick + drift of the next tick follows (not the last tick of a call)
    __shared__ double slots[32];                         // WPC > 1: one slot per broadcast value
    __shared__ double red_s[4 * WPC];                    // WPC > 1: per-warp partials of a reduction
    const int lane = Co::tid();                          // index of this thread in the group that shares a chain's rows
    const int c0 = WPC > 1 ? (int)blockIdx.x : (int)(((i64)blockIdx.x * 128 + threadIdx.x) >> 5) * TC;
    if (c0 >= v.C) return;
    const int d = v.d;
    const int c = c0 + lane;
    const bool own = lane < TC && c < v.C;
    // the chain's record: registers of its lane; several warps per chain: shared memory (only thread 0 uses it, and the
    // other 127 threads do not pay 48 registers for a copy they never read)
    ChainRecLive r_lane{};
    __shared__ ChainRecLive r_cta;
    ChainRecLive& r = WPC > 1 ? r_cta : r_lane;
    if (own) r = *static_cast<const ChainRecLive*>(v.rec + c);
    const int ph0 = own ? r.phase : (int)PH_DONE;
    if (Co::mask(ph0 != PH_DONE, slots, 0) == 0) return;
    const bool run = ph0 == PH_RUN;
    constexpr int STEP = 32 * WPC * VEC;
    const int j0 = lane * VEC;
    // tiles of several chains: this warp's ring of row stages (dynamic shared memory; see cp16)
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    constexpr int NROW = RingRows<DENSE>::value;
    [[maybe_unused]] unsigned char* const ring = dyn_smem + (size_t)(threadIdx.x >> 5) * (kRing * NROW * kRowBytes);
    // this lane's 16-byte slot of (stage 0, row 0), as a shared-space address for the asynchronous copies
    [[maybe_unused]] const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring) + (threadIdx.x & 31) * 16;
    // one chain per warp / CTA: what pass A computes or reads and pass D needs again -- p', g', q'[, imm p', imm g'] of a
    // chain that continues its sub-tree -- waits in shared memory instead of being re-read (same thread, same pieces: no
    // synchronisation); flags bit 1, set by the launcher when the rows fit
    constexpr int SROWS = DENSE ? 5 : 3;
    [[maybe_unused]] const bool stash = TC == 1 && VEC > 1 && (flags & 2) != 0;
    [[maybe_unused]] T* const stash_rows = reinterpret_cast<T*>(dyn_smem) + (size_t)(WPC > 1 ? 0 : (threadIdx.x >> 5)) * SROWS * d;
    // an inverse mass matrix shared by all chains (scalar or one diagonal): this lane's piece is loaded once
    [[maybe_unused]] const bool imm_shared = !DENSE && v.imm_sc == 0;
    [[maybe_unused]] T im_sh[VEC];
    if constexpr (TC > 1 && !DENSE) {
#pragma unroll
        for (int x = 0; x < VEC; ++x) im_sh[x] = 0;
        if (imm_shared && j0 < d) ld_imm<T, VEC>(im_sh, v.imm, j0, v.imm_sj);
    }

    // ---- S0: what the step needs before its rows arrive -------------------------------------------------------------
    int s = 0, k = 0, imax = 0, nlev = 0;
    double u_step = 0.0;
    T U_new = 0, e = 0, he = 0;
    if (run) {
        s = r.s; k = r.k;
        // progressive-sampling uniform of the step (proposals.py:96-100): independent of the energies
        if (s != 0) u_step = draw_u<8>(v.rng, DRAW_UNIFORM, c, r.t, uniform_slot(k, s), v.maxd);
        int imin;
        if (s == 0) { imin = r.imin; imax = r.imax; }            // Q2 / Q3: stale indices at step 0
        else storage_indices(s, imin, imax);
        r.imin = imin; r.imax = imax;
        nlev = (s >= 1 && imax >= imin) ? (imax - imin + 1) : 0;
        e = (T)(r.go_right ? r.eps : -r.eps);
        he = (T)0.5 * e;
        if (!(TC == 1 && VEC > 1 && v.u_center)) U_new = v.Unew[c];   // else pass A forms it from q' and g'
    }

    // ---- A: p' = p_half - (0.5 e) g', K(p'), sub-tree momentum sum, checkpoint write (even steps), U-turn dot products
    //         (odd steps: the levels imin .. imax of termination.py:164-187) ---------------------------------------
    T myK = 0;
    bool term = false;
    {
        const int fa = run ? (1 | (r.go_right ? 2 : 0) | (s == 0 ? 4 : 0) | ((s & 1) == 0 ? 8 : 0) | (imax << 8) | (nlev << 16)) : 0;
        // tiles of several chains: ring slot of (stage st, row r) for this lane; rows 0 P, 1 xb, 2 sms, 3 mck, 4 sckp,
        // 5 imm, 6 V, 7 xc, 8 vck
        [[maybe_unused]] auto slot = [&](int st, int row) -> unsigned { return ring_s + (unsigned)((st * NROW + row) * kRowBytes); };
        unsigned act = Co::mask((fa & 1) != 0, slots, 1);
        [[maybe_unused]] unsigned iss = act;
        [[maybe_unused]] auto issue = [&](int st) {
            if (iss) {
                const int i = __ffs(iss) - 1;
                iss &= iss - 1;
                const int f = __shfl_sync(kFull, fa, i);
                if (j0 < d) {
                    const bool gr = f & 2, s0 = f & 4;
                    const int imx = (f >> 8) & 0xff, nl = (f >> 16) & 0xff;
                    const int ci = c0 + i;
                    const i64 rb = (i64)ci * d + j0;
                    cp16(slot(st, 0), (gr ? v.pr : v.pl) + rb);
                    cp16(slot(st, 1), v.xb + rb);
                    if (!s0) cp16(slot(st, 2), v.sms + rb);
                    if (nl > 0) {
                        const i64 cb = (i64)ci * v.sck + (i64)imx * d + j0;
                        cp16(slot(st, 3), v.mck + cb);
                        cp16(slot(st, 4), v.sckp + cb);
                        if (DENSE) cp16(slot(st, 8), v.vck + cb);
                    }
                    if (DENSE) { cp16(slot(st, 6), (gr ? v.vr : v.vl) + rb); cp16(slot(st, 7), v.xc + rb); }
                    else if (!imm_shared) cp16(slot(st, 5), v.imm + (i64)ci * v.imm_sc + j0);
                }
            }
            cp_commit();
        };
        [[maybe_unused]] int stg = 0;
        if constexpr (TC > 1) {
#pragma unroll
            for (int st = 0; st < kRing - 1; ++st) issue(st);
        }
#pragma unroll 1
        while (act) {
            if constexpr (TC > 1) {
                issue(stg == 0 ? kRing - 1 : stg - 1);      // the stage the previous iteration consumed
                cp_wait<kRing - 1>();                       // this iteration's rows have landed
            }
            const int i = __ffs(act) - 1;
            act &= act - 1;
            const int f = Co::get(fa, i, slots, 2);
            const T hei = Co::get(he, i, slots, 3);
            const bool gr = f & 2, s0 = f & 4, even = f & 8;
            const int imx = (f >> 8) & 0xff, nl = (f >> 16) & 0xff;
            const int ci = c0 + i;
            const i64 rb = (i64)ci * d;
            const T* __restrict__ P = (gr ? v.pr : v.pl) + rb;
            const T* __restrict__ XB = v.xb + rb;
            const T* __restrict__ V = DENSE ? (gr ? v.vr : v.vl) + rb : nullptr;
            const T* __restrict__ XC = DENSE ? v.xc + rb : nullptr;
            const T* __restrict__ IM = DENSE ? nullptr : v.imm + (i64)ci * v.imm_sc;
            T* __restrict__ SMS = v.sms + rb;
            const i64 cb = (i64)ci * v.sck + (i64)imx * d;
            T* __restrict__ MCK = v.mck + cb;
            T* __restrict__ SCK = v.sckp + cb;
            T* __restrict__ VCK = DENSE ? v.vck + cb : nullptr;
            T kacc = 0, dl = 0, dr = 0;
            [[maybe_unused]] T uacc = 0;
            [[maybe_unused]] const bool ucen = TC == 1 && VEC > 1 && v.u_center != nullptr;
            // where the rows are read from: global memory, or (tiles of several chains) this stage of the ring, whose
            // row r starts at slot(stg, r) - 16 lane, so that "+ j" with j = j0 lands on this lane's piece
            const T* Pl = P;
            const T* XBl = XB;
            const T* Vl = V;
            const T* XCl = XC;
            const T* IMl = IM;
            const T* SMSl = SMS;
            const T* MCKl = MCK;
            const T* SCKl = SCK;
            const T* VCKl = VCK;
            // this stage of the ring as the source of the loads
            [[maybe_unused]] auto from_stage = [&](int st) {
                const unsigned char* sb = ring + (size_t)(st * NROW) * kRowBytes;
                Pl = reinterpret_cast<const T*>(sb);
                XBl = reinterpret_cast<const T*>(sb + 1 * kRowBytes);
                SMSl = reinterpret_cast<const T*>(sb + 2 * kRowBytes);
                MCKl = reinterpret_cast<const T*>(sb + 3 * kRowBytes);
                SCKl = reinterpret_cast<const T*>(sb + 4 * kRowBytes);
                if (!DENSE && !imm_shared) IMl = reinterpret_cast<const T*>(sb + 5 * kRowBytes);
                if (DENSE) {
                    Vl = reinterpret_cast<const T*>(sb + 6 * kRowBytes);
                    XCl = reinterpret_cast<const T*>(sb + 7 * kRowBytes);
                    VCKl = reinterpret_cast<const T*>(sb + 8 * kRowBytes);
                }
            };
            if constexpr (TC > 1) {
                from_stage(stg);
                stg = stg + 1 == kRing ? 0 : stg + 1;
            }
            for (int j = j0; j < d; j += STEP) {
                const int jl = TC > 1 ? j0 : j;                  // ring rows hold this lane's piece at the lane's own slot
                T pv[VEC], gx[VEC], so[VEC], cm[VEC], cs[VEC], cv[VEC], im[VEC], vv[VEC], wx[VEC];
                ldv<T, VEC>(pv, Pl + jl);
                ldv<T, VEC>(gx, XBl + jl);
                if (DENSE) { ldv<T, VEC>(vv, Vl + jl); ldv<T, VEC>(wx, XCl + jl); }
                else if (TC > 1 && imm_shared) {
#pragma unroll
                    for (int x = 0; x < VEC; ++x) im[x] = im_sh[x];
                } else ld_imm<T, VEC>(im, IMl, imm_shared ? j : jl, v.imm_sj);
                if (!s0) ldv<T, VEC>(so, SMSl + jl);
                if (nl > 0) {
                    ldv<T, VEC>(cm, MCKl + jl); ldv<T, VEC>(cs, SCKl + jl);
                    if (DENSE) ldv<T, VEC>(cv, VCKl + jl);
                }
                [[maybe_unused]] T qx[VEC];
                if constexpr (TC == 1 && VEC > 1) {
                    if (stash || ucen) ldv<T, VEC>(qx, v.xa + rb + j);   // for pass D (it would read it anyway) / for U
                    if (ucen) {                                          // U = 0.5 (q' - centre) . g' (Gaussian targets)
                        T mu[VEC];
                        ldv<T, VEC>(mu, v.u_center + j);
#pragma unroll
                        for (int x = 0; x < VEC; ++x) uacc += (qx[x] - mu[x]) * gx[x];
                    }
                }
                T p[VEC], vel[VEC], sm[VEC];
#pragma unroll
                for (int x = 0; x < VEC; ++x) {
                    p[x] = pv[x] - hei * gx[x];                          // integrators.py:66
                    if (DENSE) vel[x] = vv[x] - hei * wx[x];             // imm p' = imm p_half - (0.5 e) imm g'
                    else vel[x] = im[x] * p[x];
                    kacc += vel[x] * p[x];                               // metrics.py:70-73
                    sm[x] = s0 ? p[x] : so[x] + p[x];                    // trajectory.py:243,278
                    if (nl > 0) {
                        const T subsum = sm[x] - cs[x] + cm[x];
                        const T rho = subsum - (p[x] + cm[x]) / (T)2;
                        const T vleft = DENSE ? cv[x] : im[x] * cm[x];
                        dl += vleft * rho;
                        dr += vel[x] * rho;
                    }
                }
                stv<T, VEC>(SMS + j, sm);
                if (even) {                                              // termination.py:109-124
                    stv<T, VEC>(MCK + j, p);
                    stv<T, VEC>(SCK + j, sm);
                    if (DENSE) stv<T, VEC>(VCK + j, vel);
                }
                if constexpr (TC == 1 && VEC > 1) {
                    if (stash) {
                        stv<T, VEC>(stash_rows + j, p);
                        stv<T, VEC>(stash_rows + d + j, gx);
                        stv<T, VEC>(stash_rows + 2 * d + j, qx);
                        if (DENSE) { stv<T, VEC>(stash_rows + 3 * d + j, vel); stv<T, VEC>(stash_rows + 4 * d + j, wx); }
                    }
                }
            }
            double red[3] = {(double)kacc, (double)dl, (double)dr};
            [[maybe_unused]] double usum = 0.0;
            if constexpr (TC == 1 && VEC > 1) {
                if (ucen) {                                              // warp-uniform: the potential rides along
                    double r4[4] = {red[0], red[1], red[2], (double)uacc};
                    Co::template sum<4>(r4, red_s);
                    red[0] = r4[0]; red[1] = r4[1]; red[2] = r4[2]; usum = r4[3];
                } else if (nl > 0) Co::template sum<3>(red, red_s);
                else { double r1[1] = {red[0]}; Co::template sum<1>(r1, red_s); red[0] = r1[0]; }
            } else {
                if (nl > 0) Co::template sum<3>(red, red_s);
                else { double r1[1] = {red[0]}; Co::template sum<1>(r1, red_s); red[0] = r1[0]; }
            }
            bool tm = nl > 0 && ((T)red[1] <= (T)0 || (T)red[2] <= (T)0);
            // deeper levels (steps with two or more trailing one-bits); the outcome is an OR over the levels
            for (int l = 1; l < nl && !tm; ++l) {
                const T* __restrict__ MC = MCK - (i64)l * d;
                const T* __restrict__ SC = SCK - (i64)l * d;
                const T* __restrict__ VC = DENSE ? VCK - (i64)l * d : nullptr;
                T xl = 0, xr = 0;
                for (int j = j0; j < d; j += STEP) {
                    T pv[VEC], gx[VEC], sm[VEC], cm[VEC], cs[VEC], cv[VEC], im[VEC], vv[VEC], wx[VEC];
                    ldv<T, VEC>(pv, P + j);
                    ldv<T, VEC>(gx, XB + j);
                    if (DENSE) { ldv<T, VEC>(vv, V + j); ldv<T, VEC>(wx, XC + j); ldv<T, VEC>(cv, VC + j); }
                    else ld_imm<T, VEC>(im, IM, j, v.imm_sj);
                    ldv<T, VEC>(sm, SMS + j);
                    ldv<T, VEC>(cm, MC + j); ldv<T, VEC>(cs, SC + j);
#pragma unroll
                    for (int x = 0; x < VEC; ++x) {
                        const T p = pv[x] - hei * gx[x];
                        const T vright = DENSE ? vv[x] - hei * wx[x] : im[x] * p;
                        const T vleft = DENSE ? cv[x] : im[x] * cm[x];
                        const T subsum = sm[x] - cs[x] + cm[x];
                        const T rho = subsum - (p + cm[x]) / (T)2;
                        xl += vleft * rho;
                        xr += vright * rho;
                    }
                }
                double r2[2] = {(double)xl, (double)xr};
                Co::template sum<2>(r2, red_s);
                if ((T)r2[0] <= (T)0 || (T)r2[1] <= (T)0) tm = true;
            }
            if (Co::owner(i)) {
                myK = (T)0.5 * (T)red[0];
                term = tm;
                if constexpr (TC == 1 && VEC > 1) { if (ucen) U_new = (T)0.5 * (T)usum; }
            }
        }
    }

    // ---- S1: energy, divergence, progressive sampling inside the sub-tree, end of the sub-tree? ---------------------
    bool take = false, end_sub = false, div = false, expand = false;
    if (run) {
        const T E = U_new + myK;
        double delta = (double)((T)r.E0 - E);
        if (isnan(delta)) delta = -INFINITY;
        div = fabs(delta) > v.div_thr;
        const double w_new = delta;                                  // proposals.py:41-52, Q10
        const double lpa = delta > 0 ? 0.0 : delta;
        if (s == 0) {
            take = true;                                             // trajectory.py:276-277
            r.w_sub = w_new; r.slpa_sub = lpa;
        } else {
            double pa = expit(w_new - r.w_sub);                      // proposals.py:96-100, Q7
            if (isnan(pa)) pa = 0.0;
            take = bern(u_step, pa);
            r.w_sub = lae(r.w_sub, w_new);                           // proposals.py:141-144
            r.slpa_sub = lae(r.slpa_sub, lpa);
        }
        if (take) { r.E_sub = (double)E; r.U_sub = (double)U_new; }
        r.sub_len = (s == 0) ? 1 : r.sub_len + 1;
        r.nleap += 1;
        r.total_leap += 1;
        r.U_front = (double)U_new;
        const int sub_limit = v.sub_max_steps > 0 ? v.sub_max_steps : (1 << k) - (v.exact_doubling ? 1 : 0);
        end_sub = div || term || (s == sub_limit);                   // Q1
        if (!end_sub) {
            r.s = s + 1;
        } else {
            if (r.go_right) r.U_right = r.U_front; else r.U_left = r.U_front;
            r.sub_term = term ? 1 : 0;
            if (v.stop_at_subtree_end) {                             // trajectory.dynamic_integration.integrate on its own
                r.last_flags = (div ? 2 : 0) | (term ? 4 : 0);
                r.last_nleap = r.sub_len;
                r.phase = PH_DONE;
            } else {
                expand = true;
            }
        }
    }

    // ---- B: sub-tree ends (and every running chain on the last tick of a call): the front goes back to the edge
    //         arrays, sub-tree proposal if taken, trajectory momentum sum and the top-level U-turn (trajectory.py:537-553)
    bool top_turn = false;
    {
        const bool flush = run && (end_sub || !pre);
        const int fb = flush ? (1 | (take ? 2 : 0) | (expand ? 4 : 0) | (r.go_right ? 8 : 0)) : 0;
        unsigned act = Co::mask((fb & 1) != 0, slots, 4);
        {
#pragma unroll 1
            while (act) {
                const int i = __ffs(act) - 1;
                act &= act - 1;
                const int f = Co::get(fb, i, slots, 5);
                const T hei = Co::get(he, i, slots, 6);
                const bool tk = f & 2, ex = f & 4, gr = f & 8;
                const int ci = c0 + i;
                const i64 rb = (i64)ci * d;
                T* __restrict__ P = (gr ? v.pr : v.pl) + rb;
                T* __restrict__ Q = (gr ? v.qr : v.ql) + rb;
                T* __restrict__ Gd = (gr ? v.gr : v.gl) + rb;
                T* __restrict__ V = DENSE ? (gr ? v.vr : v.vl) + rb : nullptr;
                T* __restrict__ W = DENSE ? (gr ? v.wr : v.wl) + rb : nullptr;
                const T* __restrict__ PO = (gr ? v.pl : v.pr) + rb;
                const T* __restrict__ VO = DENSE ? (gr ? v.vl : v.vr) + rb : nullptr;
                const T* __restrict__ XA = v.xa + rb;
                const T* __restrict__ XB = v.xb + rb;
                const T* __restrict__ XC = DENSE ? v.xc + rb : nullptr;
                const T* __restrict__ IM = DENSE ? nullptr : v.imm + (i64)ci * v.imm_sc;
                T tl = 0, tr = 0;
                for (int j = j0; j < d; j += STEP) {
                    T pv[VEC], gx[VEC], qx[VEC], vv[VEC], wx[VEC];
                    ldv<T, VEC>(pv, P + j);
                    ldv<T, VEC>(gx, XB + j);
                    ldv<T, VEC>(qx, XA + j);
                    if (DENSE) { ldv<T, VEC>(vv, V + j); ldv<T, VEC>(wx, XC + j); }
                    T m0[VEC], m1[VEC], po[VEC], vo[VEC], im[VEC];
                    if (ex) {
                        ldv<T, VEC>(m0, v.msum + rb + j);
                        ldv<T, VEC>(m1, v.sms + rb + j);
                        ldv<T, VEC>(po, PO + j);
                        if (DENSE) ldv<T, VEC>(vo, VO + j);
                        else ld_imm<T, VEC>(im, IM, j, v.imm_sj);
                    }
                    T p[VEC], vel[VEC];
#pragma unroll
                    for (int x = 0; x < VEC; ++x) {
                        p[x] = pv[x] - hei * gx[x];
                        vel[x] = DENSE ? vv[x] - hei * wx[x] : (T)0;
                    }
                    if (tk) {
                        stv<T, VEC>(v.qs + rb + j, qx); stv<T, VEC>(v.ps + rb + j, p); stv<T, VEC>(v.gs + rb + j, gx);
                        if (DENSE) stv<T, VEC>(v.ws + rb + j, wx);
                    }
                    stv<T, VEC>(Q + j, qx); stv<T, VEC>(P + j, p); stv<T, VEC>(Gd + j, gx);
                    if (DENSE) { stv<T, VEC>(V + j, vel); stv<T, VEC>(W + j, wx); }
                    if (ex) {
                        T ms[VEC];
#pragma unroll
                        for (int x = 0; x < VEC; ++x) {
                            const T plv = gr ? po[x] : p[x], prv = gr ? p[x] : po[x];
                            T xl, xr;
                            if (DENSE) { xl = gr ? vo[x] : vel[x]; xr = gr ? vel[x] : vo[x]; }
                            else { xl = im[x] * plv; xr = im[x] * prv; }
                            ms[x] = m0[x] + m1[x];
                            const T rho = ms[x] - (prv + plv) / (T)2;
                            tl += xl * rho;
                            tr += xr * rho;
                        }
                        stv<T, VEC>(v.msum + rb + j, ms);
                    }
                }
                if (ex) {
                    double r2[2] = {(double)tl, (double)tr};
                    Co::template sum<2>(r2, red_s);
                    if (Co::owner(i)) top_turn = ((T)r2[0] <= (T)0) || ((T)r2[1] <= (T)0);
                }
            }
        }
    }

    // ---- S2: expand_once (trajectory.py:551-608): acceptance statistic, biased progressive sampling, end of the
    //          transition (nuts.py:138-151) with the dual-averaging update, or the next sub-tree's direction ----------
    bool copy_prop = false, store_draw = false, slow = false, wendv = false;
    int draw_slot = 0, wc_n = 0;
    if (expand) {
        r.accept_prob = exp(r.slpa_sub) / (double)r.sub_len;        // Q9
        const double diff = r.w_sub - r.w_prop;
        const double pb = fmin(fmax(exp(diff), 0.0), 1.0);
        const double ub = draw_u<8>(v.rng, DRAW_BIASED, c, r.t, k, v.maxd);   // Q8: always drawn
        const bool accb = bern(ub, pb);
        if (div || term) {
            r.slpa_prop = lae(r.slpa_sub, r.slpa_prop);             // trajectory.py:560-564
        } else {
            if (accb) { copy_prop = true; r.E_prop = r.E_sub; r.U_prop = r.U_sub; }
            r.w_prop = lae(r.w_prop, r.w_sub);
            r.slpa_prop = lae(r.slpa_prop, r.slpa_sub);
        }
        const int nd = k + 1;
        if (div || top_turn || term || nd >= v.maxd) {              // trajectory.py:577 / scan length
            // end_transition
            const int t = r.t;
            const int tl = t - r.t_base;
            r.last_nd = nd;
            r.last_flags = (top_turn ? 1 : 0) | (div ? 2 : 0) | (r.sub_term ? 4 : 0);
            r.last_nleap = r.nleap;
            const int thin = v.out.thin > 1 ? v.out.thin : 1;
            const int slot = tl / thin;
            if (slot < v.out.n_store && slot * thin == tl) {
                store_draw = v.out.draws != nullptr;
                draw_slot = slot;
                if (v.out.draw_stats) {
                    double* ds = v.out.draw_stats + ((i64)slot * v.C + c) * 4;
                    ds[0] = r.accept_prob; ds[1] = (double)nd; ds[2] = (double)r.nleap; ds[3] = (double)r.last_flags;
                }
            }
            if (v.adapt.enabled && t + v.adapt.step_offset < v.adapt.num_steps) {
                // adapt_update, scalar part (window_adaptation.py:194-215, algorithms.py:104-115, step_size.py:97)
                const AdaptView& ad = v.adapt;
                const int step = t + ad.step_offset;
                i64 dstep = ad.da_step[c];
                const double x_old = ad.da_x[c], xavg = ad.da_x_avg[c], gavg = ad.da_g_avg[c];
                double mu = ad.da_mu[c];
                const double grad = ad.target - r.accept_prob;
                const double eta = 1.0 / ((double)dstep + ad.t0);
                double new_gavg = (1.0 - eta) * gavg + eta * grad;
                double new_x = mu - (sqrt((double)dstep) / ad.gamma) * new_gavg;
                const double x_eta = pow((double)dstep, -ad.kappa);
                double new_xavg = x_eta * x_old + (1.0 - x_eta) * xavg;     // Q16: OLD iterate
                dstep += 1;
                double eps = exp(new_x);
                slow = ad.stage[step] != 0 && !ad.pooled;
                const bool wend = ad.window_end[step] != 0;
                i64 n = ad.pooled ? 0 : ad.wc_n[c];
                if (slow) n += 1;
                wc_n = (int)n;
                if (wend) {
                    wendv = !ad.pooled;
                    n = 0;
                    mu = eps;
                    dstep = 1; new_x = 0.0; new_xavg = 0.0; new_gavg = 0.0;
                }
                if (step == ad.num_steps - 1) eps = exp(new_xavg);
                ad.da_step[c] = dstep; ad.da_x[c] = new_x; ad.da_x_avg[c] = new_xavg; ad.da_g_avg[c] = new_gavg;
                ad.da_mu[c] = mu;
                if (!ad.pooled) ad.wc_n[c] = n;
                r.eps = eps;
            }
            r.t = t + 1;
            r.phase = (v.n_transitions > 0 && (r.t - r.t_base) >= v.n_transitions) ? PH_DONE : PH_START;
        } else {
            r.k = nd;
            const double u = draw_u<8>(v.rng, DRAW_DIR, c, r.t, r.k, v.maxd);   // trajectory.py:516-518
            r.go_right = bern(u, 0.5) ? 1 : 0;
            r.s = 0;
        }
    }

    // ---- C: proposal <- sub-tree proposal, stored draw, Welford / window end (algorithms.py:187-197,
    //         mass_matrix.py:103-116), and the start of the next transition (nuts.py:113-135) ------------------------
    const bool begin = pre && own && r.phase == PH_START;
    T K0 = 0;
    {
        int mslot = 0;
        bool have = false;
        if (DENSE && begin) {
            // queue the momentum of the transition AFTER the one that starts now (see begin_transition)
            mslot = atomicAdd(v.mom_count + v.mom_parity, 1);
            v.mom_list[(i64)v.mom_parity * v.C + mslot] = c;
            const int tn = r.t + 1;
            have = (v.rng.mode == 0) || (tn + (i64)v.rng.transition_offset < v.rng.n_injected);
        }
        const int fc = (copy_prop ? 1 : 0) | (store_draw ? 2 : 0) | (slow ? 4 : 0) | (wendv ? 8 : 0) | (begin ? 16 : 0) |
                       (have ? 32 : 0);
        unsigned act = Co::mask((fc & 31) != 0, slots, 7);
        {
#pragma unroll 1
            while (act) {
                const int i = __ffs(act) - 1;
                act &= act - 1;
                const int f = Co::get(fc, i, slots, 8);
                const bool cp = f & 1, sd = f & 2, sl = f & 4, we = f & 8, bg = f & 16, hv = f & 32;
                const int ci = c0 + i;
                const i64 rb = (i64)ci * d;
                const int ti = Co::get(r.t, i, slots, 9);
                const int dsl = Co::get(draw_slot, i, slots, 10);
                const int ni = Co::get(wc_n, i, slots, 11);
                const int msl = Co::get(mslot, i, slots, 12);
                T* __restrict__ DR = sd ? (T*)v.out.draws + ((i64)dsl * v.C + ci) * d : nullptr;
                T* __restrict__ MEAN = (T*)v.adapt.wc_mean + rb;
                T* __restrict__ M2 = (T*)v.adapt.wc_m2 + rb;
                T* __restrict__ IMW = v.imm + (i64)ci * v.imm_sc;
                T* __restrict__ ZROW = DENSE ? v.mom_z + ((i64)v.mom_parity * v.C + msl) * d : nullptr;
                const T scale = (T)((double)ni / ((double)ni + 5.0));
                const T shrink = (T)(1e-3 * (5.0 / ((double)ni + 5.0)));
                T kacc = 0;
                for (int j = j0; j < d; j += STEP) {
                    T qv[VEC], gv[VEC], wv[VEC];
                    if (cp) {
                        T pv[VEC];
                        ldv<T, VEC>(qv, v.qs + rb + j); ldv<T, VEC>(pv, v.ps + rb + j); ldv<T, VEC>(gv, v.gs + rb + j);
                        if (DENSE) ldv<T, VEC>(wv, v.ws + rb + j);
                        stv<T, VEC>(v.qp + rb + j, qv); stv<T, VEC>(v.pp + rb + j, pv); stv<T, VEC>(v.gp + rb + j, gv);
                        if (DENSE) stv<T, VEC>(v.wp + rb + j, wv);
                    } else {
                        ldv<T, VEC>(qv, v.qp + rb + j);
                        if (bg) {
                            ldv<T, VEC>(gv, v.gp + rb + j);
                            if (DENSE) ldv<T, VEC>(wv, v.wp + rb + j);
                        }
                    }
                    if (sd) stv<T, VEC>(DR + j, qv);
                    T im[VEC];
                    bool have_im = false;
                    if (sl || we) {
                        T mean[VEC], m2[VEC];
                        ldv<T, VEC>(mean, MEAN + j); ldv<T, VEC>(m2, M2 + j);
#pragma unroll
                        for (int x = 0; x < VEC; ++x) {
                            if (sl) {
                                const T delta = qv[x] - mean[x];
                                const T mn = mean[x] + delta / (T)ni;
                                const T ud = qv[x] - mn;
                                mean[x] = mn;
                                m2[x] = m2[x] + ud * delta;
                            }
                            if (we) {
                                const T cov = m2[x] / (T)(ni - 1);
                                im[x] = scale * cov + shrink;
                                mean[x] = 0;
                                m2[x] = 0;
                            }
                        }
                        stv<T, VEC>(MEAN + j, mean); stv<T, VEC>(M2 + j, m2);
                        if (we) { stv<T, VEC>(IMW + j, im); have_im = true; }
                    }
                    if (bg) {
                        T p0[VEC], vel[VEC];
                        if (DENSE) {
                            ldv<T, VEC>(p0, v.mom_p + rb + j);
                            ldv<T, VEC>(vel, v.mom_v + rb + j);
                        } else {
                            if (!have_im) ld_imm<T, VEC>(im, IMW, j, v.imm_sj);
                            T z[VEC];
                            draw_z_vec<T, VEC>(v.rng, ci, ti, j, d, z);
#pragma unroll
                            for (int x = 0; x < VEC; ++x) {
                                p0[x] = sqrt((T)1 / im[x]) * z[x];       // metrics.py:46,50,67
                                vel[x] = im[x] * p0[x];
                            }
                        }
                        if (DENSE) {
                            stv<T, VEC>(v.vl + rb + j, vel); stv<T, VEC>(v.vr + rb + j, vel);
                            stv<T, VEC>(v.wl + rb + j, wv); stv<T, VEC>(v.wr + rb + j, wv);
                        }
                        stv<T, VEC>(v.ql + rb + j, qv); stv<T, VEC>(v.qr + rb + j, qv);
                        stv<T, VEC>(v.pl + rb + j, p0); stv<T, VEC>(v.pr + rb + j, p0);
                        stv<T, VEC>(v.gl + rb + j, gv); stv<T, VEC>(v.gr + rb + j, gv);
                        stv<T, VEC>(v.pp + rb + j, p0);
                        stv<T, VEC>(v.msum + rb + j, p0);
#pragma unroll
                        for (int x = 0; x < VEC; ++x) kacc += vel[x] * p0[x];
                        if (DENSE) {
                            T z[VEC];
                            if (hv) draw_z_vec<T, VEC>(v.rng, ci, ti + 1, j, d, z);
                            else {
#pragma unroll
                                for (int x = 0; x < VEC; ++x) z[x] = 0;
                            }
                            stv<T, VEC>(ZROW + j, z);
                        }
                    }
                }
                if (bg) {
                    double r1[1] = {(double)kacc};
                    Co::template sum<1>(r1, red_s);
                    if (Co::owner(i)) K0 = (T)0.5 * (T)r1[0];
                }
            }
        }
    }

    // ---- S3: initial energy and the first sub-tree's direction (nuts.py:117-124, trajectory.py:516-518) -------------
    if (begin) {
        const T E0 = (T)r.U_prop + K0;
        r.E0 = (double)E0;
        r.U_left = r.U_prop; r.U_right = r.U_prop;
        r.E_prop = (double)E0;
        r.w_prop = 0.0;
        r.slpa_prop = -INFINITY;
        r.imin = 0; r.imax = 0;
        r.k = 0;
        r.nleap = 0;
        r.phase = PH_RUN;
        const double u = draw_u<8>(v.rng, DRAW_DIR, c, r.t, 0, v.maxd);
        r.go_right = bern(u, 0.5) ? 1 : 0;
        r.s = 0;
    }

    // ---- D: first half of the next leapfrog (integrators.py:59-62): p_half = p - (0.5 e) g, q' = q + e imm p_half.
    //         A chain that continues its sub-tree has its front in (P = p_half, xa = q', xb = g'[, V, xc]): p' is
    //         recomputed, and its sub-tree proposal (if taken) is stored from the same registers; the others take
    //         the front from the edge arrays of their (new) direction. -----------------------------------------------
    if (pre) {
        const bool go = own && r.phase == PH_RUN;
        const bool cont = run && !end_sub;
        const T en = go ? (T)(r.go_right ? r.eps : -r.eps) : (T)0;
        const T hen = (T)0.5 * en;
        const int fd = go ? (1 | (cont ? 2 : 0) | ((cont && take) ? 4 : 0) | (r.go_right ? 8 : 0)) : 0;
        // ring rows: 0 P, 1 g source (xb / edge g), 2 q source (xa / edge q), 5 imm, 6 V, 7 w source (xc / edge w)
        [[maybe_unused]] auto slot = [&](int st, int row) -> unsigned { return ring_s + (unsigned)((st * NROW + row) * kRowBytes); };
        unsigned act = Co::mask((fd & 1) != 0, slots, 13);
        [[maybe_unused]] unsigned iss = act;
        [[maybe_unused]] auto issue = [&](int st) {
            if (iss) {
                const int i = __ffs(iss) - 1;
                iss &= iss - 1;
                const int f = __shfl_sync(kFull, fd, i);
                if (j0 < d) {
                    const bool ct = f & 2, gr = f & 8;
                    const int ci = c0 + i;
                    const i64 rb = (i64)ci * d + j0;
                    cp16(slot(st, 0), (gr ? v.pr : v.pl) + rb);
                    cp16(slot(st, 1), (ct ? v.xb : (gr ? v.gr : v.gl)) + rb);
                    cp16(slot(st, 2), (ct ? v.xa : (gr ? v.qr : v.ql)) + rb);
                    if (DENSE) {
                        cp16(slot(st, 6), (gr ? v.vr : v.vl) + rb);
                        cp16(slot(st, 7), (ct ? v.xc : (gr ? v.wr : v.wl)) + rb);
                    } else if (!imm_shared) {
                        cp16(slot(st, 5), v.imm + (i64)ci * v.imm_sc + j0);
                    }
                }
            }
            cp_commit();
        };
        [[maybe_unused]] int stg = 0;
        if constexpr (TC > 1) {
#pragma unroll
            for (int st = 0; st < kRing - 1; ++st) issue(st);
        }
#pragma unroll 1
        while (act) {
            if constexpr (TC > 1) {
                issue(stg == 0 ? kRing - 1 : stg - 1);
                cp_wait<kRing - 1>();
            }
            const int i = __ffs(act) - 1;
            act &= act - 1;
            const int f = Co::get(fd, i, slots, 14);
            const T ei = Co::get(en, i, slots, 15);
            const T hei = Co::get(hen, i, slots, 16);
            const bool ct = f & 2, tk = f & 4, gr = f & 8;
            const int ci = c0 + i;
            const i64 rb = (i64)ci * d;
            T* __restrict__ P = (gr ? v.pr : v.pl) + rb;
            T* __restrict__ V = DENSE ? (gr ? v.vr : v.vl) + rb : nullptr;
            const T* __restrict__ Q = (gr ? v.qr : v.ql) + rb;
            const T* __restrict__ Gd = (gr ? v.gr : v.gl) + rb;
            const T* __restrict__ W = DENSE ? (gr ? v.wr : v.wl) + rb : nullptr;
            T* __restrict__ XA = v.xa + rb;
            const T* __restrict__ XB = v.xb + rb;
            const T* __restrict__ XC = DENSE ? v.xc + rb : nullptr;
            const T* __restrict__ IM = DENSE ? nullptr : v.imm + (i64)ci * v.imm_sc;
            // where the front is read from: continuing chains (q, g[, w]) = (xa, xb[, xc]), the others the edge arrays;
            // tiles of several chains: this stage of the ring
            const T* Pl = P;
            const T* Gl = ct ? XB : Gd;
            const T* Ql = ct ? (const T*)XA : Q;
            const T* Vl = V;
            const T* Wl = ct ? XC : W;
            const T* IMl = IM;
            [[maybe_unused]] auto from_stage = [&](int st) {
                const unsigned char* sb = ring + (size_t)(st * NROW) * kRowBytes;
                Pl = reinterpret_cast<const T*>(sb);
                Gl = reinterpret_cast<const T*>(sb + 1 * kRowBytes);
                Ql = reinterpret_cast<const T*>(sb + 2 * kRowBytes);
                if (!DENSE && !imm_shared) IMl = reinterpret_cast<const T*>(sb + 5 * kRowBytes);
                if (DENSE) {
                    Vl = reinterpret_cast<const T*>(sb + 6 * kRowBytes);
                    Wl = reinterpret_cast<const T*>(sb + 7 * kRowBytes);
                }
            };
            if constexpr (TC > 1) {
                from_stage(stg);
                stg = stg + 1 == kRing ? 0 : stg + 1;
            }
            // a continuing chain whose pass A left (p', g', q'[, imm p', imm g']) in shared memory: already kicked
            const bool stashed = TC == 1 && VEC > 1 && stash && ct;
            if (stashed) {
                Pl = stash_rows; Gl = stash_rows + d; Ql = stash_rows + 2 * d;
                if (DENSE) { Vl = stash_rows + 3 * d; Wl = stash_rows + 4 * d; }
            }
            for (int j = j0; j < d; j += STEP) {
                const int jl = TC > 1 ? j0 : j;
                T q[VEC], p[VEC], g[VEC], vel[VEC], w[VEC], im[VEC];
                ldv<T, VEC>(p, Pl + jl);
                ldv<T, VEC>(g, Gl + jl);
                ldv<T, VEC>(q, Ql + jl);
                if (DENSE) { ldv<T, VEC>(vel, Vl + jl); ldv<T, VEC>(w, Wl + jl); }
                else if (TC > 1 && imm_shared) {
#pragma unroll
                    for (int x = 0; x < VEC; ++x) im[x] = im_sh[x];
                } else ld_imm<T, VEC>(im, IMl, imm_shared ? j : jl, v.imm_sj);
                if (ct) {
                    if (!stashed) {
#pragma unroll
                        for (int x = 0; x < VEC; ++x) {
                            p[x] = p[x] - hei * g[x];                // the kick pass A applied (same e: same sub-tree)
                            if (DENSE) vel[x] = vel[x] - hei * w[x];
                        }
                    }
                    if (tk) {
                        stv<T, VEC>(v.qs + rb + j, q); stv<T, VEC>(v.ps + rb + j, p); stv<T, VEC>(v.gs + rb + j, g);
                        if (DENSE) stv<T, VEC>(v.ws + rb + j, w);
                    }
                }
                T ph[VEC], vh[VEC], qn[VEC];
#pragma unroll
                for (int x = 0; x < VEC; ++x) {
                    ph[x] = p[x] - hei * g[x];
                    if (DENSE) vh[x] = vel[x] - hei * w[x];          // imm.(p - h g) by linearity
                    else vh[x] = im[x] * ph[x];
                    qn[x] = q[x] + ei * vh[x];
                }
                stv<T, VEC>(P + j, ph);
                if (DENSE) stv<T, VEC>(V + j, vh);
                stv<T, VEC>(XA + j, qn);
            }
        }
    }

    // ---- epilogue -------------------------------------------------------------------------------------------------
    if (own && ph0 != PH_DONE) *static_cast<ChainRecLive*>(v.rec + c) = r;
    if constexpr (WPC == 1) {
        const unsigned ran = __ballot_sync(kFull, run);
        const unsigned alive = __ballot_sync(kFull, run && r.phase != PH_DONE);
        if (lane == 0) {
            if (not_done && alive) atomicAdd(not_done, __popc(alive));
            if (v.counters && ran) atomicAdd((unsigned long long*)&v.counters[3], (unsigned long long)__popc(ran));
        }
    } else if (own && run) {
        if (not_done && r.phase != PH_DONE) atomicAdd(not_done, 1);
        if (v.counters) atomicAdd((unsigned long long*)&v.counters[3], 1ull);
    }
}

}  // namespace tile

// chains per warp: enough warps to fill the machine first (16 per SM), then as many chains per warp as possible so
// that the scalar passes run with all lanes busy
static inline int tile_chains_per_warp(int C, int sm_count) {
    // B2H_TILE_TC / B2H_TILE_WPC (read at every call): force a layout, for the parity tests and A/B measurements
    const char* e = getenv("B2H_TILE_TC");
    const int forced = e ? atoi(e) : 0;
    if (forced == 1 || forced == 8 || forced == 32) return forced;
    const i64 want = (i64)sm_count * 16;
    if ((i64)C >= 32 * want) return 32;
    if ((i64)C >= 8 * want) return 8;
    return 1;
}

// Dynamic shared memory: tiles of several chains read their rows through the ring (kRing stages x rows x 512 bytes per
// warp); one chain per warp / CTA stashes the rows pass D needs again (3 or 5 rows per chain) when they fit `stash_bytes`.
template <typename T, int TC, int WPC, int VEC, bool DENSE>
static void launch_tile_smem(cudaStream_t st, const EngineView<T>& v, int* nd, int p, int grid) {
    constexpr int threads = tile::Coop<TC, WPC>::kThreads;
    int bytes = 0, flags = p;
    if (TC > 1) {
        bytes = (threads / 32) * tile::kRing * tile::RingRows<DENSE>::value * tile::kRowBytes;
    } else if (VEC > 1) {
        const char* e = getenv("B2H_TILE_STASH");                 // 0: re-read the rows instead (A/B measurements)
        const size_t need = (size_t)(WPC > 1 ? 1 : 4) * (DENSE ? 5 : 3) * v.d * sizeof(T);
        if (need <= 56 * 1024 && !(e && atoi(e) == 0)) { bytes = (int)need; flags |= 2; }
    }
    // set at every launch (cheap): the attribute belongs to the function in the CURRENT device's context
    if (bytes > 48 * 1024)
        cudaFuncSetAttribute(tile::tile_tick_kernel<T, TC, WPC, VEC, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    tile::tile_tick_kernel<T, TC, WPC, VEC, DENSE><<<grid, threads, bytes, st>>>(v, nd, flags);
}

// true when launch_tile_tick will run a one-chain layout with vector rows (TC == 1, VEC > 1): the layouts whose pass A
// can form the potential of a Gaussian target itself (EngineView::u_center)
template <typename T, bool DENSE>
static bool tile_tick_unit_chain_layout(const EngineView<T>& v, int sm_count) {
    const bool aligned = ((size_t)v.d * sizeof(T)) % 16 == 0 && ((uintptr_t)v.imm % 16 == 0 || DENSE) &&
                         (uintptr_t)v.out.draws % 16 == 0;
    const bool short_rows = (size_t)v.d * sizeof(T) <= (size_t)tile::kRowBytes;
    return aligned && (short_rows ? tile_chains_per_warp(v.C, sm_count) : 1) == 1;
}

template <typename T, bool DENSE>
static void launch_tile_tick(cudaStream_t st, const EngineView<T>& v, int* nd, bool pre, int sm_count) {
    constexpr int NV = 16 / (int)sizeof(T);
    const bool aligned = ((size_t)v.d * sizeof(T)) % 16 == 0 && ((uintptr_t)v.imm % 16 == 0 || DENSE) &&
                         (uintptr_t)v.out.draws % 16 == 0;
    // several chains per warp: rows of one 16-byte piece per lane (the ring pipelines the chains)
    const bool short_rows = (size_t)v.d * sizeof(T) <= (size_t)tile::kRowBytes;
    const int tc = aligned && short_rows ? tile_chains_per_warp(v.C, sm_count) : 1;
    const int p = pre ? 1 : 0;
    const i64 warps = ((i64)v.C + tc - 1) / tc;
    const int grid = (int)((warps + 3) / 4);
    if (!aligned) { tile::tile_tick_kernel<T, 1, 1, 1, DENSE><<<grid, 128, 0, st>>>(v, nd, p); return; }
    if (tc == 32) { launch_tile_smem<T, 32, 1, NV, DENSE>(st, v, nd, p, grid); return; }
    if (tc == 8) { launch_tile_smem<T, 8, 1, NV, DENSE>(st, v, nd, p, grid); return; }
    // one chain per warp, or -- long rows -- per CTA of 4 (8) warps, direct loads (pipelining the PIECES of a long row through
    // the ring was measured at c2: 156 us against 154 us, its shared memory costs one of the four CTAs per SM)
    const char* e = getenv("B2H_TILE_WPC");
    const int forced = e ? atoi(e) : 0;
    const int pieces = v.d / NV;
    int wpc = pieces > 64 ? 4 : 1;
    if (forced == 1 || forced == 4 || forced == 8) wpc = forced;
    if (wpc == 8) launch_tile_smem<T, 1, 8, NV, DENSE>(st, v, nd, p, v.C);
    else if (wpc == 4) launch_tile_smem<T, 1, 4, NV, DENSE>(st, v, nd, p, v.C);
    else launch_tile_smem<T, 1, 1, NV, DENSE>(st, v, nd, p, grid);
}

}  // namespace b2h
