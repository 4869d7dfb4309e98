// Shared device helpers for the B200 HMC/NUTS engine (sm_100a only).
//
// Nothing here is a port: the reference (aesara-devs/aehmc) has no native code.
// The helpers restate the scalar semantics its graphs rely on
// (proposals.py:41-52,96-100,130-144; termination.py:207-235) so that the
// kernels take the same decisions as the reference under injected draws.
#pragma once

#ifndef B2H_HOST_SIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <math.h>

#ifndef B2H_DEVINL
#define B2H_DEVINL __device__ __forceinline__
#endif
// Scalar helpers whose double-precision exp / log / Philox bodies are hundreds of instructions each: kept OUT of line
// so that the engine kernels, which call them at a dozen sites, stay small enough for the instruction cache (ncu:
// `no_instruction` was the top stall of the tick kernel and of the persistent kernel with everything inlined).  They
// take and return scalars only, so the chain record stays in registers across the calls.
#ifndef B2H_DEVCALL
#ifdef B2H_HOST_SIM
#define B2H_DEVCALL static inline
#else
#define B2H_DEVCALL static __device__ __noinline__
#endif
#endif

namespace b2h {

typedef long long i64;

// ---------------------------------------------------------------------------
// scalar semantics
// ---------------------------------------------------------------------------

// log(exp(a)+exp(b)) exactly as oracle/tree.py:logaddexp (max + log(sum exp(x-max)),
// -inf,-inf -> -inf).  Always double: weights are float64 in the reference
// whatever floatX is (nuts.py:123-124).
B2H_DEVINL double lae_inl(double a, double b) {
    if (isnan(a) || isnan(b)) return nan("");
    double m = a > b ? a : b;
    if (isinf(m)) {
        if (m < 0) return m;          // -inf + log(0) = -inf
        return m;                     // +inf
    }
    // exp(m - m) = exp(0) is exactly 1 and IEEE addition commutes, so one exponential gives the reference's
    // exp(a - m) + exp(b - m) bit for bit: half the transcendental work on the scalar critical path of every tick
    const double lo = a > b ? b : a;
    return m + log(1.0 + exp(lo - m));
}

B2H_DEVINL double expit_inl(double x) {
    if (isnan(x)) return x;
    if (x < -709.0) return 0.0;
    return 1.0 / (1.0 + exp(-x));
}

B2H_DEVCALL double lae(double a, double b) { return lae_inl(a, b); }
B2H_DEVCALL double expit(double x) { return expit_inl(x); }
// Who calls and who inlines (measured, config 4 / config 1 / config 5): the sub-warp layouts (several chains of one
// warp in different tree phases: instruction-cache bound) gain from the out-of-line copies (eight schools, 8 lanes per
// chain: 0.63 -> 0.74 G evals/s); thread-per-chain kernels reach the helpers under divergence, where a call costs more
// than it saves (0.91 -> 0.79); warp- and CTA-per-chain kernels lose the instruction-level parallelism between the
// independent draws of one lane (HMC d = 100: 1.50 -> 1.35).
// Thread per chain: with the rare paths of the persistent kernel deferred (engine_fused.inl) most lanes reach the helpers
// together, and the smaller loop body wins (eight schools c4: 89.5 -> 86.9 ms for 200 transitions, free-running 69.4 -> 66.1;
// before the deferral the calls under divergence cost more than they saved: 0.91 -> 0.79 G evals/s).
#ifndef B2H_CALL_G1
#define B2H_CALL_G1 1
#endif
template <int G> struct Helpers { static constexpr bool kCall = (G > 1 && G < 32) || (B2H_CALL_G1 && G == 1); };
template <int G> B2H_DEVINL double lae_g(double a, double b) { return Helpers<G>::kCall ? lae(a, b) : lae_inl(a, b); }
template <int G> B2H_DEVINL double expit_g(double x) { return Helpers<G>::kCall ? expit(x) : expit_inl(x); }

// Decision numpy's Generator.binomial(1, p) takes from its single uniform u
// (oracle/streams.py:bernoulli_from_uniform).  NaN p never accepts.
B2H_DEVINL bool bern(double u, double p) {
    if (p <= 0.5) return u > 1.0 - p;
    return u <= p;
}

// termination.py:207-235 in closed form: idx_max = popcount(step >> 1),
// number of sub-trees = trailing one-bits of step.
B2H_DEVINL void storage_indices(int step, int& idx_min, int& idx_max) {
    idx_max = __popc((unsigned)step >> 1);
    int trailing = __ffs(~step) - 1;
    idx_min = idx_max - trailing + 1;
}

B2H_DEVINL int uniform_slot(int expansion, int step) { return (1 << expansion) - 1 + (step - 1); }

// ---------------------------------------------------------------------------
// Philox4x32-10 counter RNG (native-RNG mode).  key = seed, counter =
// (slot, stream kind, transition, global chain id): results do not depend on
// how chains are sharded over GPUs.
// ---------------------------------------------------------------------------
enum DrawKind : uint32_t { DRAW_Z = 0, DRAW_DIR = 1, DRAW_BIASED = 2, DRAW_UNIFORM = 3, DRAW_ACCEPT = 4 };

B2H_DEVINL void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                              uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

B2H_DEVINL double u53(uint32_t lo, uint32_t hi) {          // [0, 1)
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

struct PhiloxKey { uint32_t k0, k1; };

B2H_DEVINL double philox_uniform_inl(PhiloxKey key, uint64_t chain, uint32_t transition, uint32_t kind, uint32_t slot) {
    uint32_t o[4];
    philox4x32_10(slot, kind, transition, (uint32_t)chain, key.k0 ^ (uint32_t)(chain >> 32), key.k1, o);
    return u53(o[0], o[1]);
}

// element j of the standard-normal momentum vector (Box-Muller on one Philox block per pair)
B2H_DEVINL double philox_normal_inl(PhiloxKey key, uint64_t chain, uint32_t transition, uint32_t j) {
    uint32_t o[4];
    philox4x32_10(j >> 1, DRAW_Z, transition, (uint32_t)chain, key.k0 ^ (uint32_t)(chain >> 32), key.k1, o);
    double u1 = 1.0 - u53(o[0], o[1]);                      // (0, 1]
    double u2 = u53(o[2], o[3]);
    double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    return (j & 1) ? r * s : r * c;
}

B2H_DEVCALL double philox_uniform(PhiloxKey key, uint64_t chain, uint32_t transition, uint32_t kind, uint32_t slot) {
    return philox_uniform_inl(key, chain, transition, kind, slot);
}
B2H_DEVCALL double philox_normal(PhiloxKey key, uint64_t chain, uint32_t transition, uint32_t j) {
    return philox_normal_inl(key, chain, transition, j);
}

// both normals of Box-Muller pair `pair` (elements 2 pair and 2 pair + 1 of the momentum vector): the same values
// philox_normal returns for them, from ONE Philox block
B2H_DEVCALL void philox_normal_pair(PhiloxKey key, uint64_t chain, uint32_t transition, uint32_t pair, double* z0,
                                    double* z1) {
    uint32_t o[4];
    philox4x32_10(pair, DRAW_Z, transition, (uint32_t)chain, key.k0 ^ (uint32_t)(chain >> 32), key.k1, o);
    double u1 = 1.0 - u53(o[0], o[1]);                      // (0, 1]
    double u2 = u53(o[2], o[3]);
    double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    *z0 = r * c;
    *z1 = r * s;
}

// ---------------------------------------------------------------------------
// group-of-G-threads reductions.  G in {1,2,4,8,16,32}: sub-warp groups, shuffles
// under the group's lane mask (other groups of the warp may be in another branch).
// G > 32: the group is the whole CTA (one chain per CTA), warp shuffle + smem.
// ---------------------------------------------------------------------------
template <int G>
struct Group {
    static constexpr bool kBlock = (G > 32);
    B2H_DEVINL static int lane() { return kBlock ? (int)threadIdx.x : (int)(threadIdx.x & (G - 1)); }
    B2H_DEVINL static unsigned mask() {
        if (G >= 32) return 0xffffffffu;
        unsigned l = threadIdx.x & 31u;
        return ((1u << (G & 31)) - 1u) << (l & ~(unsigned)((G & 31) - 1));
    }
    // all-reduce sum of N doubles across the group
    template <int N>
    B2H_DEVINL static void sum(double (&v)[N], double* smem /* >= N*(G/32 + 1) doubles when kBlock */) {
        if (G == 1) return;
        constexpr int W = G >= 32 ? 32 : G;
        const unsigned m = mask();
#pragma unroll
        for (int off = W / 2; off > 0; off >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(m, v[i], off);
        }
        if (kBlock) {
            constexpr int NW = G / 32;
            int w = threadIdx.x >> 5;
            __syncthreads();                       // smem reuse guard
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) smem[i * NW + w] = v[i];
            }
            __syncthreads();
            // N threads add the NW partials of one value each (same order as before); everybody reads the N totals
            if (threadIdx.x < N) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < NW; ++k) s += smem[threadIdx.x * NW + k];
                smem[N * NW + threadIdx.x] = s;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] = smem[N * NW + i];
        }
    }
    B2H_DEVINL static double sum1(double x, double* smem) {
        double v[1] = {x};
        sum<1>(v, smem);
        return v[0];
    }
    // value held by lane `src` of the group (sub-warp / warp groups only), broadcast to the whole group
    template <typename V>
    B2H_DEVINL static V shfl(V x, int src) {
        if (G == 1) return x;
        return __shfl_sync(mask(), x, (int)((threadIdx.x & 31u) & ~(unsigned)((G & 31) - 1)) + src);
    }
    // value held by the group's lane 0, broadcast to the whole group
    B2H_DEVINL static int bcast(int x, double* smem) {
        if (G == 1) return x;
        if (!kBlock) return __shfl_sync(mask(), x, (threadIdx.x & 31u) & ~(unsigned)((G & 31) - 1));
        __syncthreads();
        if (threadIdx.x == 0) reinterpret_cast<int*>(smem)[0] = x;
        __syncthreads();
        int r = reinterpret_cast<int*>(smem)[0];
        __syncthreads();
        return r;
    }
    B2H_DEVINL static void sync() {
        if (kBlock) __syncthreads();
        else if (G > 1) __syncwarp(mask());
    }
};

template <typename T> struct Num;
template <> struct Num<float> { static constexpr int dtype = 0; };
template <> struct Num<double> { static constexpr int dtype = 1; };

}  // namespace b2h
