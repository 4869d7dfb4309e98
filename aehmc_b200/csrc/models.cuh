// Built-in log-density gradients as group-cooperative device functions.
//
// The reference differentiates a user logprob_fn with aesara.grad
// (hmc.py:33-34, integrators.py:64-65); the engine's built-in targets are
// stated first in oracle/models.py and restated here operation for operation.
// q and g are this chain's rows with element stride sj.
#pragma once

#include "common.cuh"

namespace b2h {

struct ModelDev {
    int kind, dim;
    i64 n_data;
    const void *a, *b, *c;
    double s0, s1;
};

enum { MODEL_IID = 0, MODEL_CORR = 1, MODEL_FUNNEL = 2, MODEL_SCHOOLS = 3, MODEL_LOGISTIC = 4 };

// U = 1/2 sum r g + s0,  r = q - mu,  g = r * inv_var      (oracle/models.py:IIDGaussian)
template <typename T, int G>
B2H_DEVINL T grad_iid(const ModelDev& m, const T* q, T* g, i64 sj, int lane, double* red) {
    const T* mu = (const T*)m.a;
    const T* iv = (const T*)m.b;
    T acc = 0;
    for (int j = lane; j < m.dim; j += G) {
        T r = q[(i64)j * sj] - mu[j];
        T gj = r * iv[j];
        g[(i64)j * sj] = gj;
        acc += r * gj;
    }
    return (T)0.5 * (T)Group<G>::sum1((double)acc, red) + (T)m.s0;
}

// Neal's funnel                                              (oracle/models.py:NealFunnel)
template <typename T, int G>
B2H_DEVINL T grad_funnel(const ModelDev& m, const T* q, T* g, i64 sj, int lane, double* red) {
    const int d = m.dim;
    T vv = q[0];
    T acc = 0;
    for (int j = lane; j < d; j += G)
        if (j >= 1) { T x = q[(i64)j * sj]; acc += x * x; }
    T ss = (T)Group<G>::sum1((double)acc, red);
    T ev = exp(-vv);
    T n = (T)(d - 1);
    T U = vv * vv / (T)18 + (T)0.5 * ev * ss + (T)0.5 * n * vv;
    for (int j = lane; j < d; j += G) {
        if (j == 0) g[0] = vv / (T)9 - (T)0.5 * ev * ss + (T)0.5 * n;
        else g[(i64)j * sj] = q[(i64)j * sj] * ev;
    }
    return U;
}

// Non-centred eight schools, q = (mu, log tau, theta~_1..J) (oracle/models.py:EightSchools)
template <typename T, int G>
B2H_DEVINL T grad_schools(const ModelDev& m, const T* q, T* g, i64 sj, int lane, double* red) {
    const int d = m.dim;
    const T* y = (const T*)m.a;
    const T* iv = (const T*)m.b;
    T mu = q[0], t = q[sj];
    T tau = exp(t);
    T a = tau * tau / (T)25;
    T s_w = 0, s_wth = 0, s_th2 = 0, s_rw = 0;
    for (int j = lane; j < d; j += G) {
        if (j >= 2) {
            T th = q[(i64)j * sj];
            T resid = y[j - 2] - mu - tau * th;
            T w = resid * iv[j - 2];
            g[(i64)j * sj] = th - tau * w;
            s_w += w; s_wth += w * th; s_th2 += th * th; s_rw += resid * w;
        }
    }
    double r4[4] = {(double)s_w, (double)s_wth, (double)s_th2, (double)s_rw};
    Group<G>::template sum<4>(r4, red);
    T U = mu * mu / (T)50 - t + log1p(a) + (T)0.5 * (T)r4[2] + (T)0.5 * (T)r4[3];
    if (lane == 0) {
        g[0] = mu / (T)25 - (T)r4[0];
        g[sj] = (T)-1 + (T)2 * a / ((T)1 + a) - tau * (T)r4[1];
    }
    return U;
}

// ---- register-front versions (the front's q and g live in registers; see engine.cuh RegFront) -------------
template <typename T, int G, class Front>
B2H_DEVINL T grad_iid_front(const ModelDev& m, Front& f, int lane, double* red) {
    const T* mu = (const T*)m.a;
    const T* iv = (const T*)m.b;
    T acc = 0;
#pragma unroll
    for (int e = 0; e < Front::kE; ++e) {
        const int j = lane + e * G;
        if (j < m.dim) {
            T r = f.fq[e] - mu[j];
            T gj = r * iv[j];
            f.fg[e] = gj;
            acc += r * gj;
        }
    }
    return (T)0.5 * (T)Group<G>::sum1((double)acc, red) + (T)m.s0;
}

template <typename T, int G, class Front>
B2H_DEVINL T grad_funnel_front(const ModelDev& m, Front& f, int lane, double* red) {
    const int d = m.dim;
    const T vv = Group<G>::shfl(f.fq[0], 0);              // coordinate 0 lives in lane 0, element 0
    T acc = 0;
#pragma unroll
    for (int e = 0; e < Front::kE; ++e) {
        const int j = lane + e * G;
        if (j >= 1 && j < d) acc += f.fq[e] * f.fq[e];
    }
    T ss = (T)Group<G>::sum1((double)acc, red);
    T ev = exp(-vv);
    T n = (T)(d - 1);
    T U = vv * vv / (T)18 + (T)0.5 * ev * ss + (T)0.5 * n * vv;
#pragma unroll
    for (int e = 0; e < Front::kE; ++e) {
        const int j = lane + e * G;
        if (j == 0) f.fg[e] = vv / (T)9 - (T)0.5 * ev * ss + (T)0.5 * n;
        else if (j < d) f.fg[e] = f.fq[e] * ev;
    }
    return U;
}

template <typename T, int G, class Front>
B2H_DEVINL T grad_schools_front(const ModelDev& m, Front& f, int lane, double* red) {
    const int d = m.dim;
    const T* y = (const T*)m.a;
    const T* iv = (const T*)m.b;
    const T mu = Group<G>::shfl(f.fq[0], 0);
    const T t = Group<G>::shfl(f.fq[1 / G], 1 % G);       // coordinate 1: lane 1 % G, element 1 / G
    T tau = exp(t);
    T a = tau * tau / (T)25;
    T s_w = 0, s_wth = 0, s_th2 = 0, s_rw = 0;
#pragma unroll
    for (int e = 0; e < Front::kE; ++e) {
        const int j = lane + e * G;
        if (j >= 2 && j < d) {
            T th = f.fq[e];
            T resid = y[j - 2] - mu - tau * th;
            T w = resid * iv[j - 2];
            f.fg[e] = th - tau * w;
            s_w += w; s_wth += w * th; s_th2 += th * th; s_rw += resid * w;
        }
    }
    double r4[4] = {(double)s_w, (double)s_wth, (double)s_th2, (double)s_rw};
    Group<G>::template sum<4>(r4, red);
    T U = mu * mu / (T)50 - t + log1p(a) + (T)0.5 * (T)r4[2] + (T)0.5 * (T)r4[3];
#pragma unroll
    for (int e = 0; e < Front::kE; ++e) {
        const int j = lane + e * G;
        if (j == 0) f.fg[e] = mu / (T)25 - (T)r4[0];
        if (j == 1) f.fg[e] = (T)-1 + (T)2 * a / ((T)1 + a) - tau * (T)r4[1];
    }
    return U;
}

template <typename T, int G, int MODEL, class Front>
B2H_DEVINL T model_grad_front(const ModelDev& m, Front& f, int lane, double* red) {
    if (MODEL == MODEL_IID) return grad_iid_front<T, G>(m, f, lane, red);
    if (MODEL == MODEL_FUNNEL) return grad_funnel_front<T, G>(m, f, lane, red);
    return grad_schools_front<T, G>(m, f, lane, red);
}

template <typename T, int G, int MODEL>
B2H_DEVINL T model_grad(const ModelDev& m, const T* q, T* g, i64 sj, int lane, double* red) {
    if (MODEL == MODEL_IID) return grad_iid<T, G>(m, q, g, sj, lane, red);
    if (MODEL == MODEL_FUNNEL) return grad_funnel<T, G>(m, q, g, sj, lane, red);
    return grad_schools<T, G>(m, q, g, sj, lane, red);
}

}  // namespace b2h
