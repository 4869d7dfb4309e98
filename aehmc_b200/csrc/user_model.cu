// User-supplied log-densities: the reference takes an arbitrary Python `logprob_fn` and
// differentiates it with aesara.grad (reference hmc.py:33-34, integrators.py:23); device code needs the gradient as
// device code, so the user writes ONE CUDA C++ function
//
//     template <typename T>
//     __device__ T potential_and_grad(const T* q, T* g, int d, const T* data);      // returns U(q), writes dU/dq
//
// which is compiled at run time with NVRTC for sm_100a and wrapped in a thread-per-chain kernel.  The model runs in
// the engine's split mode (one gradient launch per tick).  NVRTC is opened with dlopen: the library has no link-time
// dependency on it and b2h_user_model_create fails loudly when it is missing.
#include <dlfcn.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "launch.h"

struct b2h_user_model {
    cudaLibrary_t lib;
    cudaKernel_t k32, k64;
};

namespace b2h {

namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    int (*CompileProgram)(nvrtcProgram, int, const char* const*);
    int (*GetProgramLogSize)(nvrtcProgram, size_t*);
    int (*GetProgramLog)(nvrtcProgram, char*);
    int (*GetCUBINSize)(nvrtcProgram, size_t*);
    int (*GetCUBIN)(nvrtcProgram, char*);
    int (*DestroyProgram)(nvrtcProgram*);
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    void* h = nullptr;
    for (const char* nm : names)
        if ((h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) return n;
#define B2H_SYM(field, name) *(void**)(&n.field) = dlsym(h, name); if (!n.field) return n;
    B2H_SYM(CreateProgram, "nvrtcCreateProgram")
    B2H_SYM(CompileProgram, "nvrtcCompileProgram")
    B2H_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    B2H_SYM(GetProgramLog, "nvrtcGetProgramLog")
    B2H_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    B2H_SYM(GetCUBIN, "nvrtcGetCUBIN")
    B2H_SYM(DestroyProgram, "nvrtcDestroyProgram")
#undef B2H_SYM
    n.ok = true;
    return n;
}

// Forward-mode automatic differentiation for user densities (b2h_user_model_create_ad): the user writes only
//     template <typename S, typename T> __device__ S log_density(const S* q, int d, const T* data);
// with the overloaded arithmetic below; S = Dual<T, B2H_AD_DIM> carries the value and all d partial derivatives, so
// one evaluation yields logprob and its gradient (cost O(d) per operation: meant for small models -- the role
// aesara.grad plays for the reference's logprob_fn, hmc.py:33-34).
const char* kDualHeader = R"(
template <typename T, int N>
struct Dual {
    T v;
    T g[N];
    __device__ Dual() : v(0) { for (int i = 0; i < N; ++i) g[i] = 0; }
    __device__ Dual(T x) : v(x) { for (int i = 0; i < N; ++i) g[i] = 0; }
    __device__ Dual(double x, int) : v((T)x) { for (int i = 0; i < N; ++i) g[i] = 0; }
};
#define B2H_D Dual<T, N>
#define B2H_TN template <typename T, int N> __device__ inline
B2H_TN B2H_D operator+(const B2H_D& a, const B2H_D& b) { B2H_D r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.g[i] = a.g[i] + b.g[i]; return r; }
B2H_TN B2H_D operator-(const B2H_D& a, const B2H_D& b) { B2H_D r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.g[i] = a.g[i] - b.g[i]; return r; }
B2H_TN B2H_D operator*(const B2H_D& a, const B2H_D& b) { B2H_D r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i]; return r; }
B2H_TN B2H_D operator/(const B2H_D& a, const B2H_D& b) { B2H_D r; const T ib = (T)1 / b.v; r.v = a.v * ib; for (int i = 0; i < N; ++i) r.g[i] = (a.g[i] - r.v * b.g[i]) * ib; return r; }
B2H_TN B2H_D operator-(const B2H_D& a) { B2H_D r; r.v = -a.v; for (int i = 0; i < N; ++i) r.g[i] = -a.g[i]; return r; }
B2H_TN B2H_D operator+(const B2H_D& a, T b) { B2H_D r = a; r.v += b; return r; }
B2H_TN B2H_D operator+(T a, const B2H_D& b) { return b + a; }
B2H_TN B2H_D operator-(const B2H_D& a, T b) { B2H_D r = a; r.v -= b; return r; }
B2H_TN B2H_D operator-(T a, const B2H_D& b) { return (-b) + a; }
B2H_TN B2H_D operator*(const B2H_D& a, T b) { B2H_D r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.g[i] = a.g[i] * b; return r; }
B2H_TN B2H_D operator*(T a, const B2H_D& b) { return b * a; }
B2H_TN B2H_D operator/(const B2H_D& a, T b) { return a * ((T)1 / b); }
B2H_TN B2H_D operator/(T a, const B2H_D& b) { return B2H_D(a) / b; }
B2H_TN B2H_D& operator+=(B2H_D& a, const B2H_D& b) { a = a + b; return a; }
B2H_TN B2H_D& operator-=(B2H_D& a, const B2H_D& b) { a = a - b; return a; }
B2H_TN B2H_D& operator+=(B2H_D& a, T b) { a.v += b; return a; }
B2H_TN B2H_D& operator-=(B2H_D& a, T b) { a.v -= b; return a; }
B2H_TN B2H_D chain1(const B2H_D& a, T f, T df) { B2H_D r; r.v = f; for (int i = 0; i < N; ++i) r.g[i] = df * a.g[i]; return r; }
B2H_TN B2H_D exp(const B2H_D& a) { const T e = exp(a.v); return chain1(a, e, e); }
B2H_TN B2H_D log(const B2H_D& a) { return chain1(a, log(a.v), (T)1 / a.v); }
B2H_TN B2H_D log1p(const B2H_D& a) { return chain1(a, log1p(a.v), (T)1 / ((T)1 + a.v)); }
B2H_TN B2H_D sqrt(const B2H_D& a) { const T s = sqrt(a.v); return chain1(a, s, (T)0.5 / s); }
B2H_TN B2H_D tanh(const B2H_D& a) { const T t = tanh(a.v); return chain1(a, t, (T)1 - t * t); }
B2H_TN B2H_D sin(const B2H_D& a) { return chain1(a, sin(a.v), cos(a.v)); }
B2H_TN B2H_D cos(const B2H_D& a) { return chain1(a, cos(a.v), -sin(a.v)); }
B2H_TN B2H_D square(const B2H_D& a) { return chain1(a, a.v * a.v, (T)2 * a.v); }
B2H_TN B2H_D pow(const B2H_D& a, T p) { const T f = pow(a.v, p); return chain1(a, f, p * f / a.v); }
B2H_TN B2H_D softplus(const B2H_D& a) {       // log(1 + exp(x)), stable
    const T e = exp(-fabs(a.v));
    return chain1(a, fmax(a.v, (T)0) + log1p(e), a.v >= (T)0 ? (T)1 / ((T)1 + e) : e / ((T)1 + e));
}
#undef B2H_D
#undef B2H_TN
)";

const char* kDualWrapper = R"(
template <typename T>
__device__ T potential_and_grad(const T* q, T* g, int d, const T* data) {
    typedef Dual<T, B2H_AD_DIM> S;
    S x[B2H_AD_DIM];
    for (int i = 0; i < d; ++i) { x[i].v = q[i]; x[i].g[i] = (T)1; }
    const S lp = log_density<S, T>(x, d, data);
    for (int i = 0; i < d; ++i) g[i] = -lp.g[i];
    return -lp.v;
}
)";

const char* kWrapper = R"(
extern "C" __global__ void b2h_user_kernel_f64(const double* q, double* U, double* g, long long C, int d, const double* data) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) U[c] = potential_and_grad<double>(q + c * d, g + c * d, d, data);
}
extern "C" __global__ void b2h_user_kernel_f32(const float* q, float* U, float* g, long long C, int d, const float* data) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) U[c] = potential_and_grad<float>(q + c * d, g + c * d, d, data);
}
)";

}  // namespace

template <typename T>
int user_potential_and_grad(b2h_ctx* ctx, const b2h_model* m, const T* q, T* U, T* g, i64 C) {
    const b2h_user_model* um = (const b2h_user_model*)m->a;
    if (!um) { set_error("user model: null handle (b2h_model.a must hold the b2h_user_model*)"); return B2H_ERR_ARG; }
    long long Cl = C;
    int d = m->dim;
    const void* data = m->b;
    void* args[] = {(void*)&q, (void*)&U, (void*)&g, (void*)&Cl, (void*)&d, (void*)&data};
    const int threads = 128;
    const unsigned grid = (unsigned)((C + threads - 1) / threads);
    B2H_CUDA(cudaLaunchKernel((const void*)(sizeof(T) == 8 ? um->k64 : um->k32), dim3(grid), dim3(threads), args, 0,
                              ctx->stream));
    return 0;
}
template int user_potential_and_grad<float>(b2h_ctx*, const b2h_model*, const float*, float*, float*, i64);
template int user_potential_and_grad<double>(b2h_ctx*, const b2h_model*, const double*, double*, double*, i64);

}  // namespace b2h

static int user_model_compile(const std::string& src, int ad_dim, b2h_user_model** out) {
    using namespace b2h;
    Nvrtc& rt = nvrtc();
    if (!rt.ok) { set_error("b2h_user_model_create: libnvrtc.so.12 not found (NVRTC is needed for user models)"); return B2H_ERR_UNSUPPORTED; }
    nvrtcProgram prog = nullptr;
    if (rt.CreateProgram(&prog, src.c_str(), "b2h_user_model.cu", 0, nullptr, nullptr) != 0) {
        set_error("nvrtcCreateProgram failed");
        return B2H_ERR_CUDA;
    }
    const std::string ad = "-DB2H_AD_DIM=" + std::to_string(ad_dim > 0 ? ad_dim : 1);
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", ad.c_str()};
    const int rc = rt.CompileProgram(prog, 4, opts);
    if (rc != 0) {
        size_t n = 0;
        rt.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) rt.GetProgramLog(prog, &log[0]);
        rt.DestroyProgram(&prog);
        set_error("user model does not compile:\n" + log);
        return B2H_ERR_ARG;
    }
    size_t nb = 0;
    rt.GetCUBINSize(prog, &nb);
    std::vector<char> cubin(nb);
    rt.GetCUBIN(prog, cubin.data());
    rt.DestroyProgram(&prog);
    b2h_user_model* um = new b2h_user_model();
    cudaError_t e = cudaLibraryLoadData(&um->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&um->k64, um->lib, "b2h_user_kernel_f64");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&um->k32, um->lib, "b2h_user_kernel_f32");
    if (e != cudaSuccess) {
        set_error(std::string("user model: loading the compiled code failed: ") + cudaGetErrorString(e));
        delete um;
        return B2H_ERR_CUDA;
    }
    *out = um;
    return 0;
}

extern "C" int b2h_user_model_create(const char* cuda_source, b2h_user_model** out) {
    if (!cuda_source || !out) { b2h::set_error("b2h_user_model_create: null argument"); return B2H_ERR_ARG; }
    return user_model_compile(std::string(cuda_source) + b2h::kWrapper, 0, out);
}

extern "C" int b2h_user_model_create_ad(const char* cuda_source, int32_t dim, b2h_user_model** out) {
    if (!cuda_source || !out) { b2h::set_error("b2h_user_model_create_ad: null argument"); return B2H_ERR_ARG; }
    if (dim < 1 || dim > 64) { b2h::set_error("b2h_user_model_create_ad: dim must be in [1, 64] (forward mode costs O(dim) per operation)"); return B2H_ERR_ARG; }
    return user_model_compile(std::string(b2h::kDualHeader) + cuda_source + b2h::kDualWrapper + b2h::kWrapper, dim, out);
}

extern "C" int b2h_user_model_destroy(b2h_user_model* um) {
    if (!um) return 0;
    cudaLibraryUnload(um->lib);
    delete um;
    return 0;
}
