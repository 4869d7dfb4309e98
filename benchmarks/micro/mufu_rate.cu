// Micro-benchmark: MUFU (ex2 / rcp / lg2) issue rate per SM on sm_100a, 4..32 warps per SM, 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = 1.0f + 0.001f * (threadIdx.x + i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
        }
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
    const char* names[4] = {"ex2", "rcp", "lg2", "ffma"};
    for (int op = 0; op < 4; ++op)
        for (int warps : {4, 8, 16, 32}) {
            const int iters = 2000;
            if (op == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
            if (op == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
            if (op == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
            if (op == 3) k<3><<<148, warps * 32>>>(out, iters, cyc);
            cudaDeviceSynchronize();
            double lane_ops = (double)warps * 32 * 8 * iters;
            printf("%s warps/SM=%d cycles=%lld lane-ops/clk/SM=%.1f\n", names[op], warps, *cyc, lane_ops / *cyc);
        }
    return 0;
}
