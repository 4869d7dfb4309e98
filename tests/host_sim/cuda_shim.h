// TEST TOOLING ONLY: lets g++ compile the engine's device headers (thread-per-chain
// instantiation, G = 1) so that the chain state machine can be checked against the
// oracle on a machine without a GPU.  Never linked into libb200hmc.so; the product
// has no CPU path.
#pragma once
#define B2H_HOST_SIM 1
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __align__(x) alignas(x)
#define __restrict__
#define __launch_bounds__(...)

struct sim_dim3 { unsigned x, y, z; };
static thread_local sim_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1};

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline void sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0) {}
