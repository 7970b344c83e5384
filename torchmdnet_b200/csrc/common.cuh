// Shared helpers for the sm_100a kernels of libtmdnet_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tn_math.h"

#define TMD_OK 0
#define TMD_ERR_INVALID (-1)   // bad argument (shape / alignment / null pointer)
#define TMD_ERR_CUDA (-2)      // CUDA launch error (see tmd_last_error)
#define TMD_ERR_UNSUPPORTED (-3)

void tmd_set_error(const char* msg);
int tmd_check_launch(const char* what);

#define TMD_REQUIRE(cond, msg)            \
  do {                                    \
    if (!(cond)) {                        \
      tmd_set_error(msg);                 \
      return TMD_ERR_INVALID;             \
    }                                     \
  } while (0)

static inline int tmd_blocks_for(long long work, int threads, int max_blocks = 148 * 16) {
  long long b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;  // grid-stride loops; 148 SMs x resident CTAs
  return (int)b;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Transposing warp reduction: every lane holds V partial values; afterwards lane l (< V) returns the warp-wide total of
// value l.  V-1 shuffles in total (instead of 5 per value) and no serial dependency between the values.
template <int V>
__device__ __forceinline__ float warp_reduce_transpose(float (&v)[V], int lane) {
  static_assert(V == 32 || V == 16 || V == 8, "V must be 8, 16 or 32");
#pragma unroll
  for (int s = V / 2; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  // lanes >= V (when V < 32) hold a copy of lane (l mod V)'s total after the remaining full-warp steps
  float r = v[0];
#pragma unroll
  for (int s = 16; s >= V; s >>= 1) r += __shfl_xor_sync(0xffffffffu, r, s);
  return r;
}
