// common.cuh - shared helpers for libacx (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/acx.h"

namespace acx {

void set_error(const std::string& msg);
extern uint64_t g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += (uint64_t)n; }

#define ACX_CHECK(cond, msg)                                             \
  do {                                                                   \
    if (!(cond)) {                                                       \
      acx::set_error(std::string(__func__) + ": " + (msg));              \
      return 1;                                                          \
    }                                                                    \
  } while (0)

#define ACX_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      acx::set_error(std::string(__func__) + ": " #expr ": " + cudaGetErrorString(_e));        \
      return 2;                                                                                \
    }                                                                                          \
  } while (0)

#define ACX_LAUNCH_CHECK()                                                                     \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      acx::set_error(std::string(__func__) + ": launch: " + cudaGetErrorString(_e));           \
      return 3;                                                                                \
    }                                                                                          \
    acx::count_launch();                                                                       \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// split x into up to 3 bf16 planes with x ~= p0 + p1 + p2 (round-to-nearest each time)
__device__ __forceinline__ void split3(float x, bf16& p0, bf16& p1, bf16& p2) {
  p0 = __float2bfloat16_rn(x);
  float r = x - __bfloat162float(p0);
  p1 = __float2bfloat16_rn(r);
  r = r - __bfloat162float(p1);
  p2 = __float2bfloat16_rn(r);
}

}  // namespace acx
