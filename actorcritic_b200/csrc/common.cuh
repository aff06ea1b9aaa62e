// common.cuh - shared helpers for libacx (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <cstring>
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

// Programmatic dependent launch: a kernel launched with `launch_pdl` may start (block scheduling, barrier / TMEM set-up)
// while its predecessor in the stream is still running its tail; it must call pdl_wait() before it touches global memory
// (waits until the predecessor grid has completed and its writes are visible).  pdl_trigger() in the predecessor allows the
// successor's blocks to be scheduled from that point on (otherwise: when the predecessor has finished).  ACX_PDL=0 launches
// everything with plain stream ordering.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// small kernels: let the successor be scheduled at once (it blocks in its own pdl_wait without holding much of an SM), then wait
__device__ __forceinline__ void pdl_enter() {
  pdl_trigger();
  pdl_wait();
}
void set_pdl_override(int level);   // >= 0: the level of the following launches (the acting step: one serial chain); -1: back to ACX_PDL
int pdl_level();   // ACX_PDL: 0 = off, 1 = tensor-core kernels and their finalize kernels, 2 = also the small kernels
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_at(int level, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_level() >= level ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_pdl_at(1, kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}
// launch + error check + launch count, as ACX_LAUNCH_CHECK does for <<< >>> launches
#define ACX_PDL_LAUNCH(kernel, grid, block, smem, st, ...)                                     \
  do {                                                                                         \
    ACX_CUDA(acx::launch_pdl_at(2, kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)); \
    acx::count_launch();                                                                       \
  } while (0)

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

// the same split for two values at a time with packed conversions (F2FP.BF16.F32.PACK_AB converts two floats per
// instruction on the ALU pipe; the scalar F2F.BF16.F32 goes through the conversion unit): bit-identical planes,
// element 0 in the low half of each word.
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void split3x2(float v0, float v1, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  p0 = cvt_bf16x2(v0, v1);
  float r0 = v0 - __uint_as_float(p0 << 16), r1 = v1 - __uint_as_float(p0 & 0xffff0000u);
  p1 = cvt_bf16x2(r0, r1);
  r0 -= __uint_as_float(p1 << 16);
  r1 -= __uint_as_float(p1 & 0xffff0000u);
  p2 = cvt_bf16x2(r0, r1);
}
__device__ __forceinline__ void split3x4(float v0, float v1, float v2, float v3, uint2& p0, uint2& p1, uint2& p2) {
  split3x2(v0, v1, p0.x, p1.x, p2.x);
  split3x2(v2, v3, p0.y, p1.y, p2.y);
}

}  // namespace acx
