// mma_rate.cu - micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16 x bf16 -> fp32, M = 128) issued back to
// back from shared memory operands, as a function of N, of the swizzle mode and of how the descriptors walk the tiles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I actorcritic_b200/csrc -o mma_rate tools/micro/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"
using namespace acx;

__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int reps, int layout, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sbo = layout == 2 ? 1024u : 512u;
    const uint32_t base = smem_u32(smem);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      // mode 0: same A / B k-slice every time; mode 1: walk 4 k-steps x 6 "plane pairs" like the GEMM (3 A planes, 3 B planes)
      for (int r = 0; r < reps; ++r) {
        if (mode == 0) {
          const uint64_t ad = make_smem_desc_sw(base, 16u, sbo, (uint32_t)layout);
          const uint64_t bd = make_smem_desc_sw(base + 64 * 1024, 16u, sbo, (uint32_t)layout);
          for (int i = 0; i < 24; ++i) umma_bf16(tmem, ad, bd, idesc, 1u);
        } else {
          for (int pr = 0; pr < 6; ++pr) {
            const int pa = pr % 3, pb = pr / 2;
            uint64_t ad = make_smem_desc_sw(base + pa * 16384, 16u, sbo, (uint32_t)layout);
            uint64_t bd = make_smem_desc_sw(base + 64 * 1024 + pb * (n * 128), 16u, sbo, (uint32_t)layout);
            for (int kk = 0; kk < 4; ++kk) {
              umma_bf16(tmem, ad, bd, idesc, 1u);
              ad += 2;
              bd += 2;
            }
          }
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0u, 7);
    t1 = clock64();
    if (elect_one()) { out[0] = t1 - t0; }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 200;
  for (int layout : {2, 4})
    for (int mode : {0, 1})
      for (int n : {32, 64, 96, 128, 192, 256}) {
        long long h = 0;
        for (int it = 0; it < 2; ++it) {
          rate_kernel<<<1, 128, 200 * 1024>>>(n, reps, layout, mode, d);
          cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("swizzle %3dB mode %d N %3d: %7.1f cycles per MMA (128 x N x 16)  -> %6.0f MAC/clk  %s\n", layout == 2 ? 128 : 64, mode,
               n, (double)h / (reps * 24), 128.0 * n * 16 / ((double)h / (reps * 24)), e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  // all SMs at once (shared-memory operand reads are per SM; this shows whether power/clock changes the picture)
  for (int n : {64, 128, 256}) {
    long long h = 0;
    rate_kernel<<<148, 128, 200 * 1024>>>(n, reps, 2, 1, d);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("148 CTAs, swizzle 128B mode 1 N %3d: %7.1f cycles per MMA\n", n, (double)h / (reps * 24));
  }
  return 0;
}
