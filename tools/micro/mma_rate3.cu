// mma_rate3.cu - fixed cost per k-block of the warp-specialised pipeline: M = 128, N and the number of plane pairs as arguments
// (4 MMAs per pair and k-block, issued as one asm block like the kernels do), producer warp with real empty / full hand-offs but
// no data movement, ring counters without divisions.  Prints cycles per k-block against N-dependent pipe time.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I actorcritic_b200/csrc -I include tools/micro/mma_rate3.cu -o tools/micro/mma_rate3
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"
using namespace acx;
namespace acx { void set_error(const std::string&) {} uint64_t g_launch_count = 0; int pdl_level() { return 0; } }

__global__ void __launch_bounds__(192, 1) k(int n, int pairs, int stages, int kblocks, long long* out, int flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done_bar, full_bar[8], empty_bar[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    for (int s = 0; s < 8; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < kblocks; ++it) {
      mbar_wait(&empty_bar[s], ph ^ 1u, 1);
      __syncwarp();
      if (elect_one()) mbar_expect_tx(&full_bar[s], 0u);
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a0 = make_smem_desc_sw(smem_u32(smem), 16u, 1024u, 2u);
    const uint64_t b0 = make_smem_desc_sw(smem_u32(smem) + 98304u, 16u, 1024u, 2u);
    int s = 0; uint32_t ph = 0;
    long long t0 = clock64();
    for (int it = 0; it < kblocks; ++it) {
      mbar_wait(&full_bar[s], ph, 2);
      if (!(flags & 1)) tc_fence_after();          // flag 1: no tcgen05.fence::after_thread_sync per k-block
      if (elect_one()) {
        if (pairs == 3) umma_bf16_x4_pairs3(tmem, a0, b0, 2u, 2u, idesc, it ? 1u : 0u, 0u, 0u, 0u, 1024u, 1024u, 0u);
        else umma_bf16_x4_pairs2(tmem, a0, b0, 2u, 2u, idesc, it ? 1u : 0u, 0u, 0u, 0u, 1024u);
        umma_commit(&empty_bar[s]);
      }
      if (!(flags & 2)) __syncwarp();             // flag 2: no __syncwarp per k-block
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
    mbar_wait(&done_bar, 0u, 7);
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  const int flags = argc > 1 ? atoi(argv[1]) : 0;
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int kblocks = 400;
  for (int n : {32, 64, 128})
    for (int pairs : {2, 3})
      for (int stages : {4}) {
        long long h = 0;
        for (int it = 0; it < 2; ++it) { k<<<1, 192, 200 * 1024>>>(n, pairs, stages, kblocks, d, flags); cudaDeviceSynchronize(); }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const int pipe = pairs * 4 * (n / 2 > 32 + n / 4 ? n / 2 : 32 + n / 4);
        printf("flags=%d N=%3d pairs=%d stages=%d: %6.0f cycles per k-block (tensor pipe alone: %d)\n", flags, n, pairs, stages, (double)h / kblocks, pipe);
      }
  return 0;
}
