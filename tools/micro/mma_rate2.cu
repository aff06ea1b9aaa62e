// mma_rate2.cu - what slows tcgen05.mma below its stand-alone rate inside a warp-specialised kernel?  N = 64, M = 128, 6 "plane
// pairs" x 4 k-steps = 24 MMAs per k-block, 200 k-blocks.  Variants:
//   0 back to back, one commit at the end          1 tcgen05.commit to a barrier after every k-block (nobody waits)
//   2 as 1 + four other warps stream tcgen05.ld from the other half of TMEM and stage to shared memory (epilogue traffic)
//   3 as 1 + the issuing warp polls an (already complete) mbarrier and runs fence::after_thread_sync before every k-block
//   4 as 3 + a producer warp: 2-stage ring with real empty/full hand-offs but no data movement (expect_tx 0)
//   5 as 4 with a 4-stage ring
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"
using namespace acx;

__global__ void __launch_bounds__(192, 1) k(int variant, int kblocks, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done_bar, full_bar[4], empty_bar[4], dummy_bar, commit_bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = 64;
  const int stages = variant == 5 ? 4 : 2;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    mbar_init(&dummy_bar, 1);
    mbar_init(&commit_bar, 1);
    for (int s = 0; s < 4; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    if (variant >= 4) {   // producer: hand stages back as soon as they are free
      for (int it = 0; it < kblocks; ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u, 1);
        __syncwarp();
        if (elect_one()) mbar_expect_tx(&full_bar[s], 0u);
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t base = smem_u32(smem);
    long long t0 = clock64();
    for (int it = 0; it < kblocks; ++it) {
      const int s = it % stages;
      const uint32_t ph = (uint32_t)(it / stages) & 1u;
      if (variant >= 4) {
        mbar_wait(&full_bar[s], ph, 2);
        tc_fence_after();
      } else if (variant == 3) {
        mbar_wait(&dummy_bar, 1u, 2);   // phase 0 not yet complete -> waiting on parity 1 returns at once
        tc_fence_after();
      }
      if (elect_one()) {
        for (int pr = 0; pr < 6; ++pr) {
          const int pa = pr % 3, pb = pr / 2;
          uint64_t ad = make_smem_desc_sw(base + s * 73728 + pa * 16384, 16u, 1024u, 2u);
          uint64_t bd = make_smem_desc_sw(base + s * 73728 + 49152 + pb * (n * 128), 16u, 1024u, 2u);
          for (int kk = 0; kk < 4; ++kk) {
            umma_bf16(tmem, ad, bd, idesc, 1u);
            ad += 2;
            bd += 2;
          }
        }
        if (variant >= 4) umma_commit(&empty_bar[s]);
        else if (variant >= 1) umma_commit(&commit_bar);   // arrivals nobody waits for
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
    mbar_wait(&done_bar, 0u, 7);
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; stop = 1; }
  } else if (variant == 2) {
    float* st = reinterpret_cast<float*>(smem + 150 * 1024) + (warp - 2) * 32 * 36;
    float acc = 0.f;
    while (!stop) {
      uint32_t raw[32];
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u, raw);
      for (int j = 0; j < 32; ++j) st[lane * 36 + j] = __uint_as_float(raw[j]);
      __syncwarp();
      for (int j = 0; j < 32; ++j) acc += st[j * 36 + lane];
      __syncwarp();
    }
    if (acc == 12345.f) out[1] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int kblocks = 200;
  for (int variant = 0; variant <= 5; ++variant) {
    long long h = 0;
    for (int it = 0; it < 2; ++it) {
      k<<<1, 192, 200 * 1024>>>(variant, kblocks, d);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    fflush(stdout);
    printf("variant %d: %8.1f cycles per k-block of 24 MMAs (128x64x16) = %6.1f per MMA %s\n", variant, (double)h / kblocks,
           (double)h / kblocks / 24, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  fflush(stdout);
  return 0;
}
