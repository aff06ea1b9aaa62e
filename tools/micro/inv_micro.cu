// micro-benchmark of the building blocks of kfac_inv.cu (cycles per call, one CTA per SM, no grid barriers)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. tools/micro/inv_micro.cu -o tools/micro/inv_micro
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>
namespace acx { void set_error(const std::string&) {} uint64_t g_launch_count = 0; }
#include "../../actorcritic_b200/csrc/kfac_inv.cu"
using namespace acx;

__global__ void __launch_bounds__(256, 1) micro_kernel(ResJob jb, long long* out, int reps) {
  extern __shared__ __align__(16) float smem[];
  float* stage = smem;
  float* As = stage;
  float* Bs = stage + RB * OPLD;
  float* Cs = stage + 2 * RB * OPLD;
  float (*Ds)[RB + 1] = reinterpret_cast<float (*)[RB + 1]>(stage + 3 * RB * OPLD);
  float* tiles = smem + STAGE_FLOATS;
  for (int e = threadIdx.x; e < 3 * TILE_FLOATS; e += 256) tiles[e] = 0.001f * (e % 97);
  for (int e = threadIdx.x; e < RB * RB; e += 256) Ds[e >> 5][e & 31] = (e >> 5) == (e & 31) ? 4.0f : 0.01f;
  __syncthreads();
  long long t[8];
  t[0] = clock64();
  for (int r = 0; r < reps; ++r) invert32(Ds, tiles + 2 * TILE_FLOATS);
  __syncthreads();
  t[1] = clock64();
  const int p = 20;   // pivot block 20 -> tile 10
  for (int r = 0; r < reps; ++r) panel_piece(jb, p, 10, 12 + (blockIdx.x % 8), tiles, As, Bs, Ds);
  __syncthreads();
  t[2] = clock64();
  for (int r = 0; r < reps; ++r) panel_piece(jb, p, 3 + (blockIdx.x % 4), 10, tiles, As, Bs, Ds);   // mirror
  __syncthreads();
  t[3] = clock64();
  long long u[6];
  u[0] = clock64();
  for (int r = 0; r < reps; ++r) panel_piece<1>(jb, p, 10, 12 + (blockIdx.x % 8), tiles, As, Bs, Ds);
  __syncthreads();
  u[1] = clock64();
  for (int r = 0; r < reps; ++r) panel_piece<2>(jb, p, 10, 12 + (blockIdx.x % 8), tiles, As, Bs, Ds);
  __syncthreads();
  u[2] = clock64();
  for (int r = 0; r < reps; ++r) panel_piece<4>(jb, p, 10, 12 + (blockIdx.x % 8), tiles, As, Bs, Ds);
  __syncthreads();
  u[3] = clock64();
  for (int r = 0; r < reps; ++r) panel_piece<8>(jb, p, 10, 12 + (blockIdx.x % 8), tiles, As, Bs, Ds);
  __syncthreads();
  u[4] = clock64();
  for (int r = 0; r < reps; ++r) panel_piece<15>(jb, p, 10, 12 + (blockIdx.x % 8), tiles, As, Bs, Ds);
  __syncthreads();
  u[5] = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("panel_piece variants: noload %lld nostore %lld noproduct %lld noextract %lld nothing %lld\n", (u[1] - u[0]) / reps,
           (u[2] - u[1]) / reps, (u[3] - u[2]) / reps, (u[4] - u[3]) / reps, (u[5] - u[4]) / reps);
  OpRegs regs;
  for (int r = 0; r < reps; ++r) {
    fetch_ops(jb, p, 2, 5 + (blockIdx.x % 16), regs);
    for (int s = 0; s < 3; ++s) {
      __syncthreads();
      stage_ops(regs, As, Bs, Cs, false);
      __syncthreads();
      if (s < 2) fetch_ops(jb, p, 2 + s, 6 + s + (blockIdx.x % 16), regs);
      update_tile(jb, p, 2 + s, 5 + s, tiles + s * TILE_FLOATS, As, Bs, Cs);
      __syncthreads();
      publish_pair(jb, p + 1, 2 + s, 5 + s, tiles + s * TILE_FLOATS);
    }
  }
  __syncthreads();
  t[4] = clock64();
  for (int r = 0; r < reps; ++r) {   // compute only
    update_tile(jb, p, 2, 5, tiles, As, Bs, Cs);
    __syncthreads();
  }
  t[5] = clock64();
  for (int r = 0; r < reps; ++r) {   // fetch + stage only (exposed latency)
    fetch_ops(jb, p, 2, 5 + (blockIdx.x % 16), regs);
    __syncthreads();
    stage_ops(regs, As, Bs, Cs, false);
    __syncthreads();
  }
  t[6] = clock64();
  if (threadIdx.x == 0)
    for (int i = 0; i < 6; ++i) out[blockIdx.x * 8 + i] = (t[i + 1] - t[i]) / reps;
}

int main() {
  const int n = 1569, ldp = 1572;
  ResJob jb = {};
  jb.n = n;
  jb.ldp = ldp;
  jb.nt = 25;
  float* x;
  cudaMalloc(&x, (64 * ldp + 6144) * sizeof(float));
  cudaMemset(x, 0, (64 * ldp + 6144) * sizeof(float));
  jb.x = x;
  long long* out;
  cudaMalloc(&out, 148 * 8 * sizeof(long long));
  const int smem = (STAGE_FLOATS + 3 * TILE_FLOATS) * 4;
  cudaFuncSetAttribute(micro_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148}) {
    micro_kernel<<<grid, 256, smem>>>(jb, out, 8);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(grid * 8);
    cudaMemcpy(h.data(), out, grid * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    const char* names[6] = {"invert32", "panel_piece(row)", "panel_piece(mirror)", "3-tile update loop", "update_tile compute", "fetch+stage"};
    printf("grid %d (%s)\n", grid, cudaGetErrorString(e));
    for (int i = 0; i < 6; ++i) {
      long long mx = 0, sum = 0;
      for (int b = 0; b < grid; ++b) { mx = h[b * 8 + i] > mx ? h[b * 8 + i] : mx; sum += h[b * 8 + i]; }
      printf("  %-22s mean %lld max %lld cycles\n", names[i], sum / grid, mx);
    }
  }
  return 0;
}
