// peer.cu - the data-parallel exchange as ONE kernel over NVLink peer memory: sum of the ranks' buckets
// [factor statistics | gradients | loss scalars] fused with the 1 / world_size scaling, launched from inside phase 2 (and its
// CUDA graph), instead of an eager NCCL all-reduce between the two graphs of an update.
//
// Two-shot all-reduce on peer pointers (every rank maps every other rank's arena with CUDA IPC):
//   stage 0  CTA b of rank r tells CTA b of every peer that it has started - the kernel is stream-ordered behind phase 1, so a
//            started CTA means "this rank's bucket is complete" - and waits for the same from all peers;
//   reduce   rank r sums slice r of all ranks' buckets in rank order (the same order on every rank: parameters never diverge),
//            scales the gradient part and writes the result into its own bucket;
//   stage 1  "slice r sub-range b is final";
//   gather   rank r copies slice k, sub-range b from rank k for every k != r;
//   stage 2  "done reading": nobody returns while a peer may still read its bucket.
// CTA b of every rank works on sub-range b of every slice, so all three barriers are between the CTAs with the same index
// on the different GPUs: flags[src rank][b] in each rank's arena, a monotone barrier sequence number, system-scope
// release / acquire.  The sequence number of CTA index b lives in device memory and is advanced by the kernel itself, so a
// captured graph replays correctly.  Two ranks take a shorter route: both read both buckets into registers, meet, then write
// (one pass over NVLink).  A wait that lasts longer than ~2 s records an error and gives up (no hang).
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "layers.cuh"

namespace acx {

constexpr int PEER_MAX_RANKS = 8;
constexpr int PEER_CTAS = 148;
constexpr int PEER_THREADS = 512;

struct PeerArgs {
  float* bucket[PEER_MAX_RANKS];        // the region to reduce in every rank's arena (own entry = local memory)
  unsigned int* flags[PEER_MAX_RANKS];  // every rank's flag array [PEER_MAX_RANKS][PEER_CTAS]
  unsigned int* epochs;                 // local: [PEER_CTAS]
  int rank, world;
  long long count;                      // floats in the region (multiple of 4)
  long long scale_from;                 // elements at or beyond this offset are multiplied by `scale` (multiple of 4)
  float scale;
  int one_shot;                         // two ranks: read both buckets, meet, write (ACX_PEER_TWO_SHOT=1 forces the general route)
};

static __device__ int g_peer_error = 0;

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float* p) {   // peer memory: never from a stale cache line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// barrier between the CTAs `blockIdx.x` of all ranks; every thread's earlier writes are ordered before the signal
__device__ __forceinline__ void peer_barrier(const PeerArgs& a, unsigned int value) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < (unsigned)a.world) {
    const int k = threadIdx.x;
    st_release_sys(a.flags[k] + a.rank * PEER_CTAS + blockIdx.x, value);          // my flag in rank k's array
    const unsigned int* mine = a.flags[a.rank] + k * PEER_CTAS + blockIdx.x;       // rank k's flag in my array
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - value) < 0) {
      if (clock64() - t0 > 4000000000ll) {
        atomicExch(&g_peer_error, 1);
        break;
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void add4(float4& a, const float4& b) {
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
  a.w += b.w;
}
__device__ __forceinline__ void scale4(float4& a, float s) {
  a.x *= s;
  a.y *= s;
  a.z *= s;
  a.w *= s;
}

__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_kernel(const PeerArgs a) {
  unsigned int seq = a.epochs[blockIdx.x];   // barrier sequence number of this CTA index: the same on every rank
  const int W = a.world;
  const long long quads = a.count >> 2;
  const long long sq = a.scale_from >> 2;
  float* const out = a.bucket[a.rank];
  peer_barrier(a, ++seq);                    // every rank's bucket is complete
  if (a.one_shot) {
    // one shot: both ranks read both buckets (rank 0's value + rank 1's value on either side) into registers, meet, and only
    // then overwrite their own copy - one pass over NVLink instead of reduce-scatter + all-gather
    constexpr int R = 8;
    const long long per = (quads + gridDim.x - 1) / gridDim.x;
    const long long c0 = (long long)blockIdx.x * per, c1 = min(quads, c0 + per);
    for (long long base = c0; base < c1; base += (long long)blockDim.x * R) {
      float4 x[R], y[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const long long q = base + (long long)j * blockDim.x + threadIdx.x;
        if (q < c1) x[j] = ld_peer(a.bucket[0] + (q << 2));
      }
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const long long q = base + (long long)j * blockDim.x + threadIdx.x;
        if (q < c1) y[j] = ld_peer(a.bucket[1] + (q << 2));
      }
      peer_barrier(a, ++seq);                // the peer has read this range of my bucket
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const long long q = base + (long long)j * blockDim.x + threadIdx.x;
        if (q < c1) {
          add4(x[j], y[j]);
          if (q >= sq) scale4(x[j], a.scale);
          *reinterpret_cast<float4*>(out + (q << 2)) = x[j];
        }
      }
    }
  } else {
    // two shots: slice of a rank (multiple of 4 floats), sub-range of this CTA index inside every slice
    constexpr int R = 4;
    const long long slice_q = (quads + W - 1) / W;
    const long long sub_q = (slice_q + gridDim.x - 1) / gridDim.x;
    const long long b0 = (long long)blockIdx.x * sub_q, b1 = min(slice_q, b0 + sub_q);
    {
      const long long s0 = (long long)a.rank * slice_q;
      for (long long base = b0; base < b1; base += (long long)blockDim.x * R) {
        float4 acc[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const long long q = base + (long long)j * blockDim.x + threadIdx.x;
          if (q < b1 && s0 + q < quads) acc[j] = ld_peer(a.bucket[0] + ((s0 + q) << 2));
        }
        for (int k = 1; k < W; ++k) {        // rank order: the same association on every rank
          float4 v[R];
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const long long q = base + (long long)j * blockDim.x + threadIdx.x;
            if (q < b1 && s0 + q < quads) v[j] = ld_peer(a.bucket[k] + ((s0 + q) << 2));
          }
#pragma unroll
          for (int j = 0; j < R; ++j) add4(acc[j], v[j]);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const long long q = base + (long long)j * blockDim.x + threadIdx.x;
          if (q < b1 && s0 + q < quads) {
            if (s0 + q >= sq) scale4(acc[j], a.scale);
            *reinterpret_cast<float4*>(out + ((s0 + q) << 2)) = acc[j];
          }
        }
      }
    }
    peer_barrier(a, ++seq);                  // slice r, sub-range b is final on rank r
    for (int kk = 1; kk < W; ++kk) {
      const int k = (a.rank + kk) % W;       // start with a different peer on every rank
      const long long s0 = (long long)k * slice_q;
      for (long long base = b0; base < b1; base += (long long)blockDim.x * R) {
        float4 v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const long long q = base + (long long)j * blockDim.x + threadIdx.x;
          if (q < b1 && s0 + q < quads) v[j] = ld_peer(a.bucket[k] + ((s0 + q) << 2));
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const long long q = base + (long long)j * blockDim.x + threadIdx.x;
          if (q < b1 && s0 + q < quads) *reinterpret_cast<float4*>(out + ((s0 + q) << 2)) = v[j];
        }
      }
    }
    peer_barrier(a, ++seq);                  // done reading: nobody returns while a peer may still read its bucket
  }
  if (threadIdx.x == 0) a.epochs[blockIdx.x] = seq;
}

struct PeerState {   // (declared again in learner.cu)
  int rank = 0, world = 0;
  uint8_t* base[PEER_MAX_RANKS] = {};   // arena base of every rank as mapped here
};

// `channel` 0 / 1: two exchanges may be in flight at once (phase 2's on the caller's stream, the input-factor prefix on a side
// stream); each has its own flags and sequence numbers
int peer_allreduce(const PeerState& ps, size_t region_offset_bytes, size_t flags_offset_bytes, size_t epochs_offset_bytes,
                   long long count, long long scale_from, float scale, int channel, cudaStream_t st) {
  flags_offset_bytes += (size_t)channel * PEER_MAX_RANKS * PEER_CTAS * sizeof(unsigned int);
  epochs_offset_bytes += (size_t)channel * PEER_CTAS * sizeof(unsigned int);
  ACX_CHECK(ps.world >= 2 && ps.world <= PEER_MAX_RANKS, "peer all-reduce: 2..8 ranks");
  ACX_CHECK((count & 3) == 0 && (scale_from & 3) == 0 && (region_offset_bytes & 15) == 0, "peer all-reduce: region must be 16-byte granular");
  PeerArgs a;
  memset(&a, 0, sizeof(a));
  for (int k = 0; k < ps.world; ++k) {
    a.bucket[k] = reinterpret_cast<float*>(ps.base[k] + region_offset_bytes);
    a.flags[k] = reinterpret_cast<unsigned int*>(ps.base[k] + flags_offset_bytes);
  }
  a.epochs = reinterpret_cast<unsigned int*>(ps.base[ps.rank] + epochs_offset_bytes);
  a.rank = ps.rank;
  a.world = ps.world;
  a.count = count;
  a.scale_from = scale_from;
  a.scale = scale;
  const char* e = getenv("ACX_PEER_TWO_SHOT");   // (read per launch: tests switch it between engines)
  a.one_shot = (ps.world == 2 && !(e && atoi(e))) ? 1 : 0;
  // the exchange phase 2 waits for takes every SM; the one that runs under phase 2 on a side stream (channel 1) only a few,
  // so that its spinning CTAs leave the SMs to phase 2.  ACX_PEER_CTAS / ACX_PEER_SIDE_CTAS: tuning knobs (the same on every rank)
  int ctas = channel == 0 ? PEER_CTAS : 24;
  if (const char* c = getenv(channel == 0 ? "ACX_PEER_CTAS" : "ACX_PEER_SIDE_CTAS")) {
    const int v = atoi(c);
    if (v >= 1 && v <= PEER_CTAS) ctas = v;
  }
  peer_allreduce_kernel<<<ctas, PEER_THREADS, 0, st>>>(a);
  ACX_LAUNCH_CHECK();
  return 0;
}

size_t peer_flag_bytes() { return (size_t)2 * PEER_MAX_RANKS * PEER_CTAS * sizeof(unsigned int); }
size_t peer_epoch_bytes() { return (size_t)2 * PEER_CTAS * sizeof(unsigned int); }

int peer_error_flag() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_peer_error, sizeof(int));
  return v;
}

}  // namespace acx

extern "C" {

// CUDA IPC plumbing for the peer exchange: export the allocation that contains `d_ptr` (64-byte handle + the offset of d_ptr
// inside it); import maps a peer's allocation into this process and returns the address that corresponds to the peer's d_ptr
int acx_peer_export(const void* d_ptr, unsigned char* out_handle64, unsigned long long* out_offset) {
  ACX_CHECK(d_ptr && out_handle64 && out_offset, "null argument");
  CUdeviceptr base = 0;
  size_t size = 0;
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  ACX_CHECK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && f,
            "cuMemGetAddressRange not available");
  ACX_CHECK(reinterpret_cast<RangeFn>(f)(&base, &size, reinterpret_cast<CUdeviceptr>(d_ptr)) == CUDA_SUCCESS, "cuMemGetAddressRange failed");
  cudaIpcMemHandle_t h;
  ACX_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out_handle64, &h, 64);
  *out_offset = (unsigned long long)(reinterpret_cast<CUdeviceptr>(d_ptr) - base);
  return 0;
}

void* acx_peer_import(const unsigned char* handle64, unsigned long long offset) {
  if (!handle64) return nullptr;
  // an allocation can be opened once per process: later imports of the same handle (a second engine whose arena the caching
  // allocator placed in the same segment) reuse the mapping
  static std::mutex mu;
  static std::map<std::string, void*> opened;
  std::lock_guard<std::mutex> lock(mu);
  const std::string key(reinterpret_cast<const char*>(handle64), 64);
  auto it = opened.find(key);
  void* base = it != opened.end() ? it->second : nullptr;
  if (!base) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    const cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      acx::set_error(std::string("acx_peer_import: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
      return nullptr;
    }
    opened[key] = base;
  }
  return static_cast<unsigned char*>(base) + offset;
}

int acx_peer_error(void) { return acx::peer_error_flag(); }

}  // extern "C"
