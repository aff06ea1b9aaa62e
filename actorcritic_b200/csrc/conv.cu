// conv.cu - implicit-GEMM convolution kernels for sm_100a: the patch matrix is never materialised.
//
// The im2col GEMM of nn.conv2d (nn.py:88-110; NHWC, VALID, HWIO weights) reads, for one kernel tap, a BOX of the
// activation tensor: with the kernel size a multiple of the stride (k = s*m: 4x4/2 and 3x3/1 of the Nature-CNN,
// envs/atari/model.py:180-199) the input rows/columns split as y = s*yq + y1, x = s*xq + x1 and the tensor is the
// 5-D array  [r][yq][y1][xq][(x1, c)]  whose innermost run (x1, c) has s*c_in = 64 bf16 = one 128-byte swizzle row.
// A TMA tile load of the box [hw_out yq][hw_out xq][64] at (yq, y1, xq) = (kh / s, kh % s, kw / s) delivers exactly
// the K-major A tile "all output locations of a sample x 64 patch columns" of tap (kh, kw / s), already swizzled for
// tcgen05.mma - forward needs no im2col pass and reads the activation (L2 resident) instead of the k^2/s^2 larger P.
//
// The input gradient uses the gather form of the transposed convolution (no fp32 dP matrix, no col2im pass):
//   dX[r, s*a + py, s*b + px, ci] = sum_{i,j<m} sum_co g[r, a - i, b - j, co] * W[s*i + py, s*j + px, ci, co]
// i.e. ONE GEMM with rows (r, a, b), columns (py, px, ci) and K = (i, j, co); the A tile of tap (i, j) is the box
// [hq a][hq b][c_out] of g at (-i, -j): out-of-range coordinates are zero-filled by TMA, which is the padding.  The
// epilogue scatters (pixel shuffle), applies the ReLU mask of the layer below and splits into bf16 planes.
//
// Measured on B200 at 2N = 1280 samples (ACX_CONV_DEBUG triage, tools/conv_one.py): both dgrad launches are bound by the
// tensor pipe's shared-memory operand reads - a 128 x 64 x 16 MMA costs ~86 cycles from 128-byte-swizzled tiles and ~127
// from the 64-byte-swizzled 32-channel taps of conv3 (floor 32) - not by TMA (skipping every load changes nothing),
// not by accumulator dependencies (alternating two accumulators changes nothing), and keeping the whole conv3 weight
// operand resident in shared memory (tried) gains < 2 %.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM + MMA issuer, warps 2..9 epilogue (two per TMEM lane quarter).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "layers.cuh"
#include "tc.cuh"

namespace acx {

constexpr int CV_BM = 128;
constexpr int CV_BK = 64;
constexpr int CV_A_TILE = CV_BM * CV_BK * 2;   // one plane of one k-block (1 x 64-wide or 2 x 32-wide sub-tiles)
constexpr int CV_MAX_SUB = 16;
constexpr int CV_MAX_STAGES = 4;
constexpr int CV_THREADS = 320;                 // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
constexpr int CV_EPI_BYTES = 8 * 32 * 32 * 4;   // one swizzled 32 x 32 fp32 staging tile per epilogue warp
constexpr int CV_SMEM_LIMIT = 232448;
constexpr int CV_SMEM_FIXED = CV_EPI_BYTES + 256;

struct ConvTcParams {
  int num_tiles, ts, rows_valid, samples;      // a tile = ts samples x (by x bx) cells of the location grid = rows_valid GEMM rows (<= 128)
  int gx, gy, bx, by, nx, cells, ydim;         // location grid gx x gy, cut into nx x (gy / by) boxes of bx x by cells (cells = boxes per
                                               // sample group); ydim = the box coordinate (2 or 3) that carries the grid row
  int bn;                                      // GEMM columns (= tile width: 32, 64 or 128)
  int kb_total, nsub, num_sub;                 // k-blocks of 64; sub-tiles per k-block (1 or 2); sub-tiles over all of K
  int sub_bytes;                               // bytes one sub-tile box delivers
  signed char tc1[CV_MAX_SUB], tc2[CV_MAX_SUB], tc3[CV_MAX_SUB];   // box coordinates of every sub-tile (dims 1..3)
  int num_pairs, pair_a[6], pair_b[6], npa, npb, stages;
  int lo_from_tile, num_pairs_lo;              // tiles >= lo_from_tile accumulate only the first num_pairs_lo plane pairs (tiles are
                                               // ordered sample group first)
  // epilogue
  int dgrad;                                   // 0: output row = GEMM row; 1: pixel-shuffle scatter
  int hw_in, c_in, s, hq;
  float alpha;
  const float* bias;
  int relu;
  bf16* cp[3];
  int npl, ldcp;
  const bf16* mask;
  int mask_samples;
  int b_resident;   // the whole weight operand (kb_total x npb tiles) is loaded once per CTA and stays in shared memory (conv1 forward: 32 KB;
                    // the kernel is bound by L2 -> SM traffic and re-loading the weights for each of its 2100 tiles was a third of it)
  int debug;   // ACX_CONV_DEBUG bits (performance triage only): 1 skip activation loads, 2 skip weight loads, 4 skip MMAs, 8 skip stores
};

// triage (ACX_CONV_DEBUG bit 32): cycles CTA 0's MMA warp spends waiting for operands / for a drained accumulator / issuing,
// and cycles its first epilogue warp spends waiting for an accumulator / working
__device__ long long g_conv_trace[8];

__global__ void __launch_bounds__(CV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap ta0, const __grid_constant__ CUtensorMap ta1,
               const __grid_constant__ CUtensorMap ta2, const __grid_constant__ CUtensorMap tb0,
               const __grid_constant__ CUtensorMap tb1, const __grid_constant__ CUtensorMap tb2, const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) {
    if (threadIdx.x == 0) atomicExch(&g_tc_error, 9);
    return;
  }
  const int BN = p.bn;
  const int b_tile_bytes = BN * CV_BK * 2;
  const int res_bytes = p.b_resident ? p.kb_total * p.npb * b_tile_bytes : 0;   // resident weight tiles come first
  uint8_t* const res_b = smem;
  smem += res_bytes;
  const int stage_bytes = p.npa * CV_A_TILE + (p.b_resident ? 0 : p.npb * b_tile_bytes);
  float* epi = reinterpret_cast<float*>(smem + p.stages * stage_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(epi) + CV_EPI_BYTES);
  uint64_t* empty_bar = full_bar + CV_MAX_STAGES;
  uint64_t* acc_full = empty_bar + CV_MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint64_t* b_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);   // resident weights have landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (uint32_t)(2 * BN);   // two accumulators: 64, 128 or 256 columns

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], BN > 32 ? 8 : 4);   // epilogue warps that drain an accumulator
    }
    mbar_init(b_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int sub_tile_bytes = CV_A_TILE / p.nsub;
  pdl_wait();   // the set-up above overlapped the predecessor's tail; its results are visible from here on

  if (warp == 0) {
    {
      // ===== TMA producer: per k-block one box per (plane, sub-tile) of the activation + the weight tile =====
      if (p.b_resident && elect_one()) {   // the whole weight operand, once
        mbar_expect_tx(b_bar, (uint32_t)res_bytes);
        for (int kb = 0; kb < p.kb_total; ++kb)
          for (int i = 0; i < p.npb; ++i) {
            const CUtensorMap* mb = i == 0 ? &tb0 : (i == 1 ? &tb1 : &tb2);
            tma_load_2d(res_b + (kb * p.npb + i) * b_tile_bytes, mb, b_bar, kb * CV_BK, 0);
          }
      }
      __syncwarp();
      int it = 0;
      int ring_s = 0;
      uint32_t ring_ph = 0u;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int grp = tile / p.cells, cell = tile - grp * p.cells;
        const int cy = cell / p.nx, cx = cell - cy * p.nx;
        const int sample0 = grp * p.ts, x0 = cx * p.bx, y0 = cy * p.by;
        const int y2 = p.ydim == 2 ? y0 : 0, y3 = p.ydim == 3 ? y0 : 0;
        for (int kb = 0; kb < p.kb_total; ++kb, ++it) {
          const int s = ring_s;   // ring position and phase advance without a division (this thread is the bottleneck)
          const uint32_t ph = ring_ph;
          if (++ring_s == p.stages) {
            ring_s = 0;
            ring_ph ^= 1u;
          }
          const int nload = min(p.nsub, p.num_sub - kb * p.nsub);
          mbar_wait(&empty_bar[s], ph ^ 1u, 1);
          __syncwarp();
          if (!elect_one()) continue;
          mbar_expect_tx(&full_bar[s], (uint32_t)(((p.debug & 1) ? 0 : p.npa * nload * p.sub_bytes) +
                                                  (((p.debug & 2) || p.b_resident) ? 0 : p.npb * b_tile_bytes)));
          uint8_t* a_s = smem + s * stage_bytes;
          uint8_t* b_s = a_s + p.npa * CV_A_TILE;
          for (int i = 0; i < p.npa && !(p.debug & 1); ++i) {
            const CUtensorMap* ma = i == 0 ? &ta0 : (i == 1 ? &ta1 : &ta2);
            for (int j = 0; j < nload; ++j) {
              const int t = kb * p.nsub + j;
              tma_load_5d(a_s + i * CV_A_TILE + j * sub_tile_bytes, ma, &full_bar[s], 0, p.tc1[t] + x0, p.tc2[t] + y2, p.tc3[t] + y3, sample0);
            }
          }
          for (int i = 0; i < p.npb && !(p.debug & 2) && !p.b_resident; ++i) {
            const CUtensorMap* mb = i == 0 ? &tb0 : (i == 1 ? &tb1 : &tb2);
            tma_load_2d(b_s + i * b_tile_bytes, mb, &full_bar[s], kb * CV_BK, 0);
          }
        }
      }
      __syncwarp();
      if (elect_one()) pdl_trigger();   // all loads of this CTA's last tile are issued: the successor's blocks may be scheduled
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(CV_BM >> 4) << 24);
    const uint32_t a_layout = p.nsub == 1 ? 2u : 4u;      // SWIZZLE_128B / SWIZZLE_64B
    const uint32_t a_sbo = p.nsub == 1 ? 1024u : 512u;    // 8 rows of 128 / 64 bytes
    const int ksteps = (CV_BK / p.nsub) >> 4;             // MMAs (K = 16) per sub-tile
    uint32_t a_off[6], b_off[6], a_step[4];               // descriptor offsets in 16-byte units
#pragma unroll
    for (int pr = 0; pr < 6; ++pr) {
      a_off[pr] = (uint32_t)(p.pair_a[pr] * CV_A_TILE) >> 4;
      b_off[pr] = (uint32_t)(p.pair_b[pr] * b_tile_bytes) >> 4;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) a_step[t] = (uint32_t)((t / ksteps) * sub_tile_bytes + (t % ksteps) * 32) >> 4;
    int it = 0, lt = 0;
    int ring_s = 0;
    uint32_t ring_ph = 0u;
    // descriptors of stage 0 and their step from stage to stage: built once, advanced by an addition (the low 14-bit address
    // field never carries out: every tile lies inside the 256 KB shared window)
    const uint64_t a_desc_s0 = make_smem_desc_sw(smem_u32(smem), 16u, a_sbo, a_layout);
    const uint64_t b_desc_s0 = p.b_resident ? make_smem_desc_sw(smem_u32(res_b), 16u, 1024u, 2u)
                                            : make_smem_desc_sw(smem_u32(smem) + (uint32_t)(p.npa * CV_A_TILE), 16u, 1024u, 2u);
    const uint64_t a_desc_step = (uint64_t)((uint32_t)stage_bytes >> 4);
    uint64_t a_desc_cur = a_desc_s0, b_desc_cur = b_desc_s0;
    const bool trace = (p.debug & 32) && blockIdx.x == 0;
    long long w_full = 0, w_acc = 0, t_begin = trace ? clock64() : 0;
    if (p.b_resident) {
      mbar_wait(b_bar, 0u, 5);
      tc_fence_after();
    }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
      // rows that only feed the K-FAC output factors (the Fisher-sample half of the stacked backward batch) need the
      // precision of those factors, not of the true gradient: fewer plane pairs for their tiles
      const int npairs = tile >= p.lo_from_tile ? p.num_pairs_lo : p.num_pairs;
      long long tw = trace ? clock64() : 0;
      mbar_wait(&acc_empty[buf], aph ^ 1u, 4);
      if (trace) w_acc += clock64() - tw;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
      for (int kb = 0; kb < p.kb_total; ++kb, ++it) {
        const int s = ring_s;   // ring position and phase advance without a division (this thread is the bottleneck)
        const uint32_t ph = ring_ph;
        const uint64_t a_desc0 = a_desc_cur;
        const uint64_t b_desc0 = p.b_resident ? b_desc_s0 + (uint64_t)((uint32_t)(kb * p.npb * b_tile_bytes) >> 4) : b_desc_cur;
        if (++ring_s == p.stages) {
          ring_s = 0;
          ring_ph ^= 1u;
          a_desc_cur = a_desc_s0;
          b_desc_cur = b_desc_s0;
        } else {
          a_desc_cur += a_desc_step;
          b_desc_cur += a_desc_step;
        }
        tw = trace ? clock64() : 0;
        mbar_wait(&full_bar[s], ph, 2);
        if (trace) w_full += clock64() - tw;
        tc_fence_after();
        if (elect_one()) {
          const int nload = min(p.nsub, p.num_sub - kb * p.nsub);
          uint32_t acc_flag = kb > 0 ? 1u : 0u;
          // straight-line issue: a k-block is 4 MMA steps of K = 16 per plane pair (step t: A at a_step[t] inside the plane's
          // tile - the second 32-wide sub-tile starts sub_tile_bytes in -, B 32 bytes further along its 128-byte row); pair
          // offsets are hoisted out of the tile loop.  With run-time loop bounds and per-pair parameter loads the issue loop
          // itself cost 110-150 cycles per MMA (trace, ACX_CONV_DEBUG=32) against 48-64 in the tensor pipe.
          const int nsteps = nload * ksteps;
          if (!(p.debug & 4)) {
            if (p.nsub == 1 && npairs == 3) {   // one 64-wide sub-tile: four steps of 32 bytes on both operands; the whole k-block as one block
              umma_bf16_x4_pairs3(d_tmem, a_desc0, b_desc0, 2u, 2u, idesc, acc_flag, a_off[0], b_off[0], a_off[1], b_off[1], a_off[2], b_off[2]);
            } else if (p.nsub == 1 && npairs == 2) {
              umma_bf16_x4_pairs2(d_tmem, a_desc0, b_desc0, 2u, 2u, idesc, acc_flag, a_off[0], b_off[0], a_off[1], b_off[1]);
            } else if (p.nsub == 1) {
#pragma unroll
              for (int pr = 0; pr < 6; ++pr)
                if (pr < npairs)
                  umma_bf16_x4(d_tmem, a_desc0 + (uint64_t)a_off[pr], b_desc0 + (uint64_t)b_off[pr], 2u, 2u, idesc, pr ? 1u : acc_flag);
            } else if (nsteps == 4) {   // two 32-wide sub-tiles: the four steps of a pair as one block, A offsets per step
#pragma unroll
              for (int pr = 0; pr < 6; ++pr)
                if (pr < npairs)
                  umma_bf16_x4_steps(d_tmem, a_desc0 + (uint64_t)a_off[pr], b_desc0 + (uint64_t)b_off[pr], a_step[1], a_step[2], a_step[3], idesc,
                                     pr ? 1u : acc_flag);
            } else {
#pragma unroll
              for (int pr = 0; pr < 6; ++pr) {
                if (pr < npairs) {
#pragma unroll
                  for (int t = 0; t < 4; ++t)
                    if (t < nsteps)
                      umma_bf16(d_tmem, a_desc0 + (uint64_t)(a_off[pr] + a_step[t]), b_desc0 + (uint64_t)(b_off[pr] + 2u * (uint32_t)t), idesc,
                                (pr | t) ? 1u : acc_flag);
                }
              }
            }
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
    if (trace && lane == 0) {
      g_conv_trace[0] = clock64() - t_begin;
      g_conv_trace[1] = w_full;
      g_conv_trace[2] = w_acc;
      g_conv_trace[3] = lt;
    }
  } else {
    // ===== epilogue (8 warps): TMEM -> registers -> per-warp shared-memory transpose -> bf16 planes =====
    // Warp w may read the TMEM lane quarter w % 4 (= 32 tile rows); the two warps of a quarter split the 32-column chunks
    // (chunk c goes to group c & 1).  After the transpose a lane owns 4 consecutive columns of a row: 8 lanes write 64
    // contiguous bytes per plane.  Which output element a (tile row, column) pair lands on does not depend on the tile,
    // only its sample base does: the per-lane row and column offsets are computed once.
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int nchunks = BN >> 5;
    float* st = epi + (size_t)(warp - 2) * (32 * 32);   // 32 x 32 fp32, 16-byte groups XOR-swizzled by the row (no padding)
    // a lane owns 8 consecutive columns of a row (two 16-byte staging groups): 16-byte stores / mask loads per plane, 4 lanes
    // per 32-column row, 8 rows per iteration (8-byte stores cost conv2's input gradient 15 of its 53 us: ACX_CONV_DEBUG=8)
    constexpr int CH = 32, LPR = 4, RPI = 8, NIT = 4;
    const int rsub = lane >> 2;
    const int cg = lane & 3;              // 8-column group inside the chunk
    bf16* const cp0 = p.cp[0];
    bf16* const cp1 = p.cp[1];
    bf16* const cp2 = p.cp[2];
    const int npl = p.npl;
    const float alpha = p.alpha;
    const bool relu = p.relu != 0;
    const bf16* const mask = p.mask;
    constexpr uint32_t NO_ROW = 0xffffffffu;
    uint32_t row_off[NIT];   // element offset of tile row (q*32 + rsub + t*RPI) relative to the tile's base
    int row_smp[NIT];        // sample of that row inside the tile's sample group
    const int img = p.hw_in * p.hw_in * p.c_in;   // input gradient: elements of one sample of the output tensor
#pragma unroll
    for (int t = 0; t < NIT; ++t) {
      const int i = q * 32 + rsub + t * RPI;
      row_off[t] = NO_ROW;
      row_smp[t] = 0;
      if (i < p.rows_valid && !(p.debug & 8)) {
        const int per = p.bx * p.by;
        const int sl = i / per, rem = i - sl * per;
        const int yy = rem / p.bx, xx = rem - yy * p.bx;
        row_smp[t] = sl;
        if (!p.dgrad) {   // output row = (sample, y, x) of the gx x gy grid
          row_off[t] = (uint32_t)((sl * p.gx * p.gy + yy * p.gx + xx) * p.ldcp);
        } else {          // row (a, b) of the sample -> the s x s output cell at pixel (s*a, s*b)
          row_off[t] = (uint32_t)(sl * img + ((p.s * yy) * p.hw_in + p.s * xx) * p.c_in);
        }
      }
    }
    uint32_t col_off[2];        // element offset of this lane's 8 columns in the warp's first / second chunk
    float4 bias4[2][2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int n = (grp + 2 * c) * CH + cg * 8;
      col_off[c] = (uint32_t)n;
      bias4[c][0] = bias4[c][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < BN) {
        if (p.dgrad) {   // column (py, px, ci) -> pixel (py, px) inside the cell (c_in is a multiple of 8: one pixel)
          const int pp = n / p.c_in, ci = n - pp * p.c_in;
          const int py = pp / p.s, px = pp - py * p.s;
          col_off[c] = (uint32_t)((py * p.hw_in + px) * p.c_in + ci);
        }
        if (p.bias) {
          bias4[c][0] = make_float4(__ldg(p.bias + n), __ldg(p.bias + n + 1), __ldg(p.bias + n + 2), __ldg(p.bias + n + 3));
          bias4[c][1] = make_float4(__ldg(p.bias + n + 4), __ldg(p.bias + n + 5), __ldg(p.bias + n + 6), __ldg(p.bias + n + 7));
        }
      }
    }
    if (grp < nchunks) {   // BN = 32: the second warp of each quarter has nothing to do (and is not counted by acc_empty)
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
        const bool etrace = (p.debug & 32) && blockIdx.x == 0 && warp == 2;
        const long long te = etrace ? clock64() : 0;
        mbar_wait(&acc_full[buf], aph, 3);
        if (etrace && lane == 0) g_conv_trace[5] = (lt == 0 ? 0 : g_conv_trace[5]) + (clock64() - te);
        if (etrace && lane == 0 && lt == 0) g_conv_trace[6] = te;
        if (etrace && lane == 0) g_conv_trace[7] = clock64() - g_conv_trace[6];
        tc_fence_after();
        const int sgrp = tile / p.cells, cell = tile - sgrp * p.cells;
        const int cy = cell / p.nx, cx = cell - cy * p.nx;
        const int sample0 = sgrp * p.ts, x0 = cx * p.bx, y0 = cy * p.by;
        const size_t base = p.dgrad ? (size_t)sample0 * img + (size_t)(((p.s * y0) * p.hw_in + p.s * x0) * p.c_in)
                                    : ((size_t)sample0 * p.gx * p.gy + (size_t)(y0 * p.gx + x0)) * p.ldcp;
        // rows of samples beyond the batch do not exist; the ReLU mask of sample r is that of sample r % mask_samples
        const int smp_limit = p.samples - sample0;
        bool row_ok[NIT];
        size_t mrow[NIT];
#pragma unroll
        for (int t = 0; t < NIT; ++t) {
          row_ok[t] = row_off[t] != NO_ROW && row_smp[t] < smp_limit;
          int sm = sample0 + row_smp[t], wrapped = 0;   // sample r uses the mask of sample r % mask_samples (r < 2 mask_samples as a rule)
          while (mask != nullptr && sm >= p.mask_samples) {
            sm -= p.mask_samples;
            wrapped += p.mask_samples;
          }
          mrow[t] = base + row_off[t] - (size_t)wrapped * img;
        }
        for (int ch = grp, c = 0; ch < nchunks; ch += 2, ++c) {
          uint32_t raw[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + ch * CH);
          tmem_ld32(taddr, raw);
          {
            float4* strow = reinterpret_cast<float4*>(st + lane * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              strow[j ^ (lane & 7)] = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                                  __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
          }
          if (ch + 2 >= nchunks) {   // this warp's share of the accumulator is drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
          }
          __syncwarp();
          const uint32_t coff = c == 0 ? col_off[0] : col_off[1];
          const float4 b4a = c == 0 ? bias4[0][0] : bias4[1][0];
          const float4 b4b = c == 0 ? bias4[0][1] : bias4[1][1];
          // pass 1: every ReLU-mask word of this chunk is requested before the first one is used (one L2 round trip per
          // chunk instead of one per row)
          uint4 mw[NIT];
          if (mask) {
#pragma unroll
            for (int t = 0; t < NIT; ++t) {
              mw[t] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);   // bf16 1.0: keep
              if (row_ok[t]) mw[t] = __ldg(reinterpret_cast<const uint4*>(mask + mrow[t] + coff));
            }
          }
#pragma unroll
          for (int t = 0; t < NIT; ++t) {
            if (row_ok[t]) {
              const int r = rsub + t * RPI;
              const float4 a4 = *reinterpret_cast<const float4*>(st + r * 32 + (((2 * cg) ^ (r & 7)) << 2));
              const float4 c4 = *reinterpret_cast<const float4*>(st + r * 32 + (((2 * cg + 1) ^ (r & 7)) << 2));
              float v[8] = {fmaf(alpha, a4.x, b4a.x), fmaf(alpha, a4.y, b4a.y), fmaf(alpha, a4.z, b4a.z), fmaf(alpha, a4.w, b4a.w),
                            fmaf(alpha, c4.x, b4b.x), fmaf(alpha, c4.y, b4b.y), fmaf(alpha, c4.z, b4b.z), fmaf(alpha, c4.w, b4b.w)};
              if (relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
              }
              if (mask) {
                const uint32_t mwords[4] = {mw[t].x, mw[t].y, mw[t].z, mw[t].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (!(__uint_as_float(mwords[j] << 16) > 0.0f)) v[2 * j] = 0.0f;
                  if (!(__uint_as_float(mwords[j] & 0xffff0000u) > 0.0f)) v[2 * j + 1] = 0.0f;
                }
              }
              uint2 ph0, pm0, pl0, ph1, pm1, pl1;
              split3x4(v[0], v[1], v[2], v[3], ph0, pm0, pl0);
              split3x4(v[4], v[5], v[6], v[7], ph1, pm1, pl1);
              const size_t idx = base + row_off[t] + coff;
              *reinterpret_cast<uint4*>(cp0 + idx) = make_uint4(ph0.x, ph0.y, ph1.x, ph1.y);
              if (npl > 1) *reinterpret_cast<uint4*>(cp1 + idx) = make_uint4(pm0.x, pm0.y, pm1.x, pm1.y);
              if (npl > 2) *reinterpret_cast<uint4*>(cp2 + idx) = make_uint4(pl0.x, pl0.y, pl1.x, pl1.y);
            }
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// B operand of the gather-form input gradient: rows (py, px, ci), K columns (i, j, co):
//   Bd[(py*s + px)*c_in + ci][(i*m + j)*c_out + co] = W[((s*i + py)*k + (s*j + px))*c_in + ci][co]      (k = s*m)
__global__ void __launch_bounds__(256) dgrad_weight_planes_kernel(const float* __restrict__ w, int c_in, int c_out, int k, int s,
                                                                  bf16* __restrict__ p0, bf16* __restrict__ p1, bf16* __restrict__ p2,
                                                                  int ld) {
  pdl_enter();
  const int m = k / s;
  const int rows = s * s * c_in, cols = m * m * c_out;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < rows * ld; e += gridDim.x * blockDim.x) {
    const int n = e / ld, kk = e - n * ld;
    float v = 0.0f;
    if (kk < cols) {
      const int pp = n / c_in, ci = n - pp * c_in, py = pp / s, px = pp - py * s;
      const int tap = kk / c_out, co = kk - tap * c_out, i = tap / m, j = tap - i * m;
      v = w[(size_t)(((s * i + py) * k + (s * j + px)) * c_in + ci) * c_out + co];
    }
    bf16 a, b, c;
    split3(v, a, b, c);
    p0[e] = a;
    p1[e] = b;
    p2[e] = c;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn conv_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

struct ViewKey {
  const void* ptr;
  long long dim[5], stride[4];
  int box[5], swz;
  bool operator==(const ViewKey& o) const {
    if (ptr != o.ptr || swz != o.swz) return false;
    for (int i = 0; i < 5; ++i)
      if (dim[i] != o.dim[i] || box[i] != o.box[i]) return false;
    for (int i = 0; i < 4; ++i)
      if (stride[i] != o.stride[i]) return false;
    return true;
  }
};
struct ViewKeyHash {
  size_t operator()(const ViewKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr) ^ (size_t)k.swz;
    for (int i = 0; i < 5; ++i) h = h * 1000003u ^ (size_t)(k.dim[i] * 131 + k.box[i]);
    for (int i = 0; i < 4; ++i) h = h * 1000003u ^ (size_t)k.stride[i];
    return h;
  }
};
static std::unordered_map<ViewKey, CUtensorMap, ViewKeyHash> g_views;
static std::mutex g_views_mu;

// bf16 tensor view of rank `rank` (<= 5): dim / box innermost first, stride[i] = byte stride of dim i+1
int get_view_map(const void* ptr, int rank, const long long* dim, const long long* stride_bytes, const int* box, int swizzle_bytes,
                 CUtensorMap* out) {
  ViewKey key;
  key.ptr = ptr;
  key.swz = swizzle_bytes;
  for (int i = 0; i < 5; ++i) {
    key.dim[i] = i < rank ? dim[i] : 1;
    key.box[i] = i < rank ? box[i] : 1;
  }
  for (int i = 0; i < 4; ++i) key.stride[i] = i + 1 < rank ? stride_bytes[i] : 0;
  std::lock_guard<std::mutex> lock(g_views_mu);
  auto it = g_views.find(key);
  if (it != g_views.end()) {
    *out = it->second;
    return 0;
  }
  EncodeTiledFn fn = conv_encode_fn();
  ACX_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  ACX_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "tensor pointer must be 16-byte aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstride[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dim[i];
    gbox[i] = (cuuint32_t)box[i];
    estr[i] = 1;
    ACX_CHECK(box[i] >= 1 && box[i] <= 256, "box dimension out of range");
  }
  for (int i = 0; i + 1 < rank; ++i) {
    ACX_CHECK(stride_bytes[i] > 0 && (stride_bytes[i] & 15) == 0, "tensor strides must be multiples of 16 bytes");
    gstride[i] = (cuuint64_t)stride_bytes[i];
  }
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstride, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ACX_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (conv view) failed with code " + std::to_string((int)r));
  if (g_views.size() > 1024) g_views.clear();
  g_views[key] = m;
  *out = m;
  return 0;
}

static int conv_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

bool conv_tc_supported(const ConvGeom& g, int dgrad) {
  if (g.s < 1 || g.k % g.s != 0 || g.hw_in % g.s != 0) return false;
  const int m = g.k / g.s, hq = g.hw_in / g.s;
  if (g.hw_out != (g.hw_in - g.k) / g.s + 1) return false;
  if (!dgrad) {
    // forward: innermost run (x1, c) must be one 128-byte swizzle row; a tile holds whole samples
    if (g.s * g.c_in != 64) return false;
    if (g.hw_out * g.hw_out > CV_BM || g.k * m > CV_MAX_SUB) return false;
    return g.c_out == 32 || g.c_out == 64 || g.c_out == 128;
  }
  if (g.c_out != 32 && g.c_out != 64) return false;                       // one tap = 64 or 128 bytes of K
  const int n = g.s * g.s * g.c_in;
  if (n != 32 && n != 64 && n != 128) return false;
  if (hq * hq > CV_BM) return false;
  if (hq < g.hw_out) return false;                                        // every g location must be reached
  return m * m <= CV_MAX_SUB;
}

static int fill_pairs(ConvTcParams* p, int num_pairs, const int* pair_a, const int* pair_b, int a_planes, int b_planes) {
  ACX_CHECK(num_pairs >= 1 && num_pairs <= 6, "num_pairs out of range");
  p->num_pairs = num_pairs;
  p->npa = p->npb = 1;
  for (int i = 0; i < 6; ++i) {
    p->pair_a[i] = i < num_pairs ? pair_a[i] : 0;
    p->pair_b[i] = i < num_pairs ? pair_b[i] : 0;
    if (i < num_pairs) {
      ACX_CHECK(pair_a[i] >= 0 && pair_a[i] < a_planes && pair_b[i] >= 0 && pair_b[i] < b_planes, "plane pair index");
      p->npa = std::max(p->npa, pair_a[i] + 1);
      p->npb = std::max(p->npb, pair_b[i] + 1);
    }
  }
  return 0;
}

static int launch_conv(const CUtensorMap* ta, const CUtensorMap* tb, ConvTcParams& p, cudaStream_t st) {
  const int b_tile = p.bn * CV_BK * 2;
  const int res_bytes = p.b_resident ? p.kb_total * p.npb * b_tile : 0;
  const int stage_bytes = p.npa * CV_A_TILE + (p.b_resident ? 0 : p.npb * b_tile);
  p.stages = (CV_SMEM_LIMIT - CV_SMEM_FIXED - res_bytes) / stage_bytes;
  if (p.stages > CV_MAX_STAGES) p.stages = CV_MAX_STAGES;
  ACX_CHECK(p.stages >= 2, "conv tile does not fit the shared-memory ring");
  static bool configured = false;
  if (!configured) {
    ACX_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM_LIMIT));
    configured = true;
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("ACX_CONV_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  int grid = std::min(p.num_tiles, conv_num_sms());
  if (g_cta_cap > 0 && grid > g_cta_cap) grid = g_cta_cap;
  grid = ceil_div(p.num_tiles, ceil_div(p.num_tiles, grid));   // the smallest grid with the same number of rounds: idle SMs go to other lanes
  const int smem = CV_SMEM_FIXED + res_bytes + p.stages * stage_bytes;
  ACX_CUDA(launch_pdl(conv_tc_kernel, dim3(grid), dim3(CV_THREADS), (size_t)smem, st, ta[0], ta[1], ta[2], tb[0], tb[1], tb[2], p));
  acx::count_launch();
  return 0;
}

// weight operand: K-major [rows, kcols] planes, 128B-swizzled 64-column boxes
static int weight_maps(const Planes& w, int rows, int kcols, int bn, CUtensorMap* tb) {
  ACX_CHECK((w.ld & 7) == 0, "weight plane leading dimension must be a multiple of 8");
  for (int i = 0; i < 3; ++i) {
    const bf16* ptr = w.p[i < w.n ? i : 0];
    const long long dim[2] = {kcols, rows};
    const long long stride[1] = {(long long)w.ld * 2};
    const int box[2] = {CV_BK, bn};
    int r = get_view_map(ptr, 2, dim, stride, box, 128, &tb[i]);
    if (r) return r;
  }
  return 0;
}

// How a tile's 128 GEMM rows are cut out of the (samples x gy x gx) location space: a box of bx x by grid cells (bx | gx,
// by | gy, so that no box straddles the grid edge) of ts consecutive samples.  All tiles cost the same (the MMA always runs
// 128 rows), so the cut that needs the fewest rounds of the persistent grid wins; ties go to the larger spatial box (fewer,
// longer runs per TMA box).  Whole samples per tile (bx = gx, by = gy: round 1's cut) fill only 38 - 78 % of the rows of the
// Nature-CNN's 7x7 / 9x9 / 10x10 grids.  ACX_CONV_PACK=0 restores that cut.
static void pick_tile_cut(int gx, int gy, int samples, int* bx, int* by, int* ts) {
  static int pack = -1;
  if (pack < 0) {
    const char* e = getenv("ACX_CONV_PACK");
    pack = e ? atoi(e) : 1;
  }
  if (const char* e = getenv("ACX_CONV_CUT")) {   // triage: "bx,by,ts" forces the cut (tools/conv_one.py)
    int a = 0, b = 0, c = 0;
    if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a >= 1 && b >= 1 && c >= 1 && gx % a == 0 && gy % b == 0 && a * b * c <= CV_BM) {
      *bx = a;
      *by = b;
      *ts = c < samples ? c : samples;
      return;
    }
  }
  const int sms = conv_num_sms();
  long long best_rounds = -1;
  int best_area = 0;
  for (int cy = 1; cy <= gy; ++cy) {
    if (gy % cy) continue;
    for (int cx = 1; cx <= gx; ++cx) {
      if (gx % cx) continue;
      const int area = cx * cy;
      if (area > CV_BM) continue;
      if (!pack && (cx != gx || cy != gy)) continue;
      int t = CV_BM / area;
      if (t > samples) t = samples;
      const long long tiles = (long long)(gx / cx) * (gy / cy) * ceil_div(samples, t);
      const long long rounds = (tiles + sms - 1) / sms;
      if (best_rounds < 0 || rounds < best_rounds || (rounds == best_rounds && area > best_area)) {
        best_rounds = rounds;
        best_area = area;
        *bx = cx;
        *by = cy;
        *ts = t;
      }
    }
  }
}

// y = relu(conv(x, W) + bias) as bf16 planes; x planes [samples, hw_in, hw_in, c_in], wT planes [c_out, k*k*c_in]
int conv_tc_forward(const Planes& x, const Planes& wT, const ConvGeom& g, int samples, const float* bias, int relu, const Planes& y,
                    int num_pairs, const int* pair_a, const int* pair_b, cudaStream_t st) {
  ACX_CHECK(conv_tc_supported(g, 0), "unsupported convolution geometry for the implicit-GEMM forward");
  ACX_CHECK(samples > 0 && x.n >= 1 && wT.n >= 1 && y.n >= 1 && y.ld == g.c_out, "bad operands");
  const int m = g.k / g.s, hq = g.hw_in / g.s;
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  int r = fill_pairs(&p, num_pairs, pair_a, pair_b, x.n, wT.n);
  if (r) return r;
  p.gx = p.gy = g.hw_out;
  pick_tile_cut(p.gx, p.gy, samples, &p.bx, &p.by, &p.ts);
  p.nx = p.gx / p.bx;
  p.cells = p.nx * (p.gy / p.by);
  p.ydim = 3;                            // the grid row is the yq coordinate of the 5-D view
  p.rows_valid = p.ts * p.bx * p.by;
  p.num_tiles = p.cells * ceil_div(samples, p.ts);
  p.lo_from_tile = p.num_tiles;          // every tile at full precision
  p.num_pairs_lo = num_pairs;
  p.samples = samples;
  p.bn = g.c_out;
  const int K = g.k * g.k * g.c_in;
  p.kb_total = K / CV_BK;
  p.nsub = 1;
  p.num_sub = g.k * m;
  p.sub_bytes = p.rows_valid * CV_BK * 2;
  for (int kh = 0; kh < g.k; ++kh)
    for (int kw2 = 0; kw2 < m; ++kw2) {
      const int t = kh * m + kw2;
      p.tc1[t] = (signed char)kw2;        // xq
      p.tc2[t] = (signed char)(kh % g.s); // y1
      p.tc3[t] = (signed char)(kh / g.s); // yq
    }
  p.dgrad = 0;
  p.hw_in = g.hw_in;
  p.c_in = g.c_in;
  p.s = g.s;
  p.hq = hq;
  p.alpha = 1.0f;
  p.bias = bias;
  p.relu = relu;
  for (int i = 0; i < 3; ++i) p.cp[i] = i < y.n ? y.p[i] : nullptr;
  p.npl = y.n;
  p.ldcp = y.ld;
  p.mask = nullptr;
  p.mask_samples = 1;
  ACX_CHECK((g.c_out & 7) == 0, "forward: c_out % 8");   // the epilogue moves 8 channels (16 bytes) per lane
  for (int i = 0; i < y.n; ++i) ACX_CHECK((reinterpret_cast<uintptr_t>(y.p[i]) & 15) == 0, "output planes must be 16-byte aligned");
  CUtensorMap ta[3], tb[3];
  const long long row = (long long)g.hw_in * g.c_in;   // elements of one input row
  for (int i = 0; i < 3; ++i) {
    const bf16* ptr = x.p[i < x.n ? i : 0];
    const long long dim[5] = {64, hq, g.s, hq, samples};
    const long long stride[4] = {64 * 2, row * 2, row * g.s * 2, row * g.hw_in * 2};
    const int box[5] = {64, p.bx, 1, p.by, p.ts};
    r = get_view_map(ptr, 5, dim, stride, box, 128, &ta[i]);
    if (r) return r;
  }
  r = weight_maps(wT, g.c_out, K, p.bn, tb);
  if (r) return r;
  return launch_conv(ta, tb, p, st);
}

// conv1 (8x8 / stride 4 on [84, 84, 4], envs/atari/model.py:173-179) straight from the row-pair interleaved observation copy
// (obs_pairs_bf16, layers.cu): y = relu(alpha * conv(obs, W) + bias).  In that copy the two kernel rows kh = 2j, 2j + 1 of the
// patch at (oy, ox) are one run of 64 elements at pair-row 2 oy + j, pixel 4 ox, so the K dimension is 4 chunks of 64 and the
// view  [sample][oy + j / 2][j % 2][ox][64]  has the shape of the forward views above (x stride = 4 pixels = 64 bytes: the
// runs of neighbouring locations overlap).  wT_perm: [32, 256] planes whose columns follow the copy's (kw, parity, c) order.
int conv1_pairs_forward(const bf16* obs_pairs, const Planes& wT_perm, int samples, const float* bias, float alpha, const Planes& y,
                        int num_pairs, const int* pair_a, const int* pair_b, cudaStream_t st) {
  ACX_CHECK(samples > 0 && obs_pairs != nullptr && wT_perm.n >= 1 && y.n >= 1 && y.ld == 32, "bad operands");
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  int r = fill_pairs(&p, num_pairs, pair_a, pair_b, 1, wT_perm.n);
  if (r) return r;
  p.gx = p.gy = 20;
  pick_tile_cut(p.gx, p.gy, samples, &p.bx, &p.by, &p.ts);
  p.nx = p.gx / p.bx;
  p.cells = p.nx * (p.gy / p.by);
  p.ydim = 3;
  p.rows_valid = p.ts * p.bx * p.by;
  p.num_tiles = p.cells * ceil_div(samples, p.ts);
  p.lo_from_tile = p.num_tiles;
  p.num_pairs_lo = num_pairs;
  p.samples = samples;
  p.bn = 32;
  p.kb_total = 4;
  p.nsub = 1;
  p.num_sub = 4;
  p.sub_bytes = p.rows_valid * CV_BK * 2;
  for (int j = 0; j < 4; ++j) {
    p.tc1[j] = 0;
    p.tc2[j] = (signed char)(j & 1);    // pair-row parity
    p.tc3[j] = (signed char)(j >> 1);   // + oy
  }
  p.dgrad = 0;
  p.hw_in = 84;
  p.c_in = 4;
  p.s = 4;
  p.hq = 21;
  p.alpha = alpha;
  p.bias = bias;
  p.relu = 1;
  for (int i = 0; i < 3; ++i) p.cp[i] = i < y.n ? y.p[i] : nullptr;
  p.npl = y.n;
  p.ldcp = y.ld;
  p.mask = nullptr;
  p.mask_samples = 1;
  {
    static int bres = -1;
    if (bres < 0) {
      const char* e = getenv("ACX_CONV1_BRES");   // 0: reload the weight tiles with every k-block (as the other layers do)
      bres = e ? atoi(e) : 0;   // measured: no gain (0.7237 vs 0.7244 ms/update) - the kernel is not bound by these loads
    }
    p.b_resident = bres;   // 4 k-blocks x <= 3 planes x 4 KB
  }
  for (int i = 0; i < y.n; ++i) ACX_CHECK((reinterpret_cast<uintptr_t>(y.p[i]) & 15) == 0, "output planes must be 16-byte aligned");
  CUtensorMap ta[3], tb[3];
  const long long prow = 84 * 8;   // elements of one pair-row
  const long long dim[5] = {64, 20, 2, 21, samples};
  const long long stride[4] = {32 * 2, prow * 2, 2 * prow * 2, 42 * prow * 2};
  const int box[5] = {64, p.bx, 1, p.by, p.ts};
  for (int i = 0; i < 3; ++i) {
    r = get_view_map(obs_pairs, 5, dim, stride, box, 128, &ta[i]);
    if (r) return r;
  }
  r = weight_maps(wT_perm, 32, 256, p.bn, tb);
  if (r) return r;
  return launch_conv(ta, tb, p, st);
}

// dx = relu'(act_below) * conv_transpose(gout, W) as bf16 planes [samples, hw_in, hw_in, c_in];
// gout planes [samples, hw_out, hw_out, c_out]; wD = the rearranged weight planes of conv_dgrad_weight_planes;
// mask_hi = hi plane of the forward activation below ([mask_samples, hw_in, hw_in, c_in]; sample r uses r % mask_samples)
int conv_tc_dgrad(const Planes& gout, const Planes& wD, const ConvGeom& g, int samples, const bf16* mask_hi, int mask_samples,
                  const Planes& dx, int num_pairs, const int* pair_a, const int* pair_b, cudaStream_t st, int lo_from_sample,
                  int num_pairs_lo) {
  ACX_CHECK(conv_tc_supported(g, 1), "unsupported convolution geometry for the gather-form input gradient");
  ACX_CHECK(lo_from_sample < 0 || (num_pairs_lo >= 1 && num_pairs_lo <= num_pairs), "num_pairs_lo out of range");
  ACX_CHECK(samples > 0 && gout.n >= 1 && wD.n >= 1 && dx.n >= 1 && dx.ld == g.c_in && gout.ld == g.c_out, "bad operands");
  const int m = g.k / g.s, hq = g.hw_in / g.s;
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  int r = fill_pairs(&p, num_pairs, pair_a, pair_b, gout.n, wD.n);
  if (r) return r;
  p.gx = p.gy = hq;
  pick_tile_cut(p.gx, p.gy, samples, &p.bx, &p.by, &p.ts);
  p.nx = p.gx / p.bx;
  p.cells = p.nx * (p.gy / p.by);
  p.ydim = 2;                            // the grid row is coordinate 2 (a) of the gradient view
  p.rows_valid = p.ts * p.bx * p.by;
  p.num_tiles = p.cells * ceil_div(samples, p.ts);
  // tiles are ordered sample group first: the groups that hold only samples >= lo_from_sample run at the lower precision (a
  // group that straddles the boundary keeps the full one)
  p.lo_from_tile = lo_from_sample >= 0 ? p.cells * ceil_div(lo_from_sample, p.ts) : p.num_tiles;
  p.num_pairs_lo = lo_from_sample >= 0 ? num_pairs_lo : num_pairs;
  p.samples = samples;
  p.bn = g.s * g.s * g.c_in;
  const int K = m * m * g.c_out;
  p.kb_total = ceil_div(K, CV_BK);
  p.nsub = CV_BK / g.c_out;            // c_out = 64: one tap per k-block; 32: two taps (64-byte swizzle sub-tiles)
  p.num_sub = m * m;
  p.sub_bytes = p.rows_valid * g.c_out * 2;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      const int t = i * m + j;
      p.tc1[t] = (signed char)(-j);    // b - j
      p.tc2[t] = (signed char)(-i);    // a - i
      p.tc3[t] = 0;
    }
  p.dgrad = 1;
  p.hw_in = g.hw_in;
  p.c_in = g.c_in;
  p.s = g.s;
  p.hq = hq;
  p.alpha = 1.0f;
  p.bias = nullptr;
  p.relu = 0;
  for (int i = 0; i < 3; ++i) p.cp[i] = i < dx.n ? dx.p[i] : nullptr;
  p.npl = dx.n;
  p.ldcp = dx.ld;
  p.mask = mask_hi;
  p.mask_samples = mask_samples > 0 ? mask_samples : samples;
  // the epilogue moves 8 channels (16 bytes) per lane
  ACX_CHECK((g.c_in & 7) == 0 && (reinterpret_cast<uintptr_t>(mask_hi) & 15) == 0, "input gradient: c_in % 8 / mask alignment");
  for (int i = 0; i < dx.n; ++i) ACX_CHECK((reinterpret_cast<uintptr_t>(dx.p[i]) & 15) == 0, "output planes must be 16-byte aligned");
  CUtensorMap ta[3], tb[3];
  const long long row = (long long)g.hw_out * g.c_out;
  for (int i = 0; i < 3; ++i) {
    const bf16* ptr = gout.p[i < gout.n ? i : 0];
    const long long dim[5] = {g.c_out, g.hw_out, g.hw_out, 1, samples};
    const long long stride[4] = {(long long)g.c_out * 2, row * 2, row * g.hw_out * 2, row * g.hw_out * 2};
    const int box[5] = {g.c_out, p.bx, p.by, 1, p.ts};
    r = get_view_map(ptr, 5, dim, stride, box, g.c_out == 64 ? 128 : 64, &ta[i]);
    if (r) return r;
  }
  r = weight_maps(wD, p.bn, K, p.bn, tb);
  if (r) return r;
  return launch_conv(ta, tb, p, st);
}

int conv_dgrad_weight_planes(const float* w, const ConvGeom& g, const Planes& out, cudaStream_t st) {
  ACX_CHECK(out.n == 3, "dgrad weight planes: 3 planes expected");
  const int m = g.k / g.s;
  ACX_CHECK(out.ld >= m * m * g.c_out, "dgrad weight planes: leading dimension too small");
  const int total = g.s * g.s * g.c_in * out.ld;
  ACX_PDL_LAUNCH(dgrad_weight_planes_kernel, std::min(ceil_div(total, 256), 296), 256, 0, st, w, g.c_in, g.c_out, g.k, g.s, out.p[0], out.p[1], out.p[2], out.ld);
  return 0;
}

int conv_trace(long long* out8) {
  return cudaMemcpyFromSymbol(out8, g_conv_trace, 8 * sizeof(long long)) == cudaSuccess ? 0 : 1;
}

int conv_error_flag() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_tc_error, sizeof(int));
  return v;
}

}  // namespace acx

extern "C" {

static acx::Planes to_planes(const acx_planes_t& a) {
  acx::Planes p;
  p.n = a.num_planes;
  p.ld = a.ld;
  for (int i = 0; i < a.num_planes && i < 3; ++i) p.p[i] = reinterpret_cast<acx::bf16*>(const_cast<void*>(a.planes[i]));
  return p;
}

int acx_conv_supported(const acx_conv_t* c) {
  if (!c) return 0;
  const acx::ConvGeom g = {c->hw_in, c->c_in, c->k, c->stride, c->hw_out, c->c_out};
  return acx::conv_tc_supported(g, c->dgrad) ? 1 : 0;
}

int acx_conv(const acx_conv_t* c, void* stream) {
  ACX_CHECK(c != nullptr, "null argument");
  ACX_CHECK(c->x.num_planes >= 1 && c->x.num_planes <= 3 && c->w.num_planes >= 1 && c->w.num_planes <= 3 &&
                c->out.num_planes >= 1 && c->out.num_planes <= 3,
            "plane counts");
  const acx::ConvGeom g = {c->hw_in, c->c_in, c->k, c->stride, c->hw_out, c->c_out};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (c->dgrad)
    return acx::conv_tc_dgrad(to_planes(c->x), to_planes(c->w), g, c->samples, reinterpret_cast<const acx::bf16*>(c->mask_plane),
                              c->mask_samples, to_planes(c->out), c->num_pairs, c->pair_a, c->pair_b, st);
  return acx::conv_tc_forward(to_planes(c->x), to_planes(c->w), g, c->samples, c->bias, c->relu, to_planes(c->out), c->num_pairs,
                              c->pair_a, c->pair_b, st);
}

int acx_debug_conv_trace(long long* h_out8) { return acx::conv_trace(h_out8); }

int acx_conv1_pairs_forward(const void* d_obs_pairs, const acx_planes_t* w_perm, int samples, const float* d_bias, float alpha,
                            const acx_planes_t* out, int num_pairs, const int* pair_a, const int* pair_b, void* stream) {
  ACX_CHECK(d_obs_pairs && w_perm && out && pair_a && pair_b, "null argument");
  return acx::conv1_pairs_forward(reinterpret_cast<const acx::bf16*>(d_obs_pairs), to_planes(*w_perm), samples, d_bias, alpha,
                                  to_planes(*out), num_pairs, pair_a, pair_b, reinterpret_cast<cudaStream_t>(stream));
}

int acx_conv_dgrad_weights(const float* d_w, int hw_in, int c_in, int k, int stride, int hw_out, int c_out, void* const* d_planes,
                           int ld, void* stream) {
  ACX_CHECK(d_w != nullptr && d_planes != nullptr, "null argument");
  const acx::ConvGeom g = {hw_in, c_in, k, stride, hw_out, c_out};
  acx::Planes out;
  out.n = 3;
  out.ld = ld;
  for (int i = 0; i < 3; ++i) out.p[i] = reinterpret_cast<acx::bf16*>(d_planes[i]);
  return acx::conv_dgrad_weight_planes(d_w, g, out, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
