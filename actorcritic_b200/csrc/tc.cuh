// tc.cuh - sm_100a PTX wrappers shared by the tensor-core kernels (gemm.cu, conv.cu): mbarrier, TMA tile loads,
// TMEM allocation, tcgen05.mma / commit / ld, shared-memory matrix descriptors.
#pragma once
#include <cstring>

#include "common.cuh"

namespace acx {

// first barrier-protocol error seen by a tensor-core kernel of this translation unit (0 = none); no relocatable device
// code in this build, so every .cu that includes this header owns a copy and exports a host-side reader
static __device__ int g_tc_error = 0;


// conv.cu: cached tensor map of a bf16 tensor view of rank <= 5 (dim / box innermost first, stride_bytes[i] = byte stride of
// dim i + 1); swizzle_bytes = 128 or 64
int get_view_map(const void* ptr, int rank, const long long* dim, const long long* stride_bytes, const int* box, int swizzle_bytes,
                 CUtensorMap* out);

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a wrong barrier protocol must not hang the GPU box (records an error and moves on)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int code) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {  // ~2 s
      atomicExch(&g_tc_error, code);
      return;
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// The four K = 16 steps of one 64-deep k-block of one plane pair in ONE asm statement: the descriptors advance inside the block
// (a/b step in 16-byte units), so ptxas moves the operands into uniform registers once and steps them there.  With one asm
// statement per MMA it re-materialised every operand - descriptor halves, instruction descriptor, TMEM address - from vector
// registers before each UTCHMMA: 16 instructions per MMA on the single issuing thread, which is what bounded the kernels
// (ncu source view: the MMA warp never stalls, it executes; 105 - 144 cycles per MMA against 40 - 64 in the tensor pipe).
__device__ __forceinline__ void umma_bf16_x4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t astep16, uint32_t bstep16,
                                             uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db, sa, sb;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 q, %5, %5;\n\t"
      "cvt.u64.u32 sa, %3;\n\t"
      "cvt.u64.u32 sb, %4;\n\t"
      "mov.b64 da, %1;\n\t"
      "mov.b64 db, %2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(astep16), "r"(bstep16), "r"(idesc), "r"(acc)
      : "memory");
}
// the same with explicit A offsets of steps 1..3 (16-byte units, relative to step 0): a k-block made of two 32-wide, 64-byte-swizzled
// sub-tiles (conv3's input gradient at 32 channels) steps 0, +2 inside the first sub-tile and sub, sub + 2 inside the second
__device__ __forceinline__ void umma_bf16_x4_steps(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t a1, uint32_t a2, uint32_t a3,
                                                   uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db, t;\n\t"
      "setp.ne.b32 p, %7, 0;\n\t"
      "setp.eq.b32 q, %6, %6;\n\t"
      "mov.b64 db, %2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, db, %6, p;\n\t"
      "cvt.u64.u32 t, %3;\n\t"
      "add.u64 da, %1, t;\n\t"
      "add.u64 db, db, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, q;\n\t"
      "cvt.u64.u32 t, %4;\n\t"
      "add.u64 da, %1, t;\n\t"
      "add.u64 db, db, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, q;\n\t"
      "cvt.u64.u32 t, %5;\n\t"
      "add.u64 da, %1, t;\n\t"
      "add.u64 db, db, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, q;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(a1), "r"(a2), "r"(a3), "r"(idesc), "r"(acc)
      : "memory");
}
// ... and all plane pairs of a k-block in one statement (2 or 3 pairs: the shapes the learner runs)
__device__ __forceinline__ void umma_bf16_x4_pairs2(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t astep16, uint32_t bstep16,
                                                    uint32_t idesc, uint32_t acc, uint32_t oa0, uint32_t ob0, uint32_t oa1, uint32_t ob1) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db, sa, sb, t;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 q, %5, %5;\n\t"
      "cvt.u64.u32 sa, %3;\n\t"
      "cvt.u64.u32 sb, %4;\n\t"
      "cvt.u64.u32 t, %7;\n\t"
      "add.u64 da, %1, t;\n\t"
      "cvt.u64.u32 t, %8;\n\t"
      "add.u64 db, %2, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "cvt.u64.u32 t, %9;\n\t"
      "add.u64 da, %1, t;\n\t"
      "cvt.u64.u32 t, %10;\n\t"
      "add.u64 db, %2, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "}" ::"r"(tmem_d), "l"(a0), "l"(b0), "r"(astep16), "r"(bstep16), "r"(idesc), "r"(acc), "r"(oa0), "r"(ob0), "r"(oa1), "r"(ob1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_x4_pairs3(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t astep16, uint32_t bstep16,
                                                    uint32_t idesc, uint32_t acc, uint32_t oa0, uint32_t ob0, uint32_t oa1, uint32_t ob1, uint32_t oa2, uint32_t ob2) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db, sa, sb, t;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 q, %5, %5;\n\t"
      "cvt.u64.u32 sa, %3;\n\t"
      "cvt.u64.u32 sb, %4;\n\t"
      "cvt.u64.u32 t, %7;\n\t"
      "add.u64 da, %1, t;\n\t"
      "cvt.u64.u32 t, %8;\n\t"
      "add.u64 db, %2, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "cvt.u64.u32 t, %9;\n\t"
      "add.u64 da, %1, t;\n\t"
      "cvt.u64.u32 t, %10;\n\t"
      "add.u64 db, %2, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "cvt.u64.u32 t, %11;\n\t"
      "add.u64 da, %1, t;\n\t"
      "cvt.u64.u32 t, %12;\n\t"
      "add.u64 db, %2, t;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u64 da, da, sa;\n\t"
      "add.u64 db, db, sb;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "}" ::"r"(tmem_d), "l"(a0), "l"(b0), "r"(astep16), "r"(bstep16), "r"(idesc), "r"(acc), "r"(oa0), "r"(ob0), "r"(oa1), "r"(ob1), "r"(oa2), "r"(ob2)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (SWIZZLE_128B, sm_100 "version 1")
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint2 pack4(bf16 a, bf16 b, bf16 c, bf16 d) {
  uint2 r;
  r.x = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
  r.y = (uint32_t)__bfloat16_as_ushort(c) | ((uint32_t)__bfloat16_as_ushort(d) << 16);
  return r;
}

// one lane of a converged warp (elect.sync): unlike `lane == 0`, ptxas then knows that a single thread executes the
// guarded block, so the uniform-register operands of UTMALDG / UTCHMMA need no per-instruction election loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// N-dimensional TMA tile loads (coordinates innermost first; out-of-range elements are zero-filled)
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// shared-memory matrix descriptor with an explicit swizzle mode: layout 2 = SWIZZLE_128B (8-row atoms of 1024 B),
// layout 4 = SWIZZLE_64B (8-row atoms of 512 B)
__device__ __forceinline__ uint64_t make_smem_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

}  // namespace acx
