// tc_common.cuh -- sm_100a building blocks shared by the tensor-core kernels:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), UMMA descriptors.
// Inline PTX only; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tmg {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint (ns): the hardware parks the thread until the phase flips or the hint expires, so a
// waiting warp does not burn issue slots that the arithmetic warps of the same scheduler need.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t *bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ long long clock64_volatile() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;\n" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must fault the kernel (trap -> launch error the host reports), never hang the GPU.
// The spin loop is three instructions; the clock is read only once every 256 failed probes.
#ifndef TM_MBAR_TIMEOUT_CYCLES
#define TM_MBAR_TIMEOUT_CYCLES (8ll * 1000 * 1000 * 1000)
#endif
#ifndef TM_WAIT_HINT
#define TM_WAIT_HINT 1
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  uint32_t spins = 0;
#if TM_WAIT_HINT
  while (!mbar_try_wait_hint(bar, parity, 100000u)) {
#else
  while (!mbar_try_wait(bar, parity)) {
#endif
    if ((++spins & 0xFF) == 0) {
      const long long t = clock64_volatile();
      if (t0 == 0) t0 = t;
      else if (t - t0 > TM_MBAR_TIMEOUT_CYCLES) __trap();
    }
  }
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(tmap) : "memory");
}
// 2-D tiled load: coordinates are (inner, outer) in elements of the tensor map
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const void *tmap, uint64_t *bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 1-D bulk copy global -> shared (bytes a multiple of 16, both addresses 16-byte aligned), completion on the mbarrier
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// tcgen05.commit: arrives on the mbarrier when all previously issued tcgen05 async ops of this thread complete
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], 8-bit integer operands, int32 accumulate
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A rows live in TMEM lanes, K bytes packed along columns (8 columns per K = 32 step)
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-uniform variants: the WHOLE warp executes the call, one elected lane issues.  Keeping the issuing loop free of
// lane-divergent control flow lets ptxas keep descriptors in uniform registers (a `if (lane == 0)` around the issue
// loop costs ~8 SASS instructions per MMA, more than the 32 cycles an int8 128x64x32 MMA takes).
__device__ __forceinline__ void mma_i8_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All 24 MMAs of one 128x64 int8-limb tile (HH, HL, LH, LL x 6 K-steps) under ONE elect.  The kernel owns the whole
// TMEM (512 columns, base 0 -- checked at run time), so accumulator / query-row addresses and instruction descriptors
// are immediates; only the low word of the dictionary tile's shared-memory descriptor is a register.  This takes the
// issue cost from ~12 to ~4 SASS instructions per MMA (a 128x64x32 int8 MMA occupies the tensor pipe for 32 cycles).
template <int PARITY>
__device__ __forceinline__ void mma_i8_tile_elect(uint32_t b_desc_lo) {
  if (PARITY == 0) {
    asm volatile(
#include "knn_mma_tile_p0.inc"
        ::"r"(b_desc_lo)
        : "memory");
  } else {
    asm volatile(
#include "knn_mma_tile_p1.inc"
        ::"r"(b_desc_lo)
        : "memory");
  }
}
__device__ __forceinline__ void tc_commit_elect(uint64_t *bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], f16/bf16 operands, f32 accumulate
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread t of the warp reads lane base+t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// registers -> 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// ------------------------------------------------------------------ descriptors
// K-major operand tile stored as consecutive 8-row x 128-byte swizzle atoms (TMA SWIZZLE_128B box of 128 bytes):
// start address (>>4), LBO ignored for swizzled K-major, SBO = 1024 bytes between 8-row groups, version 1 (sm_100),
// layout type 2 = SWIZZLE_128B.  Advancing along K inside the 128-byte atom = adding the byte offset to the start.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

// instruction descriptor (upper 32 bits of the "idesc" operand): see PTX ISA tcgen05 instruction descriptor
//   [4,6) D format (0 f16, 1 f32, 2 s32)  [7,10) A format  [10,13) B format  [15] A major  [16] B major
//   [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(uint32_t dfmt, uint32_t afmt, uint32_t bfmt, uint32_t M, uint32_t N) {
  return (dfmt << 4) | (afmt << 7) | (bfmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
constexpr uint32_t kFmtU8 = 0, kFmtS8 = 1;       // kind::i8
constexpr uint32_t kFmtF16 = 0, kFmtBF16 = 1;    // kind::f16
constexpr uint32_t kDFmtF32 = 1, kDFmtS32 = 2;

}  // namespace tmg
