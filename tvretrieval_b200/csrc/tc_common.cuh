// sm_100a primitives for the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Everything is inline PTX; the only host-side dependency is cuTensorMapEncodeTiled, resolved at run time through
// cudaGetDriverEntryPoint (no link against libcuda).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace tc {

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s
  }
}

// ---------------------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2D tile load global -> shared, completion signalled on an mbarrier (bytes)
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// L2 eviction-priority policies for streaming operands (evict_first) and small re-used ones (evict_last): the
// pre-encoded descriptors CUTLASS uses (cute::TMA::CacheHintSm90); `createpolicy` at run time trapped as an illegal
// instruction in these kernels on B200.
__device__ __forceinline__ uint64_t l2_policy_evict_first() { return 0x12F0000000000000ull; }
__device__ __forceinline__ uint64_t l2_policy_evict_last() { return 0x14F0000000000000ull; }
__device__ __forceinline__ void tma_load_2d_hint(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

// Row gather: FOUR rows r0..r3 of a 2-D tensor (tensor map encoded with a box of ONE row) land in four consecutive
// rows of the shared-memory tile, swizzled like a regular box load (the swizzle is a function of the destination
// address) -- measured on B200: tools/gpu notes in DESIGN.md.  Completion: 4 * box_cols * 2 bytes on the mbarrier.
__device__ __forceinline__ void tma_gather4_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int r0,
                                               int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

// ---------------------------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand whose rows are exactly one swizzle span wide
// (SWIZZLE_BYTES = 128: rows of 64 16-bit elements, 64: rows of 32): 8-row swizzle atoms stacked along M/N every
// 8 * SWIZZLE_BYTES (SBO); LBO unused for swizzled K-major.
template <int SWIZZLE_BYTES>
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t smem_addr) {
  static_assert(SWIZZLE_BYTES == 128 || SWIZZLE_BYTES == 64, "supported swizzle spans");
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // bits [0,14)  start address >> 4
  d |= (uint64_t)((8 * SWIZZLE_BYTES) >> 4) << 32;         // bits [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                                  // bits [46,48) descriptor version = 1 (sm_100)
  d |= (uint64_t)(SWIZZLE_BYTES == 128 ? 2 : 4) << 61;     // bits [61,64) layout = SWIZZLE_128B / SWIZZLE_64B
  return d;
}

// Instruction descriptor for kind::f16 (A/B fp16 or bf16, D fp32), both operands K-major.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt  [10,13) B fmt (0 = f16, 1 = bf16)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n, int is_bf16) {
  return (1u << 4) | ((uint32_t)is_bf16 << 7) | ((uint32_t)is_bf16 << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// One lane of a CONVERGED warp (all 32 lanes must execute it): true for the elected leader -- always the same lane.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Plain contiguous bulk copy global -> shared (no tensor map, no per-row processing), completion on an mbarrier
// (bytes); size and both addresses multiples of 16.
__device__ __forceinline__ void bulk_load_hint(uint32_t smem_dst, const void* src, uint32_t bytes, uint32_t bar,
                                               uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive columns (one fp32 per lane per column)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// same, 16 consecutive columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc

// ---------------------------------------------------------------------------------------------- host side
// Encodes a 2D tiled tensor map over a row-major [rows][cols] array of 16-bit elements; the swizzle span equals the
// box width: box_cols = 64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B.  Returns XMLB_OK or an error code (message set).
int xmlb_make_tmap_2d_u16(CUtensorMap* out, const void* base, unsigned long long rows, unsigned long long cols,
                          unsigned int box_rows, unsigned int box_cols);
