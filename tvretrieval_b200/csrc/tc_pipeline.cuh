// Shared warp-specialised mainloop of the split-precision tcgen05 GEMM kernels (vr_scores_tc.cu, linear_tc.cu,
// span_tc.cu).  One CTA = 192 threads:
//   warp 0 lane 0 : scheduler + TMA producer  -> tc_producer_loop
//   warp 1        : MMA issuer (converged warp, one elected lane issues) -> tc_mma_loop_warp
//   warps 2..5    : epilogue (kernel specific), driven by epi_next / epi_wait / epi_release
// A unit = one accumulation D[128 x n] = sum over k-blocks of A_tile . B_tile^T with both operands given as (hi, lo)
// 16-bit pairs: 3 MMAs per k-step (hi*lo + lo*hi + hi*hi), fp32 accumulate in TMEM.  Units alternate between two
// 256-column TMEM accumulators so that the epilogue of unit u overlaps the MMAs of unit u+1.
//
// Scheduling is DYNAMIC: only the producer thread decides what the CTA does next -- each kernel's `Sched` claims
// tiles from a global atomic counter -- and publishes one small descriptor per unit through a shared-memory queue
// to the MMA thread and the epilogue warps.  Tiles are therefore started in global order no matter how fast each
// CTA runs, which keeps the 148 in-flight tiles a contiguous window of the operand matrices: with a static
// round-robin assignment the CTAs drifted apart over a launch and the shared operand tiles fell out of L2
// (measured: 6x DRAM re-reads, tensor pipe 69 %).
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int BLOCK_M = 128;
// k-block = one swizzle row.  32 elements (64 B rows, SWIZZLE_64B) instead of 64: the same shared memory holds twice
// as many stages, so 3/4 of it can be in flight while 1/4 is being consumed (with two 96 KB stages only half could),
// which is what hides the ~2000-cycle TMA round trip at 64 B/cycle/SM of operand traffic.
constexpr int BLOCK_K = 32;
constexpr int SWIZZLE_BYTES = BLOCK_K * 2;
constexpr int MAX_STAGES = 8;
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int QUEUE_DEPTH = 4;
constexpr int PIPE_FIXED_BYTES = 512;  // barriers + TMEM slot + unit queue
constexpr int GATHER_WARPS = 4;        // extra warps of the kernels whose A or B rows are picked by an index list

struct UnitDesc {
  const CUtensorMap* a_hi;
  const CUtensorMap* a_lo;
  const CUtensorMap* b_hi;
  const CUtensorMap* b_lo;
  int a_row, b_row;   // first row of the A / B tile in their tensor maps
  int k_blocks;       // number of BLOCK_K-wide k-blocks to accumulate (> 0)
  int k_block0;       // first k-block (K-chunked accumulation: a unit may cover only a slice of K); 0 by default
  int g_count;        // gather producer only: number of valid list entries of the gathered operand's tile
  int a_bytes;        // bytes one A box (hi or lo) delivers: A_TILE_BYTES unless the unit's A maps have a shorter box
                      // (honoured on the gather-warps path, p.gather == 2, only)
  const unsigned short* a_hi_bulk;  // != null (p.gather == 2 path): A is stored as the k-blocked SHARED-MEMORY IMAGE of its
  const unsigned short* a_lo_bulk;  // tiles (a_kb_rows > 0 and the 16-byte pieces of every row already permuted like
                                    // SWIZZLE_64B would place them): a k-block's rows are fetched with ONE plain bulk
                                    // copy of a_bytes instead of a tensor box that the TMA unit walks row by row
  int b_bytes;        // the same for the B boxes (p.b_tile_bytes by default; honoured on the p.gather == 1 path only)
  int b_kb_rows;      // > 0: B is stored k-blocked (see a_kb_rows)
  int a_kb_rows;      // 0 = A is (rows, K) row-major; > 0 = A is stored K-BLOCKED, [K / BLOCK_K][a_kb_rows][BLOCK_K]
                      // (its maps describe a (K / BLOCK_K * a_kb_rows, BLOCK_K) array): the box of a k-block is one
                      // contiguous run of box_rows * 64 bytes in HBM instead of box_rows separate 64-byte pieces
  uint32_t idesc;     // instruction descriptor (carries N of this unit)
  int tag0, tag1;     // kernel-specific payload handed to the epilogue (tile index, modality, ...)
};

// shared-memory carve-up (all offsets are shared-window addresses)
struct Pipe {
  uint32_t smem_base;   // 1024-aligned start of the stage ring
  uint32_t bar_base;
  int stages, stage_bytes, b_tile_bytes;
  int probe;  // limiter experiments (XMLB_VR_PROBE): bit 1 = the producer skips the B tiles
  int epi_warps;  // epilogue warps that consume every unit (4, or 8 when two warps share a TMEM lane quadrant)
  int gather;     // 0: both operands by TMA; 1 / 2: the A / B tile is filled by GATHER_WARPS extra warps (tc_gather_loop)
  int terms;  // 3: split precision (hi*lo + lo*hi + hi*hi, stage = A_hi|A_lo|B_hi|B_lo); 1: hi*hi only (A_hi|B_hi)
  __device__ uint32_t full_bar(int s) const { return bar_base + 8u * s; }
  __device__ uint32_t empty_bar(int s) const { return bar_base + 8u * (stages + s); }
  __device__ uint32_t tfull_bar(int a) const { return bar_base + 8u * (2 * stages + a); }
  __device__ uint32_t tempty_bar(int a) const { return bar_base + 8u * (2 * stages + 2 + a); }
  __device__ uint32_t qfull_bar(int i) const { return bar_base + 8u * (2 * stages + 4 + i); }
  __device__ uint32_t qempty_bar(int i) const { return bar_base + 8u * (2 * stages + 4 + QUEUE_DEPTH + i); }
  __device__ uint32_t tmem_slot() const { return bar_base + 8u * (2 * stages + 4 + 2 * QUEUE_DEPTH); }
  __device__ uint32_t queue(int i) const { return bar_base + 256u + 16u * i; }  // int4 {tag0, tag1, k_blocks, idesc}
  __device__ uint32_t extra() const { return bar_base + PIPE_FIXED_BYTES; }       // kernel-specific scratch
};

__host__ __device__ inline int pipe_stage_bytes(int block_n, int terms = 3) {
  return (terms == 3 ? 2 : 1) * (A_TILE_BYTES + block_n * BLOCK_K * 2);
}
// number of ring stages that fit next to `extra_bytes` of kernel-specific shared memory
inline int pipe_stages(int block_n, int extra_bytes, int terms = 3) {
  const int s = (SMEM_LIMIT - 1024 - PIPE_FIXED_BYTES - extra_bytes) / pipe_stage_bytes(block_n, terms);
  return s > MAX_STAGES ? MAX_STAGES : s;
}
inline size_t pipe_smem_bytes(int block_n, int stages, int extra_bytes, int terms = 3) {
  return 1024 + (size_t)stages * pipe_stage_bytes(block_n, terms) + PIPE_FIXED_BYTES + extra_bytes;
}

// Called by all threads (64 + 32 * epi_warps) at kernel start.  Returns the TMEM base address.
__device__ __forceinline__ uint32_t pipe_setup(Pipe& p, unsigned char* smem_raw, int stages, int block_n,
                                               int terms = 3, int epi_warps = 4, int gather = 0) {
  p.smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  p.stages = stages;
  p.terms = terms;
  p.probe = 0;
  p.epi_warps = epi_warps;
  p.gather = gather;
  p.b_tile_bytes = block_n * BLOCK_K * 2;
  p.stage_bytes = (terms == 3 ? 2 : 1) * (A_TILE_BYTES + p.b_tile_bytes);
  p.bar_base = p.smem_base + stages * p.stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(p.full_bar(s), gather ? 1 + 32 * GATHER_WARPS : 1);  // TMA transaction (+ one arrival per gather thread)
      mbar_init(p.empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(p.tfull_bar(a), 1);
      mbar_init(p.tempty_bar(a), epi_warps);  // one arrival per epilogue warp
    }
    for (int i = 0; i < QUEUE_DEPTH; ++i) {
      mbar_init(p.qfull_bar(i), 1);
      mbar_init(p.qempty_bar(i), 1 + epi_warps + (gather ? GATHER_WARPS : 0));  // MMA thread + consumer warps
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(p.tmem_slot(), TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(p.tmem_slot()));
  return tmem_base;
}

// Called by all threads at kernel end.
__device__ __forceinline__ void pipe_teardown(uint32_t tmem_base) {
  fence_before_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

__device__ __forceinline__ void queue_push(const Pipe& p, uint32_t n, int tag0, int tag1, int k_blocks, uint32_t idesc) {
  const int slot = n % QUEUE_DEPTH;
  mbar_wait(p.qempty_bar(slot), ((n / QUEUE_DEPTH) & 1u) ^ 1u);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p.queue(slot)), "r"(tag0), "r"(tag1), "r"(k_blocks),
               "r"(idesc)
               : "memory");
  mbar_arrive(p.qfull_bar(slot));  // release: the entry is visible to whoever observes the phase flip
}
// consumer side; `release` tells whether this thread performs the slot's arrival (MMA thread / lane 0 of a warp)
__device__ __forceinline__ void queue_pop(const Pipe& p, uint32_t n, bool release, int& tag0, int& tag1, int& k_blocks,
                                          uint32_t& idesc) {
  const int slot = n % QUEUE_DEPTH;
  mbar_wait(p.qfull_bar(slot), (n / QUEUE_DEPTH) & 1u);
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(tag0), "=r"(tag1), "=r"(k_blocks), "=r"(idesc)
               : "r"(p.queue(slot))
               : "memory");
  if (release) mbar_arrive(p.qempty_bar(slot));
}

// Sched: bool next(UnitDesc&) -- called by the producer thread only.
template <class Sched>
__device__ __forceinline__ void tc_producer_loop(Sched sched, const Pipe& p) {
  int stage = 0;
  uint32_t phase = 0, n = 0;
  UnitDesc u;
  u.k_block0 = 0, u.a_bytes = A_TILE_BYTES, u.a_kb_rows = 0, u.b_bytes = p.b_tile_bytes, u.b_kb_rows = 0;
  u.a_hi_bulk = u.a_lo_bulk = nullptr;
  const uint64_t stream_policy = l2_policy_evict_first();
  while (sched.next(u)) {
    queue_push(p, n++, u.tag0, u.tag1, u.k_blocks, u.idesc);
    for (int kb = u.k_block0; kb < u.k_block0 + u.k_blocks; ++kb) {
      mbar_wait(p.empty_bar(stage), phase ^ 1u);
      const uint32_t sa = p.smem_base + stage * p.stage_bytes;
      if (p.gather) {  // (terms == 3) the other operand's tile is written by the gather warps
        // the corpus streams through L2 once (evict_first) so that it does not push out the query rows the gather
        // warps keep re-reading (evict_last)
        if (p.gather == 1) {
          mbar_expect_tx(p.full_bar(stage), 2u * (uint32_t)u.b_bytes);
          const int c0 = u.b_kb_rows ? 0 : kb * BLOCK_K, c1 = u.b_row + kb * u.b_kb_rows;
          tma_load_2d_hint(sa + 2 * A_TILE_BYTES, u.b_hi, p.full_bar(stage), c0, c1, stream_policy);
          tma_load_2d_hint(sa + 2 * A_TILE_BYTES + p.b_tile_bytes, u.b_lo, p.full_bar(stage), c0, c1, stream_policy);
        } else {
          mbar_expect_tx(p.full_bar(stage), 2u * (uint32_t)u.a_bytes);
          const int c0 = u.a_kb_rows ? 0 : kb * BLOCK_K, c1 = u.a_row + kb * u.a_kb_rows;
          if (u.a_hi_bulk) {
            const size_t off = (size_t)c1 * BLOCK_K;  // elements: row c1 of the (k-blocks * rows, 32) image
            bulk_load_hint(sa, u.a_hi_bulk + off, (uint32_t)u.a_bytes, p.full_bar(stage), stream_policy);
            bulk_load_hint(sa + A_TILE_BYTES, u.a_lo_bulk + off, (uint32_t)u.a_bytes, p.full_bar(stage), stream_policy);
          } else {
            tma_load_2d_hint(sa, u.a_hi, p.full_bar(stage), c0, c1, stream_policy);
            tma_load_2d_hint(sa + A_TILE_BYTES, u.a_lo, p.full_bar(stage), c0, c1, stream_policy);
          }
        }
        if (++stage == p.stages) stage = 0, phase ^= 1u;
        continue;
      }
      mbar_expect_tx(p.full_bar(stage), (p.probe & 2) ? (uint32_t)A_TILE_BYTES : (uint32_t)p.stage_bytes);
      if (p.terms == 3) {
        const int c0 = u.a_kb_rows ? 0 : kb * BLOCK_K, c1 = u.a_row + kb * u.a_kb_rows;
        tma_load_2d(sa, u.a_hi, p.full_bar(stage), c0, c1);
        tma_load_2d(sa + A_TILE_BYTES, u.a_lo, p.full_bar(stage), c0, c1);
        const int d0 = u.b_kb_rows ? 0 : kb * BLOCK_K, d1 = u.b_row + kb * u.b_kb_rows;
        tma_load_2d(sa + 2 * A_TILE_BYTES, u.b_hi, p.full_bar(stage), d0, d1);
        tma_load_2d(sa + 2 * A_TILE_BYTES + p.b_tile_bytes, u.b_lo, p.full_bar(stage), d0, d1);
      } else if (p.probe & 2) {
        tma_load_2d(sa, u.a_hi, p.full_bar(stage), kb * BLOCK_K, u.a_row);
      } else {
        tma_load_2d(sa, u.a_hi, p.full_bar(stage), kb * BLOCK_K, u.a_row);
        tma_load_2d(sa + A_TILE_BYTES, u.b_hi, p.full_bar(stage), kb * BLOCK_K, u.b_row);
      }
      if (++stage == p.stages) stage = 0, phase ^= 1u;
    }
  }
  queue_push(p, n, 0, 0, 0, 0u);  // k_blocks == 0: end of work
}

// Producer of the grouped kernels whose A (gather_a != 0) or B operand is a set of rows picked by an index list
// (the queries of a video's inverted list): the WHOLE warp 0 runs it.  Lane 0 schedules, waits for the stage and
// loads the contiguous operand exactly like tc_producer_loop; every lane then fetches four rows of the other operand
// with one TMA gather4 per half straight out of the un-gathered query array (which stays L2-resident), instead of a
// separate kernel materialising the gathered rows in HBM for a box load.  u.a_row / u.b_row of the gathered operand
// is the first list entry, u.g_count the number of entries in the tile (the rest of the tile reads row 0: ignored
// by the epilogues); its tensor maps must be encoded with box rows = 1.  rows = tile height of the gathered operand.
template <class Sched>
__device__ __forceinline__ void tc_producer_loop_gather(Sched sched, const Pipe& p, int lane,
                                                        const int* __restrict__ gather_idx, int gather_a, int rows) {
  int stage = 0;
  uint32_t phase = 0, n = 0;
  UnitDesc u;
  u.k_block0 = 0, u.g_count = 0, u.a_kb_rows = 0, u.b_kb_rows = 0;
  const uint32_t a_lo_off = A_TILE_BYTES, b_hi_off = 2 * A_TILE_BYTES, b_lo_off = 2 * A_TILE_BYTES + p.b_tile_bytes;
  for (;;) {
    int has = 0;
    if (lane == 0) has = sched.next(u) ? 1 : 0;
    if (!__shfl_sync(0xffffffffu, has, 0)) break;
    const int g0 = __shfl_sync(0xffffffffu, gather_a ? u.a_row : u.b_row, 0);
    const int g_count = __shfl_sync(0xffffffffu, u.g_count, 0);
    const int kb0 = __shfl_sync(0xffffffffu, u.k_block0, 0), kbn = __shfl_sync(0xffffffffu, u.k_blocks, 0);
    const CUtensorMap* g_hi = reinterpret_cast<const CUtensorMap*>(
        __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(gather_a ? u.a_hi : u.b_hi), 0));
    const CUtensorMap* g_lo = reinterpret_cast<const CUtensorMap*>(
        __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(gather_a ? u.a_lo : u.b_lo), 0));
    int r[4] = {0, 0, 0, 0};
    const bool mine = 4 * lane < rows;
    if (mine) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = 4 * lane + i;
        if (e < g_count) r[i] = max(0, __ldg(gather_idx + g0 + e));
      }
    }
    if (lane == 0) queue_push(p, n, u.tag0, u.tag1, u.k_blocks, u.idesc);
    ++n;
    for (int kb = kb0; kb < kb0 + kbn; ++kb) {
      const uint32_t sa = p.smem_base + stage * p.stage_bytes;
      if (lane == 0) {
        mbar_wait(p.empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(p.full_bar(stage), (uint32_t)p.stage_bytes);
        if (gather_a) {
          const int d0 = u.b_kb_rows ? 0 : kb * BLOCK_K, d1 = u.b_row + kb * u.b_kb_rows;
          tma_load_2d(sa + b_hi_off, u.b_hi, p.full_bar(stage), d0, d1);
          tma_load_2d(sa + b_lo_off, u.b_lo, p.full_bar(stage), d0, d1);
        } else {
          const int c0 = u.a_kb_rows ? 0 : kb * BLOCK_K, c1 = u.a_row + kb * u.a_kb_rows;
          tma_load_2d(sa, u.a_hi, p.full_bar(stage), c0, c1);
          tma_load_2d(sa + a_lo_off, u.a_lo, p.full_bar(stage), c0, c1);
        }
      }
      __syncwarp();
      if (mine) {
        const uint32_t dst = sa + (gather_a ? 0u : b_hi_off) + (uint32_t)lane * (4u * BLOCK_K * 2u);
        tma_gather4_2d(dst, g_hi, p.full_bar(stage), kb * BLOCK_K, r[0], r[1], r[2], r[3]);
        tma_gather4_2d(dst + (gather_a ? a_lo_off : (uint32_t)p.b_tile_bytes), g_lo, p.full_bar(stage), kb * BLOCK_K, r[0],
                       r[1], r[2], r[3]);
      }
      if (++stage == p.stages) stage = 0, phase ^= 1u;
    }
  }
  if (lane == 0) queue_push(p, n, 0, 0, 0, 0u);  // k_blocks == 0: end of work
}

// Gather warps (GATHER_WARPS warps after the epilogue warps; `t` = thread index among them, 0 .. 32 * GATHER_WARPS - 1):
// fill the A (p.gather == 1) or B (== 2) tile of every stage with rows picked by an index list -- the queries of a
// video's inverted list -- straight from the un-gathered (rows, ld) 16-bit arrays, which stay L2-resident, with
// 16-byte cp.async copies into the 64B-swizzled K-major layout TMA would have produced (piece c of row r at
// r * 64 + ((c ^ ((r >> 1) & 3)) * 16)).  Nothing is staged in registers and nothing waits for the data: as soon as a
// stage is free its copies are issued and cp.async.mbarrier.arrive.noinc makes each thread's arrival on the stage's
// full barrier happen when its copies have landed, so the gather runs ahead over all free stages.  The copies are
// generic-proxy writes: the MMA warp issues fence.proxy.async after the full-barrier wait (tc_mma_loop_warp).
// `info(tag0, tag1, e0, ne, src_hi, src_lo)` maps a unit to its first list entry, entry count and source arrays.
template <class Info>
__device__ __forceinline__ void tc_gather_loop(const Pipe& p, int t, const int* __restrict__ entry_q, int rows,
                                               int ld, Info info) {
  constexpr int NT_G = 32 * GATHER_WARPS, MAX_ITEMS = BLOCK_M * 8 / NT_G;  // 16-byte pieces per thread and stage
  const uint32_t tile_off = p.gather == 1 ? 0u : 2u * A_TILE_BYTES;
  const uint32_t lo_off = p.gather == 1 ? (uint32_t)A_TILE_BYTES : (uint32_t)p.b_tile_bytes;
  const int n_items = rows * 8;
  int stage = 0;
  uint32_t phase = 0;
  for (uint32_t unit = 0;; ++unit) {
    int tag0, tag1, k_blocks;
    uint32_t idesc;
    queue_pop(p, unit, false, tag0, tag1, k_blocks, idesc);
    __syncwarp();
    if ((t & 31) == 0) mbar_arrive(p.qempty_bar(unit % QUEUE_DEPTH));
    if (k_blocks <= 0) break;
    int e0, ne;
    const unsigned short *src_hi, *src_lo;
    info(tag0, tag1, e0, ne, src_hi, src_lo);
    const uint4* src[MAX_ITEMS];
    uint32_t dst[MAX_ITEMS];
#pragma unroll
    for (int j = 0; j < MAX_ITEMS; ++j) {
      const int i = t + j * NT_G, row = i >> 3, piece = i & 7;
      src[j] = nullptr, dst[j] = 0;
      if (i < n_items && row < ne) {  // tile rows beyond the list keep stale (finite) data: the epilogues ignore them
        const int q = max(0, __ldg(entry_q + e0 + row));
        src[j] = reinterpret_cast<const uint4*>((piece < 4 ? src_hi : src_lo) + (long long)q * ld) + (piece & 3);
        dst[j] = tile_off + (piece < 4 ? 0u : lo_off) + (uint32_t)row * 64u +
                 ((uint32_t)((piece & 3) ^ ((row >> 1) & 3)) * 16u);
      }
    }
    for (int kb = 0; kb < k_blocks; ++kb) {
      mbar_wait(p.empty_bar(stage), phase ^ 1u);
      const uint32_t sa = p.smem_base + stage * p.stage_bytes;
#pragma unroll
      for (int j = 0; j < MAX_ITEMS; ++j) {
        if (src[j])
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + dst[j]),
                       "l"(src[j] + kb * (BLOCK_K / 8))  // a k-block = 4 pieces of 8 elements
                       : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(p.full_bar(stage)) : "memory");
      if (++stage == p.stages) stage = 0, phase ^= 1u;
    }
  }
}

// MMA loop, run by ALL 32 lanes of the MMA warp: every lane follows the barriers, one elected lane issues.  When a
// single thread runs such a loop inside `if (lane == 0)`, the compiler cannot tell that only one thread is active and
// wraps every tcgen05.mma / commit in an elect-and-retry loop (ELECT, PLOP3, BRA.U.ANY around each UTCHMMA, ncu source
// view) -- comparable to the tensor time of an MMA of the grouped kernels, whose N is only the 32..80 rows of one
// video's list.  With the warp converged and elect.sync in the source the six MMAs of a k-block issue back to back.
__device__ __forceinline__ void tc_mma_loop_warp(const Pipe& p, uint32_t tmem_base) {
  const bool lane0 = (threadIdx.x & 31) == 0;
  int stage = 0;
  uint32_t phase = 0;
  for (uint32_t unit = 0;; ++unit) {
    int tag0, tag1, k_blocks;
    uint32_t idesc;
    queue_pop(p, unit, false, tag0, tag1, k_blocks, idesc);
    __syncwarp();  // every lane has read the entry before the slot is handed back
    if (lane0) mbar_arrive(p.qempty_bar(unit % QUEUE_DEPTH));
    if (k_blocks <= 0) break;
    const uint32_t acc = unit & 1u, use = unit >> 1;
    mbar_wait(p.tempty_bar(acc), (use & 1u) ^ 1u);  // the epilogue has drained this accumulator
    fence_after_sync();
    const uint32_t tmem_acc = tmem_base + acc * ACC_COLS;
    for (int kb = 0; kb < k_blocks; ++kb) {
      mbar_wait(p.full_bar(stage), phase);
      if (p.gather) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async-written tile -> async proxy
      fence_after_sync();
      __syncwarp();
      const uint32_t sa = p.smem_base + stage * p.stage_bytes;
      if (elect_one()) {
        if (p.terms == 3) {
          const uint64_t a_hi = smem_desc_kmajor<SWIZZLE_BYTES>(sa);
          const uint64_t a_lo = smem_desc_kmajor<SWIZZLE_BYTES>(sa + A_TILE_BYTES);
          const uint64_t b_hi = smem_desc_kmajor<SWIZZLE_BYTES>(sa + 2 * A_TILE_BYTES);
          const uint64_t b_lo = smem_desc_kmajor<SWIZZLE_BYTES>(sa + 2 * A_TILE_BYTES + p.b_tile_bytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t off = (uint64_t)(k * UMMA_K * 2 >> 4);  // advance 32 B inside the swizzle row
            umma_f16(tmem_acc, a_hi + off, b_lo + off, idesc, (kb | k) != 0);
            umma_f16(tmem_acc, a_lo + off, b_hi + off, idesc, 1u);
            umma_f16(tmem_acc, a_hi + off, b_hi + off, idesc, 1u);
          }
        } else {
          const uint64_t a_hi = smem_desc_kmajor<SWIZZLE_BYTES>(sa);
          const uint64_t b_hi = smem_desc_kmajor<SWIZZLE_BYTES>(sa + A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t off = (uint64_t)(k * UMMA_K * 2 >> 4);
            umma_f16(tmem_acc, a_hi + off, b_hi + off, idesc, (kb | k) != 0);
          }
        }
        umma_commit(p.empty_bar(stage));  // the smem stage is reusable once these MMAs have read it
        if (kb == k_blocks - 1) umma_commit(p.tfull_bar(acc));  // accumulator complete
      }
      if (++stage == p.stages) stage = 0, phase ^= 1u;
    }
  }
}

// ---- epilogue side (warps 2..5); `unit` counts the units this CTA has consumed -----------------------------
// Next unit of this CTA: false at end of work.  tag0 / tag1 are the scheduler's payload.
__device__ __forceinline__ bool epi_next(const Pipe& p, uint32_t unit, int& tag0, int& tag1) {
  int k_blocks;
  uint32_t idesc;
  queue_pop(p, unit, false, tag0, tag1, k_blocks, idesc);
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(p.qempty_bar(unit % QUEUE_DEPTH));
  return k_blocks > 0;
}
// Wait for the unit's accumulator; returns its TMEM address for this warp's lane quadrant.
__device__ __forceinline__ uint32_t epi_wait(const Pipe& p, uint32_t unit, uint32_t tmem_base) {
  const uint32_t acc = unit & 1u, use = unit >> 1;
  mbar_wait(p.tfull_bar(acc), use & 1u);
  fence_after_sync();
  const uint32_t quad = (threadIdx.x >> 5) & 3u;  // TMEM lane quadrant this warp may access
  return tmem_base + acc * ACC_COLS + ((quad * 32u) << 16);
}
__device__ __forceinline__ void epi_release(const Pipe& p, uint32_t unit) {  // one arrival per epilogue warp
  fence_before_sync();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(p.tempty_bar(unit & 1u));
}

}  // namespace tc
