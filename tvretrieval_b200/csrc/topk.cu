// Exact top-k selection kernels (HBM/L2-bound, one CTA per query row):
//   xmlb_topk_rows : exp(alpha * q2c) + torch.topk(k=max_n_videos)             -- reference inference.py:317,347-348
//                    (also the k-way merge of per-GPU candidate lists, see sharding.py)
//   xmlb_span_topk : einsum("qvm,qv,qvn->qvmn") * band mask, flatten, full sort, keep max_before_nms
//                                                                                 -- reference inference.py:365-386
//                    and the SVMR variant (outer product of one video)           -- inference.py:215-224,
//                                                                                    utils/tensor_utils.py:133-141
// The reference materialises and fully sorts 100*L*L floats per query; here only the ~14 in-band cells per
// (video, start clip) are evaluated, on the fly, and the k best are found with a 3-pass (12/12/8 bit)
// radix select over the score bits, ties resolved by flat index (second radix select, only when needed),
// then a bitonic sort of the k winners.  Ranking is canonical: score descending, index ascending
// (tie_desc=0) or descending (tie_desc=1, numpy's reversed argsort used by the SVMR path).
#include "common.cuh"
#include "xmlb200.h"

namespace {

constexpr int NT = 512;
constexpr int HIST = 4096;
constexpr int MAX_K = 1024;

constexpr int LIST_CAP = 3072;  // survivors of the span pre-filter kept in shared memory
constexpr int SORT_CAP = 2048;  // rows up to this length are sorted whole (HIST ints = SORT_CAP 64-bit entries)

// diagnostics (tools/gpu/kernel_bench.py): [0] span_topk rows whose survivor list overflowed (slow full selection),
// [1] sum of survivor-list lengths, [2] span_topk rows
__device__ unsigned long long g_debug_counters[4];

struct SelSmem {
  int hist[HIST];
  int warp_tot[NT / 32];
  int sel_bin, sel_above, sel_count, total;
  int count, count_eq2, n_list;
  unsigned long long buf[MAX_K];
};

__device__ __forceinline__ unsigned int float_key(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Block-wide: find the largest bin b with sum_{b' >= b} hist[b'] >= want.  Results in sm.sel_*; sm.total = sum of all.
__device__ void find_bin(SelSmem& sm, int nbins, int want) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int bpt = nbins >= NT ? nbins / NT : 1;
  const int b0 = t * bpt;
  int tsum = 0;
  if (b0 < nbins)
    for (int i = 0; i < bpt; ++i) tsum += sm.hist[b0 + i];
  int inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) sm.warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = lane < NT / 32 ? sm.warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane < NT / 32) sm.warp_tot[lane] = x;
    if (lane == NT / 32 - 1) sm.total = x;
  }
  __syncthreads();
  const int total = sm.total;
  const int prefix_incl = inc + (w ? sm.warp_tot[w - 1] : 0);
  const int suffix_after = total - prefix_incl;  // sum over threads > t
  if (b0 < nbins && suffix_after < want && want <= suffix_after + tsum) {
    int acc = suffix_after;
    for (int b = b0 + bpt - 1; b >= b0; --b) {
      const int h = sm.hist[b];
      if (acc + h >= want) {
        sm.sel_bin = b, sm.sel_above = acc, sm.sel_count = h;
        break;
      }
      acc += h;
    }
  }
  __syncthreads();
}

// Radix select of the `want`-th largest 32-bit key among the candidates for which sel(key, idk, k) is true.
// Returns false if fewer than `want` candidates exist (then `n_total` holds their number).
template <class Gen, class Sel>
__device__ bool radix_select(const Gen& gen, const Sel& sel, SelSmem& sm, int want, unsigned int& thr,
                             int& n_equal, int& n_needed_equal, int& n_total) {
  unsigned int prefix = 0, pmask = 0;
  int remaining = want;
  const int shifts[3] = {20, 8, 0};
  const int nbins[3] = {4096, 4096, 256};
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = threadIdx.x; i < nbins[pass]; i += NT) sm.hist[i] = 0;
    __syncthreads();
    const int shift = shifts[pass];
    const unsigned int bm = nbins[pass] - 1;
    gen.for_each([&](unsigned int key, unsigned int idk) {
      unsigned int k;
      if (sel(key, idk, k) && (k & pmask) == prefix) atomicAdd(&sm.hist[(k >> shift) & bm], 1);
    });
    __syncthreads();
    find_bin(sm, nbins[pass], remaining);
    if (pass == 0) {
      n_total = sm.total;
      if (sm.total < want) return false;
    }
    remaining -= sm.sel_above;
    prefix |= (unsigned int)sm.sel_bin << shift;
    pmask |= bm << shift;
    n_equal = sm.sel_count;
    __syncthreads();
  }
  thr = prefix;
  n_needed_equal = remaining;
  return true;
}

__device__ void bitonic_sort_desc(unsigned long long* buf, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool desc = (i & k) == 0;
          const unsigned long long a = buf[i], b = buf[ixj];
          if (desc ? (a < b) : (a > b)) buf[i] = b, buf[ixj] = a;
        }
      }
      __syncthreads();
    }
  }
}

// Exact top-k by (key desc, idkey desc).  On return sm.buf[0..n_out) is sorted; returns n_out (= k unless
// fewer than k candidates exist).
template <class Gen>
__device__ int block_topk(const Gen& gen, SelSmem& sm, int k) {
  unsigned int thr = 0, thr2 = 0;
  int n_equal = 0, need_equal = 0, n_total = 0;
  bool use_thr2 = false;
  int need_eq2 = 0;
  const bool full = radix_select(gen, [](unsigned int key, unsigned int, unsigned int& out) { out = key; return true; },
                                 sm, k, thr, n_equal, need_equal, n_total);
  if (full && n_equal > need_equal) {  // exact ties straddle the cut: keep the preferred ids among them
    int ne2, nt2;
    const unsigned int t1 = thr;
    radix_select(gen, [t1](unsigned int key, unsigned int idk, unsigned int& out) { out = idk; return key == t1; },
                 sm, need_equal, thr2, ne2, need_eq2, nt2);
    use_thr2 = true;  // ids are normally unique (need_eq2 == 1); duplicates are still counted exactly
  }
  int pow2 = 1;
  while (pow2 < k) pow2 <<= 1;
  for (int i = threadIdx.x; i < pow2; i += NT) sm.buf[i] = 0ull;
  if (threadIdx.x == 0) sm.count = 0, sm.count_eq2 = 0;
  __syncthreads();
  gen.for_each([&](unsigned int key, unsigned int idk) {
    bool take = !full || key > thr || (key == thr && (!use_thr2 || idk > thr2));
    if (!take && full && use_thr2 && key == thr && idk == thr2) take = atomicAdd(&sm.count_eq2, 1) < need_eq2;
    if (take) {
      const int pos = atomicAdd(&sm.count, 1);
      if (pos < MAX_K) sm.buf[pos] = ((unsigned long long)key << 32) | idk;
    }
  });
  __syncthreads();
  const int n_out = min(sm.count, k);
  bitonic_sort_desc(sm.buf, pow2);
  return n_out;
}

// ------------------------------------------------------------------------------------------------------
// Where a kernel writes its ranked lists.  mode 0: the caller's local buffers.  Modes 1 / 2 write through peer
// pointers into the symmetric workspaces of the other GPUs of a video-sharded search (NVLink stores issued by the
// kernel that produced the list -- no collective call, no staging copy; tvretrieval_b200/sharding.py):
//   mode 1 "to owner": row r (a query of the block) belongs to rank o = r / per; it lands in rank o's buffer at row
//                      self_rank * per + (r - o * per)  -> each owner ends up with [source rank][owned query] lists
//   mode 2 "to all"  : row r (an owned query) lands in EVERY rank's buffer at row self_rank * per + r
struct PeerOut {
  long long idx_ptr[8];
  long long val_ptr[8];
  int world, mode, per, self_rank;
};

struct OutSpec {
  int first, count;    // only ranks [first, first + count) of each list are written, at positions 0 .. count - 1
  int pad_to;          // positions [count, pad_to) are filled with (pad_idx, pad_val); row pitch = max(count, pad_to)
  int pad_idx;
  float pad_val;
};

__device__ __forceinline__ void out_store(const PeerOut& po, int* __restrict__ out_idx, float* __restrict__ out_val,
                                          long long row, int pitch, int pos, int idx, float val) {
  if (po.mode == 0) {
    if (out_idx) out_idx[row * pitch + pos] = idx;
    if (out_val) out_val[row * pitch + pos] = val;
  } else if (po.mode == 1) {
    const int o = (int)(row / po.per);
    const long long at = ((long long)po.self_rank * po.per + (row - (long long)o * po.per)) * pitch + pos;
    if (po.idx_ptr[o]) reinterpret_cast<int*>(po.idx_ptr[o])[at] = idx;
    if (po.val_ptr[o]) reinterpret_cast<float*>(po.val_ptr[o])[at] = val;
  } else {
    const long long at = ((long long)po.self_rank * po.per + row) * pitch + pos;
    for (int p = 0; p < po.world; ++p) {
      if (po.idx_ptr[p]) reinterpret_cast<int*>(po.idx_ptr[p])[at] = idx;
      if (po.val_ptr[p]) reinterpret_cast<float*>(po.val_ptr[p])[at] = val;
    }
  }
}

struct RowGen {
  const float* row;
  const int* ids;
  int n;
  float alpha;
  int apply_exp, tie_desc;
  template <class F>
  __device__ void for_each(F f) const {
    for (int c = threadIdx.x; c < n; c += NT) {
      const float x = __ldg(row + c);
      const float e = apply_exp ? expf(alpha * x) : x;
      const unsigned int id = ids ? (unsigned int)__ldg(ids + c) : (unsigned int)c;
      f(float_key(e), tie_desc ? id : ~id);
    }
  }
};

// seg_k > 0: the row is the concatenation of per-rank ranked lists stored as [source rank][row][seg_k] (the owner's
// side of a mode-1 exchange): column c lives at (c / seg_k) * seg_stride + r * seg_k + c % seg_k.  missing_neg: entries
// with a negative id are absent (they rank below every real entry and come out as (-1, 0)).
__global__ void __launch_bounds__(NT) topk_rows_kernel(const float* __restrict__ values, const int* __restrict__ ids,
                                                       int ids_shared, int n_cols, int k, float alpha, int apply_exp,
                                                       int tie_desc, const int* __restrict__ row_flags, int seg_k,
                                                       long long seg_stride, int missing_neg, OutSpec os,
                                                       int* __restrict__ out_idx, float* __restrict__ out_val,
                                                       const __grid_constant__ PeerOut po) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelSmem& sm = *reinterpret_cast<SelSmem*>(smem_raw);
  const long long r = blockIdx.x;
  if (row_flags && row_flags[r] == 0) return;  // restricted mode: only the flagged rows are (re)computed
  RowGen gen{values + r * n_cols, ids ? ids + (ids_shared ? 0 : r * n_cols) : nullptr, n_cols, alpha, apply_exp,
             tie_desc};
  // Short rows (candidate tables, merges of per-GPU lists): sort the whole row in shared memory -- the radix select
  // spends most of its time clearing and scanning 4096-bin histograms that such rows barely touch.
  unsigned long long* sorted = sm.buf;
  int n_out;
  if (n_cols <= SORT_CAP || seg_k > 0) {
    sorted = reinterpret_cast<unsigned long long*>(smem_raw);  // SORT_CAP entries over hist[] (same 16 KB)
    int pow2 = 32;
    while (pow2 < n_cols) pow2 <<= 1;
    for (int i = n_cols + threadIdx.x; i < pow2; i += NT) sorted[i] = 0ull;  // below every real (key, id) pair
    for (int c = threadIdx.x; c < n_cols; c += NT) {
      long long at = c;
      if (seg_k > 0) {
        const int sgm = c / seg_k;
        at = sgm * seg_stride + r * seg_k + (c - sgm * seg_k) - r * n_cols;  // relative to gen.row / gen.ids
      }
      const float x = __ldg(gen.row + at);
      float e = apply_exp ? expf(alpha * x) : x;
      unsigned int id = gen.ids ? (unsigned int)__ldg(gen.ids + at) : (unsigned int)c;
      if (missing_neg && (int)id < 0) e = -1.f, id = (unsigned int)(-1 - c);  // absent entry: distinct id, last rank
      sorted[c] = ((unsigned long long)float_key(e) << 32) | (tie_desc ? id : ~id);
    }
    __syncthreads();
    bitonic_sort_desc(sorted, pow2);
    n_out = min(n_cols, k);
  } else {
    n_out = block_topk(gen, sm, k);
  }
  const int pitch = max(os.count, os.pad_to);
  for (int i = threadIdx.x; i < pitch; i += NT) {
    int idx = os.pad_idx;
    float val = os.pad_val;
    if (i < os.count) {
      const int src = os.first + i;
      idx = -1, val = 0.f;
      if (src < n_out) {
        const unsigned long long e = sorted[src];
        const unsigned int idk = (unsigned int)(e & 0xffffffffu);
        idx = (int)(tie_desc ? idk : ~idk);
        val = key_float((unsigned int)(e >> 32));
        if (missing_neg && idx < 0) idx = -1, val = 0.f;
      }
    }
    out_store(po, out_idx, out_val, r, pitch, i, idx, val);
  }
}

// ------------------------------------------------------------------------------------------------------
// Candidate filter of the two-pass video retrieval: the scores of a row are approximations with a known error
// bound eps.  Every column whose exact score can be among the k best satisfies approx >= (k-th largest approx)
// - 2 eps: k columns have approx >= kth, hence exact >= kth - eps, hence the exact k-th largest is >= kth - eps,
// and a column with exact >= kth - eps has approx >= kth - 2 eps.
struct RawRowGen {
  const float* row;
  int n;
  template <class F>
  __device__ void for_each(F f) const {
    for (int c = threadIdx.x; c < n; c += NT) f(float_key(__ldg(row + c)), ~(unsigned int)c);
  }
};

__global__ void __launch_bounds__(NT) select_candidates_kernel(
    const float* __restrict__ values, const int* __restrict__ ids, int n_rows, int n_cols, int k,
    const float* __restrict__ err_a, const float* __restrict__ err_b, float err_scale, float err_const,
    const float* __restrict__ row_kth, int max_cand, int rows_per_group, int n_groups, int* __restrict__ cand_col,
    int* __restrict__ cand_id, float* __restrict__ cand_val, int* __restrict__ flag_ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelSmem& sm = *reinterpret_cast<SelSmem*>(smem_raw);
  const long long r = blockIdx.x;
  const float* row = values + r * n_cols;
  RawRowGen gen{row, n_cols};
  float kth;
  if (row_kth) {  // k-th largest approximate score over a LARGER set of columns (all shards of a sharded corpus)
    kth = __ldg(row_kth + r);
  } else {
    // Any LOWER bound of the k-th largest score keeps the candidate list a superset (it only lowers the cut), so the
    // three radix passes -- whose first histogram piles a whole row of similar cosines into a dozen exponent bins --
    // are replaced by: row maximum, then ONE histogram of (max - x) in HIST linear bins of width 2^-15 covering
    // [max - 1/8, max] (the top of the row spreads over hundreds of bins; scores further down are not counted), and
    // the lower edge of the bin that holds the k-th largest is the bound: at most 3.1e-5 below it, against a
    // window of 2 eps ~ 1e-3.  Rows with fewer than k scores inside the range fall back to the exact selection.
    constexpr float SCALE = 32768.f;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < n_cols; c += NT) mx = fmaxf(mx, __ldg(row + c));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) reinterpret_cast<float*>(sm.warp_tot)[threadIdx.x >> 5] = mx;
    for (int i = threadIdx.x; i < HIST; i += NT) sm.hist[i] = 0;
    __syncthreads();
    mx = reinterpret_cast<float*>(sm.warp_tot)[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) mx = fmaxf(mx, reinterpret_cast<float*>(sm.warp_tot)[w]);
    __syncthreads();  // warp_tot is reused by find_bin
    for (int c = threadIdx.x; c < n_cols; c += NT) {
      const float t = __fmul_rn(__fsub_rn(mx, __ldg(row + c)), SCALE);  // >= 0; NaN / inf fail the test below
      if (t < (float)HIST) atomicAdd(&sm.hist[HIST - 1 - (int)t], 1);
    }
    __syncthreads();
    find_bin(sm, HIST, k);
    if (sm.total >= k) {
      // every score counted in bins >= sel_bin satisfies x > mx - (HIST - sel_bin) / SCALE - 1.2e-7 (rounding of the
      // subtraction); the extra 4e-7 also covers the rounding of this expression
      kth = __fsub_rd(__fsub_rd(mx, __fdiv_ru((float)(HIST - sm.sel_bin), SCALE)), 4e-7f);
      __syncthreads();
    } else {
      __syncthreads();
      unsigned int thr = 0;
      int n_equal, need_equal, n_total;
      radix_select(gen, [](unsigned int key, unsigned int, unsigned int& out) { out = key; return true; }, sm, k, thr,
                   n_equal, need_equal, n_total);
      kth = key_float(thr);
    }
  }
  const float eps = fmaf(err_scale, __ldg(err_a + r) + (err_b ? __ldg(err_b + r) : 0.f), err_const);
  const float cut = __fmaf_rd(-2.f, eps, kth);
  if (threadIdx.x == 0) sm.count = 0;
  __syncthreads();
  int* cc = cand_col + r * max_cand;
  int* ci = cand_id + r * max_cand;
  float* cv = cand_val + r * max_cand;
  for (int c = threadIdx.x; c < n_cols; c += NT) {
    const float x = __ldg(row + c);
    if (x >= cut) {
      const int pos = atomicAdd(&sm.count, 1);
      if (pos < max_cand) cc[pos] = c, ci[pos] = ids ? __ldg(ids + c) : c, cv[pos] = x;
    }
  }
  __syncthreads();
  const int n = sm.count;
  for (int i = min(n, max_cand) + threadIdx.x; i < max_cand; i += NT) cc[i] = -1, ci[i] = 0x7fffffff, cv[i] = MASK_FILL;
  if (threadIdx.x == 0) {
    // flag_ws: [0] number of flagged groups, [1, 1+G) group flags, [1+G, 1+2G) list of flagged groups,
    // [1+2G, 1+2G+rows) row flags
    const bool overflow = n > max_cand;
    flag_ws[1 + 2 * n_groups + r] = overflow ? 1 : 0;
    if (overflow) {
      const int g = (int)(r / rows_per_group);
      if (atomicExch(flag_ws + 1 + g, 1) == 0) flag_ws[1 + n_groups + atomicAdd(flag_ws, 1)] = g;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// Cells of one query: score[j][m][n] = (st[j][m] * vr[j]) * ed[j][n] for min_l <= n - m < max_l.  The generator walks
// the rows (j, m) of a COMPACT list of slots (the valid ones of a sharded search, or those that can still reach the
// current lower bound), one row per thread, and skips a row whose best possible cell a * max_n ed[j][n] is below
// min_score -- with the bound of the pre-filter that removes ~90 % of the rows before their 14 cells are touched.
struct SpanGen {
  const float* st;  // [n_slots][L] start probabilities of this query
  const float* ed;
  const float* vr;  // [n_slots] video scores (null -> 1)
  const int* slots;     // shared memory: slot indices to visit, ascending
  const float* emax;    // shared memory: per slot, max over n of ed[j][n]
  int n_list, L, min_l, max_l, tie_desc;
  float min_score;  // cells below this known lower bound of the k-th best score are skipped (0: keep all positive)
  template <class F>
  __device__ void for_each(F f) const {
    const int rows = n_list * L;
    const int dj = NT / L, dm = NT - dj * L;  // row r = s * L + m advances by NT per iteration, without divisions
    int s = threadIdx.x / L, m = threadIdx.x - s * L;
    constexpr int U = 4;  // rows in flight per thread: their loads are issued before any of them is processed
    for (int r0 = threadIdx.x; r0 < rows; r0 += U * NT) {
      float a[U];
      int jj[U], mm[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a[u] = 0.f, jj[u] = 0, mm[u] = m;
        if (r0 + u * NT < rows) {
          jj[u] = slots[s];
          a[u] = __ldg(st + jj[u] * L + m);
        }
        s += dj, m += dm;
        if (m >= L) m -= L, ++s;
      }
      if (vr) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (a[u] > 0.f) a[u] = __fmul_rn(a[u], __ldg(vr + jj[u]));  // (st * vr) first, like torch.einsum
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = jj[u];
        if (!(a[u] > 0.f) || __fmul_ru(a[u], emax[j]) < min_score) continue;
        const int n_hi = min(mm[u] + max_l, L);
        const float* e = ed + j * L;
        const int row = j * L + mm[u];
        for (int n = mm[u] + min_l; n < n_hi; ++n) {
          const float sc = __fmul_rn(a[u], __ldg(e + n));
          if (sc > 0.f && sc >= min_score) {
            const unsigned int id = (unsigned int)(row * L + n);
            f(float_key(sc), tie_desc ? id : ~id);
          }
        }
      }
    }
  }

  // The same cells, in two phases: (1) every thread tests its rows and appends the surviving ones -- those whose best
  // possible cell can reach min_score -- to `row_list` (shared memory, `cap` entries); (2) the listed rows are shared
  // out EVENLY, one row (its ~14 cells) per thread at a time.  In for_each the cells of a surviving row are evaluated
  // by the thread that owns the row, and since the survivors are the first clips of a few short, well-ranked videos
  // they all belong to the same few warps: each of those ran ~300 dependent L1 / L2 round trips while the other warps
  // sat at the barrier that ends the pass (ncu: a third of the kernel's samples).  Rows that do not fit into the list
  // are evaluated in place as before.  Must be called by all NT threads; *row_count is zeroed by the caller.
  template <class F>
  __device__ void for_each_balanced(F f, int* row_list, int* row_count, int cap) const {
    const int rows = n_list * L;
    const int dj = NT / L, dm = NT - dj * L;
    int s = threadIdx.x / L, m = threadIdx.x - s * L;
    auto cells = [&](int j, int mm, float a) {
      const int n_hi = min(mm + max_l, L);
      const float* e = ed + j * L;
      const int row = j * L + mm;
      for (int n = mm + min_l; n < n_hi; ++n) {
        const float sc = __fmul_rn(a, __ldg(e + n));
        if (sc > 0.f && sc >= min_score) {
          const unsigned int id = (unsigned int)(row * L + n);
          f(float_key(sc), tie_desc ? id : ~id);
        }
      }
    };
    constexpr int U = 4;
    for (int r0 = threadIdx.x; r0 < rows; r0 += U * NT) {
      float a[U];
      int jj[U], mm[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a[u] = 0.f, jj[u] = 0, mm[u] = m;
        if (r0 + u * NT < rows) {
          jj[u] = slots[s];
          a[u] = __ldg(st + jj[u] * L + m);
        }
        s += dj, m += dm;
        if (m >= L) m -= L, ++s;
      }
      if (vr) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (a[u] > 0.f) a[u] = __fmul_rn(a[u], __ldg(vr + jj[u]));  // (st * vr) first, like torch.einsum
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!(a[u] > 0.f) || __fmul_ru(a[u], emax[jj[u]]) < min_score) continue;
        const int pos = atomicAdd(row_count, 1);
        if (pos < cap) row_list[pos] = jj[u] * L + mm[u];
        else cells(jj[u], mm[u], a[u]);
      }
    }
    __syncthreads();
    const int n_rows = min(*row_count, cap);
    for (int i = threadIdx.x; i < n_rows; i += NT) {
      const int r = row_list[i];
      const int j = r / L, mm_ = r - j * L;
      float a = __ldg(st + r);
      if (vr) a = __fmul_rn(a, __ldg(vr + j));
      cells(j, mm_, a);
    }
  }
};

// generator over a compact (key, id) list in shared memory
struct ListGen {
  const unsigned long long* list;
  int n;
  template <class F>
  __device__ void for_each(F f) const {
    for (int i = threadIdx.x; i < n; i += NT) {
      const unsigned long long e = list[i];
      f((unsigned int)(e >> 32), (unsigned int)(e & 0xffffffffu));
    }
  }
};

// cells with score exactly 0 (out of band, padded clips, underflow) rank after every positive cell, ordered by
// flat index like a stable sort of the reference's dense tensor would order them
__device__ void zero_fill(SelSmem& sm, int n_pos, int k, long long total_cells, int tie_desc, int* out_idx,
                          float* out_val) {
  __shared__ int filled_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) filled_s = n_pos;
  __syncthreads();
  for (long long base = 0; base < total_cells; base += NT) {
    const int filled = filled_s;
    if (filled >= k) break;
    const long long pos = base + t;
    bool is_free = pos < total_cells;
    const unsigned int cell = (unsigned int)(tie_desc ? total_cells - 1 - pos : pos);
    if (is_free) {
      const unsigned int idk = tie_desc ? cell : ~cell;
      for (int i = 0; i < n_pos; ++i)
        if ((unsigned int)(sm.buf[i] & 0xffffffffu) == idk) {
          is_free = false;
          break;
        }
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, is_free);
    if (lane == 0) sm.warp_tot[w] = __popc(bal);
    __syncthreads();
    int before = __popc(bal & ((1u << lane) - 1));
    int tot = 0;
    for (int i = 0; i < NT / 32; ++i) {
      if (i < w) before += sm.warp_tot[i];
      tot += sm.warp_tot[i];
    }
    if (is_free && filled + before < k) {
      out_idx[filled + before] = (int)cell;
      out_val[filled + before] = 0.f;
    }
    __syncthreads();
    if (t == 0) filled_s = filled + tot;
    __syncthreads();
  }
  __syncthreads();
  for (int i = filled_s + t; i < k; i += NT) out_idx[i] = -1, out_val[i] = 0.f;
}

constexpr int MAX_SPAN_SLOTS = 1024;

// Ascending compaction of the slots j < n_slots with keep(j) into list[] (warp 0; result count in *n_out).
template <class Keep>
__device__ void compact_slots(int n_slots, Keep keep, int* list, int* n_out) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int count = 0;
    for (int base = 0; base < n_slots; base += 32) {
      const int j = base + lane;
      const bool v = j < n_slots && keep(j);
      const unsigned int bal = __ballot_sync(0xffffffffu, v);
      if (v) list[count + __popc(bal & ((1u << lane) - 1u))] = j;
      count += __popc(bal);
    }
    if (lane == 0) *n_out = count;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT) span_topk_kernel(const float* __restrict__ st, const float* __restrict__ ed,
                                                       const float* __restrict__ vr,
                                                       const unsigned char* __restrict__ slot_valid, int n_slots,
                                                       int L, int min_l, int max_l, int k, int tie_desc,
                                                       int do_zero_fill, int* __restrict__ out_idx,
                                                       float* __restrict__ out_val,
                                                       const __grid_constant__ PeerOut po) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelSmem& sm = *reinterpret_cast<SelSmem*>(smem_raw);
  unsigned long long* list = reinterpret_cast<unsigned long long*>(smem_raw + sizeof(SelSmem));
  int* slots = reinterpret_cast<int*>(list + LIST_CAP);       // [n_slots] valid slots, ascending
  int* live = slots + n_slots;                                // [n_slots] slots that can reach the lower bound
  float* emax = reinterpret_cast<float*>(live + n_slots);     // [n_slots] max_n ed[j][n]
  float* amax = emax + n_slots;                               // [n_slots] vr[j] * max_m st[j][m]
  __shared__ int n_valid_s, n_live_s;
  const long long q = blockIdx.x;
  const float* st_q = st + q * n_slots * L;
  const float* ed_q = ed + q * n_slots * L;
  const float* vr_q = vr ? vr + q * n_slots : nullptr;
  const unsigned char* valid = slot_valid ? slot_valid + q * n_slots : nullptr;
  // per-slot maxima (four threads per slot, independent loads: this is the first touch of the query's rows) and the
  // compact list of valid slots
  for (int base = 0; base < n_slots; base += NT / 4) {
    const int j = base + (threadIdx.x >> 2), part = threadIdx.x & 3;
    float me = 0.f, ms = 0.f;
    if (j < n_slots && (!valid || valid[j])) {
      const float* e = ed_q + j * L;
      const float* a = st_q + j * L;
#pragma unroll 8
      for (int n = part; n < L; n += 4) {
        me = fmaxf(me, __ldg(e + n));
        ms = fmaxf(ms, __ldg(a + n));
      }
    }
    me = fmaxf(me, __shfl_xor_sync(0xffffffffu, me, 1)), ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, 1));
    me = fmaxf(me, __shfl_xor_sync(0xffffffffu, me, 2)), ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, 2));
    if (part == 0 && j < n_slots) emax[j] = me, amax[j] = vr_q ? __fmul_ru(ms, __ldg(vr_q + j)) : ms;
  }
  compact_slots(n_slots, [&](int j) { return !valid || valid[j] != 0; }, slots, &n_valid_s);
  const int n_valid = n_valid_s;
  SpanGen gen{st_q, ed_q, vr_q, slots, emax, n_valid, L, min_l, max_l, tie_desc, 0.f};
  // Pre-filter: the k-th best cell of the first few valid slots (the videos with the largest retrieval scores) is a
  // lower bound of the k-th best cell overall.  ONE pass over the slots that can still reach it then collects the few
  // cells at or above it into shared memory and the exact selection runs on that list (instead of 4 passes over ~180K
  // cells per query).
  // (a quarter of the valid slots, between 2 and 8: a rank of a sharded search sees only its share of the slots)
  const int sub_slots = min(8, max(2, n_valid / 4));
  int n_out = -1;
  if (n_valid >= 4) {
    SpanGen sub = gen;
    sub.n_list = sub_slots;
    // ONE 4096-bin pass over the top 12 key bits (sign, exponent, 3 mantissa bits): the lower edge of the bin that
    // holds the k-th largest cell of the subset is a lower bound of it (at most 12.5 % below), which is all the
    // collection pass needs -- two more passes would only tighten the bound, at 2/3 of this phase's cost.
    for (int i = threadIdx.x; i < HIST; i += NT) sm.hist[i] = 0;
    __syncthreads();
    sub.for_each([&](unsigned int key, unsigned int) { atomicAdd(&sm.hist[key >> 20], 1); });
    __syncthreads();
    find_bin(sm, HIST, k);
    const bool full = sm.total >= k;
    const unsigned int thr = (unsigned int)sm.sel_bin << 20;
    __syncthreads();
    if (full) {
      float bound = key_float(thr);
      for (int attempt = 0; attempt < 2 && n_out < 0; ++attempt) {
        // slots whose best conceivable cell amax * emax is below the bound drop out as a whole
        compact_slots(n_valid, [&](int i) { const int j = slots[i]; return __fmul_ru(amax[j], emax[j]) >= bound; },
                      live, &n_live_s);
        for (int i = threadIdx.x; i < n_live_s; i += NT) live[i] = slots[live[i]];
        if (threadIdx.x == 0) sm.n_list = 0, sm.count = 0;
        __syncthreads();
        SpanGen col = gen;
        col.slots = live, col.n_list = n_live_s, col.min_score = bound;
        // (the surviving-row list lives in the histogram array, which is idle during this pass)
        col.for_each_balanced([&](unsigned int key, unsigned int idk) {
          const int pos = atomicAdd(&sm.n_list, 1);
          if (pos < LIST_CAP) list[pos] = ((unsigned long long)key << 32) | idk;
        }, sm.hist, &sm.count, HIST);
        __syncthreads();
        const int n_list = sm.n_list;
        __syncthreads();
        if (threadIdx.x == 0 && attempt == 0) {
          atomicAdd(&g_debug_counters[1], (unsigned long long)n_list);
          atomicAdd(&g_debug_counters[2], 1ull);
        }
        if (n_list <= LIST_CAP) {
          n_out = block_topk(ListGen{list, n_list}, sm, k);
        } else {
          // Too many survivors: the LIST_CAP cells already collected are real cells, so their k-th largest is a valid
          // -- and tighter -- lower bound of the k-th best overall; collect once more with it.
          unsigned int t2 = 0;
          int ne, nn, nt;
          radix_select(ListGen{list, LIST_CAP},
                       [](unsigned int key, unsigned int, unsigned int& out) { out = key; return true; }, sm, k, t2, ne,
                       nn, nt);
          __syncthreads();
          bound = key_float(t2);
          if (threadIdx.x == 0 && attempt == 1) atomicAdd(&g_debug_counters[0], 1ull);
        }
      }
      if (n_out < 0) gen.min_score = bound;  // the slow full selection still skips what the bound excludes
    }
  }
  if (n_out < 0) n_out = block_topk(gen, sm, k);
  if (po.mode != 0) {  // sharded search: this rank's list goes straight to the rank that owns the query
    for (int i = threadIdx.x; i < k; i += NT) {
      int idx = -1;
      float val = 0.f;
      if (i < n_out) {
        const unsigned long long e = sm.buf[i];
        const unsigned int idk = (unsigned int)(e & 0xffffffffu);
        idx = (int)(tie_desc ? idk : ~idk), val = key_float((unsigned int)(e >> 32));
      }
      out_store(po, nullptr, nullptr, q, k, i, idx, val);
    }
    return;
  }
  int* oi = out_idx + q * k;
  float* ov = out_val + q * k;
  for (int i = threadIdx.x; i < k; i += NT) {
    if (i < n_out) {
      const unsigned long long e = sm.buf[i];
      const unsigned int idk = (unsigned int)(e & 0xffffffffu);
      oi[i] = (int)(tie_desc ? idk : ~idk);
      ov[i] = key_float((unsigned int)(e >> 32));
    } else if (!do_zero_fill) {
      oi[i] = -1;
      ov[i] = 0.f;
    }
  }
  if (n_out < k && do_zero_fill) {
    __syncthreads();
    zero_fill(sm, n_out, k, (long long)n_slots * L * L, tie_desc, oi, ov);
  }
}

// Zero fill for an already ranked list (multi-GPU merge): the first n_pos[r] entries of a row are the positive
// cells, the rest is replaced by the zero-score cells in canonical flat-index order.
__global__ void __launch_bounds__(NT) span_zero_fill_kernel(int* __restrict__ idx, float* __restrict__ val, int k,
                                                            long long total_cells, int tie_desc,
                                                            const __grid_constant__ PeerOut po) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelSmem& sm = *reinterpret_cast<SelSmem*>(smem_raw);
  __shared__ int n_pos_s;
  int* oi = idx + (long long)blockIdx.x * k;
  float* ov = val + (long long)blockIdx.x * k;
  if (threadIdx.x == 0) n_pos_s = k;
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += NT) {
    const bool pos = oi[i] >= 0 && ov[i] > 0.f;
    if (!pos) atomicMin(&n_pos_s, i);
    const unsigned int id = (unsigned int)oi[i];
    sm.buf[i] = tie_desc ? id : ~id;
  }
  __syncthreads();
  const int n_pos = n_pos_s;
  if (n_pos < k) zero_fill(sm, n_pos, k, total_cells, tie_desc, oi, ov);
  if (po.mode != 0) {  // publish the completed row to the other ranks (zero_fill's writes are this block's own)
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += NT) out_store(po, nullptr, nullptr, blockIdx.x, k, i, oi[i], ov[i]);
  }
}

}  // namespace

extern "C" int xmlb_debug_counters(long long* out4, int reset) {
  unsigned long long h[4];
  XMLB_CUDA(cudaMemcpyFromSymbol(h, g_debug_counters, sizeof(h)));
  for (int i = 0; i < 4; ++i) out4[i] = (long long)h[i];
  if (reset) {
    unsigned long long z[4] = {0, 0, 0, 0};
    XMLB_CUDA(cudaMemcpyToSymbol(g_debug_counters, z, sizeof(z)));
  }
  return XMLB_OK;
}

static int check_topk_args(const char* who, int k) {
  XMLB_REQUIRE(k >= 1 && k <= MAX_K, "%s: k must be in [1, %d]", who, MAX_K);
  return XMLB_OK;
}

extern "C" int xmlb_select_candidates(const float* approx, const int* ids, int n_rows, int n_cols, int k,
                                      const float* row_err_a, const float* row_err_b, float err_scale,
                                      float err_const, const float* row_kth, int max_cand, int rows_per_group,
                                      int* cand_col, int* cand_id, float* cand_val, int* flag_ws, void* stream) {
  XMLB_REQUIRE(approx && row_err_a && cand_col && cand_id && cand_val && flag_ws,
               "xmlb_select_candidates: null pointer");
  if (int rc = check_topk_args("xmlb_select_candidates", k)) return rc;
  XMLB_REQUIRE(max_cand >= 1 && (row_kth || (n_cols >= k && max_cand >= k)),
               "xmlb_select_candidates: need k <= max_cand and k <= n_cols (unless row_kth is given)");
  XMLB_REQUIRE(rows_per_group >= 1 && err_scale >= 0.f && err_const >= 0.f, "xmlb_select_candidates: bad argument");
  const int n_groups = ceil_div(n_rows, rows_per_group);
  XMLB_CUDA(cudaMemsetAsync(flag_ws, 0, sizeof(int) * (size_t)(1 + 2 * n_groups), (cudaStream_t)stream));
  if (n_rows == 0) return XMLB_OK;
  XMLB_CUDA(cudaFuncSetAttribute(select_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(SelSmem)));
  select_candidates_kernel<<<n_rows, NT, sizeof(SelSmem), (cudaStream_t)stream>>>(
      approx, ids, n_rows, n_cols, k, row_err_a, row_err_b, err_scale, err_const, row_kth, max_cand, rows_per_group,
      n_groups, cand_col, cand_id, cand_val, flag_ws);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

static int make_peer_out(PeerOut& po, const long long* peer_idx, const long long* peer_val, int world, int peer_mode,
                         int per, int self_rank, const char* who) {
  po = PeerOut{};
  po.mode = peer_mode;
  if (peer_mode == 0) return XMLB_OK;
  XMLB_REQUIRE(peer_mode == 1 || peer_mode == 2, "%s: peer_mode must be 0, 1 or 2", who);
  XMLB_REQUIRE(world >= 1 && world <= 8 && per >= 1 && self_rank >= 0 && self_rank < world && (peer_idx || peer_val),
               "%s: bad peer arguments (world <= 8)", who);
  for (int p = 0; p < world; ++p) {
    po.idx_ptr[p] = peer_idx ? peer_idx[p] : 0;
    po.val_ptr[p] = peer_val ? peer_val[p] : 0;
  }
  po.world = world, po.per = per, po.self_rank = self_rank;
  return XMLB_OK;
}

extern "C" int xmlb_topk_rows_ex(const float* values, const int* ids, int ids_shared, int n_rows, int n_cols, int k,
                                 float alpha, int apply_exp, int tie_desc, const int* row_flags, int seg_k,
                                 long long seg_stride, int missing_neg, int out_first, int out_count, int pad_to,
                                 int pad_idx, float pad_val, int* out_idx, float* out_val, const long long* peer_idx,
                                 const long long* peer_val, int world, int peer_mode, int per, int self_rank,
                                 void* stream) {
  XMLB_REQUIRE(values && (peer_mode != 0 || out_idx || out_val), "xmlb_topk_rows: null pointer");
  if (int rc = check_topk_args("xmlb_topk_rows", k)) return rc;
  XMLB_REQUIRE(n_cols >= k, "xmlb_topk_rows: selected index k out of range (k=%d > %d columns)", k, n_cols);
  XMLB_REQUIRE(out_first >= 0 && out_count >= 1 && out_first + out_count <= k && (pad_to == 0 || pad_to >= out_count),
               "xmlb_topk_rows: bad output slice");
  XMLB_REQUIRE(seg_k == 0 || (n_cols % seg_k == 0 && n_cols <= SORT_CAP),
               "xmlb_topk_rows: segmented rows must hold a whole number of segments and at most %d entries", SORT_CAP);
  PeerOut po;
  if (int rc = make_peer_out(po, peer_idx, peer_val, world, peer_mode, per, self_rank, "xmlb_topk_rows")) return rc;
  if (n_rows == 0) return XMLB_OK;
  OutSpec os{out_first, out_count, pad_to, pad_idx, pad_val};
  XMLB_CUDA(cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SelSmem)));
  topk_rows_kernel<<<n_rows, NT, sizeof(SelSmem), (cudaStream_t)stream>>>(
      values, ids, ids_shared, n_cols, k, alpha, apply_exp, tie_desc, row_flags, seg_k, seg_stride, missing_neg, os,
      out_idx, out_val, po);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_topk_rows(const float* values, const int* ids, int ids_shared, int n_rows, int n_cols, int k,
                              float alpha, int apply_exp, int tie_desc, const int* row_flags, int* out_idx,
                              float* out_val, void* stream) {
  XMLB_REQUIRE(out_idx && out_val, "xmlb_topk_rows: null pointer");
  return xmlb_topk_rows_ex(values, ids, ids_shared, n_rows, n_cols, k, alpha, apply_exp, tie_desc, row_flags, 0, 0, 0, 0,
                           k, 0, 0, 0.f, out_idx, out_val, nullptr, nullptr, 0, 0, 0, 0, stream);
}

extern "C" int xmlb_span_topk_ex(const float* st_prob, const float* ed_prob, const float* video_score,
                                 const unsigned char* slot_valid, int n_queries, int n_slots, int ctx_len, int min_l,
                                 int max_l, int k, int tie_desc, int zero_fill_missing, int* out_flat_idx,
                                 float* out_score, const long long* peer_idx, const long long* peer_val, int world,
                                 int peer_mode, int per, int self_rank, void* stream) {
  XMLB_REQUIRE(st_prob && ed_prob && (peer_mode != 0 || (out_flat_idx && out_score)), "xmlb_span_topk: null pointer");
  if (int rc = check_topk_args("xmlb_span_topk", k)) return rc;
  XMLB_REQUIRE(n_slots >= 1 && ctx_len >= 1 && (long long)n_slots * ctx_len * ctx_len < (1ll << 31),
               "xmlb_span_topk: n_slots*L*L must fit in int32");
  XMLB_REQUIRE(min_l >= 0 && max_l > min_l, "xmlb_span_topk: need 0 <= min_l < max_l");
  XMLB_REQUIRE(peer_mode == 0 || !zero_fill_missing, "xmlb_span_topk: peer output carries unfilled lists only");
  PeerOut po;
  if (int rc = make_peer_out(po, peer_idx, peer_val, world, peer_mode, per, self_rank, "xmlb_span_topk")) return rc;
  if (n_queries == 0) return XMLB_OK;
  XMLB_REQUIRE(n_slots <= MAX_SPAN_SLOTS, "xmlb_span_topk: at most %d slots per query", MAX_SPAN_SLOTS);
  const size_t smem = sizeof(SelSmem) + (size_t)LIST_CAP * sizeof(unsigned long long) + (size_t)n_slots * 16;
  XMLB_CUDA(cudaFuncSetAttribute(span_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  span_topk_kernel<<<n_queries, NT, smem, (cudaStream_t)stream>>>(
      st_prob, ed_prob, video_score, slot_valid, n_slots, ctx_len, min_l, max_l, k, tie_desc, zero_fill_missing,
      out_flat_idx, out_score, po);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_span_topk(const float* st_prob, const float* ed_prob, const float* video_score,
                              const unsigned char* slot_valid, int n_queries, int n_slots, int ctx_len, int min_l,
                              int max_l, int k, int tie_desc, int zero_fill_missing, int* out_flat_idx,
                              float* out_score, void* stream) {
  return xmlb_span_topk_ex(st_prob, ed_prob, video_score, slot_valid, n_queries, n_slots, ctx_len, min_l, max_l, k,
                           tie_desc, zero_fill_missing, out_flat_idx, out_score, nullptr, nullptr, 0, 0, 0, 0, stream);
}

extern "C" int xmlb_span_zero_fill_ex(int* flat_idx, float* score, int n_queries, int k, long long total_cells,
                                      int tie_desc, const long long* peer_idx, const long long* peer_val, int world,
                                      int peer_mode, int per, int self_rank, void* stream) {
  XMLB_REQUIRE(flat_idx && score, "xmlb_span_zero_fill: null pointer");
  if (int rc = check_topk_args("xmlb_span_zero_fill", k)) return rc;
  PeerOut po;
  if (int rc = make_peer_out(po, peer_idx, peer_val, world, peer_mode, per, self_rank, "xmlb_span_zero_fill")) return rc;
  if (n_queries == 0) return XMLB_OK;
  span_zero_fill_kernel<<<n_queries, NT, sizeof(SelSmem), (cudaStream_t)stream>>>(flat_idx, score, k, total_cells,
                                                                                 tie_desc, po);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}

extern "C" int xmlb_span_zero_fill(int* flat_idx, float* score, int n_queries, int k, long long total_cells,
                                   int tie_desc, void* stream) {
  return xmlb_span_zero_fill_ex(flat_idx, score, n_queries, k, total_cells, tie_desc, nullptr, nullptr, 0, 0, 0, 0,
                                stream);
}

// Plain copy of `bytes` bytes (multiple of 16, 16-byte aligned) from a local buffer into the same offset of every
// peer's symmetric workspace (mode "to all" for payloads no selection kernel produces: the pooled query vectors).
__global__ void __launch_bounds__(256) peer_copy_kernel(const uint4* __restrict__ src, long long n_vec,
                                                        const __grid_constant__ PeerOut po) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n_vec; i += (long long)gridDim.x * 256) {
    const uint4 v = __ldg(src + i);
    for (int p = 0; p < po.world; ++p) reinterpret_cast<uint4*>(po.val_ptr[p])[i] = v;
  }
}

extern "C" int xmlb_peer_copy(const void* src, long long bytes, const long long* peer_dst, int world, void* stream) {
  XMLB_REQUIRE(src && peer_dst && world >= 1 && world <= 8, "xmlb_peer_copy: bad argument (world <= 8)");
  XMLB_REQUIRE(bytes >= 0 && bytes % 16 == 0 && ((uintptr_t)src & 15) == 0, "xmlb_peer_copy: 16-byte granularity");
  if (bytes == 0) return XMLB_OK;
  PeerOut po = {};
  po.world = world;
  for (int p = 0; p < world; ++p) {
    XMLB_REQUIRE((peer_dst[p] & 15) == 0, "xmlb_peer_copy: destination must be 16-byte aligned");
    po.val_ptr[p] = peer_dst[p];
  }
  const long long n_vec = bytes / 16;
  const int blocks = (int)((n_vec + 255) / 256 < 1184 ? (n_vec + 255) / 256 : 1184);
  peer_copy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), n_vec, po);
  xmlb_count_launch(1);
  XMLB_LAUNCH_CHECK();
  return XMLB_OK;
}
