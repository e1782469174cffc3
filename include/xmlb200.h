/* xmlb200 -- C ABI of the B200 (sm_100a) kernels behind the XML moment-retrieval inference path.
 *
 * This is the drop-in boundary.  The reference (jayleicn/TVRetrieval) is pure
 * Python/PyTorch and has no FFI of its own; each entry point below replaces the PyTorch library calls of one
 * reference call site (cited per function, paths relative to the reference root).  A maintainer binds it with
 * ctypes exactly as tvretrieval_b200/_lib.py does (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into fp32 / int32 / uint8 row-major contiguous buffers that the caller
 *     owns (borrowed for the duration of the call); outputs are caller-allocated;
 *   - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream); all calls are asynchronous
 *     on that stream, re-entrant, and keep no global mutable state besides the last-error text;
 *   - return value: 0 = ok, < 0 = invalid argument / unsupported shape, > 0 = cudaError_t.  No call throws,
 *     aborts, or falls back to another implementation.  xmlb_last_error() returns the message of the last
 *     failing call made by the calling thread.
 */
#ifndef XMLB200_H
#define XMLB200_H

#ifdef __cplusplus
extern "C" {
#endif

const char* xmlb_last_error(void);
int xmlb_version(void);
/* number of kernel launches issued through this library since load (bench.py: gpu_launches) */
long long xmlb_launch_count(void);
/* diagnostic counters of the selection kernels (host array of 4: span_topk rows that overflowed their survivor list,
 * sum of survivor-list lengths, span_topk rows, reserved); synchronises the device */
int xmlb_debug_counters(long long* out4, int reset);

/* ---------------------------------------------------------------- encoder building blocks ---------- */

/* out[r][:] = LayerNorm(x[r][:] + add[r % add_rows][:]) * gamma + beta        (add may be NULL)
 * replaces nn.LayerNorm in LinearLayer (model_components.py:158-159), TrainablePositionalEncoding
 * (:81-88; add = position table, add_rows = sequence length), BertSelfOutput (:316; add = residual,
 * add_rows = rows) and the cross-attention residual norm (model_xml.py:371). */
int xmlb_add_layernorm(const float* x, const float* add, long long add_rows, const float* gamma,
                       const float* beta, float* out, long long rows, int dim, float eps, void* stream);

/* out[rows][out_dim] = act(x[rows][in_dim] . weight[out_dim][in_dim]^T + bias + residual)
 * replaces nn.Linear (+ReLU) in LinearLayer (model_components.py:160-163), the Q/K/V projections
 * (:278-280), BertSelfOutput.dense (:314) and {video,sub}_query_linear (model_xml.py:459-460,524).
 * bias / residual may be NULL; exact fp32 accumulation. */
int xmlb_linear(const float* x, const float* weight, const float* bias, const float* residual, float* out,
                long long rows, int out_dim, int in_dim, int relu, void* stream);

/* Multi-head attention core, replaces BertSelfAttention.forward after the projections
 * (model_components.py:277-303):  out = merge_heads(softmax(Q_h K_h^T / sqrt(dh) + (1 - mask) * -10000) V_h).
 * q (batch, len_q, hidden); k, v (batch, len_k, hidden); mask float {0,1} addressed as
 * mask[b * mask_batch_stride + i * mask_q_stride + j] (mask_q_stride = 0 broadcasts one key mask over the
 * queries, = len_k gives a full (len_q, len_k) mask as the cross attention of model_xml.py:369-370 uses).
 * scores_ws: workspace of batch * n_heads * len_q * len_k floats. */
int xmlb_attention(const float* q, const float* k, const float* v, const float* mask,
                   long long mask_batch_stride, long long mask_q_stride, float* out, float* scores_ws, int batch,
                   int len_q, int len_k, int hidden, int n_heads, void* stream);

/* Modular query pooling, replaces XML.get_modularized_queries (model_xml.py:410-423):
 * a = softmax_tokens(mask_logits(encoded . w_mod^T)); out_m[n] = sum_t a[t][m] * encoded[n][t].
 * w_mod (n_mod, hidden), n_mod in {1, 2}; out1 may be NULL when n_mod == 1. */
int xmlb_modular_pool(const float* encoded, const float* mask, const float* w_mod, float* out0, float* out1,
                      int n_queries, int len, int hidden, int n_mod, void* stream);

/* out = x / max(||x||_2, eps) per row; replaces F.normalize (model_xml.py:446-447). */
int xmlb_l2norm_rows(const float* x, float* out, long long rows, int dim, float eps, void* stream);

/* row softmax (in place allowed); replaces F.softmax(dim=-1) (inference.py:153-154,321-322). */
int xmlb_softmax_rows(const float* x, float* out, long long rows, int dim, void* stream);

/* ---------------------------------------------------------------- query x corpus scoring ----------- */

/* Video-level retrieval scores, replaces XML.get_video_level_scores x modalities + the average
 * (model_xml.py:446-452, 572-574), exact-fp32 SIMT variant:
 *   q2c[q][v] = mean over given modalities of  max_{l : mask[v][l] != 0} q_n[q] . feat1_n[v][l]   (-1e10 if none)
 * Inputs must already be L2-normalised (xmlb_l2norm_rows).  A modality is skipped when its pointers are NULL.
 * workspace: 2 * n_queries * n_videos floats. */
int xmlb_vr_scores_f32(const float* q_video_n, const float* q_sub_n, const float* feat1_video_n,
                       const float* feat1_sub_n, const float* video_mask, const float* sub_mask, float* q2c,
                       float* workspace, int n_queries, int n_videos, int ctx_len, int hidden, void* stream);

/* ---- tensor-core (tcgen05 / TMEM / TMA) variant of the video-level scores ------------------------------
 * Operands are pre-split into 16-bit hi/lo halves (x = hi + lo, fp16 when is_bf16 == 0, else bf16) so that three
 * tensor-core MMAs per k-step (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) reproduce the fp32 product.
 *
 * xmlb_split_rows: x (n_groups * group_in, k) fp32 -> hi, lo (n_groups * group_out, kpad) uint16; rows are
 * regrouped from groups of group_in (clips of a video) to zero-padded groups of group_out, columns zero-padded
 * to kpad (multiple of 64); normalize != 0 applies F.normalize (model_xml.py:446-447) first.  With row_index
 * (n_groups * group_out ints, group_in = group_out = 1) output row r is taken from source row row_index[r]
 * (negative = output row left untouched): this gathers the valid clips into the packed corpus layout and the
 * queries into inverted-list order.  Output row r is written at
 * hi/lo[r * out_ld + out_col0 ...] (out_ld >= out_col0 + kpad), which lets two streams be concatenated along K.
 * hi_err (optional, one float per output row) receives ||x - hi||_2 rounded up: by Cauchy-Schwarz the error of a
 * product evaluated with the hi halves only is <= hi_err(a) * ||b|| + ||a_hi|| * hi_err(b) (two-pass search).
 * xmlb_mask_bits: mask (n_videos, ctx_len) float {0,1} -> bits (n_videos, lp / 32), bit l%32 of word l/32. */
int xmlb_split_rows(const float* x, const int* row_index, long long n_groups, int group_in, int group_out, int k,
                    int kpad, int out_ld, int out_col0, int normalize, int is_bf16, unsigned short* hi,
                    unsigned short* lo, float* hi_err, void* stream);
int xmlb_mask_bits(const float* mask, int n_videos, int ctx_len, int lp, unsigned int* bits, void* stream);
/* Transposing split: x (rows, cols) fp32 -> hi, lo (cols, rpad) uint16 with hi[c][r] + lo[c][r] ~= x[r][c]; columns
 * r >= rows are zero (rpad multiple of 64, >= rows).  Operands of the dW = dY^T . X GEMM of the training step. */
int xmlb_split_rows_t(const float* x, long long rows, int cols, long long rpad, int is_bf16, unsigned short* hi,
                      unsigned short* lo, void* stream);
/* dst_{hi,lo}[r] = src_{hi,lo}[row_index[r]] for rows of kpad 16-bit elements (negative index: row left untouched):
 * brings operands that were split once (queries of a block) into inverted-list order for xmlb_vr_rescore_tc and
 * xmlb_span_probs_tc -- same bits as xmlb_split_rows with row_index, without re-reading the fp32 rows. */
int xmlb_gather_rows16(const unsigned short* src_hi, const unsigned short* src_lo, const int* row_index,
                       long long rows_out, int kpad, unsigned short* dst_hi, unsigned short* dst_lo, void* stream);

/* q2c[q][v] = mean over given modalities of max_{l : bit set} q[q] . c[v * lp + l]; same contract as
 * xmlb_vr_scores_f32 (model_xml.py:446-452, 572-574) on prepared operands: q_* (n_queries, kpad),
 * c_* (n_videos * lp, kpad), mask_bits_* (n_videos, lp/32).  lp multiple of 32, <= 256.  Modality b optional (NULL).
 * max_ctas: 0 = one persistent CTA per SM.  sched_ws: one int of device scratch (the dynamic tile counter; every
 * tensor-core entry point takes one and zeroes it itself). */
int xmlb_vr_scores_tc(const unsigned short* q_hi_a, const unsigned short* q_lo_a, const unsigned short* q_hi_b,
                      const unsigned short* q_lo_b, const unsigned short* c_hi_a, const unsigned short* c_lo_a,
                      const unsigned short* c_hi_b, const unsigned short* c_lo_b, const unsigned int* mask_bits_a,
                      const unsigned int* mask_bits_b, float* q2c, int* sched_ws, int n_queries, int n_videos, int lp,
                      int kpad, int is_bf16, int max_ctas, void* stream);

/* Packed ("ragged") variant of xmlb_vr_scores_tc: c_* hold only the valid clips, (n_packed_rows, kpad), videos
 * packed whole into tiles of <= 256 consecutive rows.  tile_meta (n_tiles, 4) int = {first packed row, ordinal of the
 * tile's first video, used columns, number of videos (<= 32)}; tile_starts (n_tiles, 8) = 256-bit map of columns
 * where a video starts.  q2c (n_queries, n_videos) is written in PACKED ORDINAL order: column o is the o-th packed
 * video (the caller keeps the ordinal -> video id table and passes it to xmlb_topk_rows as shared ids); columns of
 * videos that are not packed (no valid clip) are not written.  Both modalities share the packing (same masks).
 * hi_only != 0 evaluates only the hi*hi products (1 MMA per k-step, lo halves not read): an APPROXIMATE score
 * whose error is bounded by the hi_err outputs of xmlb_split_rows; it is the filter pass of the two-pass search
 * (xmlb_select_candidates, xmlb_vr_rescore_tc).
 * m_tile_list / n_m_tiles (both NULL, or device pointers): restricted mode -- only the 128-query tiles
 * m_tile_list[0 .. *n_m_tiles) are computed (the exact fallback for rows whose candidate list overflowed). */
int xmlb_vr_scores_tc_packed(const unsigned short* q_hi_a, const unsigned short* q_lo_a, const unsigned short* q_hi_b,
                             const unsigned short* q_lo_b, const unsigned short* c_hi_a, const unsigned short* c_lo_a,
                             const unsigned short* c_hi_b, const unsigned short* c_lo_b, const int* tile_meta,
                             const unsigned int* tile_starts, float* q2c, int* sched_ws, int n_queries, int n_videos,
                             long long n_packed_rows, int n_tiles, int hi_only, const int* m_tile_list,
                             const int* n_m_tiles, int kpad, int is_bf16, int max_ctas, void* stream);

/* ---- two-pass video retrieval: approximate filter, exact re-scoring of the survivors -------------------------
 * The top-k videos of inference.py:347-348 are found without evaluating all (query, video) pairs at full
 * precision: pass 1 = xmlb_vr_scores_tc_packed(hi_only = 1) (one third of the tensor-core work); then
 *
 * xmlb_select_candidates: per row r of approx (n_rows, n_cols), with error bound
 *   eps[r] = err_scale * (row_err_a[r] + row_err_b[r]) + err_const        (row_err_b may be NULL)
 * keeps every column with approx >= kth - 2 * eps[r], kth = a lower bound of the row's k-th largest approx at most
 * 3.1e-5 below it (lower edge of a 2^-15-wide bin of max - approx; the exact value when fewer than k scores lie within
 * 1/8 of the row maximum) -- a superset of the exact top-k whenever |approx - exact| <= eps[r].  row_kth (NULL, or n_rows floats): the k-th largest approx value
 * supplied by the caller instead -- for a corpus sharded over GPUs it is taken over ALL shards (one small
 * all-gather), so each shard keeps only its part of the GLOBAL candidate set (n_cols < k is then allowed).  Outputs (n_rows, max_cand): cand_col = column, cand_id = ids[column]
 * (ids NULL: column), cand_val = approx value; unused slots are (-1, INT_MAX, -1e10).  A row with more than
 * max_cand survivors is truncated and FLAGGED in flag_ws (ints, zeroed by the call): [0] = number of flagged row
 * groups (rows_per_group = 128 = query tile of the scoring kernel), [1, 1+G) group flags, [1+G, 1+2G) list of
 * flagged groups, [1+2G, 1+2G+n_rows) row flags, G = ceil(n_rows / rows_per_group).  Flagged rows are recomputed
 * by the restricted modes of xmlb_vr_scores_tc_packed and xmlb_topk_rows (no host round trip).
 *
 * xmlb_vr_rescore_tc: exact split-precision scores of the candidate pairs, same value as xmlb_vr_scores_tc_packed.
 * Pairs are grouped per packed video through xmlb_build_pair_lists(cand_col, chunk = 128) + xmlb_build_span_units;
 * qg_* (n_entries, kpad) = hi/lo of the normalised queries gathered in list order (xmlb_split_rows with
 * row_index = entry_q); c_* = packed corpus; row_start (n_packed + 1) = first packed row of each packed video;
 * cand_val[entry_out[e]] is overwritten for every list entry e.  max_len = longest video (<= 256 clips). */
int xmlb_select_candidates(const float* approx, const int* ids, int n_rows, int n_cols, int k,
                           const float* row_err_a, const float* row_err_b, float err_scale, float err_const,
                           const float* row_kth, int max_cand, int rows_per_group, int* cand_col, int* cand_id,
                           float* cand_val, int* flag_ws, void* stream);
int xmlb_vr_rescore_tc(const unsigned short* qg_hi_a, const unsigned short* qg_lo_a, const unsigned short* qg_hi_b,
                       const unsigned short* qg_lo_b, const unsigned short* c_hi_a, const unsigned short* c_lo_a,
                       const unsigned short* c_hi_b, const unsigned short* c_lo_b, const int* row_start,
                       const int* units, const int* n_units, int max_units, const int* entry_out, float* cand_val,
                       int* sched_ws, long long n_entries, long long n_packed_rows, int max_len, int kpad,
                       int is_bf16, void* stream);

/* Tensor-core variant of xmlb_linear (same contract, model_components.py:160-163,278-280,314): x_* (rows, kpad)
 * and w_* (out_dim, kpad) are the 16-bit (hi, lo) halves produced by xmlb_split_rows (normalize = 0).
 * out / bias / residual must be 16-byte aligned. */
int xmlb_linear_tc(const unsigned short* x_hi, const unsigned short* x_lo, const unsigned short* w_hi,
                   const unsigned short* w_lo, const float* bias, const float* residual, float* out, int* sched_ws,
                   long long rows, int out_dim, int kpad, int relu, int is_bf16, void* stream);

/* xmlb_linear_tc with K-chunked accumulation made explicit and fused output formats (any subset of the three):
 *   out                   fp32 (rows, out_dim), as xmlb_linear_tc;
 *   out_hi / out_lo       the 16-bit (hi, lo) split of output columns [0, out16_cols), rows of out16_ld elements
 *                         (columns beyond out16_cols are not written): the operand of the next xmlb_linear_tc /
 *                         the Q and K operands of xmlb_attention_tc;
 *   vt_hi / vt_lo         the split of output columns [vt_col0, out_dim) TRANSPOSED per sequence of vt_seq rows:
 *                         vt[(b * (out_dim - vt_col0) + c) * vt_ld + l] = out[b * vt_seq + l][vt_col0 + c] -- the V^T
 *                         operand of xmlb_attention_tc (positions l >= vt_seq are not written: zero them once).
 * k_chunk (multiple of 32; 0 = 128): elements of K summed by the tensor core between two round-to-nearest fp32
 * additions in registers -- the TMEM accumulator truncates, so its error grows with the number of updates. */
int xmlb_linear_tc_ex(const unsigned short* x_hi, const unsigned short* x_lo, const unsigned short* w_hi,
                      const unsigned short* w_lo, const float* bias, const float* residual, float* out,
                      unsigned short* out_hi, unsigned short* out_lo, int out16_ld, int out16_cols,
                      unsigned short* vt_hi, unsigned short* vt_lo, int vt_col0, int vt_seq, int vt_ld, int* sched_ws,
                      long long rows, int out_dim, int kpad, int relu, int is_bf16, int k_chunk, void* stream);

/* xmlb_add_layernorm (add_index NULL) / xmlb_add_layernorm_indexed that also (or only: out NULL) writes the 16-bit
 * (hi, lo) split of the normalised rows, kpad >= dim elements per row, zero padded -- the operand format of the
 * xmlb_linear_tc that follows every LayerNorm of the encoders (model_components.py:158-163, 81-88, 316). */
int xmlb_add_layernorm_split(const float* x, const float* add, long long add_rows, const int* add_index,
                             const float* gamma, const float* beta, float* out, unsigned short* out_hi,
                             unsigned short* out_lo, int kpad, int is_bf16, long long rows, int dim, float eps,
                             void* stream);

/* Fused multi-head attention core on the tensor cores, same function as xmlb_attention (model_components.py:277-303)
 * without the (batch, heads, len_q, len_k) workspace: S = Q K^T in TMEM, scale + additive -10000 mask + softmax in
 * registers, P written back to shared memory as (hi, lo) halves, O = P V in TMEM.  Operands are the split outputs of
 * xmlb_linear_tc_ex: q_* (batch * len_q, q_ld) with head h at columns q_col0 + h * dh, k_* (batch * len_k, k_ld)
 * likewise, vt_* (batch * hidden, vt_ld) = V transposed per sequence (row b * hidden + c, column = key position;
 * vt_ld >= max(len_k, 64), multiple of 8, columns >= len_k zero).  mask as in xmlb_attention.  Outputs: out fp32
 * (batch * len_q, hidden) and / or its (hi, lo) split with rows of out16_ld elements.  len_k <= 256,
 * head size in {64, 128, 192, 256}. */
int xmlb_attention_tc(const unsigned short* q_hi, const unsigned short* q_lo, int q_ld, int q_col0,
                      const unsigned short* k_hi, const unsigned short* k_lo, int k_ld, int k_col0,
                      const unsigned short* vt_hi, const unsigned short* vt_lo, int vt_ld, const float* mask,
                      long long mask_batch_stride, long long mask_q_stride, float* out, unsigned short* out_hi,
                      unsigned short* out_lo, int out16_ld, int batch, int len_q, int len_k, int hidden, int n_heads,
                      int is_bf16, void* stream);

/* Similarity curves + ConvSE + mask (+ softmax), replaces XML.get_merged_st_ed_prob (model_xml.py:455-502),
 * XML._get_st_ed_prob (:512-551) and the driver's softmax over clips (inference.py:321-322).
 *   sim_x[q][v][l] = q_x[q] . feat2_x[v][l]                       for stream x in {a, b}
 *   merged  : st = mask_logits(conv(w_st_a, (sim_a + sim_b) / 2), mask_a)
 *   separate: st = mean_x mask_logits(conv(w_st_x, sim_x), mask_x)           (one or two streams)
 * conv = cross-correlation, zero padding ksize/2, stride 1.  Output rows of ctx_len floats:
 *   dense (chunk_ptr == NULL): row = q * n_videos + v for every (q, v);
 *   list mode: for video v, entries e in [vid_ptr[v], vid_ptr[v+1]) give query entry_q[e] and output row
 *   entry_out[e]; chunk_ptr = exclusive scan of ceil(count/32) (see xmlb_build_pair_lists); max_chunks >= chunk_ptr[n_videos].
 * q_b/feat2_b/mask_b/w_*_b may be NULL (single stream). */
int xmlb_span_logits(const float* q_a, const float* q_b, const float* feat2_a, const float* feat2_b,
                     const float* mask_a, const float* mask_b, const float* w_st_a, const float* w_ed_a,
                     const float* w_st_b, const float* w_ed_b, int ksize, int merged, int apply_softmax,
                     int n_queries, int n_videos, int ctx_len, int hidden, const int* chunk_ptr, const int* vid_ptr,
                     const int* entry_q, const int* entry_out, int max_chunks, float* out_st, float* out_ed,
                     void* stream);

/* Inverts top_idx (n_queries, n_slots) [global video ids] into per-video lists for xmlb_span_logits.
 * Only ids in [vid_lo, vid_lo + n_videos) (this GPU's shard) and slots with slot_valid != 0 (NULL = all) are
 * listed; entry_out = q * n_slots + slot; chunk_ptr = exclusive scan of ceil(list length / chunk) (chunk = 32 for
 * xmlb_span_logits).  counts_ws, cursor_ws: n_videos ints; vid_ptr, chunk_ptr: n_videos+1;
 * entry_q, entry_out: n_queries * n_slots ints.  Replaces the advanced-index gather of inference.py:365-367. */
int xmlb_build_pair_lists(const int* top_idx, const unsigned char* slot_valid, int n_queries, int n_slots,
                          int vid_lo, int n_videos, int chunk, int* counts_ws, int* cursor_ws, int* vid_ptr,
                          int* chunk_ptr, int* entry_q, int* entry_out, void* stream);

/* ---- tensor-core variant of the similarity curves for the merged two-stream model (list mode only) -------
 * xmlb_build_span_units: units[chunk_ptr[v] + c] = {v, vid_ptr[v] + c*chunk, entries in the chunk, 0} (int4 each)
 * from lists built with the same `chunk`; the number of units is chunk_ptr[n_videos] (stays on the device).
 * xmlb_span_probs_tc: same result as xmlb_span_logits(merged = 1, list mode) (model_xml.py:459-471,496-497 +
 * inference.py:321-322): f2_* (n_videos * ctx_len, kcat) = hi/lo of [feat2_video | feat2_sub] concatenated along K
 * (each stream zero-padded to kcat/2), qg_* (n_entries, kcat) = hi/lo of [q'_video | q'_sub] gathered in list
 * order (xmlb_split_rows with row_index = entry_q).  ctx_len <= 128; block_n in {32, 64, 128} = chunk. */
int xmlb_build_span_units(const int* vid_ptr, const int* chunk_ptr, int n_videos, int chunk, int* units,
                          void* stream);
/* ... with units[].w = video_rows[v] (NULL: 0), the operand rows xmlb_span_probs_tc_clipped loads for video v */
int xmlb_build_span_units_rows(const int* vid_ptr, const int* chunk_ptr, int n_videos, int chunk,
                               const int* video_rows, int* units, void* stream);
int xmlb_span_probs_tc(const unsigned short* f2_hi, const unsigned short* f2_lo, const unsigned short* qg_hi,
                       const unsigned short* qg_lo, const float* mask, const float* w_st, const float* w_ed,
                       int ksize, int apply_softmax, int n_videos, int ctx_len, int kcat, long long n_entries,
                       int block_n, const int* units, const int* n_units, int max_units, const int* entry_out,
                       float* out_st, float* out_ed, int* sched_ws, int is_bf16, void* stream);

/* Gather variants of the two grouped kernels: entry_q (n_entries ints, the query of every list entry, as
 * xmlb_build_pair_lists returns it) is given and the q arrays are the UN-gathered (n_query_rows, k) halves of all
 * queries; the kernels fetch the listed rows themselves instead of reading a gathered copy that xmlb_gather_rows16
 * first had to write to HBM: gather_warps != 0 -- four extra warps copy 16-byte pieces (ld.global, L2-resident) into
 * the swizzled shared-memory tile; gather_warps == 0 -- the TMA producer issues cp.async.bulk.tensor ...
 * tile::gather4 (four rows per instruction; measured slower).  entry_q == NULL is the pre-gathered contract of
 * xmlb_vr_rescore_tc / xmlb_span_probs_tc.  xmlb_span_probs_tc_ex accepts ctx_len <= 256 (two 128-clip accumulator
 * halves per video). */
int xmlb_vr_rescore_tc_ex(const unsigned short* qg_hi_a, const unsigned short* qg_lo_a, const unsigned short* qg_hi_b,
                          const unsigned short* qg_lo_b, const unsigned short* c_hi_a, const unsigned short* c_lo_a,
                          const unsigned short* c_hi_b, const unsigned short* c_lo_b, const int* row_start,
                          const int* units, const int* n_units, int max_units, const int* entry_out, const int* entry_q,
                          int gather_warps, long long n_query_rows, float* cand_val, int* sched_ws, long long n_entries,
                          long long n_packed_rows, int max_len, int kpad, int is_bf16, void* stream);
/* xmlb_vr_rescore_tc_ex with c_kblocked != 0: the corpus halves c_* are stored K-BLOCKED, [kpad / 32][n_packed_rows][32]
 * (xmlb_span_probs_tc_clipped explains why).  In gather-warps mode only the quarters of the clip tile that hold the
 * video's clips are loaded. */
int xmlb_vr_rescore_tc_kb(const unsigned short* qg_hi_a, const unsigned short* qg_lo_a, const unsigned short* qg_hi_b,
                          const unsigned short* qg_lo_b, const unsigned short* c_hi_a, const unsigned short* c_lo_a,
                          const unsigned short* c_hi_b, const unsigned short* c_lo_b, int c_kblocked,
                          const int* row_start, const int* units, const int* n_units, int max_units,
                          const int* entry_out, const int* entry_q, int gather_warps, long long n_query_rows,
                          float* cand_val, int* sched_ws, long long n_entries, long long n_packed_rows, int max_len,
                          int kpad, int is_bf16, void* stream);
int xmlb_span_probs_tc_ex(const unsigned short* f2_hi, const unsigned short* f2_lo, const unsigned short* qg_hi,
                          const unsigned short* qg_lo, const float* mask, const float* w_st, const float* w_ed,
                          int ksize, int apply_softmax, int n_videos, int ctx_len, int kcat, long long n_entries,
                          int block_n, const int* units, const int* n_units, int max_units, const int* entry_out,
                          const int* entry_q, int gather_warps, long long n_query_rows, float* out_st, float* out_ed,
                          int* sched_ws, int is_bf16, void* stream);

/* xmlb_span_probs_tc_ex with clip_boxes != 0 (gather-warps mode only): units[].w (xmlb_build_span_units_rows) is the
 * number of leading clip rows of the video whose similarity the epilogue can need -- (last unmasked clip + 1 +
 * ksize / 2), at most ctx_len.  Only those rows of f2_* are loaded (TMA boxes of 16, 32, ..., 128 rows picked per
 * unit) instead of all ctx_len padded rows; masked clips never read their neighbourhood (mask_logits gives -1e10
 * for any finite logit, model_xml.py:640-641), so the results are identical.
 * f2_kblocked != 0: f2_* are stored K-BLOCKED, [kcat / 32][n_videos * ctx_len][32] -- the
 * 32 elements a k-step reads from consecutive clips are contiguous in HBM (one run of rows * 64 bytes per TMA box
 * instead of `rows` separate 64-byte pieces 2 * kcat bytes apart, each in its own DRAM page).
 * f2_kblocked == 2 (gather-warps mode, ctx_len % 8 == 0): ... and the four 16-byte pieces of every row are already
 * permuted the way SWIZZLE_64B places them in shared memory (piece c of row r at position c ^ ((r >> 1) & 3)), i.e.
 * HBM holds the shared-memory image of every tile: a k-step's rows are fetched with ONE plain cp.async.bulk instead
 * of a tensor box, which the TMA unit walks one 64-byte row at a time (the rate that bounded the kernel). */
int xmlb_span_probs_tc_clipped(const unsigned short* f2_hi, const unsigned short* f2_lo, const unsigned short* qg_hi,
                               const unsigned short* qg_lo, const float* mask, const float* w_st, const float* w_ed,
                               int ksize, int apply_softmax, int n_videos, int ctx_len, int kcat, long long n_entries,
                               int block_n, const int* units, const int* n_units, int max_units, const int* entry_out,
                               const int* entry_q, int gather_warps, long long n_query_rows, int clip_boxes,
                               int f2_kblocked, float* out_st, float* out_ed, int* sched_ws, int is_bf16,
                               void* stream);

/* Filter pass of the two-pass video retrieval on CTA pairs (tcgen05 cta_group::2, M = 256): the hi-only
 * (1 MMA per product) scores of xmlb_vr_scores_tc_packed(hi_only = 1) for all (query, video) pairs, same packed
 * corpus, tile tables and packed-ordinal output layout; approximates model_xml.py:446-452,572-574 to within the
 * bound that xmlb_select_candidates uses.  q_hi_b / c_hi_b NULL = one modality. */
int xmlb_vr_filter_pair(const unsigned short* q_hi_a, const unsigned short* q_hi_b, const unsigned short* c_hi_a,
                        const unsigned short* c_hi_b, const int* tile_meta, const unsigned int* tile_starts,
                        float* q2c, int* sched_ws, int n_queries, int n_videos, long long n_packed_rows, int n_tiles,
                        int kpad, int is_bf16, void* stream);

/* Per-row exact top-k, ranked (value desc, id asc | desc).  value = apply_exp ? exp(alpha * x) : x.
 * ids: optional explicit ids (NULL: column index), (n_rows, n_cols) when ids_shared == 0, one (n_cols) table shared
 * by all rows when ids_shared != 0.  Replaces torch.exp + torch.topk of
 * inference.py:317,347-348; returns an error when k > n_cols like torch.topk does.  k <= 1024.
 * row_flags (NULL or n_rows ints): restricted mode, rows whose flag is 0 are skipped (outputs left untouched). */
int xmlb_topk_rows(const float* values, const int* ids, int ids_shared, int n_rows, int n_cols, int k, float alpha,
                   int apply_exp, int tie_desc, const int* row_flags, int* out_idx, float* out_val, void* stream);

/* ---- peer-memory output of the ranking kernels (video-sharded search over NVLink, no counterpart in the reference) ----
 * The *_ex variants write their ranked lists through PEER POINTERS into the symmetric workspaces of the other GPUs
 * (addresses obtained from torch.distributed._symmetric_memory; host arrays of `world` <= 8 entries, 0 = skip):
 *   peer_mode 0: local out_idx / out_val only (the plain entry points);
 *   peer_mode 1 "to owner": row r belongs to rank o = r / per and is stored on rank o at row self_rank * per + r % per;
 *   peer_mode 2 "to all"  : row r (an owned query) is stored on EVERY rank at row self_rank * per + r.
 * xmlb_topk_rows_ex extras: seg_k > 0 reads the row as per-rank lists laid out [source rank][row][seg_k] with
 * seg_stride elements between source ranks (what a mode-1 exchange leaves on the owner); missing_neg treats entries
 * with negative ids as absent (returned as (-1, 0)); only ranks [out_first, out_first + out_count) of each list are
 * written, followed by (pad_idx, pad_val) up to pad_to entries per row (row pitch = max(out_count, pad_to)).
 * xmlb_peer_copy: src[0 .. bytes) -> the same bytes at peer_dst[p] for every p (16-byte granularity). */
int xmlb_topk_rows_ex(const float* values, const int* ids, int ids_shared, int n_rows, int n_cols, int k, float alpha,
                      int apply_exp, int tie_desc, const int* row_flags, int seg_k, long long seg_stride,
                      int missing_neg, int out_first, int out_count, int pad_to, int pad_idx, float pad_val,
                      int* out_idx, float* out_val, const long long* peer_idx, const long long* peer_val, int world,
                      int peer_mode, int per, int self_rank, void* stream);
int xmlb_span_topk_ex(const float* st_prob, const float* ed_prob, const float* video_score,
                      const unsigned char* slot_valid, int n_queries, int n_slots, int ctx_len, int min_l, int max_l,
                      int k, int tie_desc, int zero_fill_missing, int* out_flat_idx, float* out_score,
                      const long long* peer_idx, const long long* peer_val, int world, int peer_mode, int per,
                      int self_rank, void* stream);
int xmlb_span_zero_fill_ex(int* flat_idx, float* score, int n_queries, int k, long long total_cells, int tie_desc,
                           const long long* peer_idx, const long long* peer_val, int world, int peer_mode, int per,
                           int self_rank, void* stream);
int xmlb_peer_copy(const void* src, long long bytes, const long long* peer_dst, int world, void* stream);

/* Band-limited span scoring + exact top-k, replaces inference.py:370-386 (VCMR) and inference.py:215-224 +
 * utils/tensor_utils.py:133-141 (SVMR, n_slots = 1, video_score = NULL, tie_desc = 1):
 *   score[j][m][n] = (st[q][j][m] * video_score[q][j]) * ed[q][j][n]   for min_l <= n - m < max_l, else 0
 *   flat index = (j * ctx_len + m) * ctx_len + n;  rank by (score desc, flat index asc|desc); keep k.
 * slot_valid (n_queries, n_slots) uint8 or NULL.  When fewer than k cells are positive, the rest is filled with
 * zero-score cells in flat-index order if zero_fill_missing, else with (-1, 0). */
int xmlb_span_topk(const float* st_prob, const float* ed_prob, const float* video_score,
                   const unsigned char* slot_valid, int n_queries, int n_slots, int ctx_len, int min_l, int max_l,
                   int k, int tie_desc, int zero_fill_missing, int* out_flat_idx, float* out_score, void* stream);

/* Completes ranked lists whose tail is missing ((-1, 0) or non-positive) with zero-score cells of
 * [0, total_cells) in flat-index order (used after the multi-GPU merge). */
int xmlb_span_zero_fill(int* flat_idx, float* score, int n_queries, int k, long long total_cells, int tie_desc,
                        void* stream);

/* Greedy temporal NMS per query over ranked lists, replaces utils/temporal_nms.py:25-74 as wrapped by
 * baselines/clip_alignment_with_language/inference.py:189-265.  video_idx NULL = one group (SVMR).
 * out_idx (n_queries, max_out): indices into the input list, ranked; out_count (n_queries). n_in <= 4096. */
int xmlb_temporal_nms(const int* video_idx, const float* st, const float* ed, const float* score,
                      const int* n_valid, int n_queries, int n_in, double iou_thd, int max_per_group, int max_out,
                      int* out_idx, int* out_count, void* stream);

/* Retrieval metrics on ranked DEVICE lists, the per-query part of standalone_eval/eval.py:83-252:
 * first_hit[q][t] = rank of the first prediction of query q that is correct at IoU threshold iou_thds[t] (INT_MAX if
 * none); mode 0 = VCMR (right video and IoU >= thd), 1 = SVMR (ranks count only predictions on the ground-truth
 * video), 2 = VR (right video; n_thds = 1, spans / thresholds may be NULL).  IoU = intersection / convex hull in fp32
 * (eval.py:54-69).  pred_* (n_queries, n_pred), n_valid NULL = all n_pred; R@K = mean(first_hit < K). */
int xmlb_eval_first_hit(const int* pred_vid, const float* pred_st, const float* pred_ed, const int* n_valid,
                        const int* gt_vid, const float* gt_st, const float* gt_ed, const float* iou_thds, int n_thds,
                        int n_queries, int n_pred, int mode, int* first_hit, void* stream);

/* ---------------------------------------------------------------- packed (ragged) query encoder -------- */
/* The query encoder on a PACKED token layout: only the valid tokens of each query are stored, sequence s owns rows
 * [cu_seqlens[s], cu_seqlens[s+1]) of every (T, hidden) activation.  Padded tokens have exactly zero weight as
 * attention keys (exp(-10000 - max) == 0 in fp32) and in the modular pooling (softmax of -1e10), and their own rows
 * are never read, so the pooled query vectors equal those of the padded computation (model_xml.py:291-295,377-423)
 * at the work of the valid tokens only.  Row-wise ops (xmlb_linear*, xmlb_split_rows, xmlb_add_layernorm) run on the
 * packed rows unchanged; the three below are the sequence-aware pieces.
 *
 * xmlb_add_layernorm_indexed: out[r] = LayerNorm(x[r] + add[add_index[r]]) -- TrainablePositionalEncoding
 *   (model_components.py:81-88) with the position of every packed row given explicitly.
 * xmlb_attention_ragged: fused softmax(Q_h K_h^T / sqrt(dh)) V_h per (sequence, head), sequences of <= 32 tokens,
 *   head size a multiple of 4; q / k / v / out (T, hidden), no workspace (model_components.py:277-303).
 * xmlb_modular_pool_ragged: xmlb_modular_pool on packed rows (model_xml.py:410-423). */
int xmlb_add_layernorm_indexed(const float* x, const float* add, const int* add_index, long long add_rows,
                               const float* gamma, const float* beta, float* out, long long rows, int dim, float eps,
                               void* stream);
int xmlb_attention_ragged(const float* q, const float* k, const float* v, const int* cu_seqlens, float* out,
                          int n_seqs, int max_len, int hidden, int n_heads, void* stream);
/* xmlb_attention_ragged on the output of ONE fused Q|K|V projection: qkv (T, 3 * hidden), row = [q | k | v]. */
int xmlb_attention_ragged_qkv(const float* qkv, const int* cu_seqlens, float* out, int n_seqs, int max_len, int hidden,
                              int n_heads, void* stream);
int xmlb_modular_pool_ragged(const float* encoded, const int* cu_seqlens, const float* w_mod, float* out0, float* out1,
                             int n_queries, int max_len, int hidden, int n_mod, void* stream);

/* ---------------------------------------------------------------- training step (config #4) -------- */

/* out[i] = keep(seed, index0 + i) ? x[i] / (1 - p) : 0 -- nn.Dropout in train mode (model_components.py:77,152,
 * 263,311).  The mask is a pure function of (seed, element index), so the backward pass calls this again with the
 * same seed instead of storing a mask.  x NULL = ones (returns the scaled mask itself); in-place allowed. */
int xmlb_dropout(const float* x, float* out, long long n, float p, unsigned long long seed,
                 unsigned long long index0, void* stream);

/* xmlb_attention with dropout on the attention probabilities (model_components.py:296; train mode).  The mask of
 * probability element e (flat index in (batch, n_heads, len_q, len_k)) is keep(seed, index0 + e). */
int xmlb_attention_train(const float* q, const float* k, const float* v, const float* mask,
                         long long mask_batch_stride, long long mask_q_stride, float* out, float* scores_ws,
                         int batch, int len_q, int len_k, int hidden, int n_heads, float dropout_p,
                         unsigned long long seed, unsigned long long index0, void* stream);

/* ---- backward kernels of the training step (everything that is not a Linear layer; fp32, deterministic) ----------
 * xmlb_sum_rows: out[p][c] = sum_g in[(g * group_rows + p) * dim + c] in a fixed order (bias / LayerNorm / position
 *   table / ConvSE gradients); ws: 2 * ceil(n_groups / 64) * group_rows * dim floats (NULL when n_groups <= 64).
 * xmlb_layernorm_backward: for y = LN(x + add[r % add_rows]) * gamma + beta (xmlb_add_layernorm): dx = gradient w.r.t.
 *   (x + add), dy_xhat = dy * xhat (column sums = dgamma; column sums of dy = dbeta).
 * xmlb_l2norm_backward, xmlb_relu_backward (dx = out > 0 ? g : 0): the obvious ones.
 * xmlb_modular_pool_backward: d_encoded and dlogit (n, len, 2) of xmlb_modular_pool; xmlb_modular_mapping_grad:
 *   dW (n_mod, hidden) = dlogit^T . encoded.
 * xmlb_vr_scores_backward: one modality of xmlb_vr_scores_f32 for the in-batch (Nq, Nv) scores: g (Nq, Nv) upstream
 *   gradient, scale = 1 / number of modalities; dq (Nq, H), dc (Nv, L, H) (fully written); argmax_ws Nq * Nv ints.
 * xmlb_span_logits_diag_backward: xmlb_span_logits in diagonal list mode (item b on its own video), merged or separate
 *   streams; dw_partial (n, 2 streams, 2 predictors, 31) per-item ConvSE weight gradients (sum with xmlb_sum_rows).
 * xmlb_attention_backward: xmlb_attention / xmlb_attention_train; probabilities recomputed, dropout mask re-derived
 *   from (seed, index0); ws_p / ws_g: batch * n_heads * len_q * len_k floats each. */
int xmlb_sum_rows(const float* in, long long n_groups, int group_rows, int dim, float* out, float* ws, void* stream);
int xmlb_layernorm_backward(const float* x, const float* add, long long add_rows, const float* gamma, const float* dy,
                            long long rows, int dim, float eps, float* dx, float* dy_xhat, void* stream);
int xmlb_l2norm_backward(const float* x, const float* dy, long long rows, int dim, float eps, float* dx, void* stream);
int xmlb_relu_backward(const float* g, const float* out, long long n, float* dx, void* stream);
int xmlb_modular_pool_backward(const float* encoded, const float* mask, const float* w_mod, const float* dout0,
                               const float* dout1, int n_queries, int len, int hidden, int n_mod, float* d_encoded,
                               float* dlogit, void* stream);
int xmlb_modular_mapping_grad(const float* dlogit, const float* encoded, long long rows, int hidden, int n_mod, float* dw,
                              void* stream);
int xmlb_vr_scores_backward(const float* q_n, const float* c_n, const float* mask, const float* g, float scale,
                            int n_queries, int n_videos, int ctx_len, int hidden, int* argmax_ws, float* dq, float* dc,
                            void* stream);
int xmlb_span_logits_diag_backward(const float* q_a, const float* q_b, const float* feat2_a, const float* feat2_b,
                                   const float* mask_a, const float* mask_b, const float* w_st_a, const float* w_ed_a,
                                   const float* w_st_b, const float* w_ed_b, int ksize, int merged, int n, int ctx_len,
                                   int hidden, const float* dst, const float* ded, float* dq_a, float* dq_b,
                                   float* dfeat2_a, float* dfeat2_b, float* dw_partial, void* stream);
int xmlb_attention_backward(const float* q, const float* k, const float* v, const float* mask,
                            long long mask_batch_stride, long long mask_q_stride, const float* dout, float* dq, float* dk,
                            float* dv, float* ws_p, float* ws_g, int batch, int len_q, int len_k, int hidden, int n_heads,
                            float dropout_p, unsigned long long seed, unsigned long long index0, void* stream);

/* One BertAdam.step (optimization.py:273-338) over all parameter tensors in two launches.
 * chunk_table: n_chunks rows of 6 x int64 {param*, grad*, m*, v* (offset to the chunk), elements, tensor id};
 * tensor_table: n_tensors rows of {int first_chunk, int n_chunks, float lr_scheduled, float weight_decay};
 * partial_ws: n_chunks floats.  Per tensor: g *= min(1, max_grad_norm / (||g|| + 1e-6)) (written back, like
 * clip_grad_norm_); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g g; p -= lr (m / (sqrt(v) + eps) + wd p).
 * max_grad_norm <= 0 disables clipping.  No bias correction (as the reference). */
int xmlb_bert_adam_step(const long long* chunk_table, int n_chunks, const int* tensor_table, int n_tensors,
                        float* partial_ws, double b1, double b2, double eps, double max_grad_norm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XMLB200_H */
