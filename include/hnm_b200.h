/*
 * hnm_b200.h -- C ABI of libhnm_b200.so: the B200 (sm_100a) scoring hot path of
 * hyunlord/hnm_recommendation (LightGCN propagate -> full-catalog score -> top-k,
 * NeuralCF candidate scoring).
 *
 * The reference has no FFI of its own: its hot path is a handful of torch /
 * torch_sparse library calls inside two Python model classes.  Each entry point
 * below replaces one of those call sites (cited as file:line relative to the
 * reference repository root); INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all memory (outputs and workspaces are caller-allocated);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it
 *     unless stated otherwise and re-entrant across streams;
 *   - return value: 0 = ok, < 0 = HNM_E_* argument error, > 0 = cudaError_t;
 *   - nodes [0, num_users) are users and [num_users, num_users+num_items) items
 *     (src/models/lightgcn.py:70,161-162); embeddings are row-major fp32 [rows, dim].
 */
#ifndef HNM_B200_H_
#define HNM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HNM_API __attribute__((visibility("default")))
#else
#define HNM_API
#endif

#define HNM_ABI_VERSION 2

#define HNM_OK 0
#define HNM_E_NULL (-1)      /* required pointer is NULL */
#define HNM_E_RANGE (-2)     /* size / index argument out of range */
#define HNM_E_DIM (-3)       /* unsupported embedding dimension */
#define HNM_E_WORKSPACE (-4) /* workspace too small */
#define HNM_E_ALIGN (-5)     /* pointer not 16-byte aligned */
#define HNM_E_ARCH (-6)      /* device is not sm_100 */
#define HNM_E_DRIVER (-7)    /* driver entry point (tensor map) unavailable */

HNM_API int hnm_abi_version(void);
/* Static string for any value returned by this library. */
HNM_API const char* hnm_strerror(int code);
/* 0 when the current device can run the sm_100a kernels, HNM_E_ARCH otherwise. */
HNM_API int hnm_check_device(void);

/* ------------------------------------------------------------------------
 * LightGCN.set_graph                      src/models/lightgcn.py:81-112,114-134
 *
 * COO edge list (both directions already present, duplicates allowed, int64 as
 * the reference passes them) -> self loops appended -> CSR sorted by (row, col)
 * with duplicates retained, plus dis[i] = deg[i]^-1/2 (inf -> 0) where deg is the
 * weighted ROW sum.  nnz = num_edges + num_nodes.  The per-edge normalised value
 * dis[row]*w*dis[col] of the reference is not stored: the propagate kernels fuse
 * it.  `edge_w` NULL means all ones; then `csr_w` must be NULL too.
 * `heavy_rows` receives the ids of rows with more than `heavy_threshold`
 * entries (capacity num_nodes); their count is written to *num_heavy_host.
 * Synchronises the stream before returning (the count is read back).
 * ---------------------------------------------------------------------- */
HNM_API size_t hnm_graph_build_workspace_bytes(int64_t num_nodes, int64_t num_edges, int weighted);
HNM_API int hnm_graph_build(const int64_t* edge_row, const int64_t* edge_col, const float* edge_w,
                    int64_t num_edges, int64_t num_nodes,
                    int32_t* csr_rowptr /* [num_nodes+1] */, int32_t* csr_col /* [nnz] */,
                    float* csr_w /* [nnz] or NULL */, float* dis /* [num_nodes] */,
                    int32_t heavy_threshold, int32_t* heavy_rows, int32_t* num_heavy_host,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * LightGCN.forward                              src/models/lightgcn.py:147-158
 *
 * hnm_lightgcn_prescale:  xs[i] = dis[i] * e0[i];  acc[i] = alpha0 * e0[i]
 *     (the `final = 0; final += alpha[0] * E_0` step, :156-158, fused with the
 *     degree pre-scaling that lets the gather loop skip per-edge weights).
 * hnm_lightgcn_layer:  for rows in [row_begin, row_end):
 *         e      = dis[i] * sum_{j in row i} w_ij * xs_in[col_j]      (= (A_hat E)_i, :152)
 *         xs_out[i] = dis[i] * e            (skipped when xs_out is NULL: last layer)
 *         acc[i]   += alpha * e             (:158)
 *     Rows listed in heavy_rows (more than heavy_threshold entries) are summed by a whole
 *     thread block, the first num_huge of them (the very long ones, > HNM_HUGE_ROW entries)
 *     by a cluster of 8 thread blocks; all other rows by one warp each.  Row ranges let
 *     several GPUs own disjoint row shards.
 * ---------------------------------------------------------------------- */
HNM_API int hnm_lightgcn_prescale(const float* e0, const float* dis, float alpha0, float* xs, float* acc,
                          int64_t num_rows, int32_t dim, void* stream);
HNM_API int hnm_lightgcn_layer(const int32_t* csr_rowptr, const int32_t* csr_col, const float* csr_w,
                       const float* dis, const float* xs_in, float* xs_out, float* acc, float alpha,
                       int64_t num_nodes, int32_t dim, int64_t row_begin, int64_t row_end,
                       const int32_t* heavy_rows, int32_t num_heavy, int32_t num_huge,
                       int32_t heavy_threshold,
                       int32_t short_rows /* 1: the rows of the range are short on average (user rows): a warp takes 8
                                             consecutive rows and stages all their column indices in shared memory
                                             with one sweep; 0: one row at a time */,
                       void* stream);
#define HNM_HUGE_ROW 8192

/* User-sharded propagation (no reference counterpart; one process per GPU).  A rank that owns the
 * users [u0, u1) computes, for every item row, only the part of the neighbour sum that runs over its
 * own users -- entries [seg_begin[i - row_begin], seg_end[i - row_begin]) of the CSR row, contiguous
 * because columns are sorted -- into `partial` [row_end - row_begin, dim] (raw sums, no
 * normalisation).  The partial sums of all ranks are then added (an all-reduce of the 27 MB item
 * block instead of an all-gather of the 351 MB user block) and hnm_lightgcn_finish applies the self
 * loop, the degree normalisation and the layer sum exactly as hnm_lightgcn_layer's epilogue does:
 *     e = dis_i * (partial_i + xs_in[i]);  xs_out[i] = dis_i * e;  acc[i] += alpha * e
 * heavy_rows / num_huge classify rows by the length of the SUB-range. */
HNM_API int hnm_lightgcn_partial(const int32_t* seg_begin, const int32_t* seg_end, const int32_t* csr_col,
                         const float* csr_w, const float* xs_in, float* partial, int32_t dim,
                         int64_t row_begin, int64_t row_end, const int32_t* heavy_rows, int32_t num_heavy,
                         int32_t num_huge, int32_t heavy_threshold,
                         int32_t accumulate /* 0: partial = sum; 1: partial += sum (see below) */, void* stream);
/* The same pair also serves one GPU: the item rows gather from the user block, which at the H&M shape
 * (351 MB) does not fit the L2, so every user row was fetched from HBM ~6 times per layer.  Walking the
 * item rows once per L2-sized CHUNK of users (sub-ranges again, accumulate = 1 after the first chunk)
 * makes each user row leave HBM once. */
HNM_API int hnm_lightgcn_finish(const float* partial /* [num_rows, dim] */, const float* xs_in /* [N, dim] */,
                        const float* dis, float alpha, float* xs_out /* [N, dim] or NULL */, float* acc /* [N, dim] */,
                        int64_t row_begin, int64_t num_rows, int32_t dim, void* stream);

/* The same exchange without NCCL and without the replicated finish pass (one process per GPU, the ranks'
 * buffers mapped into each other's address space, e.g. torch.distributed._symmetric_memory): item row i is
 * OWNED by rank i / rows_per_owner.  hnm_lightgcn_partial_peer is hnm_lightgcn_partial whose epilogue stores
 * each partial sum straight into the owner's staging buffer [world][rows_per_owner][dim] over NVLink
 * (peer_stage: device array of the `world` staging-buffer addresses as seen from this rank).  After a
 * cross-rank barrier hnm_lightgcn_finish_peer adds the `world` partials of the rows this rank owns in rank
 * order (deterministic), applies self loop / normalisation / layer sum like hnm_lightgcn_finish, and writes the
 * new pre-scaled rows into every rank's xs_out (peer_xs_out, NULL on the last layer) and, when peer_acc is
 * given (last layer), the finished layer-sum rows into every other rank's acc.  A second barrier makes the rows
 * visible before the next layer gathers from them. */
HNM_API int hnm_lightgcn_partial_peer(const int32_t* seg_begin, const int32_t* seg_end, const int32_t* csr_col,
                              const float* csr_w, const float* xs_in, void* const* peer_stage /* device [world] */,
                              int32_t rows_per_owner, int32_t rank, int32_t dim, int64_t row_begin, int64_t row_end,
                              const int32_t* heavy_rows, int32_t num_heavy, int32_t num_huge,
                              int32_t heavy_threshold, void* stream);
HNM_API int hnm_lightgcn_finish_peer(const float* stage /* this rank's [world][rows_per_owner][dim] */, int32_t world,
                             int32_t rows_per_owner, int32_t rank, const float* xs_in /* [N, dim] */,
                             const float* dis, float alpha, void* const* peer_xs_out /* device [world] or NULL */,
                             float* acc /* [N, dim] */, void* const* peer_acc /* device [world] or NULL */,
                             int64_t item_row_begin, int64_t num_items, int32_t dim, void* stream);

/* ------------------------------------------------------------------------
 * LightGCN.predict                              src/models/lightgcn.py:180-184
 * out[b] = dot(user_emb[user_ids[b]], item_emb[item_ids[b]]) in fp32.  Asynchronous on `stream` (no read-back):
 * negative indices count from the end like torch indexing; a pair whose index is out of range gets NaN and
 * raises *out_of_range (device int32 zeroed by the caller, may be NULL).
 * ---------------------------------------------------------------------- */
HNM_API int hnm_pair_scores(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                    const int64_t* item_ids, int64_t batch, int32_t dim, int64_t num_users,
                    int64_t num_items, float* out, int32_t* out_of_range, void* stream);

/* ------------------------------------------------------------------------
 * LightGCN.predict_all_items                    src/models/lightgcn.py:199-202
 * scores[b, j] = dot(user_emb[user_ids[b]], item_emb[j]) for all j, fp32 FMA.
 * user_ids NULL means users 0..batch-1.
 * ---------------------------------------------------------------------- */
HNM_API int hnm_score_all_items(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                        int64_t batch, int64_t num_items, int32_t dim, float* scores /* [batch, num_items] */,
                        void* stream);

/* ------------------------------------------------------------------------
 * LightGCN.recommend, exact reference path      src/models/lightgcn.py:345-356
 *
 * For each listed user: fp64 score of every item in [item_begin, item_end)
 * (products of fp32 inputs are exact in fp64; accumulated k = 0..dim-1 in that
 * order), optional exclusion (scores[i, filter] = -inf, :349-353) and the top-k
 * by (score desc, item id asc).  Never materialises the score matrix.  This is
 * the correctness anchor and the fallback for users the tensor-core path cannot
 * certify.  excl_ptr/excl_items: CSR over the `batch` listed users of excluded
 * GLOBAL item ids (NULL = none).  When fewer than k items remain, the tail is
 * filled with score -inf and the smallest excluded ids (ascending).
 * out_ids are GLOBAL item indices (item_begin is added).
 * item_splits > 1 cuts the item range into that many equal parts handled by different
 * thread blocks (so a launch with few users still fills the GPU); the outputs are then
 * [item_splits, batch, k] partial lists to be combined with hnm_merge_topk.
 * ---------------------------------------------------------------------- */
HNM_API int hnm_topk_exact(const float* user_emb, const float* item_emb, const int64_t* user_ids, int64_t batch,
                   int64_t item_begin, int64_t item_end, int32_t dim, int32_t k,
                   const int64_t* excl_ptr, const int64_t* excl_items, int32_t item_splits,
                   int64_t* out_ids /* [item_splits, batch, k] */, double* out_scores /* same shape */,
                   void* stream);

/* ------------------------------------------------------------------------
 * Fused full-catalog score + top-k (tensor cores)   replaces lightgcn.py:202 + :356
 *
 * Stage 1  hnm_absmax + hnm_score_pack_items / hnm_score_pack_users: fp32 rows -> fp16 rows scaled by a
 *          power of two (row-major [rows_padded, 64], zero padded).  The item shard is centred on a common
 *          vector and gets ONE scale; every user row gets its OWN scale (the ranking of the items for a
 *          fixed user does not depend on that user's scale, so heavy-tailed row norms cost no precision).
 *          All scales stay on the device: no host synchronisation anywhere in the set-up.
 * Stage 2  hnm_score_topk_fused: TMA-fed tcgen05.mma (fp16 x fp16 -> fp32 in TMEM);
 *          the epilogue keeps, per user, every 32-column chunk that holds an approximate score above a
 *          running threshold tau (the kth_sel-th largest of 32 disjoint bucket maxima, a lower bound on
 *          the kth_sel-th best score).  The [users, items] score matrix never exists in memory.  Per user
 *          the kernel emits its nominations into `cand`, two arrays back to back,
 *              uint32 q[num_users * cand_cap][4]   the chunk's 8 group maxima (groups of 4 adjacent items), each
 *                            cut to its upper 16 bits (bf16 truncated toward zero), group 2j in the low half of q[j]
 *              uint32 col[num_users * cand_cap]    LOCAL index of the chunk's first item
 *          i.e. HNM_FUSED_CAND_BYTES bytes per entry.  User r owns entries [r * cand_cap, (r + 1) * cand_cap) of
 *          both arrays, as TWO lists (a row is drained by two threads, one per 64-column half of an item tile):
 *          list 0 starts at entry 0 and holds cand_count[2 r] entries, list 1 starts at cand_cap / 2 and holds
 *          cand_count[2 r + 1]; a count above cand_cap / 2 means overflow (the user is then not certifiable).
 *          When list 1 is empty list 0 may use the whole storage (the merged lists of sliced user tiles).
 *          cand_thresh[r] is the final tau: every item outside the stored chunks scored <= tau.
 * Stage 3  hnm_rescore_topk: exact fp64 scores (k = 0..dim-1 fma chain) of the items of the groups that may
 *          have ended above tau, optional exclusion (lightgcn.py:349-353), canonical
 *          (score desc, id asc) top-k, and a per-user certificate
 *              exact_kth > tau / (su*si) + eps + u.c,
 *              eps = 1.1 * 2^-10 * ||u|| * max_j ||x_j - c|| + dim * 2^-8 / (su*si)
 *          that no non-candidate can belong to the top-k given the fp16 rounding bound.
 *          Only items that can still be in the top-k get the fp64 chain: groups below the k-th best
 *          nominated group (minus 2 eps) are dropped, then an fp32 dot product with a rigorous error
 *          radius drops items below that same bound or below the cut.  The returned scores are the
 *          bits of the fp64 chain (exact fp64 products added in the order k = 0..dim-1).
 * Users whose certificate fails are re-run by the caller: the same three calls with a wider kth_sel,
 * then hnm_topk_exact for what is left.
 * ---------------------------------------------------------------------- */
#define HNM_FUSED_DIM 64            /* K chunk of the tensor-core path; supported embedding dimensions: 64, 128, 256 */
#define HNM_FUSED_USER_TILE 128     /* users per accumulator (UMMA M); users_padded must be a multiple */
#define HNM_FUSED_ITEM_TILE 128     /* items per MMA tile (UMMA N); items_padded must be a multiple */
#define HNM_FUSED_CAND_MAX 1024     /* largest cand_cap (entries per user) */
#define HNM_FUSED_CAND_BYTES 20     /* bytes per candidate entry (16 in the q array + 4 in the col array) */
#define HNM_FUSED_SIG_WORDS 32      /* uint32 words of a user's exclusion signature (1 024 bits) */

/* Column means of a [rows, dim] fp32 table (fp64 sums in a fixed order: deterministic) -> out_mean[dim].
 * The centre of the item shard; replaces item_embeddings.mean(dim=0) in eager PyTorch.  dim 64, 128 or 256;
 * workspace of hnm_column_mean_workspace_bytes(dim) bytes, 8-byte aligned. */
HNM_API int64_t hnm_column_mean_workspace_bytes(int32_t dim);
HNM_API int hnm_column_mean(const float* emb, int64_t rows, int32_t dim, float* out_mean, void* workspace,
                    int64_t workspace_bytes, void* stream);
/* max |x - center[col]| over `count` floats of a [rows, dim] table -> *out_absmax (device float,
 * zeroed by the caller).  center NULL = no centring. */
HNM_API int hnm_absmax(const float* emb, int64_t count, const float* center, int32_t dim, float* out_absmax,
               void* stream);
/* Item shard: out = fp16((emb[row] - center) * s), s = the power of two that puts params[0] (the absmax left
 * there by hnm_absmax) into [2^14, 2^15).  Item tables are centred on their mean row: for a fixed user, u.x
 * and u.(x - c) rank items identically, and the smaller magnitudes tighten the bound.
 * params (device float[4], params[2] zeroed by the caller): in [0] absmax; out [1] = s, [2] = max_j ||x_j - c||^2. */
HNM_API int hnm_score_pack_items(const float* emb, int64_t num_rows, int64_t rows_padded, int32_t dim,
                         const float* center /* [dim] or NULL */, float* params /* device float[4] */,
                         void* out_f16 /* [rows_padded, dim] __half */, void* stream);
/* Users: out[r] = fp16(emb[row_ids[r]] * s_r) with s_r the power of two that puts the row's own absmax into
 * [2^14, 2^15); out_inv_scale[r] = 1 / s_r. */
HNM_API int hnm_score_pack_users(const float* emb, const int64_t* row_ids /* NULL = identity */, int64_t num_rows,
                         int64_t rows_padded, int32_t dim, void* out_f16 /* [rows_padded, dim] __half */,
                         float* out_inv_scale /* [num_rows] */, void* stream);
HNM_API int hnm_score_topk_fused(const void* users_f16 /* [users_padded, dim] */, int64_t num_users, int64_t users_padded,
                         const void* items_f16 /* [items_padded, dim] */, int64_t num_items, int64_t items_padded,
                         int32_t dim /* 64, 128 or 256 */, int32_t kth_sel /* 1..32 */,
                         void* cand /* num_users * cand_cap * HNM_FUSED_CAND_BYTES bytes, 16-byte aligned (layout above) */,
                         int32_t cand_cap /* even */, int32_t* cand_count /* [num_users][2] */,
                         float* cand_thresh /* [num_users] final tau (scaled units) */,
                         const uint32_t* excl_sig /* [num_users][HNM_FUSED_SIG_WORDS] from hnm_exclusion_signature,
                                                     or NULL when nothing is filtered */,
                         void* workspace /* hnm_score_topk_fused_workspace_bytes(), may be NULL when that is 0 */,
                         int64_t workspace_bytes, void* stream);
/* Purchased-item filter on the tensor path (src/models/lightgcn.py:349-353; the serving default,
 * scripts/serve.py:350-352).  A user's excluded items are typically his best-scoring ones, so the nomination
 * threshold must not be built from them: hnm_exclusion_signature marks, per user, the 32-item chunks of the
 * shard that hold an excluded item (bit = chunk index mod 1 024); hnm_score_topk_fused still nominates such
 * chunks but keeps them out of the threshold estimate, and hnm_rescore_topk applies the filter exactly.
 * excl_ptr / excl_items: CSR over the `batch` users of excluded GLOBAL item ids, sorted per user. */
HNM_API int hnm_exclusion_signature(const int64_t* excl_ptr, const int64_t* excl_items, int64_t batch,
                            int64_t item_begin, int64_t num_items_local,
                            uint32_t* out_sig /* [batch][HNM_FUSED_SIG_WORDS] */, void* stream);
/* Scratch for the user tiles that do not fill a whole pass of the persistent grid: their item range is cut
 * into slices handled by different CTAs, with one candidate list per (user, slice) that a second kernel
 * merges into `cand` (at most ~240 MB at the H&M catalog).  < 0 on bad sizes. */
HNM_API int64_t hnm_score_topk_fused_workspace_bytes(int64_t users_padded, int64_t items_padded);
/* Host only: how the launch distributes its work.  out6 = {grid, full passes per CTA (T user tiles x whole
 * catalog each), first left-over user tile, left-over groups of T tiles, item slices per group, T}.  Left-over
 * unit (group j, slice s) = j * slices + s runs on CTA unit % grid (several rounds when that shortens the tail). */
HNM_API int hnm_score_topk_fused_plan(int64_t users_padded, int64_t items_padded, int32_t* out6);
HNM_API int hnm_rescore_topk(const float* user_emb, const float* item_emb /* local shard rows */,
                     const int64_t* user_ids /* NULL = identity */, int64_t batch, int32_t dim /* 64 */,
                     int64_t item_begin,
                     int64_t num_items_local /* rows of item_emb; columns past it are zero padding */,
                     const void* cand /* as written by hnm_score_topk_fused for num_users = batch */, int32_t cand_cap,
                     const int32_t* cand_count, const float* cand_thresh,
                     const float* user_inv_scale /* [batch], from hnm_score_pack_users */,
                     const float* item_params /* device float[4], from hnm_absmax + hnm_score_pack_items */,
                     const float* center /* the item centre used by hnm_score_pack_items, or NULL */,
                     const int64_t* excl_ptr, const int64_t* excl_items /* GLOBAL ids, sorted per user */,
                     int32_t k, int64_t* out_ids /* [batch,k] GLOBAL item ids */, double* out_scores /* [batch,k] */,
                     int32_t* out_certified /* [batch] 1 = provably exact; else reason bits: 2 list overflow,
                                                4 too many groups, 8 fewer than k contenders, 16 too many */,
                     void* stream);

/* ------------------------------------------------------------------------
 * Multi-GPU merge of per-shard exact top-k lists (no reference counterpart;
 * BASELINE.json north_star: item-catalog shards + allgather + merge).
 * in_ids/in_scores: [num_shards, batch, k] (1 <= num_shards <= 64), every list
 * sorted by (score desc, id asc); a short list is padded with (-inf, INT64_MAX).
 * Output: the global top-k in the same order.
 * ---------------------------------------------------------------------- */
HNM_API int hnm_merge_topk(const int64_t* in_ids, const double* in_scores, int32_t num_shards, int64_t batch,
                   int32_t k, int64_t* out_ids, double* out_scores, void* stream);

/* ------------------------------------------------------------------------
 * NeuralCF.forward / predict_all_items          src/models/neural_cf.py:125-139,167-206
 *
 * Eval-mode logits  y = wp . [gu[u] * gi[i] ; MLP([mu[u] ; mi[i]])] + bp  for any MLP depth
 * (Linear+ReLU stack, neural_cf.py:85-90).  Layer 1 is separable over the concatenated input:
 *     W1 [mu ; mi] + b1 = P[u] + Q[i],   P = mu W1[:, :h]^T,   Q = mi W1[:, h:]^T + b1,
 * hnm_ncf_precompute builds P / Q once per weight update (call it twice).  The remaining
 * layers are passed packed as `mlp_tail` = W2 (row-major [w2, w1]), b2, W3, b3, ...;
 * `widths_host` (HOST array, num_layers entries) lists the output width of every MLP
 * layer, first one included (reference default: {64, 32}).  wp = prediction_layer.weight
 * ([mf_dim + last width]), bp its bias.
 * ---------------------------------------------------------------------- */
HNM_API int hnm_ncf_precompute(const float* mlp_emb /* [rows, h] */, int64_t rows, int32_t h,
                       const float* w1 /* [h1, w1_cols] row-major */, int32_t h1, int32_t w1_cols,
                       int32_t col_offset /* 0 for users, h for items */,
                       const float* bias /* [h1] or NULL */, float* out /* [rows, h1] */, void* stream);
HNM_API int hnm_ncf_score_pairs(const float* gmf_user, const float* gmf_item, const float* pu /* [U,h1] */,
                        const float* qi /* [I,h1] */, const float* mlp_tail, const int32_t* widths_host,
                        int32_t num_layers, const float* wp, float bp, const int64_t* user_ids,
                        const int64_t* item_ids, int64_t num_pairs, int32_t mf_dim, float* out, void* stream);
/* out[r, c] for row r = user user_ids[r] (NULL = r) and candidate cand_items[r, c]
 * (NULL = item c, i.e. predict_all_items when cand_per_user == num_items). */
HNM_API int hnm_ncf_score_candidates(const float* gmf_user, const float* gmf_item, const float* pu, const float* qi,
                             const float* mlp_tail, const int32_t* widths_host, int32_t num_layers,
                             const float* wp, float bp, const int64_t* user_ids, int64_t num_rows,
                             const int32_t* cand_items, int32_t cand_per_user, int32_t mf_dim,
                             float* out /* [num_rows, cand_per_user] */, void* stream);

/* ------------------------------------------------------------------------
 * NeuralCF.recommend tail                          src/models/neural_cf.py:316-325
 * Top-k of each row of a materialised fp32 score matrix [batch, num_items] (predict_all_items), after
 * `scores[row, excluded] = -inf`, ordered by (score desc, item id asc); replaces the Python filter loop
 * + torch.topk.  excl_ptr/excl_items: CSR over the rows, item ids sorted per row (NULL = no filter).
 * out_scores may be NULL.  k <= 256.
 * ---------------------------------------------------------------------- */
HNM_API int hnm_topk_dense(const float* scores, int64_t batch, int64_t num_items, const int64_t* excl_ptr,
                   const int64_t* excl_items, int32_t k, int64_t* out_ids /* [batch, k] */,
                   float* out_scores /* [batch, k] or NULL */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HNM_B200_H_ */
