/* Plain C restatement of the scoring hot path of hyunlord/hnm_recommendation.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a second, independent CPU statement of the same
 * arithmetic as oracle/lightgcn_oracle.py and oracle/ncf_oracle.py, written as scalar loops so that nothing
 * about it depends on a tensor library's kernels.  Compiled by oracle/c_oracle.py (gcc -O2 -ffp-contract=off)
 * into oracle/_build/; only tests/ loads it.  Every function cites the reference lines it follows (paths
 * relative to the upstream repository root).  No reference source is copied: the reference is Python.
 *
 * Pinning: tests/test_oracle_c.py checks it against the golden vectors produced by the reference's own files
 * (tests/golden/make_golden.py) and against the PyTorch oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HNM_ORACLE_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------------------------------
 * set_graph: src/models/lightgcn.py:92-112 (+ _add_self_loops, :114-134).
 * in : row/col [m] (both directions already present), w [m] or NULL (-> 1, :92-93), n nodes
 * out: rowptr [n+1], out_col [m+n], out_val [m+n] sorted by (row, col), duplicates kept in input order;
 *      dis [n] = deg^-1/2 with inf -> 0 (:104-105); val = (dis[row] * w) * dis[col] (:106).
 * deg is the ROW sum with multiplicity, added in entry order (edges first, then the n self loops, :127-132).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct { int64_t row, col, pos; float w; } entry_t;

static int cmp_entry(const void* a, const void* b) {
  const entry_t* x = (const entry_t*)a;
  const entry_t* y = (const entry_t*)b;
  if (x->row != y->row) return x->row < y->row ? -1 : 1;
  if (x->col != y->col) return x->col < y->col ? -1 : 1;
  return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);   /* stable: duplicates keep their input order */
}

HNM_ORACLE_API int hnm_oracle_norm_adj(const int64_t* row, const int64_t* col, const float* w, int64_t m, int64_t n,
                                       int64_t* rowptr, int64_t* out_col, float* out_val, float* dis) {
  const int64_t nnz = m + n;
  entry_t* e = (entry_t*)malloc((size_t)nnz * sizeof(entry_t));
  float* deg = (float*)calloc((size_t)n, sizeof(float));
  if (!e || !deg) { free(e); free(deg); return -1; }
  for (int64_t i = 0; i < m; ++i) {
    if (row[i] < 0 || row[i] >= n || col[i] < 0 || col[i] >= n) { free(e); free(deg); return -2; }
    e[i].row = row[i]; e[i].col = col[i]; e[i].w = w ? w[i] : 1.0f; e[i].pos = i;
  }
  for (int64_t i = 0; i < n; ++i) { e[m + i].row = i; e[m + i].col = i; e[m + i].w = 1.0f; e[m + i].pos = m + i; }
  for (int64_t i = 0; i < nnz; ++i) deg[e[i].row] += e[i].w;                          /* :103 */
  for (int64_t i = 0; i < n; ++i) {
    const float d = 1.0f / sqrtf(deg[i]);                                               /* :104 deg.pow(-0.5) */
    dis[i] = isinf(d) ? 0.0f : d;                                                       /* :105 */
  }
  qsort(e, (size_t)nnz, sizeof(entry_t), cmp_entry);                                    /* :109-112 */
  memset(rowptr, 0, (size_t)(n + 1) * sizeof(int64_t));
  for (int64_t i = 0; i < nnz; ++i) {
    out_col[i] = e[i].col;
    out_val[i] = (dis[e[i].row] * e[i].w) * dis[e[i].col];                              /* :106, left to right */
    rowptr[e[i].row + 1]++;
  }
  for (int64_t i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
  free(e); free(deg);
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------
 * forward: src/models/lightgcn.py:147-158.  E_{l+1} = A_hat E_l (each row summed in CSR order, fp32);
 * final = 0; final += alpha_l * E_l for l = 0..L in that order, alpha_l a Python float cast to fp32 at the
 * multiply.  out: final [n, d] (users are rows [0, U), items the rest, :161-162).
 * ------------------------------------------------------------------------------------------------------- */
HNM_ORACLE_API int hnm_oracle_forward(const int64_t* rowptr, const int64_t* col, const float* val, int64_t n, int32_t d,
                                      int32_t num_layers, const double* alphas, const float* e0, float* final) {
  float* cur = (float*)malloc((size_t)n * d * sizeof(float));
  float* nxt = (float*)malloc((size_t)n * d * sizeof(float));
  if (!cur || !nxt) { free(cur); free(nxt); return -1; }
  memcpy(cur, e0, (size_t)n * d * sizeof(float));
  for (int64_t i = 0; i < n * d; ++i) final[i] = 0.0f;                                  /* :156 */
  for (int32_t l = 0; l <= num_layers; ++l) {
    const float a = (float)alphas[l];
    for (int64_t i = 0; i < n * d; ++i) final[i] += a * cur[i];                         /* :157-158 */
    if (l == num_layers) break;
    for (int64_t r = 0; r < n; ++r) {                                                   /* :151-153 */
      float* out = nxt + r * d;
      for (int32_t k = 0; k < d; ++k) out[k] = 0.0f;
      for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) {
        const float v = val[p];
        const float* x = cur + col[p] * d;
        for (int32_t k = 0; k < d; ++k) out[k] += v * x[k];
      }
    }
    float* t = cur; cur = nxt; nxt = t;
  }
  free(cur); free(nxt);
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------
 * recommend: src/models/lightgcn.py:199-202 (scores), :349-353 (filter -> -inf), :356 (top-k), with the
 * canonical order of BASELINE.json: score descending, item id ascending; scores are the exact dot products
 * of the fp32 embeddings accumulated in fp64 for k = 0..d-1 (each fp32 x fp32 product is exact in fp64).
 * excl_ptr [b+1] / excl_items: per listed user the item ids to filter (any order); NULL = no filter.
 * item_bias [num_items] fp32 or NULL: MatrixFactorization's b_i, added in fp64 after the chain
 * (src/models/matrix_factorization.py:108-131; b_u and the global bias are the same for every item of a user
 * and cannot change its ranking).
 * out_ids [b, k] int64 item indices, out_scores [b, k] fp64.  Returns -3 when k exceeds the catalog.
 * ------------------------------------------------------------------------------------------------------- */
static int before(double sa, int64_t ia, double sb, int64_t ib) { return sa > sb || (sa == sb && ia < ib); }

HNM_ORACLE_API int hnm_oracle_topk_exact(const float* user_emb, const float* item_emb, const int64_t* user_ids, int64_t b,
                                         int64_t num_items, int32_t d, int32_t k, const int64_t* excl_ptr,
                                         const int64_t* excl_items, const float* item_bias, int64_t* out_ids,
                                         double* out_scores) {
  if (k < 1 || k > num_items) return -3;                                                /* what torch.topk raises */
  double* s = (double*)malloc((size_t)num_items * sizeof(double));
  if (!s) return -1;
  for (int64_t r = 0; r < b; ++r) {
    const float* u = user_emb + user_ids[r] * d;
    for (int64_t j = 0; j < num_items; ++j) {
      const float* v = item_emb + j * d;
      double acc = 0.0;
      for (int32_t q = 0; q < d; ++q) acc += (double)u[q] * (double)v[q];
      s[j] = item_bias ? acc + (double)item_bias[j] : acc;
    }
    if (excl_ptr)
      for (int64_t p = excl_ptr[r]; p < excl_ptr[r + 1]; ++p)
        if (excl_items[p] >= 0 && excl_items[p] < num_items) s[excl_items[p]] = -INFINITY;
    int64_t* ids = out_ids + r * k;
    double* sc = out_scores + r * k;
    int32_t have = 0;
    for (int64_t j = 0; j < num_items; ++j) {                 /* insertion into the sorted head, ids ascending */
      if (have == k && !before(s[j], j, sc[k - 1], ids[k - 1])) continue;
      int32_t pos = have < k ? have : k - 1;
      while (pos > 0 && before(s[j], j, sc[pos - 1], ids[pos - 1])) { sc[pos] = sc[pos - 1]; ids[pos] = ids[pos - 1]; --pos; }
      sc[pos] = s[j]; ids[pos] = j;
      if (have < k) ++have;
    }
  }
  free(s);
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------
 * NeuralCF.forward, eval mode: src/models/neural_cf.py:125-139 (layer stack :85-90: Linear, ReLU, Dropout per
 * hidden width after the first entry of mlp_dims; dropout is the identity in eval mode).
 *   g = Gu[u] * Gi[i];  x = [Mu[u]; Mi[i]];  h = relu(W_l h + b_l) ...;  y = wp . [g; h] + bp  (logit)
 * dims [num_linear + 1]: dims[0] = 2 * mlp_emb (the concatenation), dims[l + 1] = out width of Linear l;
 * weights / biases: the num_linear matrices [out, in] row-major and vectors, back to back.
 * ------------------------------------------------------------------------------------------------------- */
HNM_ORACLE_API int hnm_oracle_ncf_forward(const float* gmf_user, const float* gmf_item, const float* mlp_user,
                                          const float* mlp_item, int32_t mf_dim, int32_t mlp_emb, int32_t num_linear,
                                          const int32_t* dims, const float* weights, const float* biases,
                                          const float* pred_w, float pred_b, const int64_t* user_ids,
                                          const int64_t* item_ids, int64_t b, float* out) {
  int32_t widest = 0;
  for (int32_t l = 0; l <= num_linear; ++l) if (dims[l] > widest) widest = dims[l];
  if (dims[0] != 2 * mlp_emb) return -2;
  float* h0 = (float*)malloc((size_t)widest * sizeof(float));
  float* h1 = (float*)malloc((size_t)widest * sizeof(float));
  if (!h0 || !h1) { free(h0); free(h1); return -1; }
  for (int64_t r = 0; r < b; ++r) {
    const int64_t u = user_ids[r], it = item_ids[r];
    for (int32_t q = 0; q < mlp_emb; ++q) {                                             /* :131-133 */
      h0[q] = mlp_user[u * mlp_emb + q];
      h0[mlp_emb + q] = mlp_item[it * mlp_emb + q];
    }
    const float* w = weights;
    const float* bias = biases;
    float* in = h0;
    float* nx = h1;
    for (int32_t l = 0; l < num_linear; ++l) {                                          /* :134 */
      const int32_t ni = dims[l], no = dims[l + 1];
      for (int32_t o = 0; o < no; ++o) {
        float acc = 0.0f;
        for (int32_t q = 0; q < ni; ++q) acc += w[o * ni + q] * in[q];
        acc += bias[o];
        nx[o] = acc > 0.0f ? acc : 0.0f;
      }
      w += (size_t)no * ni; bias += no;
      float* t = in; in = nx; nx = t;
    }
    float y = 0.0f;                                                                     /* :137-138 */
    for (int32_t q = 0; q < mf_dim; ++q) y += pred_w[q] * (gmf_user[u * mf_dim + q] * gmf_item[it * mf_dim + q]);   /* :127-129 */
    for (int32_t q = 0; q < dims[num_linear]; ++q) y += pred_w[mf_dim + q] * in[q];
    out[r] = y + pred_b;
  }
  free(h0); free(h1);
  return 0;
}
