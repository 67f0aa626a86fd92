"""CPU restatement of LightGCN's inference path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference lines it follows; paths are relative to
the upstream repository root (hyunlord/hnm_recommendation).

Conventions restated from the reference:
  * nodes [0, U) are users, [U, U+I) are items          src/models/lightgcn.py:70,161-162
  * edge_index is [2, M] int64 and already holds both directions
                                                         tests/test_models.py:178-185
  * duplicate edges are kept and contribute separately   src/models/lightgcn.py:102-112
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import warnings

import torch


def layer_weights(num_layers: int, alpha: Optional[float] = None) -> List[float]:
    """Layer-combination weights as Python floats.  src/models/lightgcn.py:59-67."""
    if alpha is None:
        return [1.0 / (num_layers + 1)] * (num_layers + 1)
    w = [alpha ** i for i in range(num_layers + 1)]
    s = sum(w)
    return [a / s for a in w]


def add_self_loops(edge_index: torch.Tensor, edge_weight: torch.Tensor,
                   num_nodes: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Append (i, i, 1) for every node.  src/models/lightgcn.py:114-134."""
    loop = torch.arange(num_nodes, dtype=edge_index.dtype)
    loop = torch.stack([loop, loop])
    return (torch.cat([edge_index, loop], dim=1),
            torch.cat([edge_weight, torch.ones(num_nodes, dtype=edge_weight.dtype)]))


def build_norm_adj(edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor],
                   num_nodes: int, dtype: torch.dtype = torch.float32):
    """Normalised adjacency  D~^-1/2 (A+I) D~^-1/2  as sorted COO + CSR row pointer.

    src/models/lightgcn.py:92-112.  ``deg`` is the ROW sum with multiplicity
    (:103, read as scatter-add -- SURVEY.md F4); ``deg.pow(-0.5)`` with inf -> 0
    (:104-105); value = (dis[row] * w) * dis[col] in that association (:106).
    The (row, col) sort mirrors what torch_sparse's SparseTensor storage does
    with the triplets it is handed (:109-112); duplicates are retained.
    Returns (rowptr[N+1] int64, col[nnz] int64, val[nnz] dtype, dis[N] dtype).
    """
    edge_index = edge_index.to(torch.int64).cpu()
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype)          # :92-93
    else:
        edge_weight = edge_weight.detach().cpu().to(dtype)
    edge_index, edge_weight = add_self_loops(edge_index, edge_weight, num_nodes)  # :97-99
    row, col = edge_index[0], edge_index[1]                                  # :102
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, row, edge_weight)  # :103
    dis = deg.pow(-0.5)                                                      # :104
    dis[torch.isinf(dis)] = 0                                                # :105
    val = dis[row] * edge_weight * dis[col]                                  # :106
    order = torch.argsort(row * num_nodes + col, stable=True)                # :109-112
    row, col, val = row[order], col[order], val[order]
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=num_nodes), 0)
    return rowptr, col, val, dis


def _spmm(rowptr, col, val, x):
    """(A x)_i = sum_{e in row i} val_e * x[col_e].  src/models/lightgcn.py:152."""
    n = rowptr.numel() - 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = torch.sparse_csr_tensor(rowptr, col, val, size=(n, n), dtype=val.dtype,
                                    check_invariants=False)
        return a @ x


def propagate(e0: torch.Tensor, rowptr, col, val, num_layers: int,
              alphas: Sequence[float]) -> torch.Tensor:
    """L rounds of E <- A_hat E, then the sequential alpha-weighted sum.

    src/models/lightgcn.py:147-158: ``final = zeros; final += alpha[i] * E_i``
    in order i = 0..L (alpha[i] is a Python float multiplied into a tensor of
    e0's dtype).
    """
    embs = [e0]
    e = e0
    for _ in range(num_layers):                       # :151-153
        e = _spmm(rowptr, col, val, e)
        embs.append(e)
    final = torch.zeros_like(e0)                      # :156
    for i, x in enumerate(embs):                      # :157-158
        final += alphas[i] * x
    return final


def forward(e0, rowptr, col, val, num_users: int, num_layers: int, alphas):
    """(user_embeddings, item_embeddings) as views of one buffer.  :136-164."""
    final = propagate(e0, rowptr, col, val, num_layers, alphas)
    return final[:num_users], final[num_users:]       # :161-162


def bpr_loss(e0, rowptr, col, val, num_users: int, num_layers: int, alphas, user_ids, pos_item_ids,
             neg_item_ids, weight_decay: float):
    """src/models/lightgcn.py:206-245, differentiable w.r.t. ``e0`` through the dense restatement of the
    propagation (small graphs only: A_hat is densified so that plain autograd applies)."""
    n = rowptr.numel() - 1
    a = torch.sparse_csr_tensor(rowptr, col, val.to(e0.dtype), size=(n, n), check_invariants=False).to_dense()
    embs, e = [e0], e0
    for _ in range(num_layers):                       # :151-153
        e = a @ e
        embs.append(e)
    final = torch.zeros_like(e0)
    for i, x in enumerate(embs):                      # :156-158
        final = final + alphas[i] * x
    ue, ie = final[:num_users], final[num_users:]
    u0, p0, n0 = e0[user_ids], e0[pos_item_ids + num_users], e0[neg_item_ids + num_users]   # :218-220
    pos = (ue[user_ids] * ie[pos_item_ids]).sum(dim=1)                                      # :225-232
    neg = (ue[user_ids] * ie[neg_item_ids]).sum(dim=1)
    loss = -torch.log(torch.sigmoid(pos - neg) + 1e-10).mean()                              # :235
    reg = weight_decay * (u0.norm(2).pow(2) + p0.norm(2).pow(2) + n0.norm(2).pow(2)) / u0.size(0)   # :238-243
    return loss + reg


def predict(user_emb, item_emb, user_ids, item_ids):
    """Row-wise dot product.  src/models/lightgcn.py:180-184."""
    return (user_emb[user_ids] * item_emb[item_ids]).sum(dim=1)


def predict_all_items(user_emb, item_emb, user_ids):
    """scores = U[user_ids] @ V^T.  src/models/lightgcn.py:199-202."""
    return torch.matmul(user_emb[user_ids], item_emb.t())


def apply_filter(scores: torch.Tensor, user_ids: torch.Tensor,
                 filter_items: Optional[Dict[int, set]]) -> torch.Tensor:
    """scores[i, filter_items[uid]] = -inf.  src/models/lightgcn.py:349-353."""
    if filter_items is not None:
        for i, uid in enumerate(user_ids.tolist()):
            if uid in filter_items:
                idx = list(filter_items[uid])
                if idx:
                    scores[i, idx] = float("-inf")
    return scores


def topk_canonical(scores: torch.Tensor, k: int) -> torch.Tensor:
    """Top-k item indices ordered by (score desc, item id asc).

    The reference calls ``torch.topk(scores, k, dim=1)`` (src/models/lightgcn.py:356)
    whose tie order is implementation-defined; BASELINE.json fixes
    the tie-break to item id ascending, which a stable descending sort provides.
    """
    if k > scores.size(1):
        raise RuntimeError("selected index k out of range")   # what torch.topk raises
    order = torch.sort(scores, dim=1, descending=True, stable=True).indices
    return order[:, :k].contiguous()


def exact_scores_fp64(user_emb: torch.Tensor, item_emb: torch.Tensor,
                      user_ids: torch.Tensor) -> torch.Tensor:
    """fp64 scores of the fp32 embeddings, accumulated k = 0..d-1 in that order.

    Each fp32*fp32 product is exact in fp64, so ``s += u_k * v_k`` in fp64 is
    one rounding per k -- the same value an fp64 ``fma`` chain produces on the
    GPU.  This is the canonical score that decides near-ties: it is the
    ordering of the real-valued dot products that the reference's fp32 sgemm
    (src/models/lightgcn.py:202) approximates with an unspecified summation order.
    """
    u = user_emb[user_ids].to(torch.float64)
    v = item_emb.to(torch.float64)
    s = torch.zeros(u.size(0), v.size(0), dtype=torch.float64)
    for k in range(u.size(1)):
        s += u[:, k:k + 1] * v[:, k].unsqueeze(0)
    return s


def recommend(user_emb, item_emb, user_ids, k: int,
              filter_items: Optional[Dict[int, set]] = None) -> torch.Tensor:
    """Reference-faithful recommend(): fp32 sgemm scores + canonical top-k.  :332-357."""
    scores = predict_all_items(user_emb, item_emb, user_ids).clone()
    scores = apply_filter(scores, user_ids, filter_items)
    return topk_canonical(scores, k)


def recommend_exact(user_emb, item_emb, user_ids, k: int,
                    filter_items: Optional[Dict[int, set]] = None,
                    chunk: int = 2048):
    """Canonical recommend(): fp64 sequential-k scores + (score desc, id asc).

    Returns (ids[B,k] int64, scores[B,k] float64).
    """
    ids, vals = [], []
    for s0 in range(0, user_ids.numel(), chunk):
        uids = user_ids[s0:s0 + chunk]
        s = exact_scores_fp64(user_emb, item_emb, uids)
        s = apply_filter(s, uids, filter_items)
        top = topk_canonical(s, k)
        ids.append(top)
        vals.append(torch.gather(s, 1, top))
    return torch.cat(ids), torch.cat(vals)


class LightGCNOracle:
    """Object form with the reference's method names (src/models/lightgcn.py:13-357)."""

    def __init__(self, num_users, num_items, embedding_dim=64, num_layers=3,
                 top_k=12, alpha=None, weight: Optional[torch.Tensor] = None,
                 dtype=torch.float32):
        self.num_users, self.num_items = num_users, num_items
        self.num_nodes = num_users + num_items
        self.embedding_dim, self.num_layers, self.top_k = embedding_dim, num_layers, top_k
        self.alpha = layer_weights(num_layers, alpha)
        self.dtype = dtype
        if weight is None:                                      # :70-71
            weight = torch.empty(self.num_nodes, embedding_dim)
            torch.nn.init.xavier_uniform_(weight)
        self.weight = weight.detach().cpu().to(dtype)
        self.graph = None

    def set_graph(self, edge_index, edge_weight=None):
        self.graph = build_norm_adj(edge_index, edge_weight, self.num_nodes, self.dtype)

    def forward(self):
        if self.graph is None:
            raise RuntimeError("Graph not set. Call set_graph() first.")   # :143-144
        rowptr, col, val, _ = self.graph
        return forward(self.weight, rowptr, col, val, self.num_users,
                       self.num_layers, self.alpha)

    def predict(self, user_ids, item_ids):
        u, v = self.forward()
        return predict(u, v, user_ids, item_ids)

    def predict_all_items(self, user_ids):
        u, v = self.forward()
        return predict_all_items(u, v, user_ids)

    def recommend(self, user_ids, filter_items=None, k=None):
        u, v = self.forward()
        return recommend(u, v, user_ids, k or self.top_k, filter_items)
