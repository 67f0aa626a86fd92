"""LightGCN with the reference's model API, computed by the B200 kernels.

Drop-in for ``src.models.LightGCN`` (src/models/lightgcn.py:13-357) on the
inference path: same constructor, attributes, ``state_dict`` key
(``embeddings.weight``) and method signatures.  ``set_graph`` / ``forward`` /
``predict`` / ``predict_all_items`` / ``recommend`` run hand-written sm_100a
kernels through the C ABI in include/hnm_b200.h; there is no CPU fallback, so
the module must be moved to a CUDA device before they are called.

Differences from the reference, all deliberate:
  * ties in ``recommend`` are ordered by item id ascending (the reference's
    ``torch.topk`` order is unspecified; BASELINE.json fixes this rule);
  * ``recommend`` ranks by the exact (fp64-accumulated) dot products of the fp32
    embeddings, i.e. the ordering the reference's fp32 sgemm approximates;
  * in eval mode ``forward()`` is cached until the weights or the graph change
    (the reference recomputes it on every call) -- results are identical.  The cache key is
    the parameter's storage and version counter: a write through ``.data`` bumps neither, so call
    ``invalidate()`` after one (``load_state_dict`` and ``.to()`` / ``.cuda()`` do it themselves), or set
    ``cache_embeddings = "verify"`` to have every hit confirmed by a device-side checksum of the table.
    In train mode nothing is cached;
  * in train mode, when autograd is recording and the table requires grad, ``forward()`` IS ``forward_with_grad()``
    (differentiable as in the reference; what ``bpr_loss`` / ``training_step`` use, lightgcn.py:206-265):
    its backward pass is the same propagate kernels on the transposed adjacency,
    ``dL/dE0 = sum_l alpha_l (A_hat^T)^l dL/dfinal``, and ``predict`` / ``predict_all_items`` then score with
    plain differentiable torch ops; under ``no_grad`` or in eval mode the kernels and the cached buffer serve;
  * ``recommend(user_ids, filter_items=None, k=None)`` accepts ``k`` (README.md:131).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import engine
from .base import ModelBase
from .metrics import RecommendationMetrics


SMALL_BATCH = 128


class _Propagate(torch.autograd.Function):
    """final = sum_l alpha_l A_hat^l E0 and its adjoint, both by engine.propagate (lightgcn.py:147-158)."""

    @staticmethod
    def forward(ctx, weight: torch.Tensor, model: "LightGCN") -> torch.Tensor:
        ctx.model = model
        return engine.propagate(model.graph, weight, model.alpha, model.num_layers,
                                item_chunks=getattr(model, "_item_chunks", None))

    @staticmethod
    def backward(ctx, grad_final: torch.Tensor):
        m = ctx.model
        grad = engine.propagate(m._adjoint_graph(), grad_final.contiguous(), m.alpha, m.num_layers)
        return grad, None


class LightGCN(ModelBase):
    def __init__(
        self,
        num_users: int,
        num_items: int,
        embedding_dim: int = 64,
        num_layers: int = 3,
        learning_rate: float = 0.001,
        weight_decay: float = 1e-4,
        top_k: int = 12,
        alpha: Optional[float] = None,
    ):
        super().__init__()
        self.save_hyperparameters()

        self.num_users = num_users
        self.num_items = num_items
        self.num_nodes = num_users + num_items
        self.embedding_dim = embedding_dim
        self.num_layers = num_layers
        self.learning_rate = learning_rate
        self.weight_decay = weight_decay
        self.top_k = top_k

        # layer-combination weights (lightgcn.py:59-67)
        if alpha is None:
            self.alpha = [1.0 / (num_layers + 1)] * (num_layers + 1)
        else:
            self.alpha = [alpha ** i for i in range(num_layers + 1)]
            alpha_sum = sum(self.alpha)
            self.alpha = [a / alpha_sum for a in self.alpha]

        # one table for users then items (lightgcn.py:70-71)
        self.embeddings = nn.Embedding(self.num_nodes, embedding_dim)
        nn.init.xavier_uniform_(self.embeddings.weight)

        self.graph: Optional[engine.Graph] = None
        self.edge_index = None
        self.edge_weight = None
        self.metrics = RecommendationMetrics(top_k=top_k)

        self.cache_embeddings = True
        self._graph_t = None
        self._cache_key = None
        self._cache_val: Optional[torch.Tensor] = None
        self._cache_sum = None
        self._scorer = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    # ------------------------------------------------------------------ graph
    def set_graph(self, edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor] = None) -> None:
        """lightgcn.py:81-112.  The CSR lives on the device the parameters are on."""
        self.edge_index = edge_index
        self.edge_weight = edge_weight
        self.graph = engine.build_graph(edge_index, edge_weight, self.num_nodes, self.embeddings.weight.device)
        if engine.is_bipartite(self.graph, self.num_users):
            self.graph.short_prefix = self.num_users        # user rows: the staged short-row kernel
        # item rows gather from the user block; when that does not fit the L2 they are walked chunk by chunk
        self._item_chunks = engine.make_item_chunks(self.graph, self.num_users, self.num_items, self.embedding_dim)
        self._cache_key = None
        self._graph_t = None

    def _adjoint_graph(self) -> engine.Graph:
        """CSR of A_hat^T: the transposed edge list with the ORIGINAL deg^-1/2 (A_hat^T_ji = dis_i w_ij dis_j).
        For the symmetric edge lists the reference is fed (tests/test_models.py:182-185) it is the graph itself."""
        if self._graph_t is None:
            g = self.graph
            ei = self.edge_index.to(g.rowptr.device)
            t = engine.build_graph(ei.flip(0), self.edge_weight, self.num_nodes, g.rowptr.device)
            same = torch.equal(t.rowptr, g.rowptr) and torch.equal(t.col, g.col) and (
                (t.w is None and g.w is None) or (t.w is not None and g.w is not None and torch.equal(t.w, g.w)))
            if same:
                self._graph_t = g
            else:
                t.dis = g.dis
                self._graph_t = t
        return self._graph_t

    # ---------------------------------------------------------------- forward
    def invalidate(self) -> None:
        """Drop the cached propagated embeddings and the packed scorer tables.  Needed after a write to the
        parameters that autograd's version counter does not see (``weight.data.copy_(...)``)."""
        self._cache_key = None
        self._cache_val = None
        self._cache_sum = None
        self._scorer = None

    def _apply(self, fn, *args, **kwargs):
        self.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def _final_embeddings(self) -> torch.Tensor:
        if self.graph is None:
            raise RuntimeError("Graph not set. Call set_graph() first.")
        w = self.embeddings.weight
        if self.graph.rowptr.device != w.device:
            raise RuntimeError("graph and parameters are on different devices; call set_graph() after .to(device)")
        key = (w.data_ptr(), w._version, id(self.graph), tuple(self.alpha), self.num_layers)
        use_cache = bool(self.cache_embeddings) and not self.training
        if use_cache and self._cache_key == key and self._cache_val is not None:
            if self.cache_embeddings != "verify" or float(w.detach().sum(dtype=torch.float64)) == self._cache_sum:
                return self._cache_val
        final = engine.propagate(self.graph, w, self.alpha, self.num_layers,
                                 item_chunks=getattr(self, "_item_chunks", None))
        self._cache_key, self._cache_val = (key, final) if use_cache else (None, None)
        self._cache_sum = float(w.detach().sum(dtype=torch.float64)) if self.cache_embeddings == "verify" else None
        self._scorer = None
        return final

    def forward(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """lightgcn.py:136-164: (user_embeddings [U,d], item_embeddings [I,d]), views of one buffer."""
        if self._wants_grad():
            return self.forward_with_grad()
        final = self._final_embeddings()
        return final[: self.num_users], final[self.num_users:]

    def _wants_grad(self) -> bool:
        """Train mode with autograd recording: the reference's differentiable behaviour.  Under ``no_grad`` or in
        eval mode the (cached, detached) kernel results are returned."""
        return (self.training and torch.is_grad_enabled() and self.embeddings.weight.requires_grad
                and self.graph is not None)

    def forward_with_grad(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """The reference's forward() as it behaves under autograd (lightgcn.py:136-164): gradients flow to
        ``embeddings.weight`` through the propagate kernels.  Not cached."""
        if self.graph is None:
            raise RuntimeError("Graph not set. Call set_graph() first.")
        final = _Propagate.apply(self.embeddings.weight, self)
        return final[: self.num_users], final[self.num_users:]

    def bpr_loss(self, user_ids: torch.Tensor, pos_item_ids: torch.Tensor, neg_item_ids: torch.Tensor) -> torch.Tensor:
        """lightgcn.py:206-245: BPR loss on the propagated embeddings + L2 on the layer-0 rows of the batch."""
        user_embeds_0 = self.embeddings(user_ids)
        pos_item_embeds_0 = self.embeddings(pos_item_ids + self.num_users)
        neg_item_embeds_0 = self.embeddings(neg_item_ids + self.num_users)
        user_embeddings, item_embeddings = self.forward_with_grad()
        user_embeds = user_embeddings[user_ids]
        pos_scores = (user_embeds * item_embeddings[pos_item_ids]).sum(dim=1)
        neg_scores = (user_embeds * item_embeddings[neg_item_ids]).sum(dim=1)
        loss = -torch.log(torch.sigmoid(pos_scores - neg_scores) + 1e-10).mean()
        reg_loss = self.weight_decay * (
            user_embeds_0.norm(2).pow(2) + pos_item_embeds_0.norm(2).pow(2) + neg_item_embeds_0.norm(2).pow(2)
        ) / user_embeds_0.size(0)
        return loss + reg_loss

    def training_step(self, batch: Dict[str, Any], batch_idx: int) -> torch.Tensor:
        """lightgcn.py:247-265."""
        loss = self.bpr_loss(batch["user_ids"], batch["pos_items"], batch["neg_items"])
        self.log("train_loss", loss, prog_bar=True)
        return loss

    def configure_optimizers(self):
        """lightgcn.py:307-330 (Adam, no optimizer weight decay, ReduceLROnPlateau on val_map_at_k)."""
        optimizer = torch.optim.Adam(self.parameters(), lr=self.learning_rate, weight_decay=0.0)
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="max", factor=0.5, patience=5)
        return {"optimizer": optimizer,
                "lr_scheduler": {"scheduler": scheduler, "monitor": "val_map_at_k", "frequency": 1}}

    def predict(self, user_ids: torch.Tensor, item_ids: torch.Tensor) -> torch.Tensor:
        """lightgcn.py:166-186."""
        ue, ie = self.forward()
        if self._wants_grad():
            return (ue[user_ids.to(ue.device)] * ie[item_ids.to(ue.device)]).sum(dim=1)      # lightgcn.py:180-184
        return engine.pair_scores(ue, ie, user_ids, item_ids)

    def predict_all_items(self, user_ids: torch.Tensor) -> torch.Tensor:
        """lightgcn.py:188-204: [batch, num_items] fp32 scores."""
        ue, ie = self.forward()
        if self._wants_grad():
            return torch.matmul(ue[user_ids.to(ue.device)], ie.t())                           # lightgcn.py:199-202
        return engine.score_all_items(ue, ie, user_ids)

    # -------------------------------------------------------------- recommend
    def recommend(self, user_ids: torch.Tensor, filter_items: Optional[Dict[int, set]] = None,
                  k: Optional[int] = None) -> torch.Tensor:
        """lightgcn.py:332-357: top-k item indices [batch, k] int64, (score desc, item id asc)."""
        self.eval()
        k = self.top_k if k is None else int(k)
        if k > self.num_items or k <= 0:
            raise RuntimeError("selected index k out of range")
        with torch.no_grad():
            ue, ie = self.forward()
            dev = ue.device
            uids = user_ids.to(dev).view(-1)
            if k > engine.EXACT_K_MAX:
                return self._recommend_by_sort(ue, ie, uids, filter_items, k)
            from .scorer import FusedScorer
            # a serving-sized batch (scripts/serve.py scores one user per request) is faster through the
            # exact kernel, which splits the catalog over thread blocks: 0.24 ms vs 0.67 ms for one user,
            # break-even near 256 users (tools/probe_latency.py)
            if uids.numel() > SMALL_BATCH and FusedScorer.supports(self.embedding_dim, k, self.num_items):
                if self._scorer is None:
                    self._scorer = FusedScorer(ue, ie)
                ids, _ = self._scorer.topk(uids, k, filter_items)
                return ids
            excl = engine.exclusion_csr(uids, filter_items, dev)
            ids, _ = engine.topk_exact(ue, ie, uids, k, excl)
        return ids

    def recommend_all(self, k: Optional[int] = None, return_scores: bool = False,
                      out_host: Optional[torch.Tensor] = None, filter_purchased: bool = False):
        """Full-catalog top-k for every user (the BASELINE.json headline path): [num_users, k].

        out_host: optional pinned int64 [num_users, k] tensor that receives the ids as they are produced
        (device-to-host copies overlap the scoring of the following users); complete when the call returns.
        filter_purchased: drop every user's own purchases, the serving default of the reference
        (scripts/serve.py:174-177,350-352); the exclusion lists come straight from the graph on the device."""
        self.eval()
        k = self.top_k if k is None else int(k)
        if k > self.num_items or k <= 0:
            raise RuntimeError("selected index k out of range")
        with torch.no_grad():
            ue, ie = self.forward()
            from .scorer import FusedScorer
            hist = engine.history_csr(self.graph, self.num_users) if filter_purchased else None
            if FusedScorer.supports(self.embedding_dim, k, self.num_items):
                if self._scorer is None:
                    self._scorer = FusedScorer(ue, ie)
                ids, sc = self._scorer.topk(None, k, hist, out_host=out_host)
            else:
                ids, sc = engine.topk_exact(ue, ie, None, k, hist if hist is not None else (None, None))
                if out_host is not None:
                    out_host.copy_(ids)
        return (ids, sc) if return_scores else ids

    def _recommend_by_sort(self, ue, ie, uids, filter_items, k):
        # k beyond the select kernels' capacity: materialise the rows and stable-sort them on the device
        scores = engine.score_all_items(ue, ie, uids).double()
        if filter_items is not None:
            for i, uid in enumerate(uids.tolist()):
                if uid in filter_items:
                    scores[i, list(filter_items[uid])] = float("-inf")
        return torch.sort(scores, dim=1, descending=True, stable=True).indices[:, :k].contiguous()

    # ------------------------------------------------- Lightning-facing hooks
    def validation_step(self, batch: Dict[str, Any], batch_idx: int):
        """lightgcn.py:267-284 with the fused top-k instead of predict_all_items + torch.topk."""
        top_k_items = self.recommend(batch["user_ids"])
        self.metrics.update(top_k_items.cpu(), batch["ground_truth"])

    def on_validation_epoch_end(self):
        metrics = self.metrics.compute()
        self.metrics.reset()
        for name, value in metrics.items():
            self.log(f"val_{name}", value, prog_bar=True)

    def test_step(self, batch: Dict[str, Any], batch_idx: int):
        self.validation_step(batch, batch_idx)

    def on_test_epoch_end(self):
        metrics = self.metrics.compute()
        self.metrics.reset()
        for name, value in metrics.items():
            self.log(f"test_{name}", value)
