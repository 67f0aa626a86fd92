"""MatrixFactorization with the reference's model API, ranked by the B200 kernels.

Drop-in for ``src.models.MatrixFactorization`` (src/models/matrix_factorization.py:10-245): same constructor,
sub-module names / ``state_dict`` keys (``user_embeddings``, ``item_embeddings``, ``user_bias``, ``item_bias``,
``global_bias``), ``forward`` / ``predict_all_items`` / ``recommend`` / ``training_step`` /
``configure_optimizers``.  SURVEY.md section 8 (f2): ``U V^T + b_u + b_i + b_g`` through the same fused
score / select kernel as LightGCN.

How the bias gets onto the tensor cores: ``b_u`` and ``b_g`` are the same for every item of a user and cannot
change his ranking; ``b_i`` can.  The tables are extended by one 64-wide K chunk,
    u' = [u, 1, 0, ...],   x' = [x, b_i, 0, ...]      so that  u'.x' = u.x + b_i,
(with a power of two beta: u' = [u, beta], x' = [x, b_i / beta], exact either way; beta balances the norms of
the two extended tables, which is what the certificate's error bound is proportional to)
and scored by the fused kernel at d' = 128 (d <= 64) or 256 (d <= 192) -- the chunked-K path of
csrc/score_fused.cu; nomination, exact fp64 rescoring (chain over u_k x_k, then + b_i), certificate and
fallback are the LightGCN ones.  Ties are ordered by item id ascending (the reference's ``torch.topk`` leaves it
unspecified).  ``forward`` and ``training_step`` are the reference's plain differentiable formulation.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from . import engine
from .base import ModelBase
from .metrics import RecommendationMetrics

SMALL_BATCH = 128


class MatrixFactorization(ModelBase):
    def __init__(
        self,
        num_users: int,
        num_items: int,
        embedding_dim: int = 64,
        learning_rate: float = 0.001,
        weight_decay: float = 0.01,
        top_k: int = 12,
        sparse: bool = True,
    ):
        super().__init__()
        self.save_hyperparameters()
        self.num_users = num_users
        self.num_items = num_items
        self.embedding_dim = embedding_dim
        self.learning_rate = learning_rate
        self.weight_decay = weight_decay
        self.top_k = top_k
        self.sparse = sparse
        self.user_embeddings = nn.Embedding(num_users, embedding_dim, sparse=sparse)      # :49-59
        self.item_embeddings = nn.Embedding(num_items, embedding_dim, sparse=sparse)
        self.user_bias = nn.Embedding(num_users, 1)                                        # :62-63
        self.item_bias = nn.Embedding(num_items, 1)
        self.global_bias = nn.Parameter(torch.zeros(1))                                    # :66
        self._init_weights()
        self.metrics = RecommendationMetrics(top_k=top_k)
        self._aug_key = None
        self._aug = None
        self._scorer = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def _init_weights(self) -> None:
        nn.init.normal_(self.user_embeddings.weight, std=0.01)                             # :74-79
        nn.init.normal_(self.item_embeddings.weight, std=0.01)
        nn.init.zeros_(self.user_bias.weight)
        nn.init.zeros_(self.item_bias.weight)

    def invalidate(self) -> None:
        """Drop the cached extended tables (needed after writing parameters through ``.data``)."""
        self._aug_key = self._aug = self._scorer = None

    def _apply(self, fn, *args, **kwargs):
        self.invalidate()
        return super()._apply(fn, *args, **kwargs)

    # ------------------------------------------------------------------ reference formulation (differentiable)
    def forward(self, user_ids: torch.Tensor, item_ids: torch.Tensor) -> torch.Tensor:
        """matrix_factorization.py:81-106."""
        dev = self.user_embeddings.weight.device
        user_ids, item_ids = user_ids.to(dev), item_ids.to(dev)
        dot = (self.user_embeddings(user_ids) * self.item_embeddings(item_ids)).sum(dim=1)
        return dot + self.user_bias(user_ids).squeeze() + self.item_bias(item_ids).squeeze() + self.global_bias

    def training_step(self, batch: Dict[str, Any], batch_idx: int) -> torch.Tensor:
        """matrix_factorization.py:133-156."""
        predictions = self(batch["user_ids"], batch["item_ids"])
        loss = nn.functional.binary_cross_entropy_with_logits(predictions, batch["labels"].float().to(predictions.device))
        self.log("train_loss", loss, prog_bar=True)
        return loss

    def configure_optimizers(self):
        """matrix_factorization.py:195-215."""
        if self.sparse:
            return torch.optim.SparseAdam(
                [{"params": self.user_embeddings.parameters()}, {"params": self.item_embeddings.parameters()},
                 {"params": self.user_bias.parameters()}, {"params": self.item_bias.parameters()},
                 {"params": [self.global_bias]}], lr=self.learning_rate)
        return torch.optim.Adam(self.parameters(), lr=self.learning_rate, weight_decay=self.weight_decay)

    # ------------------------------------------------------------------ kernels
    def _extended(self):
        """[U, d'] / [I, d'] fp32 tables with the bias chunk, rebuilt when a parameter's version changes."""
        params = [self.user_embeddings.weight, self.item_embeddings.weight, self.item_bias.weight]
        if not params[0].is_cuda:
            raise RuntimeError("MatrixFactorization parameters must live on a CUDA device; there is no CPU path")
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != self._aug_key:
            d = self.embedding_dim
            dp = 128 if d <= 64 else (256 if d <= 192 else d + 1)      # beyond 192: exact kernel only
            base = 64 if d <= 64 else (192 if d <= 192 else d)
            dev = params[0].device
            with torch.no_grad():
                ue = torch.zeros(self.num_users, dp, dtype=torch.float32, device=dev)
                ie = torch.zeros(self.num_items, dp, dtype=torch.float32, device=dev)
                ue[:, :d] = self.user_embeddings.weight
                ie[:, :d] = self.item_embeddings.weight
                # u' = [u, beta], x' = [x, b / beta]: beta a power of two (so both factors and their product are
                # exact), chosen on the device to balance ||u'|| ||x'||: beta^2 ~ ||u|| max|b| / ||x||
                b = self.item_bias.weight[:, 0].float()
                un = ue[:, :d].norm(dim=1).mean().clamp_min(1e-30)
                xn = ie[:, :d].norm(dim=1).mean().clamp_min(1e-30)
                bm = b.abs().max()
                beta = torch.where(bm > 0, torch.exp2(torch.round(0.5 * torch.log2((un * bm / xn).clamp_min(1e-30)))),
                                   torch.ones_like(bm)).clamp(2.0 ** -60, 2.0 ** 60)
                ue[:, base] = beta
                ie[:, base] = b / beta
            self._aug, self._aug_key, self._scorer = (ue, ie), key, None
        return self._aug

    def predict_all_items(self, user_ids: torch.Tensor) -> torch.Tensor:
        """matrix_factorization.py:108-131: [batch, num_items] fp32 scores (u.x + b_i by the SGEMM kernel on the
        extended tables, then + b_u + b_g)."""
        ue, ie = self._extended()
        with torch.no_grad():
            uids = user_ids.to(ue.device)
            scores = engine.score_all_items(ue, ie, uids)
            return scores + self.user_bias.weight[uids] + self.global_bias

    def recommend(self, user_ids: torch.Tensor, filter_items: Optional[Dict[int, set]] = None,
                  k: Optional[int] = None) -> torch.Tensor:
        """matrix_factorization.py:217-245: top-k item indices [batch, k] int64, (score desc, item id asc)."""
        self.eval()
        k = self.top_k if k is None else int(k)
        if k > self.num_items or k <= 0:
            raise RuntimeError("selected index k out of range")
        with torch.no_grad():
            ue, ie = self._extended()
            uids = user_ids.to(ue.device).view(-1)
            from .scorer import FusedScorer
            if uids.numel() > SMALL_BATCH and FusedScorer.supports(ue.size(1), k, self.num_items):
                if self._scorer is None:
                    self._scorer = FusedScorer(ue, ie)
                return self._scorer.topk(uids, k, filter_items)[0]
            if k > engine.EXACT_K_MAX:
                scores = self.predict_all_items(uids).double()
                if filter_items is not None:
                    for i, uid in enumerate(uids.tolist()):
                        if uid in filter_items:
                            scores[i, list(filter_items[uid])] = float("-inf")
                return torch.sort(scores, dim=1, descending=True, stable=True).indices[:, :k].contiguous()
            excl = engine.exclusion_csr(uids, filter_items, ue.device)
            return engine.topk_exact(ue, ie, uids, k, excl)[0]

    def recommend_all(self, k: Optional[int] = None) -> torch.Tensor:
        """Top-k for every user through the fused kernel: [num_users, k]."""
        return self.recommend(torch.arange(self.num_users), k=k)

    # ------------------------------------------------------------------ Lightning-facing hooks
    def validation_step(self, batch: Dict[str, Any], batch_idx: int):
        self.metrics.update(self.recommend(batch["user_ids"]).cpu(), batch["ground_truth"])

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
