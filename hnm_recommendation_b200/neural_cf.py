"""NeuralCF with the reference's model API, scored by the B200 kernels.

Drop-in for ``src.models.NeuralCF`` (src/models/neural_cf.py:9-326) on the inference
path: same constructor, sub-module names / ``state_dict`` keys, ``forward(user_ids,
item_ids)`` -> logits, ``predict_all_items``, ``recommend``.  Scoring always uses
eval-mode semantics (Dropout = identity); the reference applies dropout when the
caller forgot ``model.eval()``, which no serving caller wants.

Training (neural_cf.py:210-233, 274-298): in train mode, when autograd is recording and a
parameter requires grad, ``forward`` takes the reference's own formulation over the same
parameters (embedding lookups + ``mlp_layers`` + ``prediction_layer``, dropout live
in train mode), so ``training_step`` / ``configure_optimizers`` / ``trainer.fit`` work
on the mirror; the fused kernels serve ``no_grad`` / eval scoring.  The layer-1 tables
are rebuilt whenever a parameter's version changes; a write through ``.data`` does not
bump it -- call ``invalidate()`` after one (``load_state_dict`` and ``.to()`` do).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, engine
from ._lib import call, ptr, stream
from .base import ModelBase
from .metrics import RecommendationMetrics


class NeuralCF(ModelBase):
    def __init__(
        self,
        num_users: int,
        num_items: int,
        mf_dim: int = 64,
        mlp_dims: List[int] = [128, 64, 32],
        dropout: float = 0.1,
        learning_rate: float = 0.001,
        weight_decay: float = 0.0001,
        top_k: int = 12,
        use_pretrain: bool = False,
    ):
        super().__init__()
        self.save_hyperparameters()
        self.num_users = num_users
        self.num_items = num_items
        self.mf_dim = mf_dim
        self.mlp_dims = mlp_dims
        self.dropout = dropout
        self.learning_rate = learning_rate
        self.weight_decay = weight_decay
        self.top_k = top_k

        self.gmf_user_embedding = nn.Embedding(num_users, mf_dim)             # neural_cf.py:56-57
        self.gmf_item_embedding = nn.Embedding(num_items, mf_dim)
        self.mlp_user_embedding = nn.Embedding(num_users, mlp_dims[0] // 2)   # :60-61
        self.mlp_item_embedding = nn.Embedding(num_items, mlp_dims[0] // 2)
        self.mlp_layers = self._build_mlp(mlp_dims, dropout)                  # :64
        self.prediction_layer = nn.Linear(mf_dim + mlp_dims[-1], 1)           # :67
        self._init_weights()
        self.metrics = RecommendationMetrics(top_k=top_k)
        self._tables_key = None
        self._tables = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def invalidate(self) -> None:
        """Drop the cached layer-1 tables (needed after writing parameters through ``.data``)."""
        self._tables_key = None
        self._tables = None

    def _apply(self, fn, *args, **kwargs):
        self.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def _build_mlp(self, dims: List[int], dropout: float) -> nn.Sequential:
        layers = []                                                           # :85-90
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            layers.append(nn.ReLU())
            layers.append(nn.Dropout(dropout))
        return nn.Sequential(*layers)

    def _init_weights(self) -> None:
        nn.init.normal_(self.gmf_user_embedding.weight, std=0.01)             # :95-96
        nn.init.normal_(self.gmf_item_embedding.weight, std=0.01)
        nn.init.xavier_uniform_(self.mlp_user_embedding.weight)               # :99-100
        nn.init.xavier_uniform_(self.mlp_item_embedding.weight)
        for layer in self.mlp_layers:                                         # :103-106
            if isinstance(layer, nn.Linear):
                nn.init.xavier_uniform_(layer.weight)
                nn.init.zeros_(layer.bias)
        nn.init.xavier_uniform_(self.prediction_layer.weight)                 # :109-110
        nn.init.zeros_(self.prediction_layer.bias)

    # --------------------------------------------------------------- tables
    def _linears(self) -> List[nn.Linear]:
        return [l for l in self.mlp_layers if isinstance(l, nn.Linear)]

    def _prepared(self):
        """P/Q layer-1 tables + packed MLP tail, rebuilt only when a parameter changed."""
        _lib.require_device()
        params = list(self.parameters())
        if not params[0].is_cuda:
            raise RuntimeError("NeuralCF parameters must live on a CUDA device; there is no CPU path")
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key == self._tables_key:
            return self._tables
        lin = self._linears()
        dev = params[0].device
        h = self.mlp_dims[0] // 2
        with torch.no_grad(), torch.cuda.device(dev):
            gu = self.gmf_user_embedding.weight.detach().float().contiguous()
            gi = self.gmf_item_embedding.weight.detach().float().contiguous()
            wp = self.prediction_layer.weight.detach().float().contiguous().view(-1)
            bp = float(self.prediction_layer.bias.detach().float().item())
            if lin:
                w1 = lin[0].weight.detach().float().contiguous()
                b1 = lin[0].bias.detach().float().contiguous()
                h1 = w1.size(0)
                pu = torch.empty(self.num_users, h1, dtype=torch.float32, device=dev)
                qi = torch.empty(self.num_items, h1, dtype=torch.float32, device=dev)
                mu = self.mlp_user_embedding.weight.detach().float().contiguous()
                mi = self.mlp_item_embedding.weight.detach().float().contiguous()
                call("hnm_ncf_precompute", ptr(mu), self.num_users, h, ptr(w1), h1, w1.size(1), 0, None, ptr(pu),
                     stream())
                call("hnm_ncf_precompute", ptr(mi), self.num_items, h, ptr(w1), h1, w1.size(1), h, ptr(b1), ptr(qi),
                     stream())
                widths = [l.weight.size(0) for l in lin]
                tail_parts = []
                for l in lin[1:]:
                    tail_parts += [l.weight.detach().float().reshape(-1), l.bias.detach().float().reshape(-1)]
                tail = torch.cat(tail_parts).contiguous() if tail_parts else None
            else:
                # mlp_dims with a single entry: the "MLP output" is the raw concatenation (no Linear at all)
                raise RuntimeError("NeuralCF needs at least one MLP Linear layer (len(mlp_dims) >= 2)")
        widths_c = (C.c_int32 * len(widths))(*widths)
        self._tables = dict(gu=gu, gi=gi, pu=pu, qi=qi, tail=tail, widths=widths_c, n_layers=len(widths), wp=wp, bp=bp)
        self._tables_key = key
        return self._tables

    # -------------------------------------------------------------- scoring
    def _forward_autograd(self, user_ids: torch.Tensor, item_ids: torch.Tensor) -> torch.Tensor:
        """neural_cf.py:125-139 as written (differentiable; dropout live in train mode)."""
        dev = self.gmf_user_embedding.weight.device
        user_ids, item_ids = user_ids.to(dev), item_ids.to(dev)
        gmf_output = self.gmf_user_embedding(user_ids) * self.gmf_item_embedding(item_ids)
        mlp_input = torch.cat([self.mlp_user_embedding(user_ids), self.mlp_item_embedding(item_ids)], dim=1)
        mlp_output = self.mlp_layers(mlp_input)
        return self.prediction_layer(torch.cat([gmf_output, mlp_output], dim=1)).squeeze()

    def forward(self, user_ids: torch.Tensor, item_ids: torch.Tensor) -> torch.Tensor:
        """neural_cf.py:112-141: logits [batch] (0-dim for a single pair, as ``.squeeze()`` yields)."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_autograd(user_ids, item_ids)
        t = self._prepared()
        dev = t["gu"].device
        u = engine._norm_ids(user_ids, self.num_users, dev)
        i = engine._norm_ids(item_ids, self.num_items, dev)
        if u.numel() != i.numel():
            raise RuntimeError("user_ids and item_ids must have the same length")
        out = torch.empty(u.numel(), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call("hnm_ncf_score_pairs", ptr(t["gu"]), ptr(t["gi"]), ptr(t["pu"]), ptr(t["qi"]), ptr(t["tail"]),
                 t["widths"], t["n_layers"], ptr(t["wp"]), t["bp"], ptr(u), ptr(i), u.numel(), self.mf_dim,
                 ptr(out), stream())
        return out.squeeze()

    def score_candidates(self, user_ids: Optional[torch.Tensor], cand_items: Optional[torch.Tensor],
                         cand_per_user: Optional[int] = None) -> torch.Tensor:
        """Logits [rows, C] for C candidate items per listed user (BASELINE.json configs[3]).
        ``cand_items`` int32 [rows, C]; None means items 0..C-1."""
        t = self._prepared()
        dev = t["gu"].device
        u = engine._norm_ids(user_ids, self.num_users, dev)
        rows = u.numel() if u is not None else (cand_items.size(0) if cand_items is not None else self.num_users)
        if cand_items is not None:
            cand_items = cand_items.to(device=dev, dtype=torch.int32).contiguous()
            if cand_items.numel():
                lo, hi = int(cand_items.min()), int(cand_items.max())
                if lo < 0 or hi >= self.num_items:
                    raise IndexError("candidate item index out of range")
            c = cand_items.size(1)
        else:
            c = int(cand_per_user if cand_per_user is not None else self.num_items)
            if c > self.num_items:
                raise IndexError("candidate item index out of range")
        out = torch.empty(rows, c, dtype=torch.float32, device=dev)
        if rows == 0 or c == 0:
            return out
        with torch.cuda.device(dev):
            call("hnm_ncf_score_candidates", ptr(t["gu"]), ptr(t["gi"]), ptr(t["pu"]), ptr(t["qi"]), ptr(t["tail"]),
                 t["widths"], t["n_layers"], ptr(t["wp"]), t["bp"], ptr(u), rows, ptr(cand_items), c, self.mf_dim,
                 ptr(out), stream())
        return out

    def predict_all_items(self, user_ids: torch.Tensor) -> torch.Tensor:
        """neural_cf.py:143-208: [batch, num_items] logits (one pass instead of 1000-item chunks)."""
        return self.score_candidates(user_ids, None, self.num_items)

    def recommend(self, user_ids: torch.Tensor, filter_items: Optional[Dict[int, set]] = None,
                  k: Optional[int] = None) -> torch.Tensor:
        """neural_cf.py:300-326, ties by item id ascending."""
        self.eval()
        k = self.top_k if k is None else int(k)
        if k > self.num_items or k <= 0:
            raise RuntimeError("selected index k out of range")
        with torch.no_grad():
            scores = self.predict_all_items(user_ids)                     # [B, I] fp32 logits (the kernels above)
            dev = scores.device
            uids = user_ids.to(dev).view(-1)
            if k > engine.EXACT_K_MAX:
                if filter_items is not None:
                    for i, user_id in enumerate(uids.tolist()):
                        if user_id in filter_items:
                            scores[i, list(filter_items[user_id])] = float("-inf")
                return torch.sort(scores, dim=1, descending=True, stable=True).indices[:, :k].contiguous()
            # filter + top-k in one select kernel (neural_cf.py:316-325), ties by item id ascending
            excl = engine.exclusion_csr(uids, filter_items, dev)
            ids = torch.empty(scores.size(0), k, dtype=torch.int64, device=dev)
            with torch.cuda.device(dev):
                call("hnm_topk_dense", ptr(scores), scores.size(0), scores.size(1), ptr(excl[0]), ptr(excl[1]), k,
                     ptr(ids), None, stream())
            return ids

    def training_step(self, batch: Dict[str, Any], batch_idx: int) -> torch.Tensor:
        """neural_cf.py:210-233: binary cross entropy with logits on (user, item, label) triples."""
        predictions = self(batch["user_ids"], batch["item_ids"])
        labels = batch["labels"].float().to(predictions.device)
        loss = nn.functional.binary_cross_entropy_with_logits(predictions, labels)
        self.log("train_loss", loss, prog_bar=True)
        return loss

    def configure_optimizers(self):
        """neural_cf.py:274-298 (Adam with weight decay, ReduceLROnPlateau on val_map_at_k)."""
        optimizer = torch.optim.Adam(self.parameters(), lr=self.learning_rate, weight_decay=self.weight_decay)
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="max", factor=0.5, patience=5)
        return {"optimizer": optimizer,
                "lr_scheduler": {"scheduler": scheduler, "monitor": "val_map_at_k", "frequency": 1}}

    def validation_step(self, batch: Dict[str, Any], batch_idx: int):
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
