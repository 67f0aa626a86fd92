"""Data contract of the hot path: the interactions table the reference's (absent) data module serves.

The reference's ``src/data`` is missing from the repository; what its callers need is
inferred from them:
  * ``data_module.get_graph() -> (edge_index [2, 2E] int64, edge_weight | None)``   scripts/train.py:221
    -- both directions, item node ids offset by num_users (tests/test_models.py:178-185), duplicates kept;
  * ``data/processed/train.parquet`` with integer columns ``customer_idx`` and ``article_idx``
    (scripts/serve.py:174-177), from which the server builds ``{customer_idx: set(article_idx)}`` for the
    purchased-item filter (serve.py:350-352).
``InteractionData`` is that contract on top of one parquet file (or of arrays / the synthetic generator);
nothing here touches the GPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

TRAIN_FILE = os.path.join("processed", "train.parquet")
USER_COL, ITEM_COL = "customer_idx", "article_idx"


@dataclass
class InteractionData:
    num_users: int
    num_items: int
    users: np.ndarray          # int64 [E] customer_idx
    items: np.ndarray          # int64 [E] article_idx (0-based item index, not node id)

    # ------------------------------------------------------------------ constructors
    @staticmethod
    def from_parquet(path: str, num_users: Optional[int] = None, num_items: Optional[int] = None) -> "InteractionData":
        """`path` is the parquet file or the data directory that holds processed/train.parquet."""
        import pyarrow.parquet as pq
        if os.path.isdir(path):
            path = os.path.join(path, TRAIN_FILE)
        table = pq.read_table(path, columns=[USER_COL, ITEM_COL])
        users = table.column(USER_COL).to_numpy().astype(np.int64, copy=False)
        items = table.column(ITEM_COL).to_numpy().astype(np.int64, copy=False)
        return InteractionData.from_arrays(users, items, num_users, num_items)

    @staticmethod
    def from_arrays(users, items, num_users: Optional[int] = None, num_items: Optional[int] = None) -> "InteractionData":
        users = np.ascontiguousarray(users, dtype=np.int64)
        items = np.ascontiguousarray(items, dtype=np.int64)
        if users.shape != items.shape or users.ndim != 1:
            raise ValueError("customer_idx and article_idx must be 1-d arrays of the same length")
        if users.size and (users.min() < 0 or items.min() < 0):
            raise ValueError("negative index in the interactions table")
        nu = int(users.max()) + 1 if users.size else 0
        ni = int(items.max()) + 1 if items.size else 0
        if num_users is not None and num_users < nu or num_items is not None and num_items < ni:
            raise ValueError("index beyond num_users / num_items in the interactions table")
        return InteractionData(num_users if num_users is not None else nu, num_items if num_items is not None else ni,
                               users, items)

    @staticmethod
    def synthetic(num_users: int, num_items: int, num_edges: int, seed: int = 42) -> "InteractionData":
        from . import synth
        d = synth.interactions(num_users, num_items, num_edges, seed=seed)
        return InteractionData(num_users, num_items, d.users, d.items)

    # ------------------------------------------------------------------ the contract
    def get_graph(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """scripts/train.py:221: (edge_index [2, 2E] int64, None); model.set_graph(*data.get_graph())."""
        u = torch.from_numpy(self.users)
        i = torch.from_numpy(self.items) + self.num_users
        return torch.stack([torch.cat([u, i]), torch.cat([i, u])]), None

    def user_history(self) -> Dict[int, set]:
        """scripts/serve.py:174-177: {customer_idx: set(article_idx)} -- the `filter_items` argument of recommend()."""
        order = np.argsort(self.users, kind="stable")
        u, it = self.users[order], self.items[order]
        cuts = np.flatnonzero(np.diff(u)) + 1
        return {int(g[0]): set(int(x) for x in v) for g, v in zip(np.split(u, cuts), np.split(it, cuts)) if g.size}

    def history_csr(self, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """The same history as a CSR over ALL users (int64 row pointer [U + 1], item ids sorted per user, repeat
        purchases kept once): the exclusion lists hnm_rescore_topk / hnm_topk_exact take, without a Python dict."""
        key = np.unique(self.users * np.int64(self.num_items) + self.items)
        u, it = key // self.num_items, key % self.num_items
        ptr = np.zeros(self.num_users + 1, dtype=np.int64)
        np.cumsum(np.bincount(u, minlength=self.num_users), out=ptr[1:])
        p, x = torch.from_numpy(ptr), torch.from_numpy(it.astype(np.int64))
        return (p.to(device), x.to(device)) if device is not None else (p, x)

    def to_parquet(self, path: str) -> str:
        """Write processed/train.parquet under `path` (a directory) in the layout serve.py reads."""
        import pyarrow as pa
        import pyarrow.parquet as pq
        if not path.endswith(".parquet"):
            os.makedirs(os.path.join(path, "processed"), exist_ok=True)
            path = os.path.join(path, TRAIN_FILE)
        pq.write_table(pa.table({USER_COL: self.users, ITEM_COL: self.items}), path)
        return path
