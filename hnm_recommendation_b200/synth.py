"""Synthetic H&M-shaped interaction data (the reference's ``src/data`` is absent).

Shape facts come from the reference's EDA (CLAUDE.md:12-14, scripts/deep_data_analysis.py:430-431):
~1.37 M customers, ~105 k articles, ~31 M transactions, power-law item popularity.
Generated with numpy's PCG64 so the same seed gives the same graph on every host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

HM_USERS, HM_ITEMS, HM_EDGES = 1_371_980, 105_542, 31_788_324     # Kaggle H&M counts
CONFIG1 = (10_000, 5_000, 200_000)                                  # BASELINE.json configs[0]


@dataclass
class Interactions:
    num_users: int
    num_items: int
    users: np.ndarray     # int64 [E]
    items: np.ndarray     # int64 [E]   0-based item index (not node id)

    def edge_index(self) -> torch.Tensor:
        """[2, 2E] int64 in the layout of tests/test_models.py:178-185: both directions, items offset by U."""
        u = torch.from_numpy(self.users)
        i = torch.from_numpy(self.items) + self.num_users
        return torch.stack([torch.cat([u, i]), torch.cat([i, u])])


def interactions(num_users: int, num_items: int, num_edges: int, seed: int = 42,
                 uniform: bool = False, pop_offset: float = 100.0, pop_exponent: float = 1.0) -> Interactions:
    """Users: every user buys at least once, the rest follows a log-normal activity (mean E/U).
    Items: popularity proportional to (rank + pop_offset)^-pop_exponent over a random permutation,
    which at the H&M shape makes the best seller ~45 k purchases (the real one has ~50 k).
    Repeat purchases (duplicate edges) are kept, as the reference keeps them.
    """
    if num_edges < num_users:
        raise ValueError("need at least one interaction per user")
    rng = np.random.Generator(np.random.PCG64(seed))
    if uniform:
        extra = rng.multinomial(num_edges - num_users, np.full(num_users, 1.0 / num_users))
    else:
        act = rng.lognormal(mean=0.0, sigma=1.1, size=num_users)
        extra = rng.multinomial(num_edges - num_users, act / act.sum())
    deg = extra.astype(np.int64) + 1
    users = np.repeat(np.arange(num_users, dtype=np.int64), deg)
    if uniform:
        items = rng.integers(0, num_items, size=num_edges, dtype=np.int64)
    else:
        pop = (np.arange(num_items, dtype=np.float64) + pop_offset) ** (-pop_exponent)
        cdf = np.cumsum(pop / pop.sum())
        cdf[-1] = 1.0
        rank = np.searchsorted(cdf, rng.random(num_edges), side="right").astype(np.int64)
        perm = rng.permutation(num_items).astype(np.int64)
        items = perm[np.minimum(rank, num_items - 1)]
    return Interactions(num_users, num_items, users, items)


def xavier_embeddings(num_nodes: int, dim: int, seed: int = 42) -> torch.Tensor:
    """nn.init.xavier_uniform_ on [N, d] (lightgcn.py:70-71) with RANDOM_SEED (src/utils/constants.py:39)."""
    g = torch.Generator().manual_seed(seed)
    bound = (6.0 / (num_nodes + dim)) ** 0.5
    return (torch.rand(num_nodes, dim, generator=g) * 2 - 1) * bound


def trained_like_embeddings(num_nodes: int, dim: int, seed: int = 42, scale: float = 0.1) -> torch.Tensor:
    """N(0, scale^2): closer in magnitude to trained embeddings than the Xavier init of a 1.5 M-row table."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(num_nodes, dim, generator=g) * scale
