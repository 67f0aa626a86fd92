"""CPU restatement of MatrixFactorization's scoring path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows src/models/matrix_factorization.py of hyunlord/hnm_recommendation.  ``p`` uses the reference's
state_dict keys (user_embeddings.weight, item_embeddings.weight, user_bias.weight, item_bias.weight,
global_bias).  Pinned by tests/golden/mf_*.npz, produced by the reference's own file (make_golden_mf.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .lightgcn_oracle import apply_filter, topk_canonical


def mf_forward(p: Dict[str, torch.Tensor], user_ids: torch.Tensor, item_ids: torch.Tensor) -> torch.Tensor:
    """prediction = u.x + b_u + b_i + b_g            matrix_factorization.py:92-106"""
    dot = (p["user_embeddings.weight"][user_ids] * p["item_embeddings.weight"][item_ids]).sum(dim=1)
    return dot + p["user_bias.weight"][user_ids].squeeze(-1) + p["item_bias.weight"][item_ids].squeeze(-1) + \
        p["global_bias"]


def mf_predict_all_items(p: Dict[str, torch.Tensor], user_ids: torch.Tensor) -> torch.Tensor:
    """scores = U_b V^T + b_u + b_i^T + b_g           matrix_factorization.py:108-131"""
    scores = torch.matmul(p["user_embeddings.weight"][user_ids], p["item_embeddings.weight"].t())
    return scores + p["user_bias.weight"][user_ids] + p["item_bias.weight"].t() + p["global_bias"]


def mf_rank_scores_fp64(p: Dict[str, torch.Tensor], user_ids: torch.Tensor) -> torch.Tensor:
    """What decides the ranking of a user's items, exactly: fp64 chain of u_k x_k (k ascending) then + b_i.
    b_u and b_g are the same for every item of a user, so they cannot change the order (:217-232)."""
    u = p["user_embeddings.weight"][user_ids].to(torch.float64)
    v = p["item_embeddings.weight"].to(torch.float64)
    s = torch.zeros(u.size(0), v.size(0), dtype=torch.float64)
    for k in range(u.size(1)):
        s += u[:, k:k + 1] * v[:, k].unsqueeze(0)
    return s + p["item_bias.weight"].to(torch.float64).t()


def mf_recommend_exact(p: Dict[str, torch.Tensor], user_ids: torch.Tensor, k: int,
                       filter_items: Optional[Dict[int, set]] = None) -> torch.Tensor:
    """Canonical recommend(): (exact score desc, item id asc)   matrix_factorization.py:217-245."""
    s = apply_filter(mf_rank_scores_fp64(p, user_ids), user_ids, filter_items)
    return topk_canonical(s, k)
