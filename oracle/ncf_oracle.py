"""CPU restatement of NeuralCF's inference path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows src/models/neural_cf.py of hyunlord/hnm_recommendation (eval mode:
Dropout is the identity).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


def ncf_forward(p: Dict[str, torch.Tensor], user_ids: torch.Tensor,
                item_ids: torch.Tensor) -> torch.Tensor:
    """Logit per (user, item) pair.  src/models/neural_cf.py:125-139.

    ``p`` uses the reference's state_dict keys.  ``mlp_layers`` is
    Linear, ReLU, Dropout repeated len(mlp_dims)-1 times (:85-90), so the
    Linear modules sit at indices 0, 3, 6, ...
    """
    g = p["gmf_user_embedding.weight"][user_ids] * p["gmf_item_embedding.weight"][item_ids]  # :125-127
    x = torch.cat([p["mlp_user_embedding.weight"][user_ids],
                   p["mlp_item_embedding.weight"][item_ids]], dim=1)                        # :130-132
    i = 0
    while f"mlp_layers.{i}.weight" in p:                                                    # :133
        x = torch.relu(torch.nn.functional.linear(x, p[f"mlp_layers.{i}.weight"],
                                                  p[f"mlp_layers.{i}.bias"]))
        i += 3
    z = torch.cat([g, x], dim=1)                                                             # :136
    y = torch.nn.functional.linear(z, p["prediction_layer.weight"], p["prediction_layer.bias"])
    return y.squeeze()                                                                       # :139


def ncf_predict_all_items(p: Dict[str, torch.Tensor], user_ids: torch.Tensor,
                          item_batch_size: int = 1000) -> torch.Tensor:
    """[B, I] logits, item-chunked like src/models/neural_cf.py:167-206."""
    num_items = p["gmf_item_embedding.weight"].size(0)
    b = user_ids.numel()
    out = []
    for i in range(0, num_items, item_batch_size):
        items = torch.arange(i, min(i + item_batch_size, num_items))
        uu = user_ids.repeat_interleave(items.numel())
        ii = items.repeat(b)
        out.append(ncf_forward(p, uu, ii).view(b, -1))
    return torch.cat(out, dim=1)


class NeuralCFOracle:
    """Holds a reference-keyed state_dict; builds one with the reference's init when none is given."""

    def __init__(self, num_users, num_items, mf_dim=64, mlp_dims: List[int] = (128, 64, 32),
                 top_k=12, state: Optional[Dict[str, torch.Tensor]] = None, dtype=torch.float32):
        self.num_users, self.num_items, self.top_k = num_users, num_items, top_k
        if state is None:
            state = {}
            h = mlp_dims[0] // 2
            state["gmf_user_embedding.weight"] = torch.empty(num_users, mf_dim).normal_(std=0.01)   # :95
            state["gmf_item_embedding.weight"] = torch.empty(num_items, mf_dim).normal_(std=0.01)   # :96
            state["mlp_user_embedding.weight"] = torch.nn.init.xavier_uniform_(torch.empty(num_users, h))  # :99
            state["mlp_item_embedding.weight"] = torch.nn.init.xavier_uniform_(torch.empty(num_items, h))  # :100
            for j in range(len(mlp_dims) - 1):                                                     # :103-106
                state[f"mlp_layers.{3 * j}.weight"] = torch.nn.init.xavier_uniform_(
                    torch.empty(mlp_dims[j + 1], mlp_dims[j]))
                state[f"mlp_layers.{3 * j}.bias"] = torch.zeros(mlp_dims[j + 1])
            state["prediction_layer.weight"] = torch.nn.init.xavier_uniform_(
                torch.empty(1, mf_dim + mlp_dims[-1]))                                             # :109
            state["prediction_layer.bias"] = torch.zeros(1)                                        # :110
        self.state = {k: v.detach().cpu().to(dtype) for k, v in state.items()}

    def forward(self, user_ids, item_ids):
        return ncf_forward(self.state, user_ids, item_ids)

    def predict_all_items(self, user_ids):
        return ncf_predict_all_items(self.state, user_ids)
