"""CPU oracle for the LightGCN / NeuralCF scoring hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, in plain CPU PyTorch, the
arithmetic of hyunlord/hnm_recommendation's scoring path so that the CUDA
path can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it;
nothing under ``hnm_recommendation_b200/`` does (the product path fails loudly
when the CUDA library is missing -- there is no CPU fallback).

Pinning status: the reference ships no golden vectors and cannot be imported
as shipped.  The oracle is pinned instead against
outputs of the reference's *own* ``src/models/lightgcn.py`` and
``src/models/neural_cf.py`` executed in the authoring container with minimal
stand-ins for the three absent third-party packages
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).  See DESIGN.md
section "Oracle" for exactly what those stand-ins supply.

``oracle/c/hnm_oracle.c`` (wrapper: ``oracle.c_oracle``) is a second, independent restatement of the same path
in plain C -- scalar loops, no tensor library -- pinned against the same golden vectors and required to agree
with this package bit for bit on index work and exact scores (tests/test_oracle_c.py).
"""
from .lightgcn_oracle import (  # noqa: F401
    layer_weights, add_self_loops, build_norm_adj, propagate, forward,
    predict, predict_all_items, apply_filter, topk_canonical, exact_scores_fp64,
    recommend, recommend_exact, LightGCNOracle, bpr_loss,
)
from .ncf_oracle import ncf_forward, ncf_predict_all_items, NeuralCFOracle  # noqa: F401
from .mf_oracle import mf_forward, mf_predict_all_items, mf_rank_scores_fp64, mf_recommend_exact  # noqa: F401
