"""Stand-in for the reference's missing ``RecommendationMetrics``.

The reference imports ``src.evaluation.RecommendationMetrics`` (src/models/lightgcn.py:10,79)
but defines it nowhere.  The contract inferred from its callers
(lightgcn.py:284-292, scripts/benchmark_models.py:64,149,167,203-206):
``update(top_k_items, ground_truth)``, ``compute() -> {map_at_k, recall_at_k,
precision_at_k, ndcg_at_k}``, ``reset()``.  Not on the hot path.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence


class RecommendationMetrics:
    def __init__(self, top_k: int = 12):
        self.top_k = top_k
        self.reset()

    def reset(self) -> None:
        self._n = 0
        self._map = self._rec = self._prec = self._ndcg = 0.0

    def update(self, top_k_items, ground_truth: Sequence[Iterable[int]]) -> None:
        rows: List[List[int]] = top_k_items.tolist() if hasattr(top_k_items, "tolist") else list(top_k_items)
        for rec, truth in zip(rows, ground_truth):
            truth = set(int(t) for t in (truth.tolist() if hasattr(truth, "tolist") else truth))
            if not truth:
                continue
            rec = rec[: self.top_k]
            hits, ap, dcg = 0, 0.0, 0.0
            for rank, item in enumerate(rec):
                if item in truth:
                    hits += 1
                    ap += hits / (rank + 1)
                    dcg += 1.0 / math.log2(rank + 2)
            idcg = sum(1.0 / math.log2(r + 2) for r in range(min(len(truth), self.top_k)))
            self._map += ap / min(len(truth), self.top_k)
            self._rec += hits / len(truth)
            self._prec += hits / self.top_k
            self._ndcg += dcg / idcg
            self._n += 1

    def compute(self) -> Dict[str, float]:
        n = max(self._n, 1)
        return {"map_at_k": self._map / n, "recall_at_k": self._rec / n,
                "precision_at_k": self._prec / n, "ndcg_at_k": self._ndcg / n}
