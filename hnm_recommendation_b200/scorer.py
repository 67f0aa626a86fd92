"""Tensor-core full-catalog scorer (placeholder until csrc/score_fused.cu lands)."""


class FusedScorer:
    @staticmethod
    def supports(dim: int, k: int, num_items: int) -> bool:
        return False
