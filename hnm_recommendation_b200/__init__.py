"""B200-native scoring hot path of hyunlord/hnm_recommendation (LightGCN / NeuralCF)."""
from .lightgcn import LightGCN  # noqa: F401
from .neural_cf import NeuralCF  # noqa: F401
from .matrix_factorization import MatrixFactorization  # noqa: F401

__all__ = ["LightGCN", "NeuralCF", "MatrixFactorization"]
