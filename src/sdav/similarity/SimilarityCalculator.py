"""`from src.sdav.similarity.SimilarityCalculator import SimilarityCalculator` (reference create_similarity_matrix.py:12)."""
from deeploopcloser_b200.similarity import SimilarityCalculator  # noqa: F401
