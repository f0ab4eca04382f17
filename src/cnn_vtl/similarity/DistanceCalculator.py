"""`from src.cnn_vtl.similarity.DistanceCalculator import DistanceCalculator` (reference create_distance_matrix.py:9)."""
from deeploopcloser_b200.distance import DistanceCalculator  # noqa: F401
