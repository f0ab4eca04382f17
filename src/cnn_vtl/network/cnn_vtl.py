"""`from src.cnn_vtl.network.cnn_vtl import CnnVtl` (reference create_distance_matrix.py:8, basic_example.py:3)."""
from deeploopcloser_b200.cnn_vtl import CnnVtl  # noqa: F401
