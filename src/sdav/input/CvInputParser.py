"""`from src.sdav.input.CvInputParser import CvInputParser` (reference create_similarity_matrix.py:10)."""
from deeploopcloser_b200.input_parser import (CvInputParser, get_1d_boundaries, get_2d_boundaries,  # noqa: F401
                                              get_top_n_key_points, get_vectorized_patches_from_key_points)
