"""`from src.sdav.network.SDAV import SDAV` as in the reference's train.py:3 / create_similarity_matrix.py:11."""
from deeploopcloser_b200.sdav import SDAV  # noqa: F401
