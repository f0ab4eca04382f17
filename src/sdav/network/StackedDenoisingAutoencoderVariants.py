"""`from src.sdav.network.StackedDenoisingAutoencoderVariants import SDA` (reference train-sdav.py:3)."""
from deeploopcloser_b200.da import SDA  # noqa: F401
