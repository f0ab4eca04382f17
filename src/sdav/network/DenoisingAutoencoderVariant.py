"""`from src.sdav.network.DenoisingAutoencoderVariant import DA` (reference StackedDenoisingAutoencoderVariants.py:7)."""
from deeploopcloser_b200.da import DA  # noqa: F401
