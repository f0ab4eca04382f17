"""`from src.utils.MathUtils import MathUtils` (reference cnn_vtl.py:8)."""
from deeploopcloser_b200.cnn_vtl import compressed_size as _compressed_size


class MathUtils:
    @staticmethod
    def compressed_size(value: int, compression: float):
        return _compressed_size(value, compression)
