"""`from src.sdav.input.InputGenerator import get_generator` (reference SDAV.py:9)."""
from deeploopcloser_b200.input_parser import get_generator  # noqa: F401
