"""`import src.utils.TensorflowWrapper as tw` (reference SDAV.py:8, test/TensorflowWrapperTest.py:6)."""
from deeploopcloser_b200.tensorwrapper import *  # noqa: F401,F403
from deeploopcloser_b200.tensorwrapper import (Session, TensorWrapper, constant, float32, float64, int32,  # noqa: F401
                                               int64, ones, parameter_guard, placeholder, random_mask, zeros)
