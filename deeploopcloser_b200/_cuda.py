"""Device-memory plumbing on top of PyTorch (allocation, streams, host<->device copies). No compute happens here."""
import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("deeploopcloser_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    _lib.call("dlc_device_check")


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def empty(shape, dtype, device=None):
    return torch.empty(shape, dtype=dtype, device=device or "cuda")


def to_device(a, dtype=None):
    """numpy / torch -> contiguous CUDA tensor (pinned staging for large host arrays)."""
    if isinstance(a, torch.Tensor):
        t = a
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous().cuda()
    arr = np.ascontiguousarray(a)
    t = torch.from_numpy(arr)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.cuda()


class Workspace:
    """Grow-only scratch buffer handed to the C ABI (which never allocates on the hot path)."""

    def __init__(self):
        self._buf = None

    def get(self, nbytes):
        nbytes = max(int(nbytes), 256)
        if self._buf is None or self._buf.numel() < nbytes:
            self._buf = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
        base = self._buf.data_ptr()
        off = (-base) % 256
        return base + off, self._buf.numel() - off
