"""GPU experiment: where does the encoder-layer kernel lose tensor-pipe time? Toggles epilogue pieces."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200 import _lib, ops  # noqa: E402

m, k, n = 31890, 2500, 2500
A = torch.rand((m, k), device="cuda")
B = torch.randn((k, n), device="cuda")
a_hi, a_lo = ops.split_planes(A)
b_hi, b_lo = ops.pack_weight_planes(B, n_pad=2560)


def t(prec, planes=True, f32=False, reps=5):
    kw = dict(want_f32=f32, want_planes=planes)
    for _ in range(2):
        ops.gemm_planes(a_hi, a_lo, b_hi, b_lo, m, n, None, "sigmoid", prec, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gemm_planes(a_hi, a_lo, b_hi, b_lo, m, n, None, "sigmoid", prec, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for pk in (256, 1 << 20):
    _lib.call("dlc_debug_set", 2, pk)
    for flags, name in ((0, "full epilogue"), (1, "no stores"), (2, "no activation"), (3, "no stores, no activation")):
        _lib.call("dlc_debug_set", 3, flags)
        print("promote_k=%8d %-26s fp16x2: %.3f ms   fp16: %.3f ms" % (pk, name, t("fp16x2"), t("fp16")), flush=True)
_lib.call("dlc_debug_set", 3, 0)
_lib.call("dlc_debug_set", 2, 256)
print("f32 output only (last layer):", t("fp16x2", planes=False, f32=True))
