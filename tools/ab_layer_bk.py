"""A/B of the K block (32 / 64, dlc_debug_set key 0) of the encoder layers: layer-0 shape on exact pixel planes
(two products) and an inner layer (three products), M = 31890, plane outputs like the real layers."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200 import _lib, ops  # noqa: E402

M, N, NPAD = 31890, 2500, 2560
g = torch.Generator(device="cuda")
g.manual_seed(0)


def planes(rows, ld, k, scale):
    x = torch.zeros((rows, ld), dtype=torch.float16, device="cuda")
    x[:, :k] = (torch.rand((rows, k), device="cuda", generator=g) * scale).half()
    return x


bias = torch.zeros(NPAD, dtype=torch.float32, device="cuda")
cases = {"layer0 (K 1681, A exact: 2 products)": (planes(M, 1728, 1681, 255.0).round(), None, planes(NPAD, 1728, 1681, 0.1), planes(NPAD, 1728, 1681, 1e-4)),
         "inner (K 2500, 3 products)": (planes(M, 2560, 2500, 1.0), planes(M, 2560, 2500, 1e-3), planes(NPAD, 2560, 2500, 0.1), planes(NPAD, 2560, 2500, 1e-4))}
for name, (ah, al, bh, bl) in cases.items():
    ref = None
    for bk in (32, 64, 32, 64):
        _lib.call("dlc_debug_set", 0, bk)
        out = ops.gemm_planes(ah, al, bh, bl, M, N, bias, "sigmoid", "fp16x2", want_f32=False, want_planes=True)[1][0]
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        same = bool(torch.equal(out, ref))
        o_hi = torch.zeros((M, NPAD), dtype=torch.float16, device="cuda")
        o_lo = torch.zeros_like(o_hi)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _lib.call("dlc_gemm_planes", ops.ptr(ah), ops.ptr(al), ops.ptr(bh), ops.ptr(bl), M, N, NPAD, ah.shape[1],
                      ops.ptr(bias), _lib.ACT_SIGMOID, _lib.PREC_FP16X2, None, N, ops.ptr(o_hi), ops.ptr(o_lo), NPAD,
                      ops.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"case": name, "bk": bk, "ms": round(e0.elapsed_time(e1) / 10, 4), "same_bits_as_bk32": same}))
_lib.call("dlc_debug_set", 0, 32)
