"""GPU diagnostic: accuracy and speed of the 3-product kernel as a function of the promotion interval
(K elements accumulated inside the tensor core before the partial sum is added in fp32 registers)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200 import _lib, ops  # noqa: E402
from oracle import sda as o_sda  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/diag_promote.txt", "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    out.write(s + "\n")
    out.flush()


rng = np.random.default_rng(0)
a = rng.uniform(0, 1, (777, 2500))
b = rng.standard_normal((2500, 2500))
ref = a @ b
dims = [1681, 2500, 2500, 2500, 2500, 2500]
ws, bs = o_sda.make_weights(dims, seed=1, scale="normal")
x = rng.integers(0, 256, (4 * 30, 1681)).astype(np.float64) / 255.0
sda_ref = o_sda.sda_forward(x, ws, bs)
enc = ops.SdaEncoder(dims, "fp16x2")
for l, (w, bb) in enumerate(zip(ws, bs)):
    enc.set_layer(l, w, bb)

m, k, n = 31890, 2500, 2500
A = torch.rand((m, k), device="cuda")
B = torch.randn((k, n), device="cuda")
a_hi, a_lo = ops.split_planes(A)
b_hi, b_lo = ops.pack_weight_planes(B, n_pad=2560)

for pk in (64, 128, 256, 512, 1024, 1 << 20):
    _lib.call("dlc_debug_set", 2, pk)
    got = ops.matmul(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).double().cpu().numpy()
    nerr = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    o = enc.encode(torch.from_numpy(x).cuda()).cpu().numpy()
    serr = np.max(np.abs(o - sda_ref))
    for _ in range(2):
        ops.gemm_planes(a_hi, a_lo, b_hi, b_lo, m, n, None, "sigmoid", "fp16x2", want_f32=False, want_planes=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.gemm_planes(a_hi, a_lo, b_hi, b_lo, m, n, None, "sigmoid", "fp16x2", want_f32=False, want_planes=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    log("promote_k=%7d  gemm normwise err %.3e   sda(N(0,1) weights) max-abs err %.3e   layer %.3f ms (%.0f alg TFLOP/s)" % (
        pk, nerr, serr, ms, 2.0 * m * k * n / ms / 1e9))
_lib.call("dlc_debug_set", 2, 256)
log("done")
