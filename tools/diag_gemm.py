"""GPU diagnostic for the tcgen05 GEMM core: structured inputs that expose layout / descriptor mistakes.
Writes a report to gpurun_out/diag_gemm.txt. Not part of the product or the tests."""
import os
import sys
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200 import _lib, ops  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/diag_gemm.txt", "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    out.write(s + "\n")
    out.flush()


def run(m, k, n, precision, a=None, b=None, tag=""):
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, (m, k)) if a is None else a
    b = rng.uniform(-1, 1, (k, n)) if b is None else b
    ref = a @ b
    try:
        got = ops.matmul(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), precision=precision)
        torch.cuda.synchronize()
        got = got.double().cpu().numpy()
    except Exception as e:  # noqa: BLE001
        log("FAIL", tag, (m, k, n), precision, repr(e))
        traceback.print_exc()
        return None
    err = np.abs(got - ref)
    log("%-10s m=%d k=%d n=%d %s: max_abs_err=%.3e normwise=%.3e" % (
        tag, m, k, n, precision, err.max(), np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)))
    if err.max() > 1e-2:
        bad = np.argwhere(err > 1e-2)
        log("   bad elements:", len(bad), "of", err.size, "rows", np.unique(bad[:, 0])[:16], "cols",
            np.unique(bad[:, 1])[:16])
        log("   got[0,:8]", got[0, :8], "\n   ref[0,:8]", ref[0, :8])
    return got


log("device", torch.cuda.get_device_name(0), "sms", _lib.call("dlc_sm_count"))
for prec in ("fp16", "fp16x2"):
    # identity-like: B = I (k = n) -> output must equal A rounded; exposes K-slice / swizzle mistakes
    for k in (64, 128, 256):
        a = np.arange(128 * k, dtype=np.float64).reshape(128, k) % 97 / 97.0
        run(128, k, k, prec, a=a, b=np.eye(k), tag="identity")
    run(128, 64, 32, prec, tag="tiny")
    run(256, 192, 256, prec, tag="2tiles")
    run(1000, 1681, 2500, prec, tag="layer0")
    run(777, 2500, 2500, prec, tag="layerN")
for bk in (64, 32):
    _lib.call("dlc_debug_set", 0, bk)
    run(300, 777, 520, "fp16x2", tag="bk%d" % bk)
_lib.call("dlc_debug_set", 0, 32)

# quick timing of the encoder-shaped GEMM (not a benchmark: sanity of the pipeline)
for prec in ("fp16", "fp16x2"):
    m, k, n = 31890, 2500, 2500
    a = torch.rand((m, k), device="cuda")
    b = torch.randn((k, n), device="cuda")
    split = prec == "fp16x2"
    a_hi, a_lo = ops.split_planes(a, need_lo=split)
    b_hi, b_lo = ops.pack_weight_planes(b, n_pad=2560, need_lo=split)
    for _ in range(2):
        ops.gemm_planes(a_hi, a_lo, b_hi, b_lo, m, n, None, "sigmoid", prec, want_f32=False, want_planes=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.gemm_planes(a_hi, a_lo, b_hi, b_lo, m, n, None, "sigmoid", prec, want_f32=False, want_planes=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    log("encoder-layer GEMM %s: %.3f ms  -> %.1f algorithmic TFLOP/s" % (prec, ms, 2.0 * m * k * n / ms / 1e9))
log("done")
