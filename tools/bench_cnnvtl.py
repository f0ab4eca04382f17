"""cnn_vtl timing (BASELINE.json config 3 shape: 1063 frames 192x240x3, fused conv head + Hamming matrix).

    python tools/bench_cnnvtl.py [chunk] [precision]

Prints one JSON line: conv-head time per 1063 frames (frames resident in HBM), algorithmic TFLOP/s (1.748 GFLOP per
frame) against the measured sustained tensor peak, and the exact Hamming matrix time."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deeploopcloser_b200 import ops  # noqa: E402
from deeploopcloser_b200.cnn_vtl import CnnVtl  # noqa: E402

N, H, W = 1063, 192, 240
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 1063
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16x2"
if os.environ.get("DLC_CTA_PAIR"):            # A/B: 0 = single-CTA conv kernels, 1 = CTA pairs (default)
    from deeploopcloser_b200 import _lib
    _lib.call("dlc_debug_set", 6, int(os.environ["DLC_CTA_PAIR"]))
net = CnnVtl(input_shape=[N, H, W, 3], weights="synthetic", seed=4, precision=precision)
x = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device="cuda")


def forward():
    outs = []
    for s in range(0, N, chunk):
        outs.append(net._forward_chunk(x[s:s + chunk]))
    return torch.cat(outs)


for _ in range(3):
    d = forward()
torch.cuda.synchronize()
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    d = forward()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flop = 1.748e9 * N
for _ in range(2):
    D = ops.hamming_matrix(d)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    D = ops.hamming_matrix(d)
e1.record()
torch.cuda.synchronize()
hms = e0.elapsed_time(e1) / 5
peak = 1407.6
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["bf16_tflops_sustained"]
tf = flop / ms / 1e9
print(json.dumps({"bench": "cnn_vtl", "frames": N, "chunk": chunk, "precision": precision, "conv_head_ms": ms,
                  "frames_per_s": N / ms * 1e3, "algorithmic_tflops": tf,
                  "frac_of_measured_sustained_tensor_peak": tf / peak,
                  "tensor_products_per_algorithmic_flop": 3 if precision == "fp16x2" else 1,
                  "descriptor_len": int(d.shape[1]), "hamming_ms": hms, "hamming_pairs_per_s": N * N / hms * 1e3}))
