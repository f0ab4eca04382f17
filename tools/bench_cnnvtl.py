"""cnn_vtl timing (BASELINE.json config 3 shape: 1063 frames 192x240x3, conv head + Hamming matrix)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200 import ops  # noqa: E402
from deeploopcloser_b200.cnn_vtl import CnnVtl  # noqa: E402

N, H, W = 1063, 192, 240
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 128
net = CnnVtl(input_shape=[N, H, W, 3], batch_size=chunk, weights="synthetic", seed=4)
x = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device="cuda")


def forward():
    outs = []
    for s in range(0, N, chunk):
        outs.append(net._forward_chunk(x[s:s + chunk]))
    return torch.cat(outs)


for _ in range(2):
    d = forward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    d = forward()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
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
print(json.dumps({"bench": "cnn_vtl", "frames": N, "chunk": chunk, "conv_head_ms": ms, "frames_per_s": N / ms * 1e3,
                  "algorithmic_tflops": flop / ms / 1e9, "descriptor_len": int(d.shape[1]), "hamming_ms": hms,
                  "hamming_pairs_per_s": N * N / hms * 1e3}))
