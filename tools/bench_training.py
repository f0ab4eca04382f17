"""Training-step timing at the reference's sizes (batch 10 frames x 30 patches, 1681 -> 5 x 2500): one SGD step of
SDAV's train_steps[i] on the B200 next to the float64 oracle step on the host cores.

    python tools/bench_training.py [--layers 0,4] [--cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200.training import DaeStackTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", default="0,4")
ap.add_argument("--cpu", action="store_true", help="also time the float64 oracle step on the host")
args = ap.parse_args()
dims, B, P = [1681, 2500, 2500, 2500, 2500, 2500], 10, 30
rng = np.random.default_rng(0)
x = rng.uniform(0, 1, (B, P, dims[0]))
Ws = [rng.standard_normal((k, n)) / np.sqrt(k) for k, n in zip(dims[:-1], dims[1:])]
bs = [np.zeros(n) for n in dims[1:]]
bds = [np.zeros(k) for k in dims[:-1]]
tr = DaeStackTrainer(dims, patches=P)
tr.set_weights(Ws, bs, bds)
xd = torch.from_numpy(x).float().cuda()
for top in [int(v) for v in args.layers.split(",")]:
    masks = tr.sdav_masks(top, 0.3)
    for _ in range(3):
        tr.step(xd, top, masks)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        tr.step(xd, top, masks)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    g = tr.graphed_step(xd, top, masks)
    for _ in range(3):
        g(xd, masks)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g(xd, masks)
    e1.record()
    torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1) / reps
    # forward + decoder + 2 gradient GEMMs at the top, forward + input-gradient below, weight gradients everywhere
    R = B * P
    flop = 0
    for l in range(top + 1):
        kn = dims[l] * dims[l + 1]
        flop += 2 * R * kn * (2 if l < top else 1)      # forward, d/dx (not needed for layer 0 but cheap to count exactly)
        flop += 2 * R * kn * (2 if l == top else 1)     # weight gradient (two contractions at the top)
    flop += 2 * R * dims[top] * dims[top + 1] * 2        # decoder forward, gradient into the hidden layer
    flop -= 2 * R * dims[0] * dims[1] if top > 0 else 0  # layer 0 has no input gradient
    out = {"bench": "train_step", "loss_layer": top, "batch_frames": B, "rows": R, "gpu_ms_per_step": ms, "gpu_ms_per_step_cuda_graph": ms_graph,
           "algorithmic_gflop_per_step": flop / 1e9, "gpu_tflops": flop / ms / 1e9}
    if args.cpu:
        from oracle import train as o_train
        m_np = [m.cpu().numpy().astype(np.float64) for m in masks]
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            o_train.sdav_train_step(x, Ws, bs, bds, top, m_np)
        out["cpu_oracle_ms_per_step"] = (time.perf_counter() - t0) / n * 1e3
        out["cpu_cores"] = os.cpu_count()
    print(json.dumps(out))
