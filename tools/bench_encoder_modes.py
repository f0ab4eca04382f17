#!/usr/bin/env python
"""Encoder arithmetic modes on the config-2 workload (1063 frames): time per encode and descriptor error against the
float64 oracle on sampled frames, for the reference's N(0,1) initialisation and Xavier-scaled (trained-like) weights.
Developer tool (uses the oracle as the checker)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deeploopcloser_b200.pipeline import LoopClosurePipeline  # noqa: E402
from oracle import patches as o_patch  # noqa: E402
from oracle import sda as o_sda  # noqa: E402

frames, xy = bench.synthetic_inputs(100)
f_d, x_d = torch.from_numpy(frames).cuda(), torch.from_numpy(xy).cuda()
picks = [0, 377, 1062]
x = np.concatenate([o_patch.extract_patches(frames[i], xy[i]) for i in picks])
for wname in ("normal", "xavier"):
    ws, bs = o_sda.make_weights(bench.DIMS, seed=1 if wname == "normal" else 2, scale=wname)
    want = o_sda.sda_forward(x, ws, bs)
    for prec in ("fp16x2", "fp16x2a16", "fp16", "auto"):
        pipe = LoopClosurePipeline(bench.DIMS, precision=prec)
        pipe.set_weights(ws, bs)
        for _ in range(3):
            d = pipe.encode(f_d, x_d)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            d = pipe.encode(f_d, x_d)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        got = torch.cat([d[i * 30:(i + 1) * 30] for i in picks]).cpu().numpy()
        err = float(np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))))
        print(json.dumps({"weights": wname, "precision": prec, "chosen": pipe.encoder.chosen_precision(),
                          "probe_err_1prod_2prod": pipe.encoder.probe_stats(), "ms_per_encode": round(ms, 3),
                          "algorithmic_tflops": round(1063 * bench.ENC_FLOP_PER_FRAME / ms / 1e9, 1),
                          "max_rel_err_vs_oracle_3_frames": err}), flush=True)
        del pipe
