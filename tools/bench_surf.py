"""Keypoint detector timing (dlc_surf_detect): the kennedylong-shaped sequence and the streaming batch."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200 import ops  # noqa: E402

PEAK_GBS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
for name, B, H, W, chunk in (("sequence 1063 x 192x240", 1063, 192, 240, 1063), ("streaming batch 256 x 480x640", 256, 480, 640, 64),
                             ("streaming batch 256 x 480x640, one chunk", 256, 480, 640, 256)):
    rng = np.random.default_rng(7)
    # smooth synthetic texture (noise frames have thousands of maxima per frame, natural images hundreds)
    base = rng.integers(0, 256, (B, H // 8 + 2, W // 8 + 2)).astype(np.float32)
    img = torch.nn.functional.interpolate(torch.from_numpy(base)[:, None], size=(H, W), mode="bicubic")[:, 0]
    img = img.clamp(0, 255).round().to(torch.uint8).cuda().contiguous()
    for _ in range(3):
        xy, info, found = ops.surf_detect(img, 30, chunk=chunk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ops.surf_detect(img, 30, chunk=chunk)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # compulsory bytes per frame: pixels read once, integral image written (twice: row pass, column pass) and read
    # once from HBM (the 32 lookups per sample and layer hit L1/L2); the determinant layers live in shared memory
    alg = B * (H * W + 3 * 4 * (H + 1) * (W + 1))
    print(json.dumps({"workload": name, "ms": round(ms, 4), "frames_per_s": round(B / ms * 1e3),
                      "keypoints_per_frame_mean": float(found.float().mean()), "min_found": int(found.min()),
                      "algorithmic_GB": round(alg / 1e9, 3), "achieved_GBs": round(alg / ms / 1e6, 1),
                      "frac_of_hbm_peak": round(alg / ms / 1e6 / PEAK_GBS, 3)}))
