"""A/B of the whole bench step with the SDAV Gram kernel on CTA pairs vs single CTAs (sustained: 20 steps per sample,
interleaved rounds)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deeploopcloser_b200 import _lib  # noqa: E402
from deeploopcloser_b200.pipeline import LoopClosurePipeline  # noqa: E402

frames, xy = bench.synthetic_inputs(100)
ws, bs = bench.reference_weights()
pipe = LoopClosurePipeline(bench.DIMS)
pipe.set_weights(ws, bs)
f_d, x_d = torch.from_numpy(frames).cuda(), torch.from_numpy(xy).cuda()
res = {(1, 1): [], (1, 0): [], (0, 0): []}
for _ in range(4):
    for enc_pair, gram_pair in res:
        _lib.call("dlc_debug_set", 6, enc_pair)
        _lib.call("dlc_debug_set", 7, gram_pair)
        for _ in range(3):
            pipe.run(f_d, x_d, k=10)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            pipe.run(f_d, x_d, k=10)
        e1.record()
        torch.cuda.synchronize()
        res[(enc_pair, gram_pair)].append(e0.elapsed_time(e1) / 20)
for (e, g), v in res.items():
    print(json.dumps({"encoder_pair": e, "gram_pair": g, "step_ms": [round(x, 3) for x in v]}))
