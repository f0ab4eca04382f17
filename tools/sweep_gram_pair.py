"""A/B of the SDAV Gram kernel on CTA pairs vs single CTAs on the bench workload (interleaved rounds, best of 4)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deeploopcloser_b200 import _lib, ops  # noqa: E402
from deeploopcloser_b200.pipeline import LoopClosurePipeline  # noqa: E402

frames, xy = bench.synthetic_inputs(100)
ws, bs = bench.reference_weights()
pipe = LoopClosurePipeline(bench.DIMS)
pipe.set_weights(ws, bs)
desc = pipe.encode(torch.from_numpy(frames).cuda(), torch.from_numpy(xy).cuda()).view(bench.N_FRAMES, bench.P, -1)
best = {0: 1e9, 1: 1e9}
ref = None
for _ in range(4):
    for pair in (1, 0):
        _lib.call("dlc_debug_set", 6, pair)
        S = ops.sdav_similarity(desc)
        torch.cuda.synchronize()
        if ref is None:
            ref = S.clone()
        assert torch.equal(S, ref)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.sdav_similarity(desc)
        e1.record()
        torch.cuda.synchronize()
        best[pair] = min(best[pair], e0.elapsed_time(e1) / 10)
for pair, ms in best.items():
    print(json.dumps({"cta_pair": pair, "similarity_ms_best_of_4": ms}))
