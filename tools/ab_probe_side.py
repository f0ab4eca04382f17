"""A/B: similarity call with the precision probe on its side stream (dlc_debug_set key 9) vs serial."""
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
best, ref = {}, None
for _ in range(4):
    for side in (1, 0):
        _lib.call("dlc_debug_set", 9, side)
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
        best[side] = min(best.get(side, 1e9), e0.elapsed_time(e1) / 10)
_lib.call("dlc_debug_set", 9, 1)
for side, ms in best.items():
    print(json.dumps({"probe_side_stream": side, "similarity_call_ms_best_of_4": round(ms, 4), "identical": True}))
