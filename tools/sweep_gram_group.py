"""Sweep the L2 super-block size (M tiles per group) of the SDAV Gram/score kernel's tile order on the bench workload."""
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
pipe = LoopClosurePipeline(bench.DIMS, precision="fp16x2")
pipe.set_weights(ws, bs)
desc = pipe.encode(torch.from_numpy(frames).cuda(), torch.from_numpy(xy).cuda()).view(bench.N_FRAMES, bench.P, -1)
PREC = os.environ.get("SIM_PREC", "auto")
groups = [int(g) for g in sys.argv[1:]] or [4, 8, 16, 24, 32, 48, 64, 96, 133, 8]
ref = None
for g in groups:
    _lib.call("dlc_debug_set", 4, g)
    S = ops.sdav_similarity(desc, precision=PREC)
    torch.cuda.synchronize()
    if ref is None:
        ref = S.clone()
    same = bool(torch.equal(S, ref))
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.sdav_similarity(desc, precision=PREC)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"mgroup": g, "similarity_ms": e0.elapsed_time(e1) / reps, "identical_to_first": same}))
