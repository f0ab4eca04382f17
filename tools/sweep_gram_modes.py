"""Gram/score kernel alone on the bench workload: every similarity precision mode on single CTAs and on CTA pairs
(dlc_debug_set key 7). Interleaved rounds, best of R; checks that pair and single-CTA results are bit-identical."""
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
modes = sys.argv[1:] or ["fp16", "fp16r", "fp16x2"]
R, REPS = 3, 10
FLOP = 2.0 * bench.P * bench.P * bench.DIMS[-1] * bench.N_FRAMES * (bench.N_FRAMES - 1) / 2
best, ref, whole = {}, {}, {}
for _ in range(R):
    for mode in modes:
        for pair in (1, 0):
            _lib.call("dlc_debug_set", 7, pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            S = ops.sdav_similarity(desc, precision=mode)  # full call: planes, statistics, probe, tile list
            torch.cuda.synchronize()
            e0.record()
            for _ in range(REPS):
                ops.sdav_similarity(desc, precision=mode)
            e1.record()
            torch.cuda.synchronize()
            whole[(mode, pair)] = min(whole.get((mode, pair), 1e9), e0.elapsed_time(e1) / REPS)
            if mode not in ref:
                ref[mode] = S.clone()
            assert torch.equal(S, ref[mode]), (mode, pair)
            _lib.call("dlc_sdav_debug_gram_only", 1)
            try:
                ops.sdav_similarity(desc, precision=mode)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(REPS):
                    ops.sdav_similarity(desc, precision=mode)
                e1.record()
                torch.cuda.synchronize()
            finally:
                _lib.call("dlc_sdav_debug_gram_only", 0)
            best[(mode, pair)] = min(best.get((mode, pair), 1e9), e0.elapsed_time(e1) / REPS)
_lib.call("dlc_debug_set", 7, 1)
for (mode, pair), ms in best.items():
    print(json.dumps({"sim_precision": mode, "cta_pair": pair, "gram_ms": round(ms, 4),
                      "similarity_call_ms": round(whole[(mode, pair)], 4),
                      "algorithmic_tflops": round(FLOP / ms / 1e9, 1), "identical_across_pair_modes": True}))
