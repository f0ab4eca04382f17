"""One GPU: does running the encoder of sequence i+1 on a second stream under the score-matrix stage of sequence i
(memory-bound second pass / preparation / top-k under tensor-bound GEMMs) beat run() after run()?
ShardedSequencePipeline with world 1 is that two-stream pipeline (no collectives are issued).

    python tools/bench_overlap_1gpu.py [--weights normal|xavier] [--steps 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deeploopcloser_b200.pipeline import LoopClosurePipeline, ShardedSequencePipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--weights", default="normal")
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
frames, xy = bench.synthetic_inputs(100)
ws, bs = bench.reference_weights() if args.weights == "normal" else bench.xavier_weights()
f_d, x_d = torch.from_numpy(frames).cuda(), torch.from_numpy(xy).cuda()
seqs = [(f_d, x_d)] * args.steps


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


out = {"weights": args.weights, "steps": args.steps}
ref = LoopClosurePipeline(bench.DIMS, precision="auto")
ref.set_weights(ws, bs)
want = ref.run(f_d, x_d, k=bench.K_CAND)["candidates"]
out["serial_ms"] = timed(lambda: ref.run_many(seqs, k=bench.K_CAND))
pipe = ShardedSequencePipeline(bench.DIMS, precision="auto")
pipe.set_weights(ws, bs)
got = pipe.run_many(seqs[:3], k=bench.K_CAND)
torch.cuda.synchronize()
out["same_candidates"] = all(torch.equal(g[1], want[1]) and torch.equal(g[0], want[0]) for g in got)
out["staged_serial_ms"] = timed(lambda: [pipe.run(f_d, x_d, k=bench.K_CAND) for _ in range(args.steps)])
out["two_stream_ms"] = timed(lambda: pipe.run_many(seqs, k=bench.K_CAND))
print(json.dumps(out))
