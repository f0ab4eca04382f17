"""Which arithmetic is right where `auto` (one product + refinement) and `fp16x2` (three products) disagree?
Evaluates the float64 oracle on the disagreeing frame pairs of the bench workload."""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deeploopcloser_b200 import ops  # noqa: E402
from deeploopcloser_b200.pipeline import LoopClosurePipeline  # noqa: E402
from oracle import similarity as o_sim  # noqa: E402

frames, xy = bench.synthetic_inputs(100)
ws, bs = bench.reference_weights()
pipe = LoopClosurePipeline(bench.DIMS)
pipe.set_weights(ws, bs)
desc = pipe.encode(torch.from_numpy(frames).cuda(), torch.from_numpy(xy).cuda()).view(bench.N_FRAMES, bench.P, -1)
from deeploopcloser_b200 import _lib
if len(sys.argv) > 2:
    _lib.call("dlc_debug_set", 7, int(sys.argv[2]))
S3 = ops.sdav_similarity(desc, precision="fp16x2").cpu().numpy().astype(np.float64)
Sa = ops.sdav_similarity(desc, precision=sys.argv[1] if len(sys.argv) > 1 else "auto").cpu().numpy().astype(np.float64)
print(ops.sdav_similarity_stats(bench.N_FRAMES, bench.P, bench.DIMS[-1]))
rel = np.abs(Sa - S3) / np.maximum(1.0, np.abs(S3))
ii, jj = np.nonzero(np.triu(rel > 1e-3, 1))
print("disagreeing pairs:", len(ii))
d = desc.cpu().numpy().astype(np.float64)
w = o_sim.distinctive_weights(d)
a_ok = x_ok = 0
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for i, j in list(zip(ii, jj))[:40]:
        want = o_sim.similarity_score(d[i], d[j], w)
        ea, e3 = abs(Sa[i, j] - want) / max(1, abs(want)), abs(S3[i, j] - want) / max(1, abs(want))
        a_ok += ea < 1e-3
        x_ok += e3 < 1e-3
        if e3 >= 1e-3 and x_ok + a_ok < 12:
            dist = ((d[i][:, None, :] - d[j][None, :, :]) ** 2).sum(-1)
            srt = np.sort(dist, axis=1)
            gaps = srt[:, 1] - srt[:, 0]
            print("x2 wrong: pair", i, j, "oracle %.4f auto %.4f x2 %.4f" % (want, Sa[i, j], S3[i, j]), "smallest NN gaps of the 30 rows",
                  np.sort(gaps)[:3], "dup rows in j:", 30 - len(np.unique(d[j], axis=0)), "in i:", 30 - len(np.unique(d[i], axis=0)))
        if ea >= 1e-3:
            # which rows differ: nearest-neighbour gaps of the rows of frame i against frame j
            dist = ((d[i][:, None, :] - d[j][None, :, :]) ** 2).sum(-1)
            srt = np.sort(dist, axis=1)
            gaps = srt[:, 1] - srt[:, 0]
            print("pair", i, j, "oracle", want, "auto", Sa[i, j], "x2", S3[i, j], "smallest gaps", np.sort(gaps)[:3])
print("of the first 40 disagreements: auto right %d, three-product right %d" % (a_ok, x_ok))
