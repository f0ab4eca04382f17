"""Two passes of the fused cnn_vtl head over a batch (for ncu captures: `-k regex:gemm_tc_kernel -s 5 -c 5`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200.cnn_vtl import CnnVtl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
if os.environ.get("DLC_DEBUG_SET"):            # A/B switches, e.g. DLC_DEBUG_SET="6=0,5=0" (see dlc_debug_set)
    from deeploopcloser_b200 import _lib
    for kv in os.environ["DLC_DEBUG_SET"].split(","):
        k, v = kv.split("=")
        _lib.call("dlc_debug_set", int(k), int(v))
net = CnnVtl(input_shape=[n, 192, 240, 3], batch_size=n, weights="synthetic", seed=4)
x = torch.randint(0, 256, (n, 192, 240, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    d = net._forward_chunk(x)
torch.cuda.synchronize()
print("ok", tuple(d.shape))
