"""GPU diagnostic of the fused cnn_vtl head: per-layer error of the implicit-GEMM convolutions against the float64
oracle, and agreement of the int8 descriptors with the explicit-im2col formulation."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200.cnn_vtl import CnnVtl  # noqa: E402
from oracle import cnnvtl as o_cnn  # noqa: E402

for hw in [(67, 83), (192, 240)]:
    for prec in ("fp16x2", "fp16"):
        H, W = hw
        n = 3
        rng = np.random.default_rng(5)
        x = rng.integers(0, 256, (n, H, W, 3)).astype(np.uint8)
        params = o_cnn.make_weights(3)
        sizes = o_cnn.layer_sizes(hw)
        keep = o_cnn.make_keep_columns(sizes, compress_factor=99.0, seed=4)
        net = CnnVtl(input_shape=[n, H, W, 3], batch_size=n, weights=params, keep_cols=keep, precision=prec)
        want = o_cnn.conv_outputs(x.astype(np.float64), params)
        try:
            got = net.conv_outputs(torch.from_numpy(x).cuda())
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(hw, prec, "FAILED:", e)
            continue
        for l, (g, w) in enumerate(zip(got, want)):
            g = g.cpu().numpy().astype(np.float64)
            err = np.abs(g - w) / np.maximum(1.0, np.abs(w))
            print("%s %s conv%d shape %s max err %.3e (at %s) mean %.3e" % (
                hw, prec, l + 1, g.shape, err.max(), np.unravel_index(err.argmax(), err.shape), err.mean()))
        d_f = net._forward_chunk(torch.from_numpy(x).cuda()).cpu().numpy()
        d_e = net._forward_chunk_explicit(torch.from_numpy(x).cuda()).cpu().numpy()
        d_o, _ = o_cnn.descriptors_from_outputs(want, keep)
        print("%s %s descriptors: fused vs explicit differ at %d of %d; fused vs oracle %d; explicit vs oracle %d" % (
            hw, prec, int((d_f != d_e).sum()), d_f.size, int((d_f != d_o).sum()), int((d_e != d_o).sum())))
