"""The loop-closure hot path as one object: patch gather -> SDA encode -> SDAV score matrix -> loop candidates.
This is what bench.py times and what smoke() exercises; every stage is a libdlc kernel launch on the current stream."""
import torch

from . import ops


class LoopClosurePipeline:
    def __init__(self, dims=(1681, 2500, 2500, 2500, 2500, 2500), precision="fp16x2", patch=41, swap_xy_quirk=True,
                 mu=0.5, sigma=0.2, a=10.0, b=-10.0):
        self.dims = list(dims)
        self.precision = precision
        self.patch = patch
        self.swap_xy_quirk = swap_xy_quirk
        self.sim_args = dict(mu=mu, sigma=sigma, a=a, b=b)
        self.encoder = ops.SdaEncoder(self.dims, precision)

    def set_weights(self, weights, biases):
        for l, (w, b) in enumerate(zip(weights, biases)):
            self.encoder.set_layer(l, w, b)

    def encode(self, frames, xy):
        """frames uint8 [B,H,W] (CUDA), xy float32 [B,P,2] (CUDA) -> float32 [B*P, D] descriptors."""
        split = self.precision == "fp16x2"
        hi, lo = ops.patch_gather(frames, xy, self.patch, self.swap_xy_quirk, need_lo=split)
        return self.encoder.encode_planes(hi, lo, hi.shape[0])

    def match(self, desc, n_frames, k=10, exclude_band=0):
        P = desc.shape[0] // n_frames
        S = ops.sdav_similarity(desc.view(n_frames, P, -1), precision=self.precision, **self.sim_args)
        cand = ops.topk_rows(S, min(k, max(n_frames - 1, 1)), largest=True, exclude_band=exclude_band)
        return S, cand

    def run(self, frames, xy, k=10, exclude_band=0):
        desc = self.encode(frames, xy)
        S, cand = self.match(desc, frames.shape[0], k, exclude_band)
        return {"descriptors": desc, "similarity": S, "candidates": cand}
