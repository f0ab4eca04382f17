"""The loop-closure hot path as one object: patch gather -> SDA encode -> SDAV score matrix -> loop candidates.
This is what bench.py times and what smoke() exercises; every stage is a libdlc kernel launch on the current stream."""
import torch

from . import ops


class LoopClosurePipeline:
    def __init__(self, dims=(1681, 2500, 2500, 2500, 2500, 2500), precision="fp16x2", patch=41, swap_xy_quirk=True,
                 mu=0.5, sigma=0.2, a=10.0, b=-10.0, sim_precision="auto", raw_pixels=True):
        self.dims = list(dims)
        # encoder arithmetic: "fp16x2" (three tensor products: holds 1e-3 on the reference's N(0,1) initialisation),
        # "fp16" (one), "fp16x2a16" (two) or "auto" (the cheapest of them that stays inside the descriptor tolerance
        # on a sample of the first batch - one product for trained-like weights, see dlc_sda_probe)
        self.precision = precision
        # "auto" (default): a device-side probe picks one fp16 product + exact refinement of the ambiguous rows when
        # few rows need it (< 1.2 %; bit-identical duplicate patches never do), else the three-product kernel.
        # "fp16x2": always three products. "fp16r": always refine (most exact, slower when many rows are near-ties).
        self.sim_precision = sim_precision
        self.patch = patch
        self.swap_xy_quirk = swap_xy_quirk
        self.sim_args = dict(mu=mu, sigma=sigma, a=a, b=b)
        # raw_pixels: the patch planes hold pixel values 0..255 (exact in fp16) and layer 0 absorbs the /255 of
        # CvInputParser.py:27 - no input rounding, and two tensor-core products for layer 0 instead of three
        self.raw_pixels = bool(raw_pixels)
        self.encoder = ops.SdaEncoder(self.dims, precision, input_u8=self.raw_pixels)

    def set_weights(self, weights, biases):
        for l, (w, b) in enumerate(zip(weights, biases)):
            self.encoder.set_layer(l, w, b)

    def detect(self, frames, n=30, **detector_args):
        """frames uint8 [B,H,W] (CUDA) -> float32 [B,n,2] keypoint centres (x, y), best response first: the
        fast-Hessian detector in front of the patch gather (get_top_n_key_points, CvInputParser.py:36-46). Frames
        with fewer than n keypoints are padded with the image centre (the reference would hand the encoder fewer
        than 30 patches and fail)."""
        xy, _, _ = ops.surf_detect(frames, top_n=n, **detector_args)
        return xy

    def encode(self, frames, xy=None):
        """frames uint8 [B,H,W] (CUDA), xy float32 [B,P,2] (CUDA; None: detect P = 30 keypoints per frame on the
        device) -> float32 [B*P, D] descriptors."""
        if xy is None:
            xy = self.detect(frames, 30)
        if self.raw_pixels:
            hi, lo = ops.patch_gather_u8(frames, xy, self.patch, self.swap_xy_quirk), None
        else:
            hi, lo = ops.patch_gather(frames, xy, self.patch, self.swap_xy_quirk, need_lo=self.encoder.needs_lo_input())
        return self.encoder.encode_planes(hi, lo, hi.shape[0])

    def similarity(self, desc, n_frames):
        """float32 [n_frames*P, D] descriptors -> float32 [n_frames, n_frames] SDAV score matrix (i<j mirrored,
        diagonal -1: create_similarity_matrix.py:31-38)."""
        P = desc.shape[0] // n_frames
        self.last_similarity = ops.sdav_similarity(desc.view(n_frames, P, -1), precision=self.sim_precision,
                                                   **self.sim_args)
        return self.last_similarity

    def match(self, desc, n_frames, k=10, exclude_band=0):
        S = self.similarity(desc, n_frames)
        cand = ops.topk_rows(S, min(k, max(n_frames - 1, 1)), largest=True, exclude_band=exclude_band)
        return S, cand

    @staticmethod
    def host_bytes_per_step(frames_h, xy_h, k):
        """(host->device, device->host) bytes one run_host_stream step moves: the frames and keypoints in, the
        [N, k] candidate scores (float32) and indices (int64) out."""
        n = frames_h.shape[0]
        kk = min(k, max(n - 1, 1))
        return int(frames_h.numel() * frames_h.element_size() + xy_h.numel() * xy_h.element_size()), int(n * kk * 12)

    def run(self, frames, xy=None, k=10, exclude_band=0):
        desc = self.encode(frames, xy)
        S, cand = self.match(desc, frames.shape[0], k, exclude_band)
        return {"descriptors": desc, "similarity": S, "candidates": cand}

    def run_host_stream(self, batches, k=10, exclude_band=0):
        """Process a stream of HOST batches [(frames uint8 [B,H,W], xy float32 [B,P,2]), ...] (pinned torch tensors)
        end to end and return the candidate lists [(scores [B,k], idx [B,k]), ...] in pinned host memory (views of a
        buffer owned by the pipeline: valid until the next call).
        The upload of batch i+1 runs on a copy stream while batch i computes (double-buffered device inputs), the
        candidate lists come back with asynchronous D2H copies; one synchronisation at the end."""
        batches = list(batches)
        if not batches:
            return []
        compute = torch.cuda.current_stream()
        copy = self._copy_stream = getattr(self, "_copy_stream", None) or torch.cuda.Stream()
        bufs, ready, consumed = [None, None], [None, None], [None, None]

        def upload(i):
            slot = i & 1
            f_h, x_h = batches[i]
            if consumed[slot] is not None:
                copy.wait_event(consumed[slot])      # the compute stream is done reading this slot
            with torch.cuda.stream(copy):
                if bufs[slot] is None or bufs[slot][0].shape != f_h.shape or bufs[slot][1].shape != x_h.shape:
                    bufs[slot] = (torch.empty(f_h.shape, dtype=f_h.dtype, device="cuda"),
                                  torch.empty(x_h.shape, dtype=x_h.dtype, device="cuda"))
                bufs[slot][0].copy_(f_h, non_blocking=True)
                bufs[slot][1].copy_(x_h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            ready[slot] = ev

        copy.wait_stream(compute)
        upload(0)
        outs = []
        for i in range(len(batches)):
            slot = i & 1
            if i + 1 < len(batches):
                upload(i + 1)
            compute.wait_event(ready[slot])
            f_d, x_d = bufs[slot]
            r = self.run(f_d, x_d, k, exclude_band)
            ev = torch.cuda.Event()
            ev.record(compute)
            consumed[slot] = ev
            # pinned result buffers for the whole stream, allocated once per shape (cudaHostAlloc per batch would
            # serialise the pipeline)
            per = tuple(r["candidates"][0].shape)
            pool = getattr(self, "_pinned_out", None)
            if pool is None or tuple(pool[0].shape[1:]) != per or pool[0].shape[0] < len(batches):
                slots = max(64, len(batches))            # grow-only, so a longer stream does not re-allocate each call
                pool = self._pinned_out = (torch.empty((slots,) + per, dtype=torch.float32).pin_memory(),
                                           torch.empty((slots,) + per, dtype=torch.int64).pin_memory())
            s_h, i_h = pool[0][i], pool[1][i]
            s_h.copy_(r["candidates"][0], non_blocking=True)
            i_h.copy_(r["candidates"][1], non_blocking=True)
            outs.append((s_h, i_h))
        compute.synchronize()
        return outs


class ShardedSequencePipeline(LoopClosurePipeline):
    """ONE sequence over the GPUs of a box (strong scaling of BASELINE config 2): one process per GPU
    (torch.distributed, NCCL). Frames are dealt in contiguous blocks: every rank gathers + encodes its block, the
    descriptors are all-gathered (every rank needs all of them for the score matrix), every rank evaluates its
    interleaved tile rows of the SDAV score matrix (`dlc_sdav_similarity_part`), the parts are summed with an
    all-reduce, and the candidate lists are selected on every rank (identical everywhere). Two exchange steps:
    N*30*D*4 bytes of descriptors and N*N*4 bytes of scores; everything else is rank-local."""

    def __init__(self, *args, group=None, **kwargs):
        super().__init__(*args, **kwargs)
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    @staticmethod
    def frame_block(n_frames, rank, world):
        """(start, end, per): frames [start, end) belong to `rank`; `per` = block length used for the gather."""
        per = -(-n_frames // world)
        start = min(rank * per, n_frames)
        return start, min(start + per, n_frames), per

    def run(self, frames, xy, k=10, exclude_band=0):
        """frames uint8 [N,H,W] and xy float32 [N,P,2]: the WHOLE sequence, resident on every rank."""
        n, P = frames.shape[0], xy.shape[1]
        D = self.dims[-1]
        start, end, per = self.frame_block(n, self.rank, self.world)
        local = torch.zeros((per * P, D), dtype=torch.float32, device=frames.device)
        if end > start:
            local[:(end - start) * P] = self.encode(frames[start:end], xy[start:end])
        if self.world > 1:
            gathered = torch.empty((self.world * per * P, D), dtype=torch.float32, device=frames.device)
            self.dist.all_gather_into_tensor(gathered, local, group=self.group)
            desc = gathered[:n * P]
        else:
            desc = local[:n * P]
        S = ops.sdav_similarity_part(desc.view(n, P, D), self.rank, self.world, precision=self.sim_precision,
                                     **self.sim_args)
        if self.world > 1:
            self.dist.all_reduce(S, group=self.group)
        cand = ops.topk_rows(S, min(k, max(n - 1, 1)), largest=True, exclude_band=exclude_band)
        return {"descriptors": desc, "similarity": S, "candidates": cand}
