"""The loop-closure hot path as one object: patch gather -> SDA encode -> SDAV score matrix -> loop candidates.
This is what bench.py times and what smoke() exercises; every stage is a libdlc kernel launch on the current stream."""
import contextlib
import os

import torch

from . import _lib, ops

# NVTX ranges around the stages of a step (tracing row of SURVEY 5): visible in an Nsight Systems / ncu --nvtx
# timeline. Off unless DLC_NVTX=1 - range pushes are host calls on the launch path.
_NVTX = os.environ.get("DLC_NVTX", "0") not in ("", "0")


@contextlib.contextmanager
def nvtx_range(name):
    if _NVTX:
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


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
        with nvtx_range("dlc.encode"):
            desc = self.encode(frames, xy)
        with nvtx_range("dlc.match"):
            S, cand = self.match(desc, frames.shape[0], k, exclude_band)
        return {"descriptors": desc, "similarity": S, "candidates": cand}

    def run_many(self, sequences, k=10, exclude_band=0):
        """A stream of device-resident sequences [(frames, xy), ...] -> [(scores [N,k], idx [N,k]), ...]: on one GPU
        simply run() after run() (the sharded pipeline overlaps the encoder of the next sequence with the exchange
        stage of the current one)."""
        return [self.run(f, x, k, exclude_band)["candidates"] for f, x in sequences]

    def run_host_stream(self, batches, k=10, exclude_band=0):
        """Process a stream of HOST batches [(frames uint8 [B,H,W], xy float32 [B,P,2]), ...] (pinned torch tensors)
        end to end and return the candidate lists [(scores [B,k], idx [B,k]), ...] in pinned host memory (views of a
        buffer owned by the pipeline: valid until the next call).
        The upload of batch i+1 runs on a copy stream while batch i computes (double-buffered device inputs), the
        candidate lists come back with asynchronous D2H copies; one synchronisation at the end."""
        return self._host_stream([(f, x, f.shape[0]) for f, x in batches], k, exclude_band,
                                 lambda f, x, n: self.run(f, x, k, exclude_band))

    def _host_stream(self, batches, k, exclude_band, run):
        batches = list(batches)
        if not batches:
            return []
        compute = torch.cuda.current_stream()
        copy = self._copy_stream = getattr(self, "_copy_stream", None) or torch.cuda.Stream()
        bufs, ready, consumed = [None, None], [None, None], [None, None]

        def upload(i):
            slot = i & 1
            f_h, x_h, _ = batches[i]
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
            r = run(f_d, x_d, batches[i][2])
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


SM_RESERVE_DEFAULT = 0


class _SmReserve:
    """Context manager: dlc_set_sm_reserve(n) for the duration of a sharded step (process-wide knob, restored to 0)."""

    def __init__(self, n):
        self.n = n

    def __enter__(self):
        if self.n:
            _lib.call("dlc_set_sm_reserve", int(self.n))

    def __exit__(self, *exc):
        if self.n:
            _lib.call("dlc_set_sm_reserve", 0)
        return False


class ShardedSequencePipeline(LoopClosurePipeline):
    """ONE sequence over the GPUs of a box (strong scaling of BASELINE config 2): one process per
    GPU (torch.distributed, NCCL). Frames are dealt in contiguous blocks of `per = ceil(N / world)`; every rank gathers
    and encodes ITS block, and the score matrix is evaluated through the staged C ABI (dlc_sdav_stage_*) so that no
    work is replicated:

      encode block -> [fp32 descriptors: all-gather, in the background on a second communicator]
                   -> column sums -> all-gather (world x D doubles) -> mean, weights
                   -> centred fp16 operand planes + row statistics + precision probe of the block
                   -> all-gather planes (N x P x ld x 2 bytes) and stats blocks
                   -> Gram + argmin + score of this rank's interleaved tile rows (one tensor product; ambiguous pairs
                      are listed) -> wait for the descriptors -> exact second pass over the listed pairs
                   -> all-reduce (sum) of the N x N scores -> loop candidates on every rank.

    The exchanged operand is the fp16 plane (half the bytes of the descriptors); only the second pass needs the
    float32 descriptors, whose all-gather overlaps everything up to it."""

    def __init__(self, *args, group=None, **kwargs):
        super().__init__(*args, **kwargs)
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._bg_group = None       # second communicator: the descriptor all-gather runs next to the other collectives
        self._buf_key = None
        # pipelined streams (run_many / run_host_stream): 2 = the encoder of step i+1 on a side stream under the exchange
        # + score stages of step i; 3 = encoder, exchange and score each on a stream of their own (steps i+2, i+1, i).
        # Measured on 8 GPUs, see DESIGN.md 6.
        self.pipeline_stages = 2
        # SMs the library's persistent kernels leave free while this pipeline runs (dlc_set_sm_reserve), so that NCCL's
        # kernels of the exchange stage start at once next to a running encoder layer. None: leave the process-wide
        # setting alone. Measured on 8 GPUs, see DESIGN.md 6.
        self.sm_reserve = int(os.environ.get("DLC_SM_RESERVE", SM_RESERVE_DEFAULT)) if self.world > 1 else None

    @staticmethod
    def frame_block(n_frames, rank, world):
        """(start, end, per): frames [start, end) belong to `rank`; `per` = block length used for the gathers."""
        per = -(-n_frames // world)
        start = min(rank * per, n_frames)
        return start, min(start + per, n_frames), per

    # ---- device stages (overridden by CPU stand-ins in tests/test_distributed_cpu.py)
    def _alloc(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device="cuda")

    def _buffers(self, n, P):
        """Device buffers for an n-frame sequence: the score matrix and two SLOTS of everything that is exchanged
        (descriptors, planes, stats blocks, column sums, weights, centring vector) - slot i & 1 belongs to step i of a
        pipelined stream, so the exchange stage of step i+1 can run while step i is in its score-matrix stage. The second
        slot is allocated on first use."""
        D = self.dims[-1]
        key = (n, P, D, self.world)
        if self._buf_key != key:
            self._b = {"S": self._alloc((n, n), torch.float32), "slots": [self._slot(n, P), None]}
            self._b.update(self._b["slots"][0])      # run_block / tests address slot 0 directly
            self._buf_key = key
        return self._b

    def _slot(self, n, P):
        D = self.dims[-1]
        per = -(-n // self.world)
        ld = ops.plane_ld(D)
        rows = self.world * per * P
        return {
            "desc": self._alloc((rows, D), torch.float32),
            "plane": self._alloc((rows, ld), torch.float16),
            "plane_lo": self._alloc((rows, ld), torch.float16) if self.sim_precision == "fp16x2" else None,
            "stats": self._alloc((self.world, ops.sdav_stage_stats_bytes(per)), torch.uint8),
            "colsums": self._alloc((self.world, 2 * D), torch.float64),
            "w": self._alloc((D,), torch.float64), "mean": self._alloc((D,), torch.float32),
            "work": None,
        }

    def _encode_into(self, frames, xy, out):
        if xy is None:
            xy = self.detect(frames, 30)
        if self.raw_pixels:
            hi, lo = ops.patch_gather_u8(frames, xy, self.patch, self.swap_xy_quirk), None
        else:
            hi, lo = ops.patch_gather(frames, xy, self.patch, self.swap_xy_quirk, need_lo=self.encoder.needs_lo_input())
        self.encoder.encode_planes(hi, lo, hi.shape[0], out=out)

    def _stage_colsum(self, desc_local, out):
        ops.sdav_stage_colsum(desc_local, out)

    def _stage_weights(self, colsums, rows_total, w, mean):
        ops.sdav_stage_weights(colsums, rows_total, w, mean, self.sim_args.get("mu", 0.5), self.sim_args.get("sigma", 0.2))

    def _stage_prepare(self, desc_local, n_local, per, P, w, mean, plane_local, plane_lo_local, stats_local):
        ops.sdav_stage_prepare(desc_local, n_local, per, P, w, mean, self.sim_precision, plane_local, plane_lo_local,
                               stats_local)

    def _stage_gram(self, b, per, n, P):
        """b: one slot + the score matrix "S"."""
        ops.sdav_stage_gram(b["plane"], b["plane_lo"], b["stats"], self.world, per, n, P, self.dims[-1], self.sim_precision,
                            self.rank, b["S"], a=self.sim_args.get("a", 10.0), b=self.sim_args.get("b", -10.0))

    def _stage_fix(self, b, n, P):
        ops.sdav_stage_fix(b["plane"], b["desc"], n, P, self.dims[-1], self.sim_precision, self.rank, self.world, b["S"],
                           a=self.sim_args.get("a", 10.0), b=self.sim_args.get("b", -10.0))

    def _all_gather(self, full, rank_slice, group, async_op=False):
        return self.dist.all_gather_into_tensor(full, rank_slice, group=group, async_op=async_op)

    def run(self, frames, xy, k=10, exclude_band=0):
        """frames uint8 [N,H,W] and xy float32 [N,P,2]: the WHOLE sequence, resident on every rank (each rank reads
        only its block)."""
        n = frames.shape[0]
        start, end, _ = self.frame_block(n, self.rank, self.world)
        return self.run_block(frames[start:end], None if xy is None else xy[start:end], n, k, exclude_band)

    def run_block(self, frames_local, xy_local, n, k=10, exclude_band=0, P=None):
        """frames_local / xy_local: this rank's block of the N-frame sequence (frame_block(N, rank, world))."""
        P = P or (xy_local.shape[1] if xy_local is not None else 30)
        b = self._buffers(n, P)
        slot = b["slots"][0]
        with _SmReserve(getattr(self, "sm_reserve", None)):
            self._encode_block(frames_local, xy_local, n, P, slot["desc"])
            self._exchange_block(slot, n, P)
            return self._score_block(b, slot, n, P, k, exclude_band)

    def _encode_block(self, frames_local, xy_local, n, P, desc_all):
        """Stage A (rank-local, no collective): patch gather + encoder of this rank's frames into its slice of the
        float32 descriptor buffer."""
        start, end, per = self.frame_block(n, self.rank, self.world)
        n_local = end - start
        assert frames_local.shape[0] == n_local, "expected this rank's block of %d frames" % n_local
        lo_r, hi_r = self.rank * per * P, (self.rank + 1) * per * P
        desc_local = desc_all[lo_r:hi_r]
        if n_local < per:
            desc_local[n_local * P:].zero_()          # padded tail of the last block(s): gathered but never read
        if n_local:
            with nvtx_range("dlc.encode_block"):
                self._encode_into(frames_local, xy_local, desc_local[:n_local * P])

    def _exchange_block(self, slot, n, P):
        """Stage B (the collectives): mean / weights, centred planes + statistics + probe of this rank's block, and
        their all-gathers; the float32 descriptors start travelling in the background (second communicator)."""
        start, end, per = self.frame_block(n, self.rank, self.world)
        n_local = end - start
        lo_r, hi_r = self.rank * per * P, (self.rank + 1) * per * P
        desc_local = slot["desc"][lo_r:hi_r]
        slot["work"] = None
        with nvtx_range("dlc.exchange_block"):
            if self.world > 1:
                if self._bg_group is None:
                    self._bg_group = self.dist.new_group(list(range(self.world))) if self.group is None else self.group
                # float32 descriptors of all frames: only the second pass reads them -> gathered in the background
                slot["work"] = self._all_gather(slot["desc"], desc_local, self._bg_group, async_op=True)
            self._stage_colsum(desc_local[:n_local * P], slot["colsums"][self.rank])
            if self.world > 1:
                self._all_gather(slot["colsums"].view(-1), slot["colsums"][self.rank], self.group)
            self._stage_weights(slot["colsums"], n * P, slot["w"], slot["mean"])
            plane_lo_local = None if slot["plane_lo"] is None else slot["plane_lo"][lo_r:hi_r]
            self._stage_prepare(desc_local, n_local, per, P, slot["w"], slot["mean"], slot["plane"][lo_r:hi_r],
                                plane_lo_local, slot["stats"][self.rank])
            if self.world > 1:
                self._all_gather(slot["plane"].view(-1), slot["plane"][lo_r:hi_r].view(-1), self.group)
                if plane_lo_local is not None:
                    self._all_gather(slot["plane_lo"].view(-1), plane_lo_local.view(-1), self.group)
                self._all_gather(slot["stats"].view(-1), slot["stats"][self.rank], self.group)

    def _score_block(self, b, slot, n, P, k, exclude_band):
        """Stage C: this rank's tile rows of the score matrix, the exact second pass (waits for the descriptor
        gather), the all-reduce of the parts and the loop candidates."""
        per = -(-n // self.world)
        view = dict(slot, S=b["S"])
        with nvtx_range("dlc.score_block"):
            self._stage_gram(view, per, n, P)
            if slot["work"] is not None:
                slot["work"].wait()                    # the current stream waits for the descriptor gather
                slot["work"] = None
            self._stage_fix(view, n, P)
            S = b["S"]
            if self.world > 1:
                self.dist.all_reduce(S, group=self.group)
            self.last_similarity = S
            cand = ops.topk_rows(S, min(k, max(n - 1, 1)), largest=True, exclude_band=exclude_band)
        return {"descriptors": slot["desc"][:n * P], "similarity": S, "candidates": cand}

    def _pipelined(self, sequences, k, exclude_band, host):
        with _SmReserve(getattr(self, "sm_reserve", None)):
            return self._pipelined_steps(sequences, k, exclude_band, host)

    def run_many(self, sequences, k=10, exclude_band=0):
        """A stream of sequences [(frames uint8 [N,H,W], xy float32 [N,P,2]), ...] resident on the device (every rank
        holds them; each reads its block) -> [(scores [N,k], idx [N,k]), ...]. Successive sequences are PIPELINED: the
        encoder of sequence i+1 (rank-local, on its own stream, into the other descriptor slot) runs while sequence i
        is in its exchange / score-matrix stage, so the GPU works through the NCCL waits that dominate a split
        sequence. Per-sequence results are the same as run()."""
        return self._pipelined(list(sequences), k, exclude_band, host=False)

    def run_host_stream(self, batches, k=10, exclude_band=0):
        """Stream of HOST sequences [(frames uint8 [N,H,W], xy float32 [N,P,2]), ...] (pinned, the same on every rank):
        each rank uploads only ITS block of every sequence (on the encoder stream, ahead of the encoder that reads
        it), the steps are pipelined like run_many, and the candidate lists come back with asynchronous D2H copies
        into pinned buffers owned by the pipeline (valid until the next call); one synchronisation at the end."""
        return self._pipelined(list(batches), k, exclude_band, host=True)

    def _pipelined_steps(self, sequences, k, exclude_band, host):
        """Three stages per step: A = upload + encode (rank-local), B = exchange (collectives), C = score matrix +
        candidates. A of step i+1 runs on a side stream, into slot (i+1) & 1, while B and C of step i run on the
        caller's stream (pipeline_stages = 2) or B has a stream of its own as well (3). Collectives are issued in one
        program order on every rank (C_i's all-reduce before B_{i+1}'s all-gathers), as NCCL requires."""
        if not sequences:
            return []
        main = torch.cuda.current_stream()
        side = self._enc_stream = getattr(self, "_enc_stream", None) or torch.cuda.Stream()
        xch = self._xch_stream = getattr(self, "_xch_stream", None) or torch.cuda.Stream()
        n0, P0 = sequences[0][0].shape[0], sequences[0][1].shape[1]
        b = self._buffers(n0, P0)
        if b["slots"][1] is None:
            b["slots"][1] = self._slot(n0, P0)
        encoded, exchanged, freed, outs = [None, None], [None, None], [None, None], []
        side.wait_stream(main)
        xch.wait_stream(main)
        start, end, _ = self.frame_block(n0, self.rank, self.world)

        def encode(i):
            f, x = sequences[i]
            assert f.shape[0] == n0 and x.shape[1] == P0, "a pipelined stream expects sequences of one shape"
            slot = i & 1
            if freed[slot] is not None:
                side.wait_event(freed[slot])          # step i-2 is done with this slot
            with torch.cuda.stream(side):
                f_loc, x_loc = f[start:end], x[start:end]
                if host:                              # this rank's block only; stream-ordered ahead of its encoder
                    f_loc, x_loc = f_loc.cuda(non_blocking=True), x_loc.cuda(non_blocking=True)
                self._encode_block(f_loc, x_loc, n0, P0, b["slots"][slot]["desc"])
                ev = torch.cuda.Event()
                ev.record(side)
            encoded[slot] = ev

        def exchange(i):
            slot = i & 1
            if getattr(self, "pipeline_stages", 2) < 3:      # two stages: the exchange stays on the caller's stream
                main.wait_event(encoded[slot])
                self._exchange_block(b["slots"][slot], n0, P0)
                exchanged[slot] = None
                return
            xch.wait_event(encoded[slot])                    # three stages: the exchange has a stream of its own
            with torch.cuda.stream(xch):
                self._exchange_block(b["slots"][slot], n0, P0)
                ev = torch.cuda.Event()
                ev.record(xch)
            exchanged[slot] = ev

        encode(0)
        exchange(0)
        for i in range(len(sequences)):
            slot = i & 1
            if i + 1 < len(sequences):
                encode(i + 1)
            if exchanged[slot] is not None:
                main.wait_event(exchanged[slot])
            r = self._score_block(b, b["slots"][slot], n0, P0, k, exclude_band)
            ev = torch.cuda.Event()
            ev.record(main)
            freed[slot] = ev
            if i + 1 < len(sequences):
                exchange(i + 1)                       # issued after C_i: its all-gathers queue behind C_i's all-reduce
            if not host:
                outs.append(r["candidates"])
                continue
            per = tuple(r["candidates"][0].shape)
            pool = getattr(self, "_pinned_out", None)
            if pool is None or tuple(pool[0].shape[1:]) != per or pool[0].shape[0] < len(sequences):
                slots = max(64, len(sequences))
                pool = self._pinned_out = (torch.empty((slots,) + per, dtype=torch.float32).pin_memory(),
                                           torch.empty((slots,) + per, dtype=torch.int64).pin_memory())
            s_h, i_h = pool[0][i], pool[1][i]
            s_h.copy_(r["candidates"][0], non_blocking=True)
            i_h.copy_(r["candidates"][1], non_blocking=True)
            outs.append((s_h, i_h))
        if host:
            main.synchronize()
        return outs

    def host_bytes_per_step(self, frames_h, xy_h, k):
        n = frames_h.shape[0]
        start, end, _ = self.frame_block(n, self.rank, self.world)
        per_frame = frames_h[0].numel() * frames_h.element_size() + xy_h[0].numel() * xy_h.element_size()
        return int((end - start) * per_frame), int(n * min(k, max(n - 1, 1)) * 12)
