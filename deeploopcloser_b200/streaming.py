"""Streaming loop closing (BASELINE.json config 5): batches of frames -> SDA descriptors -> one L2-normalised
descriptor per frame -> top-k against a keyframe database that grows with every batch and is row-sharded over the
GPUs of one box.

Per batch and rank (one process per GPU):
  1. this rank's slice of the batch: patch gather + SDA encode            (data parallel, no collective)
  2. frame descriptor = mean of the 30 patch descriptors                   (new definition)
  3. all_gather of the frame descriptors -> every rank holds all queries   (B x D x 4 bytes, NCCL)
  4. fused similarity + top-k on the local shard                           (tcgen05 kernel)
  5. all_gather of the [B, k] partial lists + deterministic merge          (B x k x 12 bytes per rank, NCCL)
  6. append this rank's slice to its shard (global row = rank * capacity + local row)
The reference has no database, top-k or incremental insertion; this module is north-star capability built from the
same kernels as the parity-checked path."""
import torch

from . import ops
from .matcher import ShardedKeyframeDatabase
from .pipeline import LoopClosurePipeline


class StreamingLoopCloser:
    def __init__(self, capacity_per_rank, dims=(1681, 2500, 2500, 2500, 2500, 2500), precision="fp16x2",
                 metric="cos", db_dtype="fp16", k=10, rows_per_frame=30, group=None):
        self.pipe = LoopClosurePipeline(dims, precision=precision)
        self.k = k
        self.rows_per_frame = rows_per_frame
        self.db = ShardedKeyframeDatabase(dims[-1], capacity_per_rank, metric, db_dtype, group)
        self.group = group

    def set_weights(self, weights, biases):
        self.pipe.set_weights(weights, biases)

    @property
    def rank(self):
        return self.db.rank

    @property
    def world(self):
        return self.db.world

    def frame_descriptors(self, frames, xy=None):
        desc = self.pipe.encode(frames, xy)                       # [b*P, D]; xy None: keypoints detected on the device
        return ops.mean_pool_rows(desc, self.rows_per_frame)      # [b, D]

    def step(self, frames_local, xy_local=None):
        """frames_local uint8 [b, H, W], xy_local float32 [b, P, 2] (None: the fast-Hessian detector finds the 30
        keypoints of every frame on the device): this rank's slice of the batch (same b on every rank). Returns (scores [B, k], global indices [B, k]) for the whole batch, identical on every rank, matched
        against the database as it was BEFORE this batch is inserted."""
        q_local = self.frame_descriptors(frames_local, xy_local)
        if self.world > 1:
            import torch.distributed as dist
            q = torch.empty((self.world * q_local.shape[0], q_local.shape[1]), dtype=q_local.dtype, device=q_local.device)
            dist.all_gather_into_tensor(q, q_local, group=self.group)
        else:
            q = q_local
        # fused similarity + top-k on the shard, one NCCL all-gather of the packed lists, merge (one C-ABI call)
        s, i = self.db.topk(q, self.k)
        self.db.append_local(q_local)
        return s, i
