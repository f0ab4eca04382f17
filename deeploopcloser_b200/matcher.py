"""Keyframe database + global matcher (cosine / dot / squared-L2 top-k and threshold selection), single GPU and
row-sharded over the GPUs of one box. New capability asked for by the north star - the reference only builds dense
N x N matrices (src/sdav/create_similarity_matrix.py:31-38, src/cnn_vtl/create_distance_matrix.py:31-36).

Sharding (one process per GPU, torch.distributed): database rows are dealt to ranks; a query batch is replicated,
every rank runs the fused similarity+top-k kernel on its shard, the [B, k] partial lists (score, global index) are
all-gathered (B*k*12 bytes per rank - latency bound) and merged on every rank by the deterministic merge kernel
(best score first, ties -> lowest global index), so all ranks end with identical candidate lists."""
import ctypes as C

import torch

from . import _lib
from ._cuda import Workspace, ptr, stream_ptr
from .ops import METRICS, _DT, topk_rows

_STORE = {"fp16": _lib.F16, "f16": _lib.F16, "bf16": _lib.BF16}


class KeyframeDatabase:
    def __init__(self, dim, capacity, metric="cos", dtype="fp16"):
        self.dim = int(dim)
        self.capacity = int(capacity)
        self.metric = metric
        self.dtype = dtype
        self._h = C.c_void_p()
        _lib.call("dlc_db_create", C.byref(self._h), self.dim, self.capacity, METRICS[metric], _STORE[dtype])
        self._ws = Workspace()

    def __len__(self):
        return int(_lib.call("dlc_db_size", self._h))

    @property
    def smaller_is_better(self):
        return METRICS[self.metric] == _lib.METRIC_L2

    def clear(self):
        _lib.call("dlc_db_clear", self._h)

    def append(self, rows):
        """rows: CUDA tensor [n, dim] float32 / float16 / bfloat16."""
        if not rows.is_cuda or not rows.is_contiguous() or rows.shape[1] != self.dim:
            raise ValueError("expected a contiguous CUDA tensor [n, %d]" % self.dim)
        _lib.call("dlc_db_append", self._h, ptr(rows), _DT[rows.dtype], rows.shape[0], stream_ptr())

    def _prep(self, q, k):
        if q.dtype != torch.float32 or not q.is_cuda or not q.is_contiguous() or q.shape[1] != self.dim:
            raise ValueError("queries must be a contiguous float32 CUDA tensor [B, %d]" % self.dim)
        B = q.shape[0]
        ws, ws_bytes = self._ws.get(_lib.call("dlc_match_workspace_bytes", self._h, B, k))
        scores = torch.empty((B, k), dtype=torch.float32, device=q.device)
        idx = torch.empty((B, k), dtype=torch.int64, device=q.device)
        return B, ws, ws_bytes, scores, idx

    def topk(self, q, k=10, idx_offset=0):
        B, ws, ws_bytes, scores, idx = self._prep(q, k)
        _lib.call("dlc_match_topk", self._h, ptr(q), B, k, int(idx_offset), ptr(scores), ptr(idx), ws, ws_bytes,
                  stream_ptr())
        return scores, idx

    def threshold(self, q, thr, max_per_row=32, idx_offset=0):
        B, ws, ws_bytes, scores, idx = self._prep(q, max_per_row)
        counts = torch.empty(B, dtype=torch.int32, device=q.device)
        _lib.call("dlc_match_threshold", self._h, ptr(q), B, float(thr), max_per_row, int(idx_offset), ptr(counts),
                  ptr(scores), ptr(idx), ws, ws_bytes, stream_ptr())
        return counts, scores, idx

    def close(self):
        if self._h:
            _lib.call("dlc_db_destroy", self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge_partial_lists(scores, idx, k, smaller_is_better=False):
    """scores/idx [B, L] candidates (idx = -1 marks padding) -> deterministic top-k (score, lowest index)."""
    return topk_rows(scores.contiguous(), k, largest=not smaller_is_better, cand_idx=idx.contiguous())


def gather_partial_lists(scores, idx, group=None):
    """All-gather per-rank partial top-k lists [B, k] -> candidate matrices [B, world * k] (rank-major columns),
    identical on every rank. Pure torch.distributed plumbing: NCCL on the GPUs, gloo in the CPU tests."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    B, k = scores.shape
    gs = torch.empty((world * B, k), dtype=scores.dtype, device=scores.device)  # rank-major concatenation
    gi = torch.empty((world * B, k), dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(gs, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(gi, idx.contiguous(), group=group)
    cs = gs.view(world, B, k).permute(1, 0, 2).reshape(B, world * k).contiguous()
    ci = gi.view(world, B, k).permute(1, 0, 2).reshape(B, world * k).contiguous()
    return cs, ci


class ShardedKeyframeDatabase:
    """Row-sharded database: rank r owns the global rows it was given (contiguous block `row_offset + local`).

    With the NCCL backend the whole query runs behind ONE C-ABI call (`dlc_match_topk_sharded`): fused kernel on the
    shard, one ncclAllGather of the packed (index, score) lists issued by the library on the caller's stream, merge
    kernel reading the gathered blocks in place. The library's communicator is created here from a unique id that rank
    0 draws and torch.distributed broadcasts. Without NCCL (the gloo CPU tests) the exchange is two torch all-gathers
    and the generic merge (`gather_partial_lists` / `merge_partial_lists`)."""

    def __init__(self, dim, capacity_per_rank, metric="cos", dtype="fp16", group=None, native_comm=True):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local = KeyframeDatabase(dim, capacity_per_rank, metric, dtype)
        self.row_offset = self.rank * int(capacity_per_rank)
        self._comm = None
        if native_comm and self.world > 1 and group is None and dist.get_backend() == "nccl":
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                buf = (C.c_char * 128)()
                _lib.call("dlc_comm_unique_id", C.cast(buf, C.c_void_p))
                uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            raw = bytes(uid.cpu().numpy().tobytes())
            self._comm = C.c_void_p()
            _lib.call("dlc_comm_create", C.byref(self._comm), raw, self.rank, self.world)
        self._ws = Workspace()

    def append_local(self, rows):
        self.local.append(rows)

    def topk(self, q, k=10):
        """q must be identical on every rank (broadcast it first if it is produced on one rank)."""
        if self._comm is not None:
            B, _, _, s, i = self.local._prep(q, k)
            ws, ws_bytes = self._ws.get(_lib.call("dlc_match_sharded_workspace_bytes", self.local._h, B, k, self.world))
            _lib.call("dlc_match_topk_sharded", self.local._h, self._comm, ptr(q), B, k, int(self.row_offset), ptr(s),
                      ptr(i), ws, ws_bytes, stream_ptr())
            return s, i
        s, i = self.local.topk(q, k, idx_offset=self.row_offset)
        if self.world == 1:
            return s, i
        cs, ci = gather_partial_lists(s, i, self.group)
        return merge_partial_lists(cs, ci, k, self.local.smaller_is_better)

    def close(self):
        if self._comm is not None:
            _lib.call("dlc_comm_destroy", self._comm)
            self._comm = None
        self.local.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
