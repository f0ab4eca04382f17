"""CPU, world_size 2 over gloo: the exchange step of the sharded matcher (all-gather of per-rank partial top-k lists
and their layout) plus the global-index bookkeeping. The CUDA merge kernel itself is covered by `-m gpu` tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import matcher as o_match


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deeploopcloser_b200.matcher import gather_partial_lists
        rng = np.random.default_rng(7)           # same stream on both ranks: a replicated query batch
        B, k, shard = 5, 4, 50
        full = rng.standard_normal((B, world * shard))
        local = full[:, rank * shard:(rank + 1) * shard]
        ls, li = o_match.topk(local, k)          # stands in for the fused GPU kernel on this rank's shard
        li = li + rank * shard                   # global index = row_offset + local row
        cs, ci = gather_partial_lists(torch.from_numpy(ls), torch.from_numpy(li))
        assert cs.shape == (B, world * k)
        # rank-major layout: columns [r*k, (r+1)*k) hold rank r's list
        assert torch.equal(cs[:, rank * k:(rank + 1) * k], torch.from_numpy(ls))
        ms, mi = o_match.topk(cs.numpy(), k)     # deterministic merge (the CUDA kernel implements the same order)
        merged_idx = np.take_along_axis(ci.numpy(), mi, axis=1)
        want_s, want_i = o_match.topk(full, k)
        assert np.array_equal(merged_idx, want_i) and np.allclose(ms, want_s)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_exchange_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def _seq_worker(rank, world, port, q):
    """ShardedSequencePipeline.run_block with the device stages replaced by CPU stand-ins: the host logic under test is
    the frame blocks (padded last block), the in-place all-gathers of descriptors / column sums / planes / stats
    blocks, the background gather on the second communicator being complete before the second pass, the parts-sum
    all-reduce, and that every rank ends with the same matrix."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deeploopcloser_b200 import ops
        from deeploopcloser_b200.pipeline import ShardedSequencePipeline
        from oracle import similarity as o_sim
        n, P, D = 7, 3, 5                                  # 7 frames over 2 ranks: blocks of 4 and 3 (+1 pad)
        rng = np.random.default_rng(3)
        table = rng.uniform(0, 1, (n, P, D)).astype(np.float32)   # "descriptor" of frame f = table[f]
        seen = {}

        class Pipe(ShardedSequencePipeline):
            def _alloc(self, shape, dtype):
                return torch.zeros(shape, dtype=dtype)

            def _encode_into(self, frames, xy, out):
                out.copy_(torch.from_numpy(table[frames.numpy()[:, 0, 0]].reshape(-1, D)))

            def _stage_colsum(self, desc_local, out):
                out[:D].copy_(desc_local.double().sum(0))
                out[D:].copy_((desc_local.double() ** 2).sum(0))

            def _stage_weights(self, colsums, rows_total, w, mean):
                m = colsums[:, :D].sum(0) / rows_total
                mean.copy_(m.float())
                w.copy_(torch.exp(-(m - 0.5) ** 2 / (2 * 0.2 ** 2)))

            def _stage_prepare(self, desc_local, n_local, per, P_, w, mean, plane_local, plane_lo_local, stats_local):
                plane_local.zero_()
                plane_local[:n_local * P_, :D] = (desc_local[:n_local * P_] - mean).half()
                stats_local.fill_(self.rank + 1)

            def _stage_gram(self, b, per, n_, P_):
                # every rank sees every rank's centred planes and stats block
                want = torch.from_numpy(table.reshape(-1, D)).double() - b["mean"].double()
                assert torch.allclose(b["plane"][:n_ * P_, :D].double(), want, atol=2e-3)
                assert [int(b["stats"][r, 0]) for r in range(self.world)] == [r + 1 for r in range(self.world)]
                w_ref = o_sim.distinctive_weights(table.astype(np.float64))
                assert np.allclose(b["w"].numpy(), w_ref, rtol=1e-12)
                S = o_sim.similarity_matrix(table.astype(np.float64))
                own = np.zeros_like(S)
                for i, j in zip(*np.triu_indices(n_, 1)):
                    if i % self.world == self.rank:
                        own[i, j] = own[j, i] = S[i, j]
                if self.rank == 0:
                    np.fill_diagonal(own, -1.0)
                b["S"].copy_(torch.from_numpy(own.astype(np.float32)))

            def _stage_fix(self, b, n_, P_):
                # the background all-gather of the float32 descriptors has completed when the second pass starts
                seen["desc_complete"] = bool(np.array_equal(b["desc"][:n_ * P_].numpy(), table.reshape(-1, D)))

        pipe = Pipe.__new__(Pipe)
        pipe.dims, pipe.dist, pipe.group, pipe.rank, pipe.world = [1, D], dist, None, rank, world
        pipe.sim_precision, pipe.sim_args, pipe._bg_group, pipe._buf_key = "fp16r", {}, None, None
        ops.topk_rows = lambda S, k, largest=True, exclude_band=-1: torch.topk(S, k, dim=1)
        frames = torch.arange(n).reshape(n, 1, 1).repeat(1, 2, 2)          # frame id stored in its pixels
        xy = torch.zeros((n, P, 2))
        for _ in range(2):                                                   # second call reuses the buffers
            res = pipe.run(frames, xy, k=2)
        want = o_sim.similarity_matrix(table.astype(np.float64))
        assert np.allclose(res["similarity"].numpy(), want, rtol=1e-5, atol=1e-5)
        assert np.array_equal(res["descriptors"].numpy(), table.reshape(-1, D))
        assert seen["desc_complete"]
        start, end, per = Pipe.frame_block(n, rank, world)
        assert (start, end, per) == ((0, 4, 4) if rank == 0 else (4, 7, 4))
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()[-900:] or repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_sequence_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_seq_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
