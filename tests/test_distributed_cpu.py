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
