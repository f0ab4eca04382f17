"""Multi-GPU parity check of the sharded matcher (run under torchrun, one rank per GPU): the merged candidate lists
must equal the oracle's top-k over the WHOLE database and be identical on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200.matcher import ShardedKeyframeDatabase  # noqa: E402
from oracle import matcher as o_match  # noqa: E402

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ok = True
for metric in ("cos", "l2"):
    rng = np.random.default_rng(11)
    N, D, B, k = 4000 * world, 256, 70, 10
    rows = rng.standard_normal((N, D)).astype(np.float32) * (1.0 if metric == "cos" else 0.25)
    q = rng.standard_normal((B, D)).astype(np.float32) * (1.0 if metric == "cos" else 0.25)
    q[:20] = rows[rng.choice(N, 20, replace=False)] + 0.02 * rng.standard_normal((20, D)).astype(np.float32)
    shard = N // world
    db = ShardedKeyframeDatabase(D, shard, metric, "fp16")
    db.append_local(torch.from_numpy(rows[rank * shard:(rank + 1) * shard]).cuda())
    s, i = db.topk(torch.from_numpy(q).cuda(), k)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    t = torch.from_numpy(rows)
    if metric == "cos":
        t = t / t.norm(dim=1, keepdim=True)
        qq = torch.from_numpy(q)
        qq = (qq / qq.norm(dim=1, keepdim=True)).half().double().numpy()
        ref = qq @ t.half().double().numpy().T
        rs, ri = o_match.topk(ref, k)
    else:
        qq = torch.from_numpy(q).half().double().numpy()
        ref = o_match.score_matrix(qq, t.half().double().numpy(), o_match.L2)
        rs, ri = o_match.topk(ref, k, largest=False)
    same = np.array_equal(i, ri)
    mism = int((i != ri).sum())
    # remaining differences must be ties inside the tolerance
    tie_ok = True
    for r, c in zip(*np.nonzero(i != ri)):
        tie_ok &= abs(ref[r, i[r, c]] - rs[r, c]) <= 1e-3 * max(1.0, abs(rs[r, c])) + 1e-4 * 40
    gathered = [None] * world
    dist.all_gather_object(gathered, i.tobytes())
    identical = all(g == gathered[0] for g in gathered)
    if rank == 0:
        print("sharded %s world=%d: lists equal oracle: %s (%d index differences, ties ok: %s), identical on all ranks: %s, "
              "max score err %.2e" % (metric, world, same, mism, tie_ok, identical, np.max(np.abs(s - rs))), flush=True)
    ok &= tie_ok and identical
dist.destroy_process_group()
sys.exit(0 if ok else 1)
