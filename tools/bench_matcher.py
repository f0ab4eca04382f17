"""Global matcher benchmark (BASELINE.json config 4: synthetic N x 4096 database, batched top-10 queries, database
row-sharded over the ranks). Run single-process for 1 GPU or under torchrun for 2/4/8 GPUs:

    python tools/bench_matcher.py --rows 1000000 --dim 4096 --batches 32,256,1024
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_matcher.py ...

Prints one JSON line per query batch size: queries/s, algorithmic TFLOP/s (2*B*D*N) and GB/s (N*D*2 bytes, the
database streamed once), both against the measured peaks. Top-1 of the planted near-duplicates is checked."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200.matcher import ShardedKeyframeDatabase  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1000000, help="total database rows over all ranks")
    ap.add_argument("--dim", type=int, default=4096)
    ap.add_argument("--batches", default="32,256,1024")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--metric", default="cos")
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shard = args.rows // world
    db = ShardedKeyframeDatabase(args.dim, shard, args.metric, args.dtype)
    g = torch.Generator(device="cuda")
    g.manual_seed(5 + rank)
    chunk = 65536
    keep = None
    for s in range(0, shard, chunk):
        n = min(chunk, shard - s)
        rows = torch.randn((n, args.dim), device="cuda", generator=g)
        if s == 0:
            keep = rows[:4096].clone()      # planted-neighbour sources live in every rank's first rows
        db.append_local(rows)
    peaks = {"tflops": 1407.6, "gbs": 6537.6}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        p = json.load(open(pk))
        peaks = {"tflops": p["bf16_tflops_sustained"], "gbs": p["hbm_gbs"]}
    for B in [int(b) for b in args.batches.split(",")]:
        gq = torch.Generator(device="cuda")
        gq.manual_seed(6)
        q = torch.randn((B, args.dim), device="cuda", generator=gq)
        nplant = max(B // 10, 1)
        # planted near-duplicates of rank 0's first rows (every rank builds the same queries)
        src = keep[:nplant]
        if world > 1:
            src = src.clone()
            dist.broadcast(src, 0)
        q[:nplant] = src + 0.05 * torch.randn((nplant, args.dim), device="cuda", generator=gq)
        for _ in range(2):
            s, i = db.topk(q, args.k)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            s, i = db.topk(q, args.k)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        ok = bool((i[:nplant, 0].cpu() == torch.arange(nplant)).all())
        if rank == 0:
            flop = 2.0 * B * args.dim * args.rows
            byts = float(args.rows) * args.dim * 2
            print(json.dumps({"bench": "matcher", "n_gpus": world, "rows": args.rows, "dim": args.dim, "B": B,
                              "k": args.k, "metric": args.metric, "dtype": args.dtype, "ms": ms,
                              "queries_per_s": B / ms * 1e3, "tflops": flop / ms / 1e9,
                              "tflops_frac_of_measured_sustained": flop / ms / 1e9 / peaks["tflops"] / world,
                              "db_gbs": byts / ms / 1e6, "hbm_frac_of_measured": byts / ms / 1e6 / peaks["gbs"] / world,
                              "planted_top1_ok": ok}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
