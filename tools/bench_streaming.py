"""Streaming SLAM benchmark (BASELINE.json config 5 shape): 640x480 frames at batch 256 through SDA encode +
incremental match against a keyframe database sharded over the ranks. Run single-process or under torchrun.

    python tools/bench_streaming.py --db-rows 1000000 --batch 256 --steps 5"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeploopcloser_b200.streaming import StreamingLoopCloser  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--db-rows", type=int, default=1000000, help="database rows over all ranks before streaming starts")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--detect", action="store_true", help="find the 30 keypoints per frame on the device (fast-Hessian "
                    "detector) instead of feeding seeded keypoints; frames are then a smooth synthetic texture")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = [1681, 2500, 2500, 2500, 2500, 2500]
    shard = args.db_rows // world
    b_local = args.batch // world
    cap = shard + (args.steps + args.warmup + 1) * b_local
    sl = StreamingLoopCloser(cap, dims)
    rng = np.random.default_rng(1)
    sl.set_weights([rng.standard_normal((k, n)) for k, n in zip(dims[:-1], dims[1:])], [np.zeros(n) for n in dims[1:]])
    g = torch.Generator(device="cuda")
    g.manual_seed(5 + rank)
    for s in range(0, shard, 65536):
        sl.db.append_local(torch.rand((min(65536, shard - s), dims[-1]), device="cuda", generator=g))
    frames = torch.randint(0, 256, (b_local, args.height, args.width), dtype=torch.uint8, device="cuda", generator=g)
    xy = torch.stack([torch.rand((b_local, 30), device="cuda", generator=g) * args.width,
                      torch.rand((b_local, 30), device="cuda", generator=g) * args.height], -1).contiguous()
    if args.detect:
        base = torch.rand((b_local, 1, args.height // 8 + 2, args.width // 8 + 2), device="cuda", generator=g) * 255
        frames = torch.nn.functional.interpolate(base, size=(args.height, args.width), mode="bicubic")[:, 0]
        frames = frames.clamp(0, 255).round().to(torch.uint8).contiguous()
        xy = None
    for _ in range(args.warmup):
        s, i = sl.step(frames, xy)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        s, i = sl.step(frames, xy)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # the frames of the previous step were inserted: each query's best match must be its own earlier copy
    own = int(((i[:, 0] - sl.db.row_offset) >= shard).sum()) if world == 1 else None
    if rank == 0:
        print(json.dumps({"bench": "streaming", "n_gpus": world, "db_rows": args.db_rows, "batch": args.batch,
                          "frame": [args.height, args.width], "keypoints": "detected on the device" if args.detect else "seeded",
                          "ms_per_batch": float(ms.item()),
                          "frames_per_s": args.batch / float(ms.item()) * 1e3,
                          "top1_is_previously_inserted_copy": own, "db_rows_after": len(sl.db.local) * world}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
