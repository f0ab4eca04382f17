"""Multi-GPU check + timing of ONE sequence split over the ranks (run under torchrun, one rank per GPU): the score
matrix and candidate lists of ShardedSequencePipeline against the single-GPU pipeline on every rank: descriptors
bit-identical, candidate indices identical, scores within float32 rounding (the dataset mean is summed in another order
across ranks, so the last bits of a score may differ; whether they are bit-identical is reported).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_sequence.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deeploopcloser_b200.pipeline import LoopClosurePipeline, ShardedSequencePipeline  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
frames, xy = bench.synthetic_inputs(100)
ws, bs = bench.reference_weights()
f_d, x_d = torch.from_numpy(frames).cuda(), torch.from_numpy(xy).cuda()
ref = LoopClosurePipeline(bench.DIMS)
ref.set_weights(ws, bs)
want = ref.run(f_d, x_d, k=bench.K_CAND)
pipe = ShardedSequencePipeline(bench.DIMS)
pipe.set_weights(ws, bs)
got = pipe.run(f_d, x_d, k=bench.K_CAND)
torch.cuda.synchronize()
rel = ((got["similarity"] - want["similarity"]).abs() / want["similarity"].abs().clamp_min(1.0)).max()
same = torch.tensor([int(bool(rel <= 1e-5) and torch.equal(got["candidates"][1], want["candidates"][1]) and
                         torch.equal(got["descriptors"], want["descriptors"])),
                     int(torch.equal(got["similarity"], want["similarity"]))], device="cuda")
dist.all_reduce(same, op=dist.ReduceOp.MIN)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


t_one = timed(lambda: ref.run(f_d, x_d, k=bench.K_CAND))
t_all = timed(lambda: pipe.run(f_d, x_d, k=bench.K_CAND))
many_ok, t_many = True, {}
for stages in (2, 3):                      # both depths of the step pipeline give run()'s lists
    pipe.pipeline_stages = stages
    many = pipe.run_many([(f_d, x_d)] * 5, k=bench.K_CAND)
    torch.cuda.synchronize()
    many_ok = many_ok and all(torch.equal(m[1], want["candidates"][1]) and torch.equal(m[0], want["candidates"][0])
                              for m in many)
    t_many[stages] = timed(lambda: pipe.run_many([(f_d, x_d)] * 10, k=bench.K_CAND), reps=2) / 10
ok_t = torch.tensor([int(many_ok)], device="cuda")
dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
many_ok = bool(ok_t.item())
if rank == 0:
    print(json.dumps({"check": "sharded_sequence", "n_gpus": world, "frames": bench.N_FRAMES,
                      "matches_single_gpu_on_all_ranks": bool(same[0].item()),
                      "scores_bit_identical_on_all_ranks": bool(same[1].item()), "max_rel_score_diff": float(rel),
                      "pipelined_run_many_matches_on_all_ranks": many_ok, "pipelined_ms_per_sequence_2_stages": t_many[2],
                      "pipelined_ms_per_sequence_3_stages": t_many[3],
                      "single_gpu_ms": t_one, "sharded_ms": t_all, "speedup": t_one / t_all,
                      "frames_per_s": bench.N_FRAMES / t_all * 1e3}))
dist.destroy_process_group()
