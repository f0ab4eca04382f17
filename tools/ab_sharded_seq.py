"""A/B of the sequence split over the ranks, all variants in ONE process group (run under torchrun): SMs left free for
the NCCL kernels (dlc_set_sm_reserve), the encoder's per-call GEMM width (dlc_debug_set 11) and the pipeline depth.
Prints one JSON line per variant (max over ranks of the device time per sequence, pipelined run_many).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/ab_sharded_seq.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deeploopcloser_b200 import _lib  # noqa: E402
from deeploopcloser_b200.pipeline import ShardedSequencePipeline  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
frames, xy = bench.synthetic_inputs(100)
f_d, x_d = torch.from_numpy(frames).cuda(), torch.from_numpy(xy).cuda()
STEPS = 20


def timed(fn):
    for _ in range(2):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / STEPS], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


for weights in ("normal", "xavier"):
    ws, bs = bench.reference_weights() if weights == "normal" else bench.xavier_weights()
    pipe = ShardedSequencePipeline(bench.DIMS, precision="auto")
    pipe.set_weights(ws, bs)
    pipe.sm_reserve = None                      # this script sets the reserve itself
    want = None
    for reserve, wave_pad, stages in ((0, 1, 2), (8, 1, 2), (16, 1, 2), (24, 1, 2), (8, 0, 2), (0, 0, 2), (8, 1, 3), (4, 1, 2)):
        _lib.call("dlc_set_sm_reserve", reserve)
        _lib.call("dlc_debug_set", 11, wave_pad)
        pipe.pipeline_stages = stages
        got = pipe.run_many([(f_d, x_d)] * 2, k=bench.K_CAND)[-1]
        torch.cuda.synchronize()
        same = True if want is None else bool(torch.equal(got[1], want[1]) and torch.equal(got[0], want[0]))
        want = want or (got[0].clone(), got[1].clone())
        ms = timed(lambda: pipe.run_many([(f_d, x_d)] * STEPS, k=bench.K_CAND))
        if rank == 0:
            print(json.dumps({"ab": "sharded_sequence", "n_gpus": world, "weights": weights, "sm_reserve": reserve,
                              "wave_pad": wave_pad, "pipeline_stages": stages, "ms_per_sequence": round(ms, 4),
                              "frames_per_s": round(bench.N_FRAMES / ms * 1e3), "same_lists_as_first_variant": same}),
                  flush=True)
    _lib.call("dlc_set_sm_reserve", 0)
    _lib.call("dlc_debug_set", 11, 1)
dist.destroy_process_group()
