#!/usr/bin/env python
"""Headline benchmark: loop-query frames/sec (encode + match) on BASELINE.json's config[1] workload:
SDA descriptors + loop detection over an outdoor_kennedylong-shaped sequence (1063 frames, 240x192, 30 keypoints per
frame) on B200. One step = one pass of the whole hot path over the sequence:

    patch gather -> 5 x (GEMM + bias + sigmoid) -> SDAV score matrix (Gram + argmin + score, i<j) -> per-row top-10

`value`   : frames/s with the frames and keypoints already resident in HBM (CUDA events, max over ranks).
`e2e`     : the same through the public pipeline with HOST (pinned) frames/keypoints copied in and the candidate lists
            copied out inside the timed region.
`roofline`: the kernel with the largest share of the step (the encoder layer kernel), algorithmic FLOPs per launch over
            its average launch time; `roofline_sim` is the score-matrix kernel alone, with its second pass, and with
            the preparation kernels.
`arms`    : the same step with trained-like (Xavier-scaled) weights next to the reference's N(0,1) initialisation
            (SDAV.py:189-217 vs :232-240): the encoder's precision probe then needs one tensor product, not three.
`other_configs` (N = 1): short runs of BASELINE configs 3, 4 and 5 (`--config 3|4|5` runs one of them as the headline).
N > 1     : STRONG scaling - the ONE sequence is split over the ranks (frames dealt in blocks, descriptors exchanged
            with NCCL all-gather, every rank evaluates its tile rows of the score matrix, scores all-reduced); plus
            BASELINE config 4 (1 M x 4096 database row-sharded 1 M / N per rank, NCCL merge of the top-k lists).
--impl reference : the reference's CPU path, timed on the host cores on a fixed sub-workload run to completion
            (96 frames encoded + all 4560 frame pairs scored, dataset mean hoisted and un-hoisted). The reference needs
            TensorFlow 1.x (not installable here) so this runs the float64 oracle port (oracle/), the only place outside
            tests/smoke where oracle/ is executed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FRAMES, H, W, P, PATCH = 1063, 192, 240, 30, 41
DIMS = [1681, 2500, 2500, 2500, 2500, 2500]
K_CAND = 10
METRIC = "loop-query frames/sec (encode+match)"
WORKLOAD = "SDA descriptors + loop detection, outdoor_kennedylong-shaped sequence (1063 frames 240x192, 30 kp/frame)"
ENC_FLOP_PER_FRAME = 30 * 2 * (1681 * 2500 + 4 * 2500 * 2500)          # 1.75215 GFLOP
GRAM_FLOP = 2.0 * 30 * 30 * 2500 * (N_FRAMES * (N_FRAMES - 1) / 2)       # i<j pairs, 2*30*30*2500 each = 2.54 TFLOP
WEIGHT_NOTES = {"normal": "N(0,1) init (reference default, no checkpoint shipped)",
                "xavier": "Xavier-scaled N(0, 1/fan_in), biases 0.1 N(0,1) (trained-like; no trained checkpoint exists)"}


def workload_config(weights, world):
    """`config` of the JSON line: names the workload only, and is THE SAME dict in both arms (`--impl ours` and
    `--impl reference`), so the driver can see that they measured the same thing. How our arm computes it (arithmetic
    modes, input layout) is in `config_detail`."""
    return {"workload": WORKLOAD, "weights": WEIGHT_NOTES[weights], "k": K_CAND,
            "l2": "per-step working set (~1.4 GB of operand planes and descriptors) exceeds the 126 MB L2; no explicit "
                  "flush between steps",
            "multi_gpu": "single GPU" if world == 1 else
                         "strong scaling: ONE sequence split over %d ranks, frames dealt in blocks; NCCL all-gather of "
                         "the fp16 descriptor planes, the row statistics and (in the background) the fp32 descriptors, "
                         "every rank evaluates its interleaved tile rows of the score matrix, NCCL all-reduce (sum) of "
                         "the scores" % world}


def synthetic_inputs(seed):
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (N_FRAMES, H, W), dtype=np.uint8)
    xy = np.stack([rng.uniform(0, W, (N_FRAMES, P)), rng.uniform(0, H, (N_FRAMES, P))], -1).astype(np.float32)
    return frames, xy


def reference_weights():
    """The reference's initialisation when no checkpoint exists: N(0,1) weights, zero biases (SDAV.py:189-217)."""
    rng = np.random.default_rng(1)
    ws = [rng.standard_normal((k, n)) for k, n in zip(DIMS[:-1], DIMS[1:])]
    bs = [np.zeros(n) for n in DIMS[1:]]
    return ws, bs


def xavier_weights():
    """Trained-like magnitudes (what restoring a checkpoint, SDAV.py:232-240, would give): sigma = 1/sqrt(fan_in)."""
    rng = np.random.default_rng(2)
    ws = [rng.standard_normal((k, n)) / np.sqrt(k) for k, n in zip(DIMS[:-1], DIMS[1:])]
    bs = [0.1 * rng.standard_normal(n) for n in DIMS[1:]]
    return ws, bs


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops_sustained": p["bf16_tflops_sustained"], "tflops_burst": p["bf16_tflops"],
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def kernel_traffic(name, arm):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of one kernel, from the round's `ncu --set full`
    capture of this weights arm (profiles/r2_traffic.json, written by tools/ncu_traffic.py from the capture's raw page);
    None when not captured."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)
    e = t.get(name + "@" + arm) or t.get(name)
    return None if not e else float(e["dram_bytes_read"] + e["dram_bytes_write"])


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML in-process, ~every 5 ms;
    falls back to polling nvidia-smi when pynvml is unavailable)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.max_mhz = None
        self.how = None

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; map through the UUID-free way: NVML index = visible list entry
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                idx = int(ids[self.index])
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        self.how = "nvml"
        while not self.stop_flag.is_set():
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)
            self.stop_flag.wait(0.005)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.how = "nvidia-smi"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:  # noqa: BLE001
            self._run_smi()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "sm_mhz_min": float(np.min(self.samples)) if self.samples else None,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "how": self.how}


# ---------------------------------------------------------------------------------------------- CPU reference arm
_PAIR_STATE = {}


def _score_pair_rows(args):
    """Worker: score frame i against every j > i of the sub-sequence; hoisted = dataset weights computed once,
    un-hoisted = recomputed for every pair like the literal reference (SimilarityCalculator.py:13-14)."""
    from oracle import similarity as o_sim
    i, hoisted = args
    desc, w = _PAIR_STATE["desc"], _PAIR_STATE["w"]
    with np.errstate(all="ignore"):
        for j in range(i + 1, len(desc)):
            o_sim.similarity_score(desc[i], desc[j], w if hoisted else o_sim.distinctive_weights(desc))
    return len(desc) - i - 1


def cpu_step_measured(n_frames=96, seed=0, unhoisted=True):
    """One step of the reference CPU path (float64 oracle port) on a FIXED sub-workload run to completion on all host
    cores: the first n_frames frames encoded (OpenBLAS threads) and ALL n_frames (n_frames - 1) / 2 frame pairs scored
    (one process per core), with the dataset mean hoisted out of the pair loop and - the literal reference - recomputed
    per pair. Returns measured times and the extrapolation to the 1063-frame step. Must run in a process that has not
    initialised CUDA (it forks)."""
    import multiprocessing as mp

    from oracle import patches as o_patch
    from oracle import sda as o_sda
    from oracle import similarity as o_sim
    cores = os.cpu_count() or 1
    frames, xy = synthetic_inputs(seed)
    ws, bs = reference_weights()
    t0 = time.perf_counter()
    x = np.concatenate([o_patch.extract_patches(frames[i], xy[i], PATCH) for i in range(n_frames)])
    desc = o_sda.sda_forward(x, ws, bs).reshape(n_frames, P, -1)
    t_enc = time.perf_counter() - t0
    _PAIR_STATE["desc"] = desc
    _PAIR_STATE["w"] = o_sim.distinctive_weights(desc)
    n_pairs = n_frames * (n_frames - 1) // 2
    times = {}
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_score_pair_rows, [(n_frames - 2, True)] * cores)        # untimed: every worker imports and runs once
        for label, hoisted in (("hoisted", True),) + ((("unhoisted", False),) if unhoisted else ()):
            t0 = time.perf_counter()
            done = sum(pool.imap_unordered(_score_pair_rows, [(i, hoisted) for i in range(n_frames - 1)], chunksize=2))
            times[label] = time.perf_counter() - t0
            assert done == n_pairs
    full_pairs = N_FRAMES * (N_FRAMES - 1) // 2
    scale = N_FRAMES / n_frames
    t_full_h = t_enc * scale + times["hoisted"] * full_pairs / n_pairs
    out = {"frames": n_frames, "pairs": n_pairs, "cores": cores, "encode_s": t_enc, "pairs_hoisted_s": times["hoisted"],
           "measured_step_s": t_enc + times["hoisted"], "extrapolated_full_step_s": t_full_h}
    if unhoisted:
        # per pair the literal path adds one dataset mean, whose cost grows with the dataset (x scale)
        t_mean = max(times["unhoisted"] - times["hoisted"], 0.0) / n_pairs
        out["pairs_unhoisted_s"] = times["unhoisted"]
        out["extrapolated_full_step_unhoisted_s"] = t_full_h + full_pairs * t_mean * scale
    return out


def reference_sample_text(m):
    s = ("oracle port, float64, %d cores, sub-workload run to completion: %d frames encoded in %.2f s (OpenBLAS) + all "
         "%d frame pairs scored in %.2f s (dataset mean hoisted, one process per core) = %.2f s measured; extrapolated "
         "to the 1063-frame step (x%.2f frames, x%.1f pairs): %.1f s" % (
             m["cores"], m["frames"], m["encode_s"], m["pairs"], m["pairs_hoisted_s"], m["measured_step_s"],
             N_FRAMES / m["frames"], (N_FRAMES * (N_FRAMES - 1) // 2) / m["pairs"], m["extrapolated_full_step_s"]))
    if "pairs_unhoisted_s" in m:
        s += ("; literal reference (mean recomputed per pair, SimilarityCalculator.py:13): %.2f s for the same pairs, "
              "%.0f s extrapolated" % (m["pairs_unhoisted_s"], m["extrapolated_full_step_unhoisted_s"]))
    return s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ms, last = [], None
    for it in range(args.warmup + args.steps):
        last = cpu_step_measured(args.cpu_frames, unhoisted=(it == args.warmup + args.steps - 1))
        if it >= args.warmup:
            ms.append(last)
    full_s = float(np.mean([m["extrapolated_full_step_s"] for m in ms]))
    meas_s = float(np.mean([m["measured_step_s"] for m in ms]))
    v = N_FRAMES / full_s
    line = {"metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * meas_s, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config("normal", int(os.environ.get("WORLD_SIZE", "1"))),
            "measured": {"what": "sub-workload of the step, run to completion every step: %d frames encoded + all %d "
                                 "pairs scored (mean hoisted)" % (last["frames"], last["pairs"]),
                         "ms_per_step": 1e3 * meas_s, "frames_per_s_of_the_sub_workload": last["frames"] / meas_s},
            "extrapolated": {"what": "the 1063-frame step: encode time x 1063/%d, pair time x 564453/%d; `value` = "
                                     "1063 / this" % (last["frames"], last["pairs"]),
                             "ms_per_full_step": 1e3 * full_s,
                             "ms_per_full_step_literal_unhoisted": 1e3 * last.get("extrapolated_full_step_unhoisted_s", 0.0)},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": last["cores"], "kind": "port",
                             "sample": reference_sample_text(last)},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- B200 path
class Ctx:
    """Process-wide state of one bench run: ranks, barrier, timed loop."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        from deeploopcloser_b200 import _cuda
        _cuda.require_cuda()
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return float(ms)
        t = self.torch.tensor([ms], device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup, sampler=None):
        """W untimed steps, barrier + synchronize, EXACTLY `steps` timed steps between CUDA events on the launch
        stream, barrier + synchronize, max over ranks -> ms per step."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        if sampler:
            sampler.start()          # clocks are sampled during the timed region only
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        if sampler:
            sampler.stop_flag.set()
            sampler.join()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps

    def timed_stream(self, fn_n, steps, warmup, sampler=None):
        """Like timed(), for a call fn_n(n) that pushes n steps through a streaming API (successive steps may overlap
        inside it): fn_n(warmup) untimed, then EXACTLY `steps` steps in one fn_n(steps) between the events."""
        torch = self.torch
        if warmup:
            fn_n(warmup)
        self.barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn_n(steps)
        e1.record()
        self.barrier()
        if sampler:
            sampler.stop_flag.set()
            sampler.join()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def make_pipeline(ctx, weights, args):
    from deeploopcloser_b200.pipeline import LoopClosurePipeline, ShardedSequencePipeline
    ws, bs = reference_weights() if weights == "normal" else xavier_weights()
    cls = ShardedSequencePipeline if ctx.world > 1 else LoopClosurePipeline
    pipe = cls(DIMS, precision=args.precision, sim_precision=args.sim_precision, raw_pixels=not args.split_pixel_input)
    pipe.set_weights(ws, bs)
    if args.pipeline_stages and ctx.world > 1:
        pipe.pipeline_stages = args.pipeline_stages
    return pipe


def step_launches(sim_precision, world):
    """OUR kernel launches of one step on one rank (NCCL's kernels are not counted): patch gather, one kernel per layer,
    the similarity call - column sums, their reduction, weights / centring, row preparation, Gram; with a precision
    probe also the duplicate-row mask, the probe, its finalisation, [auto: the gated residual planes], the second pass
    and [auto, one GPU] the gated three-product twin; a split sequence adds the stats unpack - and the per-row top-k.
    Checked against the ncu launch list (profiles/r2_launches_summary.txt)."""
    if world > 1:
        sim = {"auto": 10, "fp16r": 10}.get(sim_precision, 6)
    else:
        sim = {"auto": 11, "fp16r": 10}.get(sim_precision, 5)
    return 1 + len(DIMS) - 1 + sim + 1


def run_config2(ctx, args):
    torch = ctx.torch
    from deeploopcloser_b200 import _lib, ops
    world, rank = ctx.world, ctx.rank
    # strong scaling: every rank holds the same sequence (seed 100) and processes its share of it
    frames_h, xy_h = synthetic_inputs(100)
    frames_pin = torch.from_numpy(frames_h).pin_memory()
    xy_pin = torch.from_numpy(xy_h).pin_memory()
    frames_d, xy_d = frames_pin.cuda(), xy_pin.cuda()
    out_s_pin = torch.empty((N_FRAMES, K_CAND), dtype=torch.float32).pin_memory()
    out_i_pin = torch.empty((N_FRAMES, K_CAND), dtype=torch.int64).pin_memory()
    pk = peaks()

    def measure_arm(weights, sampler=None, detail=False):
        pipe = make_pipeline(ctx, weights, args)

        def steps_device(n):
            """n steps through the device-resident streaming call (one GPU: run() after run(); a split sequence
            overlaps the encoder of step i+1 with the exchange stage of step i)."""
            return pipe.run_many([(frames_d, xy_d)] * n, k=K_CAND, exclude_band=0)

        def e2e_steps(n):
            """n sequences through the public streaming call: host (pinned) frames + keypoints in, candidate lists
            out; every step's H2D and D2H copies are issued inside the timed region."""
            outs = pipe.run_host_stream([(frames_pin, xy_pin)] * n, k=K_CAND, exclude_band=0)
            out_s_pin.copy_(outs[-1][0])
            out_i_pin.copy_(outs[-1][1])

        ms_step = ctx.timed_stream(steps_device, args.steps, args.warmup, sampler)
        ms_latency = ctx.timed(lambda: pipe.run(frames_d, xy_d, k=K_CAND, exclude_band=0), max(args.steps // 2, 2), 1)
        e2e_steps(max(args.warmup, 1))
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_steps(args.steps)
        e1.record()
        ctx.barrier()
        ms_e2e = ctx.max_over_ranks(e0.elapsed_time(e1)) / args.steps
        h2d, d2h = pipe.host_bytes_per_step(frames_pin, xy_pin, K_CAND)
        arm = {"weights": WEIGHT_NOTES[weights], "value": N_FRAMES / (ms_step * 1e-3), "ms_per_step": ms_step,
               "ms_per_step_unpipelined": ms_latency,
               "e2e": {"value": N_FRAMES / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
               "encoder_precision": pipe.encoder.chosen_precision(),
               "encoder_probe_err_1prod_2prod_vs_3prod": list(pipe.encoder.probe_stats())}
        if rank == 0 and detail:
            arm.update(single_gpu_detail(pipe, weights))
        elif rank == 0 and world > 1:
            arm.update(sharded_detail(pipe, weights))
        return arm, pipe

    def t_loop(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def sharded_detail(pipe, weights):
        """N > 1, rank 0: roofline of the dominant kernel on THIS rank's share (the encoder layers over the rank's
        block of frames, timed in a loop of their own)."""
        reps = max(args.steps, 5)
        enc = pipe.encoder
        start, end, _ = pipe.frame_block(N_FRAMES, ctx.rank, world)
        f_loc, x_loc = frames_d[start:end], xy_d[start:end]
        if args.split_pixel_input:
            hi, lo = ops.patch_gather(f_loc, x_loc, PATCH, True, need_lo=enc.needs_lo_input())
        else:
            hi, lo = ops.patch_gather_u8(f_loc, x_loc, PATCH, True), None
        rows = hi.shape[0]
        products = {"fp16": [1] * 5, "fp16x2a16": [2] * 5, "fp16x2": [3] * 5}[enc.chosen_precision()]
        if not args.split_pixel_input and products[0] == 3:
            products[0] = 2
        exec_mult = (products[0] * 1681 * 2500 + sum(products[1:]) * 2500 * 2500) / (1681 * 2500 + 4 * 2500 * 2500)
        ms = t_loop(lambda: enc.encode_planes(hi, lo, rows), reps)
        flop = (end - start) * ENC_FLOP_PER_FRAME
        layer_ms = ms / (len(DIMS) - 1)
        ach = flop / (len(DIMS) - 1) / (layer_ms * 1e-3) / 1e12
        kname = "gemm_pair_kernel<BiasActPolicy<%s>>" % ("64,1" if products[-1] == 1 else "32,3")
        m_pairs = -(-(-(-rows // 128)) // 2)
        return {"roofline": {
            "kernel": kname + " (the five encoder layers over this rank's %d frames)" % (end - start), "bound": "tensor",
            "achieved": ach, "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tflops_sustained"],
            "traffic": None, "peak_source": pk["source"] + ", sustained bf16", "ms_per_launch": layer_ms,
            "launches_per_step": len(DIMS) - 1, "algorithmic_flop_per_launch": flop / (len(DIMS) - 1),
            "tensor_products_per_layer": products, "executed_over_algorithmic_flop": exec_mult,
            "frac_executed": ach * exec_mult / pk["tflops_sustained"],
            "note": "rank 0's share; %d pair-tiles per layer on %d CTA pairs = %.2f waves (the partial last wave is "
                    "what a split sequence loses in the encoder)" % (m_pairs * 10, 74, m_pairs * 10 / 74.0)}}

    def single_gpu_detail(pipe, weights):
        """Stage split + rooflines of this rank's kernels (each stage timed alone with CUDA events on the launch
        stream; informational - `value` is the whole step)."""
        reps = max(args.steps, 5)
        enc = pipe.encoder
        if args.split_pixel_input:
            gather = lambda: ops.patch_gather(frames_d, xy_d, PATCH, True, need_lo=enc.needs_lo_input())  # noqa: E731
        else:
            gather = lambda: (ops.patch_gather_u8(frames_d, xy_d, PATCH, True), None)  # noqa: E731
        hi, lo = gather()
        rows = hi.shape[0]
        products = {"fp16": [1] * 5, "fp16x2a16": [2] * 5, "fp16x2": [3] * 5}[enc.chosen_precision()]
        if not args.split_pixel_input and products[0] == 3:
            products[0] = 2                       # exact 8-bit pixel plane: no residual product in layer 0
        exec_mult = (products[0] * 1681 * 2500 + sum(products[1:]) * 2500 * 2500) / (1681 * 2500 + 4 * 2500 * 2500)
        st = {"patch_gather": t_loop(gather, reps)}
        st["encoder_layers"] = t_loop(lambda: enc.encode_planes(hi, lo, rows), reps)
        desc = pipe.encode(frames_d, xy_d)
        dview = desc.view(N_FRAMES, P, -1)
        sim = lambda: pipe.similarity(desc, N_FRAMES)  # noqa: E731
        sim()
        torch.cuda.synchronize()
        probe = args.sim_precision in ("auto", "fp16r")
        sim_stats = ops.sdav_similarity_stats(N_FRAMES, P, DIMS[-1]) if probe else None
        sim_products = 3 if (args.sim_precision == "fp16x2" or (sim_stats and not sim_stats["use_refine"])) else 1

        def time_gram(mode):
            _lib.call("dlc_sdav_debug_gram_only", mode)
            try:
                return t_loop(sim, reps)
            finally:
                _lib.call("dlc_sdav_debug_gram_only", 0)

        # mode 1: Gram kernel + (one-product mode) the second pass over the deferred pairs; mode 2: Gram kernel alone
        gram_total = time_gram(1)
        gram = time_gram(2) if sim_products == 1 else gram_total
        st["similarity_total"] = t_loop(sim, reps)
        st["gram_kernel"] = gram
        st["gram_plus_refinement_pass"] = gram_total
        # the preparation kernels on their own, through the staged ABI (the same kernels on one part): a difference of
        # the two loops above would mix two power states
        D_ = DIMS[-1]
        colsums = torch.empty((1, 2 * D_), dtype=torch.float64, device="cuda")
        w_d = torch.empty(D_, dtype=torch.float64, device="cuda")
        c_d = torch.empty(D_, dtype=torch.float32, device="cuda")
        plane = torch.empty((N_FRAMES * P, ops.plane_ld(D_)), dtype=torch.float16, device="cuda")
        stats = torch.empty(ops.sdav_stage_stats_bytes(N_FRAMES), dtype=torch.uint8, device="cuda")
        prep_prec = "fp16r" if args.sim_precision == "auto" else args.sim_precision
        plane_lo = torch.empty_like(plane) if prep_prec == "fp16x2" else None

        def prep():
            ops.sdav_stage_colsum(desc, colsums[0])
            ops.sdav_stage_weights(colsums, N_FRAMES * P, w_d, c_d, pipe.sim_args["mu"], pipe.sim_args["sigma"])
            ops.sdav_stage_prepare(desc, N_FRAMES, N_FRAMES, P, w_d, c_d, prep_prec, plane, plane_lo, stats)

        st["similarity_preparation"] = t_loop(prep, reps)
        del plane, plane_lo, stats
        st["topk"] = t_loop(lambda: ops.topk_rows(pipe.last_similarity, K_CAND, largest=True, exclude_band=0), reps)
        enc_flop = N_FRAMES * ENC_FLOP_PER_FRAME
        layer_ms = st["encoder_layers"] / (len(DIMS) - 1)
        ach = enc_flop / (len(DIMS) - 1) / (layer_ms * 1e-3) / 1e12
        kname = "gemm_pair_kernel<BiasActPolicy<%s>>" % ("64,1" if products[-1] == 1 else "32,3")
        roof = {"kernel": kname + " (the five encoder layers; largest share of the step)", "bound": "tensor",
                "achieved": ach, "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tflops_sustained"],
                "traffic": kernel_traffic(kname, weights),
                "peak_source": pk["source"] + ", sustained bf16 (five launches per step inside a long loop); burst %.1f"
                               % pk["tflops_burst"],
                "ms_per_launch": layer_ms, "launches_per_step": len(DIMS) - 1,
                "algorithmic_flop_per_launch": enc_flop / (len(DIMS) - 1),
                "tensor_products_per_layer": products, "executed_over_algorithmic_flop": exec_mult,
                "frac_executed": ach * exec_mult / pk["tflops_sustained"],
                "note": "algorithmic FLOPs 2*M*K*N of the layers as the reference computes them; the arithmetic mode "
                        "the precision probe chose issues `tensor_products_per_layer` fp16 products per algorithmic "
                        "one (three are needed for 1e-3 parity on N(0,1) weights, SURVEY 7), so `frac` is capped at "
                        "1/executed_over_algorithmic_flop and `frac_executed` is the tensor-pipe view"}
        gname = "GramRefinePolicy" if sim_products == 1 else "GramPolicy<32,3>"
        g_ach = GRAM_FLOP / (gram * 1e-3) / 1e12
        roof_sim = {"kernel": "gemm_pair_kernel<%s> (SDAV Gram + argmin + score)" % gname, "bound": "tensor",
                    "achieved": g_ach, "peak": pk["tflops_burst"], "unit": "TFLOP/s", "frac": g_ach / pk["tflops_burst"],
                    "frac_of_sustained_peak": g_ach / pk["tflops_sustained"],
                    "traffic": kernel_traffic("gemm_pair_kernel<%s>" % gname, weights), "ms_per_launch": gram,
                    "algorithmic_flop_per_launch": GRAM_FLOP, "tensor_products": sim_products,
                    "with_refinement_pass": {"ms": gram_total,
                                             "frac_of_burst": GRAM_FLOP / (gram_total * 1e-3) / 1e12 / pk["tflops_burst"],
                                             "frac_of_sustained": GRAM_FLOP / (gram_total * 1e-3) / 1e12 / pk["tflops_sustained"]},
                    "with_preparation_and_refinement": {
                        "ms": st["similarity_total"],
                        "frac_of_sustained": GRAM_FLOP / (st["similarity_total"] * 1e-3) / 1e12 / pk["tflops_sustained"]},
                    "precision_probe": sim_stats,
                    "peak_source": pk["source"] + ", burst bf16 for the kernel timed alone (back-to-back launches of "
                                   "itself), sustained for the figures that include the other passes",
                    "note": "algorithmic FLOPs of the i<j pairs (2*30*30*2500 each). The descriptors of the %s arm are "
                            "%s: saturated operands draw less power, so the clock under this kernel is higher than "
                            "under the cuBLAS run that set the peak" % (
                                weights, "89 % saturated 0/1 values" if weights == "normal" else "dense mid-range values")}
        return {"stages_ms": st, "roofline": roof, "roofline_sim": roof_sim,
                "step_tflops_algorithmic": (enc_flop + GRAM_FLOP) / 1e12}

    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    main_arm, pipe = measure_arm(args.weights, sampler, detail=(world == 1))
    other = "xavier" if args.weights == "normal" else "normal"
    arms = {args.weights: main_arm}
    del pipe
    if not args.one_arm:
        arms[other], _ = measure_arm(other, None, detail=(world == 1))

    extra = {}
    if world > 1 and not args.headline_only:
        extra["config4_sharded"] = bench_config4(ctx, args, batches=(32, 256, 1024))
    if world == 1 and not args.headline_only:
        extra["other_configs"] = {"config3": bench_config3(ctx, args), "config4": bench_config4(ctx, args, (32, 1024)),
                                  "config5": bench_config5(ctx, args)}
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # timed in a fresh process (it forks one worker per core, which a CUDA-initialised process must not do)
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                  "--warmup", "0"], capture_output=True, text=True, timeout=900).stdout
            cpu_base = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])["cpu_baseline"]
        except Exception as e:  # noqa: BLE001
            cpu_base = {"value": None, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": "failed: %r" % (e,)}
    if rank != 0:
        return
    m = main_arm
    prod = {"fp16": "f16 operands", "fp16x2a16": "f16 activations x f16 hi/lo split weights",
            "fp16x2": "f16 hi/lo split operands"}.get(m["encoder_precision"], m["encoder_precision"])
    line = {"metric": METRIC, "value": m["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "ms_per_step_unpipelined": m["ms_per_step_unpipelined"],
            "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": prod + ", f32 accumulate", "data": "synthetic",
            "config": workload_config(args.weights, world),
            "config_detail": {"precision": args.precision, "encoder_precision_chosen": m["encoder_precision"],
                              "sim_precision": args.sim_precision,
                              "layer0_input": "pixel/255 hi/lo planes" if args.split_pixel_input else
                                              "exact 8-bit pixel plane, 1/255 folded into layer 0",
                              "pipelining": "none (one GPU: run() after run())" if world == 1 else
                                            "successive steps pipelined: encoder of step i+1 on a side stream under the "
                                            "exchange + score stages of step i; ms_per_step_unpipelined = one step at "
                                            "a time"},
            "e2e": m["e2e"], "gpu_launches": step_launches(args.sim_precision, world) * args.steps,
            "clocks": sampler.summary() if sampler else None,
            "roofline": m.get("roofline"), "roofline_sim": m.get("roofline_sim"), "cpu_baseline": cpu_base,
            "stages_ms": m.get("stages_ms"),
            "arms": {k: {kk: vv for kk, vv in v.items()} for k, v in arms.items() if k != args.weights}}
    line.update(extra)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- configs 3 / 4 / 5
def bench_config3(ctx, args, reps=5):
    """BASELINE config 3: cnn_vtl descriptors of 1063 frames 192x240x3 + exact Hamming matrix + cosine top-10."""
    torch = ctx.torch
    from deeploopcloser_b200 import ops
    from deeploopcloser_b200.cnn_vtl import CnnVtl
    N = N_FRAMES
    net = CnnVtl(input_shape=[N, H, W, 3], weights="synthetic", seed=3, precision="fp16x2")
    net.DEVICE_CHUNK = N            # one device pass, like the reference's single session.run over all N images (7.4 GB)
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    x = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
    state = {}

    def step():
        d = torch.cat([net._forward_chunk(x[s:s + net.DEVICE_CHUNK]) for s in range(0, N, net.DEVICE_CHUNK)])
        state["D"] = ops.hamming_matrix(d)
        state["cand"] = net.cosine_candidates(d, k=K_CAND)
        state["d"] = d

    def head_only():
        for s in range(0, N, net.DEVICE_CHUNK):
            net._forward_chunk(x[s:s + net.DEVICE_CHUNK])

    ms = ctx.timed(step, reps, 3)
    ms_head = ctx.timed(head_only, reps, 1)
    tf = 1.748e9 * N / ms_head / 1e9
    ok = bool((state["cand"][1][:, 0] != torch.arange(N, device="cuda")).all())
    return {"workload": "cnn_vtl descriptors (conv head, fp16x2) + Hamming matrix + cosine top-10, 1063 frames 192x240x3",
            "frames_per_s": N / ms * 1e3, "ms_per_step": ms, "conv_head_ms": ms_head, "device_chunk": net.DEVICE_CHUNK,
            "conv_head_algorithmic_tflops": tf, "conv_head_frac_of_sustained_tensor_peak": tf / peaks()["tflops_sustained"],
            "descriptor_len": int(state["d"].shape[1]), "self_excluded_from_candidates": ok}


def bench_config4(ctx, args, batches=(32, 256, 1024), rows=1_000_000, dim=4096, reps=5):
    """BASELINE config 4: synthetic 1 M x 4096 database (L2-normalised N(0,1) rows, fp16), batched top-10 queries with
    planted near-duplicates; at N > 1 the database is row-sharded rows / N per rank (strong scaling) and the per-shard
    lists are merged over NCCL."""
    torch = ctx.torch
    from deeploopcloser_b200.matcher import ShardedKeyframeDatabase
    world, rank = ctx.world, ctx.rank
    shard = rows // world
    db = ShardedKeyframeDatabase(dim, shard, "cos", "fp16")
    g = torch.Generator(device="cuda")
    g.manual_seed(5 + rank)
    keep = None
    for s in range(0, shard, 65536):
        chunk = torch.randn((min(65536, shard - s), dim), device="cuda", generator=g)
        if s == 0:
            keep = chunk[:128].clone()
        db.append_local(chunk)
    pk = peaks()
    out = []
    for B in batches:
        gq = torch.Generator(device="cuda")
        gq.manual_seed(6)
        q = torch.randn((B, dim), device="cuda", generator=gq)
        nplant = min(max(B // 10, 1), 128)
        src = keep[:nplant].clone()
        if world > 1:
            ctx.dist.broadcast(src, 0)          # planted near-duplicates of rank 0's first rows, same queries everywhere
        q[:nplant] = src + 0.05 * torch.randn((nplant, dim), device="cuda", generator=gq)
        res = {}

        def step():
            res["s"], res["i"] = db.topk(q, K_CAND)

        ms = ctx.timed(step, reps, 3)
        flop, byts = 2.0 * B * dim * rows, float(rows) * dim * 2
        out.append({"B": B, "ms": ms, "queries_per_s": B / ms * 1e3, "tflops": flop / ms / 1e9,
                    "frac_of_sustained_tensor_peak_x_gpus": flop / ms / 1e9 / pk["tflops_sustained"] / world,
                    "db_gbs": byts / ms / 1e6, "frac_of_hbm_peak_x_gpus": byts / ms / 1e6 / pk["hbm_gbs"] / world,
                    "planted_top1_ok": bool((res["i"][:nplant, 0].cpu() == torch.arange(nplant)).all())})
    del db
    torch.cuda.empty_cache()
    return {"workload": "1M x 4096 fp16 cosine database, top-10, row-sharded %d rows per rank%s" % (
        shard, "" if world == 1 else ", NCCL all-gather + merge of the [B,k] lists"), "n_gpus": world, "batches": out}


def bench_config5(ctx, args, db_rows=1_000_000, batch=256, reps=3):
    """BASELINE config 5 shape on the available GPUs: 640x480 frames at batch 256 through SDA encode + incremental
    match against a keyframe database (1 M rows per GPU here; the 10 M-row / 8-GPU run is tools/bench_streaming.py)."""
    torch = ctx.torch
    from deeploopcloser_b200.streaming import StreamingLoopCloser
    world, rank = ctx.world, ctx.rank
    b_local = batch // world
    shard = db_rows // world
    sl = StreamingLoopCloser(shard + (reps + 4) * b_local, DIMS, precision=args.precision)
    sl.set_weights(*reference_weights())
    g = torch.Generator(device="cuda")
    g.manual_seed(50 + rank)
    for s in range(0, shard, 65536):
        sl.db.append_local(torch.rand((min(65536, shard - s), DIMS[-1]), device="cuda", generator=g))
    frames = torch.randint(0, 256, (b_local, 480, 640), dtype=torch.uint8, device="cuda", generator=g)
    xy = torch.stack([torch.rand((b_local, P), device="cuda", generator=g) * 640,
                      torch.rand((b_local, P), device="cuda", generator=g) * 480], -1).contiguous()
    ms = ctx.timed(lambda: sl.step(frames, xy), reps, 3)
    del sl
    torch.cuda.empty_cache()
    return {"workload": "streaming: 640x480 frames, batch %d, SDA encode + top-10 against a %d-row database + insertion"
                        % (batch, db_rows), "n_gpus": world, "ms_per_batch": ms, "frames_per_s": batch / ms * 1e3}


def run_other_config(ctx, args):
    fn = {3: bench_config3, 4: bench_config4, 5: bench_config5}[args.config]
    sampler = ClockSampler(ctx.local_rank) if ctx.rank == 0 else None
    if sampler:
        sampler.start()
    r = fn(ctx, args)
    if sampler:
        sampler.stop_flag.set()
        sampler.join()
    if ctx.rank != 0:
        return
    if args.config == 4:
        best = max(r["batches"], key=lambda b: b["queries_per_s"])
        value, ms = best["queries_per_s"], best["ms"]
    elif args.config == 3:
        value, ms = r["frames_per_s"], r["ms_per_step"]
    else:
        value, ms = r["frames_per_s"], r["ms_per_batch"]
    print(json.dumps({"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": ctx.world, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
                      "config": {"workload": r["workload"], "baseline_config": args.config},
                      "clocks": sampler.summary() if sampler else None, "detail": r}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json config (1-based) to run as "
                    "the headline; 2 = the config the metric is quoted on")
    ap.add_argument("--weights", default="normal", choices=["normal", "xavier"],
                    help="headline arm: the reference's N(0,1) initialisation, or trained-like Xavier-scaled weights")
    ap.add_argument("--one-arm", action="store_true", help="skip the other weights arm")
    ap.add_argument("--headline-only", action="store_true", help="skip the config 3/4/5 side measurements")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp16x2", "fp16x2a16", "fp16"],
                    help="encoder arithmetic (auto: a probe on the weights picks the cheapest that holds 1e-3)")
    ap.add_argument("--sim-precision", default="auto", choices=["auto", "fp16r", "fp16x2", "fp16"],
                    help="SDAV score-matrix arithmetic")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split-pixel-input", action="store_true",
                    help="A/B: feed layer 0 pixel/255 as hi/lo planes instead of exact pixel values")
    ap.add_argument("--cpu-frames", default=96, type=int, help="reference arm: frames of the sub-workload")
    ap.add_argument("--pipeline-stages", type=int, default=0, choices=[0, 2, 3],
                    help="N > 1: stages of the step pipeline (0 = the pipeline's default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    ctx = Ctx()
    try:
        if args.config == 2:
            run_config2(ctx, args)
        else:
            run_other_config(ctx, args)
    finally:
        ctx.close()


if __name__ == "__main__":
    main()
