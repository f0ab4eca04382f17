#!/usr/bin/env python
"""Whole-matrix parity of BASELINE config 2 (1063 frames, 30 keypoints, reference-default N(0,1) weights), FROM PIXELS:
all 564 453 frame pairs and all 1063 loop-candidate lists of the device pipeline against the float64 oracle on the
host cores (test infrastructure; run through tools/check_config2_full.py on a GPU box, ~4 minutes on 16 cores).

Three comparisons, written to profiles/r2_config2_full_parity.json:
  stage       device scores vs the float64 oracle evaluated on the DEVICE's descriptors: checks every arg-min and score
              of the matcher (the `auto` arithmetic: one tensor product + exact refinement inside a statistical
              margin) - any difference beyond 1e-3 is a matcher failure;
  from_pixels device scores vs pixels -> float64 patches -> float64 encoder -> float64 scores; every pair outside 1e-3
              is classified (tests/parity_report.py): nearest-neighbour tie inside the measured descriptor error,
              conditioning of b*ln(s_k), or unexplained;
  modes       the pairs where `auto` and the three-product `fp16x2` similarity differ by more than 1e-3, each checked
              against the stage oracle: which arithmetic is the one that agrees.
The device part runs in a child process (this one forks a worker per core and must not hold a CUDA context)."""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
TOL = 1e-3
K = 10
_STATE = {}


def device_part(args):
    """Child process: the device pipeline on the bench inputs -> descriptors, S (auto), S (fp16x2), candidates."""
    import torch

    import bench
    from deeploopcloser_b200 import ops
    from deeploopcloser_b200.pipeline import LoopClosurePipeline
    frames, xy = bench.synthetic_inputs(100)
    frames, xy = frames[:args.frames], xy[:args.frames]
    ws, bs = bench.reference_weights()
    pipe = LoopClosurePipeline(bench.DIMS, precision=args.precision, sim_precision="auto")
    pipe.set_weights(ws, bs)
    res = pipe.run(torch.from_numpy(frames).cuda(), torch.from_numpy(xy).cuda(), k=K, exclude_band=0)
    torch.cuda.synchronize()
    stats = ops.sdav_similarity_stats(args.frames, 30, bench.DIMS[-1])
    dview = res["descriptors"].view(args.frames, 30, -1)
    S3 = ops.sdav_similarity(dview, precision="fp16x2")
    np.save(os.path.join(args.tmp, "desc.npy"), res["descriptors"].cpu().numpy())
    np.save(os.path.join(args.tmp, "S_auto.npy"), res["similarity"].cpu().numpy())
    np.save(os.path.join(args.tmp, "S_x2.npy"), S3.cpu().numpy())
    np.save(os.path.join(args.tmp, "cand_idx.npy"), res["candidates"][1].cpu().numpy())
    with open(os.path.join(args.tmp, "stats.json"), "w") as f:
        json.dump({"probe": stats, "encoder_precision": pipe.encoder.chosen_precision()}, f)


def standin_part(args):
    """No GPU (dry run of this script's logic): a float32 NumPy encoder + float64 scores stand in for the device."""
    import bench
    from oracle import matcher as o_m
    from oracle import patches as o_patch
    from oracle import similarity as o_sim
    frames, xy = bench.synthetic_inputs(100)
    ws, bs = bench.reference_weights()
    h = np.concatenate([o_patch.extract_patches(frames[i], xy[i]) for i in range(args.frames)]).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for w, b in zip(ws, bs):
            h = (1.0 / (1.0 + np.exp(-(h @ w.astype(np.float32) + b.astype(np.float32))))).astype(np.float32)
        S = o_sim.similarity_matrix(h.reshape(args.frames, 30, -1).astype(np.float64)).astype(np.float32)
    np.save(os.path.join(args.tmp, "desc.npy"), h)
    np.save(os.path.join(args.tmp, "S_auto.npy"), S)
    np.save(os.path.join(args.tmp, "S_x2.npy"), S)
    Sm = S.astype(np.float64)
    np.fill_diagonal(Sm, -np.inf)
    np.save(os.path.join(args.tmp, "cand_idx.npy"), o_m.topk(Sm, min(K, args.frames - 1))[1])
    with open(os.path.join(args.tmp, "stats.json"), "w") as f:
        json.dump({"probe": None, "encoder_precision": "float32 NumPy stand-in (no device)"}, f)


def _rows(args):
    """Worker: oracle scores of frame i against every j > i, on descriptor set `which`."""
    from oracle import similarity as o_sim
    which, i = args
    d, w = _STATE[which], _STATE["w_" + which]
    out = np.empty(len(d) - i - 1)
    with np.errstate(all="ignore"):
        for n, j in enumerate(range(i + 1, len(d))):
            out[n] = o_sim.similarity_score(d[i], d[j], w)
    return i, out


def _classify(pairs):
    """Worker: class of each frame pair outside the tolerance (tests/parity_report.py)."""
    import parity_report as pr
    out = []
    for i, j in pairs:
        c, det = pr.classify_pair(_STATE["S_auto"][i, j], _STATE["dev"][i], _STATE["dev"][j], _STATE["ref"][i],
                                  _STATE["ref"][j], _STATE["w_dev"], _STATE["w_ref"], tol=TOL, s_ref=None)
        out.append((i, j, c, det if c == "unexplained" else None))
    return out


def oracle_matrix(which, pool, n):
    S = np.full((n, n), -1.0)
    order = sorted(range(n - 1), key=lambda i: i)            # long rows first: balanced tail
    for i, row in pool.imap_unordered(_rows, [(which, i) for i in order], chunksize=4):
        S[i, i + 1:] = row
        S[i + 1:, i] = row
    return S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1063)
    ap.add_argument("--precision", default="fp16x2", help="encoder arithmetic of the device pipeline")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_config2_full_parity.json"))
    ap.add_argument("--standin", action="store_true", help="dry run without a GPU (float32 NumPy stand-in)")
    ap.add_argument("--device-part", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--tmp", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.device_part:
        return device_part(args)

    import bench
    import parity_report as pr
    from oracle import patches as o_patch
    from oracle import sda as o_sda
    from oracle import similarity as o_sim
    n = args.frames
    t_all = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        args.tmp = tmp
        if args.standin:
            standin_part(args)
        else:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--device-part", "--tmp", tmp,
                                   "--frames", str(n), "--precision", args.precision])
        desc_dev = np.load(os.path.join(tmp, "desc.npy")).astype(np.float64).reshape(n, 30, -1)
        S_auto = np.load(os.path.join(tmp, "S_auto.npy")).astype(np.float64)
        S_x2 = np.load(os.path.join(tmp, "S_x2.npy")).astype(np.float64)
        cand_idx = np.load(os.path.join(tmp, "cand_idx.npy"))
        with open(os.path.join(tmp, "stats.json")) as f:
            dev_stats = json.load(f)

    # ---- oracle from pixels
    frames, xy = bench.synthetic_inputs(100)
    ws, bs = bench.reference_weights()
    t0 = time.perf_counter()
    x = np.concatenate([o_patch.extract_patches(frames[i], xy[i]) for i in range(n)])
    desc_ref = o_sda.sda_forward(x, ws, bs).reshape(n, 30, -1)
    t_enc = time.perf_counter() - t0
    _STATE.update(dev=desc_dev, ref=desc_ref, w_dev=o_sim.distinctive_weights(desc_dev),
                  w_ref=o_sim.distinctive_weights(desc_ref), S_auto=S_auto)
    cores = os.cpu_count() or 1
    iu = np.triu_indices(n, 1)
    n_pairs = len(iu[0])
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        S_stage = oracle_matrix("dev", pool, n)
        S_ref = oracle_matrix("ref", pool, n)
        t_scores = time.perf_counter() - t0
        # from pixels: classes of every pair outside the tolerance (in the pool: ~6 ms of float64 work per pair)
        with np.errstate(invalid="ignore"):
            rel_px = np.abs(S_auto - S_ref) / np.maximum(1.0, np.abs(S_ref))
        rel_px[~np.isfinite(S_ref) & ~np.isfinite(S_auto)] = 0.0
        outside = [(int(i), int(j)) for i, j in zip(*iu) if not rel_px[i, j] <= TOL]
        chunks = [outside[c:c + 64] for c in range(0, len(outside), 64)]
        classified = [r for part in pool.imap_unordered(_classify, chunks) for r in part]

    # ---- stage: the matcher on its own descriptors
    with np.errstate(invalid="ignore"):
        rel_stage = np.abs(S_auto - S_stage) / np.maximum(1.0, np.abs(S_stage))
    rel_stage[~np.isfinite(S_stage) & ~np.isfinite(S_auto)] = 0.0
    bad_stage = [(int(i), int(j)) for i, j in zip(*iu) if not rel_stage[i, j] <= TOL]
    stage_detail = []
    for i, j in bad_stage[:20]:
        gap = o_sim.nn_margin(desc_dev[i], desc_dev[j])
        stage_detail.append({"pair": [i, j], "device": float(S_auto[i, j]), "oracle_on_device_descriptors":
                             float(S_stage[i, j]), "smallest_nn_gap": float(gap.min())})

    # ---- modes: auto vs three products, every differing pair against the stage oracle
    with np.errstate(invalid="ignore"):
        rel_modes = np.abs(S_auto - S_x2) / np.maximum(1.0, np.abs(S_x2))
        rel_x2 = np.abs(S_x2 - S_stage) / np.maximum(1.0, np.abs(S_stage))
    differing = [(int(i), int(j)) for i, j in zip(*iu) if rel_modes[i, j] > TOL]
    auto_right = sum(1 for i, j in differing if rel_stage[i, j] <= TOL)
    x2_right = sum(1 for i, j in differing if rel_x2[i, j] <= TOL)
    x2_bad_total = int(np.sum(~(rel_x2[iu] <= TOL) & np.isfinite(S_stage[iu])))

    # ---- from pixels: classes of every pair outside the tolerance
    counts = {"ok": n_pairs - len(outside), "tie": 0, "conditioning": 0, "unexplained": 0}
    cls = np.full((n, n), "ok", dtype=object)
    unexplained = []
    for i, j, c, det in classified:
        counts[c] += 1
        cls[i, j] = cls[j, i] = c
        if c == "unexplained":
            unexplained.append({"pair": [i, j], **det})
    ok_mask = np.ones(n_pairs, dtype=bool)
    ok_mask[[k for k, (i, j) in enumerate(zip(*iu)) if cls[i, j] != "ok"]] = False
    finite = np.isfinite(S_ref[iu]) & ok_mask
    crep = pr.candidate_report(cand_idx, S_ref, cls, cand_idx.shape[1], TOL)
    # the same lists against the stage oracle: must be identical except exact score ties
    crep_stage = pr.candidate_report(cand_idx, S_stage, np.full((n, n), "ok", dtype=object), cand_idx.shape[1], TOL)

    out = {
        "workload": "BASELINE config 2: %d frames 240x192, 30 keypoints/frame, N(0,1) weights (bench.py inputs, seed 100)" % n,
        "device": dev_stats, "tolerance": TOL, "pairs": n_pairs, "host_cores": cores,
        "oracle_seconds": {"encode_f64": round(t_enc, 1), "two_score_matrices": round(t_scores, 1),
                           "total": round(time.perf_counter() - t_all, 1)},
        "descriptors_from_pixels": {"max_rel_err": float(pr.rel_err(desc_dev, desc_ref).max()),
                                    "normwise_rel_err": float(np.linalg.norm(desc_dev - desc_ref) / np.linalg.norm(desc_ref))},
        "stage": {"pairs_outside_tol": len(bad_stage), "max_rel_err": float(np.nanmax(rel_stage[iu])),
                  "detail": stage_detail,
                  "candidate_lists": crep_stage},
        "modes_auto_vs_fp16x2": {"pairs_differing": len(differing), "auto_agrees_with_oracle": auto_right,
                                 "fp16x2_agrees_with_oracle": x2_right,
                                 "fp16x2_pairs_outside_tol_vs_oracle_total": x2_bad_total},
        "from_pixels": {**counts, "max_rel_err_of_ok_pairs": float(np.max(rel_px[iu][finite])) if finite.any() else 0.0,
                        "unexplained_detail": unexplained[:10], "candidate_lists": crep},
    }
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
    ok = not bad_stage and counts["unexplained"] == 0 and crep["positions_unexplained"] == 0
    print("FULL PARITY %s" % ("OK" if ok else "FAILED"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
