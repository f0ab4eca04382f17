"""The staged score-matrix ABI (dlc_sdav_stage_*: one sequence split over several GPUs) on ONE GPU:
the parts are run one after the other in a single process - plain copies stand in for the NCCL all-gathers - and their
sum must reproduce the monolithic dlc_sdav_similarity and the float64 oracle. `-m gpu`."""
import warnings

import numpy as np
import pytest
import torch

from oracle import similarity as o_sim

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _descriptors(n, P, D, seed, saturated):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, P, D)) * (8.0 if saturated else 1.0)
    return (1.0 / (1.0 + np.exp(-x))).astype(np.float32)


@pytest.mark.parametrize("n,P,D,parts,precision,saturated", [
    (37, 30, 200, 1, "fp16r", True), (37, 30, 200, 3, "fp16r", True), (50, 30, 2500, 4, "fp16r", True),
    (41, 30, 320, 3, "fp16x2", False), (23, 7, 64, 2, "fp16r", False), (64, 32, 128, 8, "auto", True),
    (9, 30, 96, 4, "fp16r", False),           # more parts than tile rows: some parts own nothing
])
def test_staged_parts_sum_to_the_matrix(cuda, n, P, D, parts, precision, saturated):
    from deeploopcloser_b200 import ops
    desc = _descriptors(n, P, D, n + parts, saturated)
    d_all = torch.from_numpy(desc).cuda()
    per = -(-n // parts)
    ld = ops.plane_ld(D)
    rows = parts * per * P
    desc_flat = torch.zeros((rows, D), dtype=torch.float32, device="cuda")
    desc_flat[:n * P] = d_all.view(n * P, D)
    plane = torch.zeros((rows, ld), dtype=torch.float16, device="cuda")
    plane_lo = torch.zeros_like(plane) if precision == "fp16x2" else None
    stats = torch.zeros((parts, ops.sdav_stage_stats_bytes(per)), dtype=torch.uint8, device="cuda")
    colsums = torch.zeros((parts, 2 * D), dtype=torch.float64, device="cuda")
    w = torch.empty(D, dtype=torch.float64, device="cuda")
    mean = torch.empty(D, dtype=torch.float32, device="cuda")             # centring vector (dataset mean or zero)
    blocks = [(min(r * per, n), min((r + 1) * per, n)) for r in range(parts)]
    for r, (s, e) in enumerate(blocks):                          # stage 1 on every "rank"
        ops.sdav_stage_colsum(desc_flat[s * P:e * P], colsums[r])
    ops.sdav_stage_weights(colsums, n * P, w, mean)              # stage 2 (after the "all-gather" of colsums)
    assert np.allclose(w.cpu().numpy(), o_sim.distinctive_weights(desc.astype(np.float64)), rtol=1e-12)
    for r, (s, e) in enumerate(blocks):                          # stage 3: each rank fills its slices
        lo_r, hi_r = r * per * P, (r + 1) * per * P
        ops.sdav_stage_prepare(desc_flat[lo_r:hi_r], e - s, per, P, w, mean, precision, plane[lo_r:hi_r],
                               None if plane_lo is None else plane_lo[lo_r:hi_r], stats[r])
    total = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    for r in range(parts):                                       # stages 4 + 5 on every "rank", then the "all-reduce"
        S = torch.empty((n, n), dtype=torch.float32, device="cuda")
        ops.sdav_stage_gram(plane, plane_lo, stats, parts, per, n, P, D, precision, r, S)
        ops.sdav_stage_fix(plane, desc_flat, n, P, D, precision, r, parts, S)
        if parts == 1:
            total = S
        else:
            total += S
    torch.cuda.synchronize()
    got = total.cpu().numpy().astype(np.float64)
    mono = ops.sdav_similarity(d_all, precision="fp16r" if precision == "auto" else precision).cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = o_sim.similarity_matrix(desc.astype(np.float64))
    rel = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    print("staged n=%d P=%d D=%d parts=%d %s: max rel err vs oracle %.2e, identical to the monolithic call: %s" % (
        n, P, D, parts, precision, rel.max(), np.array_equal(got, mono.astype(np.float64))))
    assert rel.max() <= TOL
    assert np.array_equal(got, got.T) and np.all(np.diag(got) == -1.0)
    assert np.max(np.abs(got - mono) / np.maximum(1.0, np.abs(mono))) <= 1e-5


def test_sharded_pipeline_single_rank_equals_monolithic(cuda):
    """ShardedSequencePipeline without a process group (world 1) runs the staged ABI end to end: same descriptors,
    scores within float32 rounding of the monolithic pipeline's, same candidates."""
    from deeploopcloser_b200.pipeline import LoopClosurePipeline, ShardedSequencePipeline
    from oracle import sda as o_sda
    rng = np.random.default_rng(5)
    n, H, W, P = 45, 96, 128, 30
    dims = [1681, 320, 256]
    frames = torch.from_numpy(rng.integers(0, 256, (n, H, W), dtype=np.uint8)).cuda()
    xy = torch.from_numpy(np.stack([rng.uniform(0, W, (n, P)), rng.uniform(0, H, (n, P))], -1).astype(np.float32)).cuda()
    ws, bs = o_sda.make_weights(dims, seed=1, scale="normal")
    a, b = LoopClosurePipeline(dims), ShardedSequencePipeline(dims)
    a.set_weights(ws, bs)
    b.set_weights(ws, bs)
    ra, rb = a.run(frames, xy, k=5), b.run(frames, xy, k=5)
    torch.cuda.synchronize()
    assert torch.equal(ra["descriptors"], rb["descriptors"])
    sa, sb = ra["similarity"].cpu().numpy(), rb["similarity"].cpu().numpy()
    assert np.max(np.abs(sa - sb) / np.maximum(1.0, np.abs(sa))) <= 1e-5
    assert torch.equal(ra["candidates"][1], rb["candidates"][1])


def test_run_many_pipelining_equals_run(cuda):
    """run_many (the encoder of sequence i+1 on its own stream, into the other descriptor slot, under the score-matrix
    stage of sequence i) returns exactly what run() returns for every sequence, also when the sequences differ."""
    from deeploopcloser_b200.pipeline import ShardedSequencePipeline
    from oracle import sda as o_sda
    rng = np.random.default_rng(9)
    n, H, W, P = 40, 96, 128, 30
    dims = [1681, 320, 256]
    seqs = []
    for _ in range(4):
        f = torch.from_numpy(rng.integers(0, 256, (n, H, W), dtype=np.uint8)).cuda()
        x = torch.from_numpy(np.stack([rng.uniform(0, W, (n, P)), rng.uniform(0, H, (n, P))], -1).astype(np.float32)).cuda()
        seqs.append((f, x))
    ws, bs = o_sda.make_weights(dims, seed=1, scale="normal")
    pipe = ShardedSequencePipeline(dims)
    pipe.set_weights(ws, bs)
    want = []
    for f, x in seqs:
        r = pipe.run(f, x, k=5)
        want.append((r["candidates"][0].clone(), r["candidates"][1].clone()))
    for _ in range(2):
        got = pipe.run_many(seqs, k=5)
        torch.cuda.synchronize()
        for (ws_, wi), (gs, gi) in zip(want, got):
            assert torch.equal(wi, gi) and torch.equal(ws_, gs)
    # the same stream from pinned HOST memory: per-step upload on the encoder stream, results in pinned buffers
    host = [(f.cpu().pin_memory(), x.cpu().pin_memory()) for f, x in seqs]
    for (ws_, wi), (gs, gi) in zip(want, pipe.run_host_stream(host, k=5)):
        assert not gs.is_cuda and torch.equal(wi.cpu(), gi) and torch.equal(ws_.cpu(), gs)


@pytest.mark.parametrize("metric,B,k", [("cos", 70, 10), ("l2", 5, 32), ("dot", 300, 16)])
def test_match_topk_sharded_world1_equals_local(cuda, metric, B, k):
    """dlc_match_topk_sharded on a one-rank NCCL communicator created through the C ABI (dlc_comm_unique_id /
    dlc_comm_create; libnccl resolved with dlopen): fused kernel -> packed list -> (no exchange at world 1) -> rank
    merge kernel must reproduce dlc_match_topk on the same database, offsets included. The multi-rank exchange is
    checked under torchrun by tools/check_sharded.py and on every bench.py --gpus N run (planted neighbours)."""
    import ctypes as C

    from deeploopcloser_b200 import _lib
    from deeploopcloser_b200._cuda import Workspace, ptr, stream_ptr
    from deeploopcloser_b200.matcher import KeyframeDatabase
    rng = np.random.default_rng(B + k)
    N, D = 5000, 192
    scale = 1.0 if metric == "cos" else 0.25
    db = KeyframeDatabase(D, N, metric, "fp16")
    db.append(torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32) * scale).cuda())
    q = torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32) * scale).cuda()
    want_s, want_i = db.topk(q, k, idx_offset=1000)
    uid = (C.c_char * 128)()
    _lib.call("dlc_comm_unique_id", C.cast(uid, C.c_void_p))
    comm = C.c_void_p()
    _lib.call("dlc_comm_create", C.byref(comm), bytes(uid.raw), 0, 1)
    try:
        s = torch.empty((B, k), dtype=torch.float32, device="cuda")
        i = torch.empty((B, k), dtype=torch.int64, device="cuda")
        ws = Workspace()
        w, wb = ws.get(_lib.call("dlc_match_sharded_workspace_bytes", db._h, B, k, 1))
        _lib.call("dlc_match_topk_sharded", db._h, comm, ptr(q), B, k, 1000, ptr(s), ptr(i), w, wb, stream_ptr())
        torch.cuda.synchronize()
        assert torch.equal(i, want_i) and torch.equal(s, want_s)
    finally:
        _lib.call("dlc_comm_destroy", comm)
