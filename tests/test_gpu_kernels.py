"""Parity of the CUDA kernels (called through the C ABI via ctypes) against the float64 oracle. `-m gpu`."""
import warnings

import numpy as np
import pytest
import torch

from oracle import hamming as o_ham
from oracle import matcher as o_match
from oracle import patches as o_patch
from oracle import sda as o_sda
from oracle import similarity as o_sim

pytestmark = pytest.mark.gpu

TOL = 1e-3  # north star: descriptors and scores within 1e-3 relative:  |a-b| <= 1e-3 * max(1, |b|)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def norm_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# ------------------------------------------------------------------------------------------- tensor-core GEMM
@pytest.mark.parametrize("m,k,n", [(3, 2, 2), (128, 64, 32), (130, 200, 96), (600, 1681, 2500), (257, 2500, 300)])
@pytest.mark.parametrize("precision", ["fp16x2", "fp16"])
def test_matmul_vs_oracle(cuda, m, k, n, precision):
    from deeploopcloser_b200 import ops
    rng = np.random.default_rng(m * 7 + k)
    a = rng.uniform(0, 1, (m, k))
    b = rng.standard_normal((k, n))
    bias = rng.standard_normal(n)
    ref = a @ b + bias
    out = ops.matmul(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(bias).cuda(),
                     act="none", precision=precision).cpu().numpy()
    err = norm_err(out, ref)
    print("matmul", (m, k, n), precision, "normwise", err, "max-abs", np.max(np.abs(out - ref)))
    assert err < (3e-6 if precision == "fp16x2" else 2e-3)


def test_matmul_known_answer(cuda, golden_dir):
    """The reference's only test vector (test/TensorflowWrapperTest.py:11-21): exact."""
    from deeploopcloser_b200 import ops
    g = np.load(golden_dir + "/misc.npz")
    x = torch.from_numpy(g["tw_x"]).cuda()
    w = torch.from_numpy(g["tw_w"]).cuda()
    y = ops.matmul(x.reshape(-1, 2), w).reshape(3, 2, 2).double().cpu().numpy()
    assert np.array_equal(y, g["tw_expected"])


@pytest.mark.parametrize("bk", [32, 64])
def test_split_kernel_both_ring_shapes(cuda, bk):
    from deeploopcloser_b200 import _lib, ops
    _lib.call("dlc_debug_set", 0, bk)
    try:
        rng = np.random.default_rng(bk)
        a = rng.uniform(0, 1, (300, 777))
        b = rng.standard_normal((777, 520))
        out = ops.matmul(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
        assert norm_err(out, a @ b) < 3e-6
    finally:
        _lib.call("dlc_debug_set", 0, 32)


# ------------------------------------------------------------------------------------------- SDA encoder
@pytest.mark.parametrize("scale,precision,tol", [("normal", "fp16x2", TOL), ("xavier", "fp16x2", TOL),
                                                 ("xavier", "fp16", 5e-3)])
def test_sda_encode_vs_oracle(cuda, scale, precision, tol):
    from deeploopcloser_b200 import ops
    dims = [1681, 2500, 2500, 2500, 2500, 2500]
    ws, bs = o_sda.make_weights(dims, seed=1, scale=scale)
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, (4 * 30, 1681)).astype(np.float64) / 255.0  # patch-like inputs
    ref = o_sda.sda_forward(x, ws, bs)
    enc = ops.SdaEncoder(dims, precision)
    for l, (w, b) in enumerate(zip(ws, bs)):
        enc.set_layer(l, w, b)
    out = enc.encode(torch.from_numpy(x).cuda()).cpu().numpy()
    e_rel, e_norm = rel_err(out, ref), norm_err(out, ref)
    print("sda", scale, precision, "max |a-b|/max(1,|b|)", e_rel, "normwise", e_norm)
    assert out.shape == (120, 2500)
    assert e_rel <= tol and e_norm <= tol


def test_da_single_layer(cuda):
    from deeploopcloser_b200 import ops
    ws, bs = o_sda.make_weights([1681, 2500], seed=5)
    x = np.random.default_rng(3).uniform(0, 1, (30, 1681))
    enc = ops.SdaEncoder([1681, 2500], "fp16x2")
    enc.set_layer(0, ws[0], bs[0])
    out = enc.encode(torch.from_numpy(x).cuda()).cpu().numpy()
    assert rel_err(out, o_sda.sda_forward(x, ws, bs)) <= TOL


# ------------------------------------------------------------------------------------------- patch gather
def test_patch_gather_golden(cuda, golden_dir):
    from deeploopcloser_b200 import ops
    g = np.load(golden_dir + "/patches.npz")
    img = torch.from_numpy(g["img"]).cuda()
    xy = torch.from_numpy(g["xy"]).cuda()
    out = ops.patch_gather_f64(img, xy).cpu().numpy().reshape(g["out"].shape)
    assert np.array_equal(out, g["out"])  # bit-exact vs the reference's own output
    hi, lo = ops.patch_gather(img, xy)
    v = (hi.double() + lo.double()).cpu().numpy()
    assert v.shape == (60, 1728)
    assert np.max(np.abs(v[:, :1681] - g["out"].reshape(60, 1681))) < 2e-7
    assert np.all(v[:, 1681:] == 0)


@pytest.mark.parametrize("quirk", [True, False])
def test_patch_gather_random(cuda, quirk):
    from deeploopcloser_b200 import ops
    rng = np.random.default_rng(11)
    B, H, W, P = 3, 480, 640, 30
    img = rng.integers(0, 256, (B, H, W), dtype=np.uint8)
    xy = np.stack([rng.uniform(-3, W + 3, (B, P)), rng.uniform(-3, H + 3, (B, P))], axis=-1).astype(np.float32)
    xy[0, 0] = [0.5, 1.5]
    xy[0, 1] = [2.5, 3.5]
    out = ops.patch_gather_f64(torch.from_numpy(img).cuda(), torch.from_numpy(xy).cuda(), swap_xy_quirk=quirk)
    ref = np.concatenate([o_patch.extract_patches(img[b], xy[b], 41, quirk) for b in range(B)])
    assert np.array_equal(out.cpu().numpy(), ref)


# ------------------------------------------------------------------------------------------- Hamming
def test_hamming_golden(cuda, golden_dir):
    from deeploopcloser_b200 import ops
    g = np.load(golden_dir + "/hamming.npz")
    D = ops.hamming_matrix(torch.from_numpy(g["desc"]).cuda()).cpu().numpy()
    assert np.array_equal(D, g["D"])


@pytest.mark.parametrize("n,m,quirk", [(1, 1, True), (70, 2243, True), (130, 515, False)])
def test_hamming_random(cuda, n, m, quirk):
    from deeploopcloser_b200 import ops
    desc = np.random.default_rng(n).integers(-128, 128, (n, m)).astype(np.int8)
    D = ops.hamming_matrix(torch.from_numpy(desc).cuda(), quirk).cpu().numpy()
    assert np.array_equal(D, o_ham.distance_matrix(desc, quirk))
    assert np.all(np.diag(D) == 0)


# ------------------------------------------------------------------------------------------- SDAV similarity
def _check_similarity(S, desc, full):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, det = o_sim.similarity_matrix(desc.astype(np.float64), full_asymmetric=full, return_details=True)
    n = len(desc)
    bad = []
    for i in range(n):
        for j in range(n):
            if i == j:
                assert S[i, j] == -1.0
                continue
            if abs(S[i, j] - ref[i, j]) > TOL * max(1.0, abs(ref[i, j])):
                bad.append((i, j, S[i, j], ref[i, j]))
    # a mismatch is only acceptable when the oracle's nearest-neighbour choice was a tie within tolerance
    ties = 0
    for (i, j, got, want) in bad:
        a, b = (i, j) if (full or i < j) else (j, i)
        margin = o_sim.nn_margin(desc[a].astype(np.float64), desc[b].astype(np.float64))
        assert margin.min() < 1e-3, "score mismatch at %s: got %r want %r (no NN tie, min margin %g)" % (
            (i, j), got, want, margin.min())
        ties += 1
    print("similarity: %d pairs, %d tie-induced mismatches" % (n * (n - 1), ties))


def test_similarity_golden(cuda, golden_dir):
    from deeploopcloser_b200 import ops
    g = np.load(golden_dir + "/similarity.npz")
    for name in "abc":
        desc = g["desc_" + name].astype(np.float32)
        ref = g["S_" + name]  # produced by the reference's SimilarityCalculator
        for precision in ("fp16x2", "fp16r", "auto"):
            S = ops.sdav_similarity(torch.from_numpy(desc).cuda(), full_asymmetric=True, precision=precision).cpu().numpy()
            m = ~np.eye(len(ref), dtype=bool)
            assert np.max(np.abs(S[m] - ref[m]) / np.maximum(1, np.abs(ref[m]))) <= TOL, precision
        m = ~np.eye(len(ref), dtype=bool)
        err = np.max(np.abs(S[m] - ref[m]) / np.maximum(1, np.abs(ref[m])))
        print("similarity golden", name, err)
        assert err <= TOL
        assert np.all(np.diag(S) == -1)


@pytest.mark.parametrize("precision", ["fp16x2", "fp16r", "auto"])
@pytest.mark.parametrize("n,p,d,full", [(9, 30, 2500, False), (21, 30, 300, True), (2, 5, 64, False), (1, 30, 64, False)])
def test_similarity_random(cuda, n, p, d, full, precision):
    from deeploopcloser_b200 import ops
    rng = np.random.default_rng(n)
    desc = (1.0 / (1.0 + np.exp(-4.0 * rng.standard_normal((n, p, d))))).astype(np.float32)
    S = ops.sdav_similarity(torch.from_numpy(desc).cuda(), full_asymmetric=full, precision=precision).cpu().numpy()
    if precision != "fp16x2":
        print(precision, ops.sdav_similarity_stats(n, p, d))
    _check_similarity(S, desc, full)
    if not full:
        assert np.array_equal(S, S.T)


@pytest.mark.parametrize("precision", ["fp16r", "auto"])
def test_similarity_refinement_handles_near_ties(cuda, precision):
    """Frames built so that many rows have two almost equidistant candidates (gap far below the fp16 rounding error of
    a single-product Gram entry): the refinement must still pick the reference's argmin."""
    from deeploopcloser_b200 import ops
    rng = np.random.default_rng(42)
    n, p, d = 12, 30, 1024
    desc = rng.uniform(0.05, 0.95, (n, p, d)).astype(np.float32)
    for f in range(1, n):           # frame f holds near-copies of patches of frame 0, in pairs 1e-4 apart
        for k in range(0, p, 2):
            base = desc[0, (k + f) % p]
            desc[f, k] = base + 3e-3 * rng.standard_normal(d).astype(np.float32)
            desc[f, k + 1] = desc[f, k] + 2e-5 * rng.standard_normal(d).astype(np.float32)
    desc = np.clip(desc, 0, 1)
    S = ops.sdav_similarity(torch.from_numpy(desc).cuda(), precision=precision).cpu().numpy()
    st = ops.sdav_similarity_stats(n, p, d)
    print(precision, st)
    _check_similarity(S, desc, False)
    if precision == "fp16r":
        assert st["use_refine"] == 1 and st["flagged_rows"] > 0


# ------------------------------------------------------------------------------------------- row top-k
@pytest.mark.parametrize("largest", [True, False])
def test_topk_rows(cuda, largest):
    from deeploopcloser_b200 import ops
    rng = np.random.default_rng(5)
    s = rng.standard_normal((37, 211)).astype(np.float32)
    s[:, 5] = s[:, 9]  # ties -> lowest index
    s[3, 7] = np.inf
    s[4, 8] = -np.inf
    s[5, 2] = np.nan
    for band in (-1, 0, 3):
        gs, gi = ops.topk_rows(torch.from_numpy(s).cuda(), 10, largest, band)
        rs, ri = o_match.topk(s, 10, largest, band)
        assert np.array_equal(gi.cpu().numpy(), ri)
        assert np.array_equal(gs.cpu().numpy().astype(np.float64), rs)
    # fewer candidates than k
    gs, gi = ops.topk_rows(torch.from_numpy(s[:, :4].copy()).cuda(), 6, largest)
    rs, ri = o_match.topk(s[:, :4], 6, largest)
    assert np.array_equal(gi.cpu().numpy(), ri)


# ------------------------------------------------------------------------------------------- global matcher
def _stored(db_rows, metric, dtype):
    x = torch.from_numpy(db_rows)
    if metric == "cos":
        x = x / x.norm(dim=1, keepdim=True).clamp_min(1e-30)
    td = torch.float16 if dtype == "fp16" else torch.bfloat16
    return x.to(td).double().numpy()


def _check_topk(scores, idx, ref_scores, k, smaller, scale=None):
    """indices identical except for ties within tolerance (reported); scores within TOL.
    `scale` (L2 only): |q|^2 + |d|^2 per pair. The squared distance is formed as |q|^2 + |d|^2 - 2 q.d in fp32, so its
    absolute error is relative to those norms, not to the (possibly tiny) distance itself."""
    ties = 0

    def tol(r, cols, want):
        t = TOL * np.maximum(1, np.abs(want))
        if scale is not None:
            t = t + 1e-4 * scale[r, cols]
        return t

    for r in range(len(scores)):
        rs, ri = o_match.topk(ref_scores[r:r + 1], k, largest=not smaller)
        if not np.array_equal(idx[r], ri[0]):
            for t in range(k):
                if idx[r, t] != ri[0, t]:
                    assert idx[r, t] >= 0
                    got = ref_scores[r, idx[r, t]]
                    assert abs(got - rs[0, t]) <= 2 * tol(r, idx[r, t], rs[0, t]), (r, t, idx[r], ri[0])
                    ties += 1
        valid = idx[r] >= 0
        want = ref_scores[r, idx[r][valid]]
        assert np.all(np.abs(scores[r][valid] - want) <= tol(r, idx[r][valid], want))
    return ties


@pytest.mark.parametrize("metric", ["cos", "dot", "l2"])
@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
# B = 130 / 256 / 512: an even number of 128-query tiles -> the CTA-pair kernel (two lists per query and CTA group)
@pytest.mark.parametrize("B,N,D,k", [(5, 1000, 128, 10), (130, 5000, 2500, 10), (300, 70000, 256, 32), (1, 37, 64, 5),
                                     (256, 9000, 192, 32), (512, 40000, 128, 10)])
def test_match_topk(cuda, metric, dtype, B, N, D, k):
    from deeploopcloser_b200.matcher import KeyframeDatabase
    rng = np.random.default_rng(B + N)
    db_rows = rng.standard_normal((N, D)).astype(np.float32)
    q = rng.standard_normal((B, D)).astype(np.float32)
    nplant = min(B, N) // 2
    q[:nplant] = db_rows[rng.choice(N, nplant, replace=False)] + 0.05 * rng.standard_normal((nplant, D)).astype(np.float32)
    if metric != "cos":  # keep dot/L2 magnitudes inside fp16 range and scores O(1..100)
        db_rows *= 0.25
        q *= 0.25
    db = KeyframeDatabase(D, N + 7, metric, dtype)
    half = N // 2
    db.append(torch.from_numpy(db_rows[:half]).cuda())   # incremental insertion
    db.append(torch.from_numpy(db_rows[half:]).cuda())
    assert len(db) == N
    scores, idx = db.topk(torch.from_numpy(q).cuda(), k)
    scores, idx = scores.cpu().numpy().astype(np.float64), idx.cpu().numpy()
    code = {"cos": o_match.COS, "dot": o_match.DOT, "l2": o_match.L2}[metric]
    qr = torch.from_numpy(q)
    if metric == "cos":
        qr = qr / qr.norm(dim=1, keepdim=True)
    # oracle on the values the kernel contracts: stored database rows, queries rounded to the storage type
    td = torch.float16 if dtype == "fp16" else torch.bfloat16
    q_used = qr.to(td).double().numpy() if metric != "cos" else qr.to(td).double().numpy()
    ref = o_match.score_matrix(q_used, _stored(db_rows, metric, dtype), o_match.DOT if metric == "cos" else code)
    scale = None
    if metric == "l2":
        st = _stored(db_rows, metric, dtype)
        scale = (q_used ** 2).sum(1)[:, None] + (st ** 2).sum(1)[None, :]
    ties = _check_topk(scores, idx, ref, min(k, N), metric == "l2", scale)
    print("match", metric, dtype, (B, N, D, k), "ties within tol:", ties)
    # end-to-end accuracy vs the unrounded float64 definition
    full = o_match.score_matrix(q.astype(np.float64), _stored(db_rows, metric, dtype), code)
    top1 = full.argmin(1) if metric == "l2" else full.argmax(1)
    agree = np.mean(idx[:, 0] == top1)
    assert agree > (0.97 if dtype == "fp16" else 0.9)


def test_match_threshold(cuda):
    from deeploopcloser_b200.matcher import KeyframeDatabase
    rng = np.random.default_rng(9)
    N, D, B = 3000, 96, 40
    db_rows = rng.standard_normal((N, D)).astype(np.float32)
    q = rng.standard_normal((B, D)).astype(np.float32)
    q[:10] = db_rows[:10] + 0.01
    db = KeyframeDatabase(D, N, "cos", "fp16")
    db.append(torch.from_numpy(db_rows).cuda())
    thr = 0.25
    counts, scores, idx = db.threshold(torch.from_numpy(q).cuda(), thr, 16)
    qn = torch.from_numpy(q)
    qn = (qn / qn.norm(dim=1, keepdim=True)).half().double().numpy()
    ref = qn @ _stored(db_rows, "cos", "fp16").T
    margin = np.abs(ref - thr).min()
    rc, rs, ri = o_match.threshold(ref, thr, 16)
    if margin > 1e-4:
        assert np.array_equal(counts.cpu().numpy(), rc)
        assert np.array_equal(idx.cpu().numpy(), ri)


def test_match_empty_and_errors(cuda):
    from deeploopcloser_b200 import _lib
    from deeploopcloser_b200.matcher import KeyframeDatabase
    db = KeyframeDatabase(64, 10, "cos", "fp16")
    s, i = db.topk(torch.zeros((3, 64), device="cuda"), 4)
    assert np.all(i.cpu().numpy() == -1)
    db.append(torch.randn(10, 64, device="cuda"))
    with pytest.raises(_lib.DlcError):
        db.append(torch.randn(1, 64, device="cuda"))  # over capacity
    with pytest.raises(_lib.DlcError):
        db.topk(torch.zeros((3, 64), device="cuda"), 33)  # k > 32


# ------------------------------------------------------------------------------------------- edge cases
def test_gemm_bf16_planes_and_odd_tiles(cuda):
    """bf16 operand planes (DLC_PREC_BF16) and accumulator widths other than 256 (96 / 192 / 160 columns)."""
    from deeploopcloser_b200 import _lib, ops
    from deeploopcloser_b200._cuda import ptr, stream_ptr
    rng = np.random.default_rng(1)
    for m, k, n, n_pad in ((200, 130, 96, 96), (77, 300, 384, 384), (129, 64, 150, 160)):
        a = rng.standard_normal((m, k)).astype(np.float32)
        b = rng.standard_normal((k, n)).astype(np.float32)
        ld = _lib.plane_ld(k)
        A = torch.zeros((m, ld), dtype=torch.bfloat16, device="cuda")
        A[:, :k] = torch.from_numpy(a).cuda().to(torch.bfloat16)
        Bt = torch.zeros((n_pad, ld), dtype=torch.bfloat16, device="cuda")
        Bt[:n, :k] = torch.from_numpy(b.T.copy()).cuda().to(torch.bfloat16)
        out = torch.empty((m, n), dtype=torch.float32, device="cuda")
        _lib.call("dlc_gemm_planes", ptr(A), None, ptr(Bt), None, m, n, n_pad, ld, None, _lib.ACT_RELU, _lib.PREC_BF16,
                  ptr(out), n, None, None, 0, stream_ptr())
        ref = np.maximum(A[:, :k].double().cpu().numpy() @ Bt[:n, :k].double().cpu().numpy().T, 0)
        assert norm_err(out.cpu().numpy(), ref) < 1e-5, (m, k, n)
        # fp16x2 on the same shapes through the generic wrapper
        got = ops.matmul(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), act="relu").cpu().numpy()
        assert norm_err(got, np.maximum(a.astype(np.float64) @ b.astype(np.float64), 0)) < 3e-6


def test_empty_and_degenerate_inputs(cuda):
    from deeploopcloser_b200 import _lib, ops
    # zero patches / zero rows are accepted and produce empty outputs
    img = torch.zeros((0, 64, 64), dtype=torch.uint8, device="cuda")
    xy = torch.zeros((0, 30, 2), dtype=torch.float32, device="cuda")
    hi, lo = ops.patch_gather(img, xy)
    assert hi.shape == (0, 1728)
    hi, lo = ops.split_planes(torch.zeros((0, 10), dtype=torch.float64, device="cuda"))
    assert hi.shape == (0, 64)
    assert ops.hamming_matrix(torch.zeros((0, 8), dtype=torch.int8, device="cuda")).shape == (0, 0)
    # image smaller than the patch, even patch size, k = 0: loud errors, not silent garbage
    with pytest.raises(_lib.DlcError):
        ops.patch_gather(torch.zeros((1, 40, 64), dtype=torch.uint8, device="cuda"),
                         torch.zeros((1, 3, 2), dtype=torch.float32, device="cuda"))
    with pytest.raises(_lib.DlcError):
        ops.patch_gather(torch.zeros((1, 64, 64), dtype=torch.uint8, device="cuda"),
                         torch.zeros((1, 3, 2), dtype=torch.float32, device="cuda"), patch=40)
    with pytest.raises(_lib.DlcError):
        ops.topk_rows(torch.zeros((2, 4), device="cuda"), 0)
    with pytest.raises(_lib.DlcError):
        ops.sdav_similarity(torch.zeros((2, 33, 8), device="cuda"))      # more than 32 patch rows per frame
    enc = ops.SdaEncoder([64, 32], "fp16x2")
    with pytest.raises(_lib.DlcError):
        enc.encode(torch.zeros((4, 64), device="cuda"))                   # weights never set


def test_similarity_identical_patches_give_inf(cuda):
    """log(0): a matched pair of identical patches makes the reference score +inf (SimilarityCalculator.py:48 with
    b < 0); the kernel must reproduce that, not a NaN or a large finite number."""
    from deeploopcloser_b200 import ops
    rng = np.random.default_rng(3)
    desc = rng.uniform(0.1, 0.9, (3, 30, 128)).astype(np.float32)
    desc[1, 4] = desc[0, 7]                      # frame 0 patch 7 has an exact twin in frame 1
    S = ops.sdav_similarity(torch.from_numpy(desc).cuda()).cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = o_sim.similarity_matrix(desc.astype(np.float64))
    assert np.isposinf(ref[0, 1]) and np.isposinf(S[0, 1]) and np.isposinf(S[1, 0])
    assert np.isfinite(S[0, 2]) and abs(S[0, 2] - ref[0, 2]) <= TOL * abs(ref[0, 2])


def test_match_threshold_l2(cuda):
    from deeploopcloser_b200.matcher import KeyframeDatabase
    rng = np.random.default_rng(2)
    N, D, B = 800, 64, 9
    rows = (rng.standard_normal((N, D)) * 0.25).astype(np.float32)
    q = rows[:B] + 0.01
    db = KeyframeDatabase(D, N, "l2", "fp16")
    db.append(torch.from_numpy(rows).cuda())
    counts, scores, idx = db.threshold(torch.from_numpy(q).cuda(), 0.5, 8)
    st = torch.from_numpy(rows).half().double().numpy()
    ref = o_match.score_matrix(torch.from_numpy(q).half().double().numpy(), st, o_match.L2)
    rc, rs, ri = o_match.threshold(ref, 0.5, 8, smaller_is_better=True)
    if np.abs(ref - 0.5).min() > 1e-3:
        assert np.array_equal(counts.cpu().numpy(), rc) and np.array_equal(idx.cpu().numpy(), ri)
    assert np.array_equal(idx[:, 0].cpu().numpy(), np.arange(B))


# ------------------------------------------------------------------------------------------- fused cnn_vtl head
@pytest.mark.parametrize("hw,n", [((67, 83), 3), ((192, 240), 2), ((224, 224), 1), ((35, 43), 5)])
@pytest.mark.parametrize("dtype", ["uint8", "float64"])
def test_cnnvtl_fused_layers_vs_oracle(cuda, hw, n, dtype):
    """Every convolution of the fused head (conv1 from the image, conv2..5 as implicit GEMMs over im2col-mode TMA)
    against the float64 oracle, layer by layer; M tiles cross image boundaries (130 / 12 pixels per image)."""
    from deeploopcloser_b200.cnn_vtl import CnnVtl
    from oracle import cnnvtl as o_cnn
    H, W = hw
    rng = np.random.default_rng(H + n)
    x = rng.integers(0, 256, (n, H, W, 3)).astype(np.uint8)
    if dtype == "float64":
        x = x.astype(np.float64) + rng.uniform(0, 0.5, x.shape)     # non-integer pixels exercise the lo plane
    params = o_cnn.make_weights(3)
    net = CnnVtl(input_shape=[n, H, W, 3], batch_size=n, weights=params, keep_cols=[0])
    want = o_cnn.conv_outputs(x.astype(np.float64), params)
    got = net.conv_outputs(torch.from_numpy(x).cuda())
    for l, (g, w) in enumerate(zip(got, want)):
        g = g.cpu().numpy()
        assert g.shape == w.shape
        # activations reach several hundred (0..255 inputs, He-scaled filters): tolerance relative to the layer scale
        scale = max(1.0, float(np.abs(w).max()))
        err = float(np.max(np.abs(g - w))) / scale
        print("conv%d %s n=%d %s: max|err|/max|act| = %.2e, normwise %.2e" % (l + 1, hw, n, dtype, err, norm_err(g, w)))
        assert err <= 1e-4 and norm_err(g, w) <= 1e-5


@pytest.mark.parametrize("precision", ["fp16x2", "fp16"])
def test_cnnvtl_fused_equals_explicit(cuda, precision):
    """The fused head (implicit GEMM, min/max + kept-column gather in the epilogue) produces the same int8
    descriptors as the explicit-im2col formulation built from the library's building blocks."""
    from deeploopcloser_b200.cnn_vtl import CnnVtl
    from oracle import cnnvtl as o_cnn
    n, H, W = 4, 96, 128
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.integers(0, 256, (n, H, W, 3)).astype(np.uint8)).cuda()
    params = o_cnn.make_weights(7)
    keep = o_cnn.make_keep_columns(o_cnn.layer_sizes((H, W)), compress_factor=98.0, seed=1)
    net = CnnVtl(input_shape=[n, H, W, 3], batch_size=n, weights=params, keep_cols=keep, precision=precision)
    fused = net._forward_chunk(x).cpu().numpy()
    explicit = net._forward_chunk_explicit(x).cpu().numpy()
    ndiff = int((fused != explicit).sum())
    print("fused vs explicit (%s): %d of %d bytes differ" % (precision, ndiff, fused.size))
    assert fused.shape == explicit.shape == (n, keep.size)
    d = (fused.astype(np.int16) - explicit.astype(np.int16) + 128) % 256 - 128      # wrap-aware
    assert np.all(np.abs(d) <= 1)
    # fp16x2: identical up to the last bit of a float32 (pooling on hi+lo vs on float32). fp16: the two formulations
    # order K differently (conv1 runs on the space-to-depth image), so single-product rounding noise flips a few
    # values that sit on an integer boundary of the 0..255 scale.
    assert ndiff <= (0.001 if precision == "fp16x2" else 0.01) * fused.size


def test_cnnvtl_handle_errors(cuda):
    from deeploopcloser_b200 import _lib, ops
    head = ops.CnnVtlHead(64, 64)
    x = torch.zeros((1, 64, 64, 3), dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.DlcError):          # no weights yet
        head.forward(x)
    with pytest.raises(_lib.DlcError):          # columns must be strictly increasing and in range
        head.set_keep_cols([5, 5])
    with pytest.raises(_lib.DlcError):
        head.set_keep_cols([head.descriptor_len])
    with pytest.raises(_lib.DlcError):          # too small for conv1 + two pools
        ops.CnnVtlHead(8, 8)


@pytest.mark.parametrize("n,n_parts", [(21, 2), (21, 3), (70, 8), (5, 4)])
def test_similarity_parts_sum_to_the_whole(cuda, n, n_parts):
    """dlc_sdav_similarity_part: the parts of the score matrix (interleaved tile rows; what each GPU of a box
    evaluates when ONE sequence is split over the GPUs) are disjoint and sum to the full matrix bit for bit."""
    from deeploopcloser_b200 import ops
    rng = np.random.default_rng(n)
    desc = torch.from_numpy(rng.uniform(0, 1, (n, 30, 300)).astype(np.float32)).cuda()
    whole = ops.sdav_similarity(desc)
    parts = [ops.sdav_similarity_part(desc, p, n_parts).clone() for p in range(n_parts)]
    nz = torch.stack([(p != 0) for p in parts]).sum(0)
    assert int(nz.max()) <= 1                                   # every entry is owned by at most one part
    assert torch.equal(torch.stack(parts).sum(0), whole)


def test_sharded_sequence_pipeline_single_rank(cuda):
    """ShardedSequencePipeline with one rank is the plain pipeline (the multi-rank exchange is covered by
    tools/check_sharded_sequence.py under torchrun and by the gloo test of the host logic)."""
    from deeploopcloser_b200.pipeline import LoopClosurePipeline, ShardedSequencePipeline
    rng = np.random.default_rng(2)
    dims = [1681, 128, 64]
    ws, bs = o_sda.make_weights(dims, seed=1, scale="xavier")
    frames = torch.from_numpy(rng.integers(0, 256, (9, 120, 160), dtype=np.uint8)).cuda()
    xy = torch.from_numpy(np.stack([rng.uniform(0, 160, (9, 30)), rng.uniform(0, 120, (9, 30))], -1).astype(np.float32)).cuda()
    a, b = LoopClosurePipeline(dims), ShardedSequencePipeline(dims)
    a.set_weights(ws, bs)
    b.set_weights(ws, bs)
    ra, rb = a.run(frames, xy, k=3), b.run(frames, xy, k=3)
    assert torch.equal(ra["similarity"], rb["similarity"]) and torch.equal(ra["candidates"][1], rb["candidates"][1])
    assert ShardedSequencePipeline.frame_block(1063, 7, 8) == (931, 1063, 133)
    assert ShardedSequencePipeline.frame_block(5, 3, 4) == (5, 5, 2)        # more ranks than needed: empty block


@pytest.mark.parametrize("m,k,n", [(3, 2, 2), (130, 200, 96), (257, 2500, 300), (700, 1681, 2500), (1000, 70, 520)])
@pytest.mark.parametrize("bk", [32, 64])
def test_cta_pair_kernel_vs_oracle(cuda, m, k, n, bk):
    """The cta_group::2 kernel (two CTAs share one 256-row MMA, each staging half of the B tile) forced onto small and
    ragged shapes: odd numbers of 128-row tiles (the peer CTA's tile is entirely out of range), N tiles of 96 / 64 /
    256 columns, K not a multiple of the block."""
    from deeploopcloser_b200 import _lib, ops
    rng = np.random.default_rng(m + k + n)
    a = rng.uniform(0, 1, (m, k))
    b = rng.standard_normal((k, n))
    bias = rng.standard_normal(n)
    ref = 1.0 / (1.0 + np.exp(-(a @ b + bias)))
    args = (torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(bias).cuda())
    _lib.call("dlc_debug_set", 0, bk)
    try:
        _lib.call("dlc_debug_set", 6, 0)
        single = ops.matmul(*args, act="sigmoid", precision="fp16x2").cpu().numpy()
        _lib.call("dlc_debug_set", 6, 2)
        out = ops.matmul(*args, act="sigmoid", precision="fp16x2").cpu().numpy()
    finally:
        _lib.call("dlc_debug_set", 6, 1)
        _lib.call("dlc_debug_set", 0, 32)
    assert np.array_equal(out, single)        # same K order and accumulation chunks as the single-CTA kernel
    err = float(np.max(np.abs(out - ref)))
    print("pair kernel", (m, k, n), "bk", bk, "max abs err", err)
    assert err < 5e-5          # float32 accumulation of a sigmoid argument of magnitude up to ~50 * 30


def _near_tie_frames(seed, n=12, p=30, d=1024):
    rng = np.random.default_rng(seed)
    desc = rng.uniform(0.05, 0.95, (n, p, d)).astype(np.float32)
    for f in range(1, n):           # frame f holds near-copies of patches of frame 0, in pairs 1e-4 apart
        for k in range(0, p, 2):
            base = desc[0, (k + f) % p]
            desc[f, k] = base + 3e-3 * rng.standard_normal(d).astype(np.float32)
            desc[f, k + 1] = desc[f, k] + 2e-5 * rng.standard_normal(d).astype(np.float32)
    return np.clip(desc, 0, 1)


@pytest.mark.parametrize("pair_mode", [0, 2])
def test_similarity_deferred_refinement_equals_in_place(cuda, pair_mode):
    """fp16r: ambiguous frame pairs are handed to a second kernel through a work list; a full list falls back to the
    re-evaluation inside the Gram epilogue. All three routes (deferred, in place, list of 3 entries) must give the
    same bits and the same flagged-row count, on single CTAs and on CTA pairs."""
    from deeploopcloser_b200 import _lib, ops
    desc = torch.from_numpy(_near_tie_frames(7)).cuda()
    n, p, d = desc.shape
    out, flagged = {}, {}
    try:
        _lib.call("dlc_debug_set", 6, pair_mode)
        for cap in (-1, 0, 3):
            _lib.call("dlc_debug_set", 8, cap)
            out[cap] = ops.sdav_similarity(desc, precision="fp16r").clone()
            flagged[cap] = ops.sdav_similarity_stats(n, p, d)["flagged_rows"]
    finally:
        _lib.call("dlc_debug_set", 8, -1)
        _lib.call("dlc_debug_set", 6, 1)
    assert flagged[-1] > 3 and flagged[0] == flagged[-1] and flagged[3] == flagged[-1]
    assert torch.equal(out[0], out[-1]) and torch.equal(out[3], out[-1])
    _check_similarity(out[-1].cpu().numpy(), desc.cpu().numpy(), False)


@pytest.mark.parametrize("precision", ["fp16", "fp16r"])
def test_similarity_one_product_cta_pair_equals_single(cuda, precision):
    from deeploopcloser_b200 import _lib, ops
    rng = np.random.default_rng(11)
    desc = torch.from_numpy(rng.uniform(0, 1, (37, 30, 300)).astype(np.float32)).cuda()
    try:
        _lib.call("dlc_debug_set", 6, 0)
        single = ops.sdav_similarity(desc, precision=precision).clone()
        _lib.call("dlc_debug_set", 6, 2)
        pair = ops.sdav_similarity(desc, precision=precision).clone()
    finally:
        _lib.call("dlc_debug_set", 6, 1)
    assert torch.equal(pair, single)


@pytest.mark.parametrize("n,p,d", [(9, 30, 2500), (21, 30, 300), (2, 30, 64), (33, 6, 128), (17, 5, 96)])
def test_similarity_cta_pair_equals_single(cuda, n, p, d):
    """The SDAV Gram kernel on CTA pairs (each CTA stages half of the 240-column N tile) gives bit-identical scores
    to the single-CTA kernel, including odd numbers of M tiles and the 32-stride fallback (odd P)."""
    from deeploopcloser_b200 import _lib, ops
    rng = np.random.default_rng(n * 7 + p)
    desc = torch.from_numpy(rng.uniform(0, 1, (n, p, d)).astype(np.float32)).cuda()
    try:
        _lib.call("dlc_debug_set", 6, 0)
        single = ops.sdav_similarity(desc).clone()
        _lib.call("dlc_debug_set", 6, 2)
        pair = ops.sdav_similarity(desc).clone()
        parts = torch.stack([ops.sdav_similarity_part(desc, q, 3).clone() for q in range(3)]).sum(0)
    finally:
        _lib.call("dlc_debug_set", 6, 1)
    assert torch.equal(pair, single) and torch.equal(parts, single)


# ------------------------------------------------------------------------------------------- raw-pixel encoder input
def test_patch_gather_u8_golden(cuda, golden_dir):
    """dlc_patch_gather_u8: the plane holds the pixel VALUE of the reference's patches, exactly."""
    from deeploopcloser_b200 import ops
    g = np.load(golden_dir + "/patches.npz")
    out = ops.patch_gather_u8(torch.from_numpy(g["img"]).cuda(), torch.from_numpy(g["xy"]).cuda()).double().cpu().numpy()
    assert out.shape == (60, 1728)
    assert np.array_equal(out[:, :1681], np.rint(g["out"].reshape(60, 1681) * 255.0))
    assert np.all(out[:, 1681:] == 0)


@pytest.mark.parametrize("dims,frames,pair_mode", [([1681, 256, 128], 3, 1), ([1681, 2500, 300], 9, 2),
                                                    ([1681, 256, 128], 10, 2), ([1681, 128], 2, 1)])
def test_encoder_raw_pixel_mode(cuda, dims, frames, pair_mode):
    """SdaEncoder(input_u8=True): sigmoid((p / 255) W + b) evaluated as sigmoid((p (W 256/255)) / 256 + b) with the
    pixel plane exact in fp16 and TWO tensor-core products in layer 0. Same tolerance as the split-input path (1e-3 on
    the reference's saturating N(0,1) weights) and at least as accurate; single-CTA and CTA-pair kernels."""
    from deeploopcloser_b200 import _lib, ops
    rng = np.random.default_rng(len(dims) * 100 + frames)
    H, W, P = 96, 128, 30
    img = rng.integers(0, 256, (frames, H, W), dtype=np.uint8)
    xy = np.stack([rng.uniform(0, W, (frames, P)), rng.uniform(0, H, (frames, P))], -1).astype(np.float32)
    ws, bs = o_sda.make_weights(dims, seed=3, scale="normal")
    bs = [b + 0.1 * rng.standard_normal(b.shape) for b in bs]
    x = np.concatenate([o_patch.extract_patches(img[i], xy[i]) for i in range(frames)])
    ref = o_sda.sda_forward(x, ws, bs)
    img_d, xy_d = torch.from_numpy(img).cuda(), torch.from_numpy(xy).cuda()
    errs = {}
    try:
        _lib.call("dlc_debug_set", 6, pair_mode)
        for raw in (True, False):
            enc = ops.SdaEncoder(dims, "fp16x2", input_u8=raw)
            for l, (w, b) in enumerate(zip(ws, bs)):
                enc.set_layer(l, w, b)
            if raw:
                out = enc.encode_planes(ops.patch_gather_u8(img_d, xy_d), None, frames * P)
            else:
                hi, lo = ops.patch_gather(img_d, xy_d)
                out = enc.encode_planes(hi, lo, frames * P)
            errs[raw] = float(np.max(np.abs(out.cpu().numpy() - ref)))
            enc.close()
    finally:
        _lib.call("dlc_debug_set", 6, 1)
    print("raw-pixel encoder", dims, "max abs err raw / split input:", errs[True], errs[False])
    assert errs[True] <= TOL and errs[True] <= 2.0 * errs[False] + 1e-6


@pytest.mark.parametrize("precision", ["fp16x2", "fp16"])
def test_encoder_wave_aware_width_gives_the_same_bits(cuda, precision):
    """dlc_sda_encode picks the GEMM width per call from the row count (sda.cu gemm_pad: 3 990 rows = one rank's
    block of config 2 split over 8 GPUs and 7 680 rows = a streaming batch of 256 frames use 224-wide accumulators,
    the full sequence 256-wide ones). The tile shape must not change a single bit of a row's descriptor."""
    from deeploopcloser_b200 import ops
    dims = [1681, 2500, 2500, 2500]
    ws, bs = o_sda.make_weights(dims, seed=2, scale="xavier")
    rng = np.random.default_rng(5)
    rows = 31890
    x = torch.from_numpy(rng.integers(0, 256, (rows, 1728)).astype(np.float16))
    x[:, 1681:] = 0
    x = x.cuda()
    enc = ops.SdaEncoder(dims, precision, input_u8=True)
    for l, (w, b) in enumerate(zip(ws, bs)):
        enc.set_layer(l, w, b)
    full = enc.encode_planes(x, None, rows).clone()
    for r in (3990, 7680, 12000):
        part = enc.encode_planes(x[:r].contiguous(), None, r)
        assert torch.equal(part, full[:r]), r
    ref = o_sda.sda_forward(x[:60, :1681].double().cpu().numpy() / 255.0, ws, bs)
    assert rel_err(full[:60].cpu().numpy(), ref) <= (TOL if precision == "fp16x2" else 5e-3)
    enc.close()


def test_sm_reserve_leaves_results_unchanged(cuda):
    """dlc_set_sm_reserve(n): the persistent kernels size their grids for (SM count - n) - fewer CTAs walk the same
    tiles, so encoder descriptors and similarity scores keep their bits."""
    from deeploopcloser_b200 import _lib, ops
    dims = [1681, 2500, 2500]
    ws, bs = o_sda.make_weights(dims, seed=6, scale="xavier")
    rng = np.random.default_rng(8)
    rows = 40 * 30
    x = torch.from_numpy(rng.integers(0, 256, (rows, 1728)).astype(np.float16))
    x[:, 1681:] = 0
    x = x.cuda()
    enc = ops.SdaEncoder(dims, "fp16x2", input_u8=True)
    for l, (w, b) in enumerate(zip(ws, bs)):
        enc.set_layer(l, w, b)
    want = enc.encode_planes(x, None, rows).clone()
    s_want = ops.sdav_similarity(want.view(40, 30, -1), precision="fp16r").clone()
    try:
        for n in (8, 24, 64):
            _lib.call("dlc_set_sm_reserve", n)
            assert _lib.call("dlc_sm_count") >= 2
            got = enc.encode_planes(x, None, rows)
            assert torch.equal(got, want), n
            assert torch.equal(ops.sdav_similarity(got.view(40, 30, -1), precision="fp16r"), s_want), n
    finally:
        _lib.call("dlc_set_sm_reserve", 0)
    with pytest.raises(RuntimeError):
        _lib.call("dlc_set_sm_reserve", 65)
    enc.close()


def test_encoder_raw_pixel_mode_needs_layer0_again(cuda):
    from deeploopcloser_b200 import _lib, ops
    enc = ops.SdaEncoder([1681, 64], "fp16x2")
    enc.set_layer(0, np.zeros((1681, 64)), np.zeros(64))
    _lib.call("dlc_sda_set_input_u8", enc._h, 1)          # layer 0 is packed differently: must be set again
    x = torch.zeros((30, 1728), dtype=torch.float16, device="cuda")
    with pytest.raises(RuntimeError):
        enc.encode_planes(x, None, 30)
    enc.close()


# ------------------------------------------------------------------------------------------- keypoint detector (8f rank 3)
def _surf_frames(kind, B, H, W, seed):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (B, H, W), dtype=np.uint8)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    out = []
    for _ in range(B):
        img = np.full((H, W), 90.0) + rng.normal(0, 2.0, (H, W))
        for _ in range(14):
            s = rng.uniform(2.5, 9.0)
            img += rng.choice([-1, 1]) * rng.uniform(40, 140) * np.exp(
                -((xx - rng.uniform(0, W)) ** 2 + (yy - rng.uniform(0, H)) ** 2) / (2 * s * s))
        out.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
    return np.stack(out)


@pytest.mark.parametrize("kind,B,H,W", [("blobs", 3, 192, 240), ("noise", 2, 192, 240), ("blobs", 2, 101, 135),
                                        ("noise", 1, 50, 61), ("blobs", 1, 480, 640)])
@pytest.mark.parametrize("fast", [1, 0])
def test_surf_detect_equals_oracle(cuda, kind, B, H, W, fast):
    """dlc_surf_detect == oracle/surf.py bit for bit: positions, sizes, responses, order and the keypoint count, for
    blob and noise frames, sizes that are not multiples of the sampling steps and frames too small for the upper
    octaves. fast = 1: octaves 0 and 1 run the shared-memory kernel with compile-time box tables (the default);
    fast = 0: every octave runs the generic kernel."""
    from deeploopcloser_b200 import _lib, ops
    from oracle import surf
    frames = _surf_frames(kind, B, H, W, seed=H + B)
    n = 30
    _lib.call("dlc_debug_set", 10, fast)
    try:
        xy, info, found = ops.surf_detect(torch.from_numpy(frames).cuda(), top_n=n)
        xy, info, found = xy.cpu().numpy(), info.cpu().numpy(), found.cpu().numpy()
    finally:
        _lib.call("dlc_debug_set", 10, 1)
    for b in range(B):
        all_kp = surf.detect(frames[b])
        ref = surf.top_n(all_kp, n)
        assert found[b] == len(all_kp) <= ops.SURF_CANDIDATE_CAP, (found[b], len(all_kp))
        m = len(ref)
        assert np.array_equal(xy[b, :m], ref[:, 0:2].astype(np.float32))
        assert np.array_equal(info[b, :m], ref[:, 2:4].astype(np.float32))
        assert np.all(xy[b, m:] == [0.5 * (W - 1), 0.5 * (H - 1)]) and np.all(info[b, m:] == 0)
        print("surf", kind, (H, W), "keypoints", len(all_kp), "top response", ref[0, 3] if m else None)


def test_surf_feeds_the_parser(cuda):
    """CvInputParser.parse(image) without injected keypoints: GPU detector -> GPU patch gather == the oracle chain."""
    from oracle import surf
    from src.sdav.input.CvInputParser import CvInputParser, get_top_n_key_points
    img = _surf_frames("blobs", 1, 192, 240, seed=9)[0]
    kps = get_top_n_key_points(img, 30)
    ref = surf.top_n(surf.detect(img), 30)
    assert len(kps) == len(ref) > 0
    assert np.array_equal(np.array([kp.pt for kp in kps], dtype=np.float32), ref[:, :2].astype(np.float32))
    assert all(kps[i].response >= kps[i + 1].response for i in range(len(kps) - 1))
    out = CvInputParser(30, 41).parse(img)
    want = o_patch.extract_patches(img, ref[:, :2].astype(np.float32))
    assert np.array_equal(out, want)
