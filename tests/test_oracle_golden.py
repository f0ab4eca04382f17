"""CPU: the oracle against the committed golden vectors (made by running the reference's own code, see
tests/golden/make_golden.py) and against its own algebraic restatements."""
import warnings

import numpy as np
import pytest

from oracle import cnnvtl as o_cnn
from oracle import hamming as o_ham
from oracle import matcher as o_match
from oracle import patches as o_patch
from oracle import sda as o_sda
from oracle import similarity as o_sim


def test_patches_match_reference(golden_dir):
    g = np.load(golden_dir + "/patches.npz")
    for i in range(len(g["img"])):
        assert np.array_equal(o_patch.extract_patches(g["img"][i], g["xy"][i], 41, True), g["out"][i])


def test_patch_quirk_documented_case():
    """keypoint (x=230.5, y=10.5) on a 192x240 image reads rows 151..191, cols 0..40."""
    img = np.arange(192 * 240, dtype=np.int64).reshape(192, 240) % 251
    out = o_patch.extract_patches(img.astype(np.uint8), np.array([[230.5, 10.5]]), 41, True)
    want = img[151:192, 0:41].astype(np.uint8).reshape(-1) / 255.0
    assert np.array_equal(out[0], want)
    lo, hi = o_patch.window_bounds(192, [0, 20, 21, 100, 171, 172, 191, 230], 41)
    assert list(lo) == [0, 0, 1, 80, 151, 151, 151, 151] and list(hi - lo) == [40] * 8


def test_similarity_matches_reference(golden_dir):
    g = np.load(golden_dir + "/similarity.npz")
    for name in "abc":
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            S = o_sim.similarity_matrix(g["desc_" + name], full_asymmetric=True)
        ref = g["S_" + name]
        m = ~np.eye(len(ref), dtype=bool)
        assert np.max(np.abs(S[m] - ref[m]) / np.maximum(1, np.abs(ref[m]))) < 1e-12


def test_similarity_gram_restatement():
    """The CUDA kernel's algebra: argmin_j (n_j - 2 G_kj) and s_k = |p_ik - p_jj*| reproduce similarity_score."""
    rng = np.random.default_rng(3)
    d = 1.0 / (1.0 + np.exp(-3 * rng.standard_normal((4, 30, 120))))
    w = o_sim.distinctive_weights(d)
    p = d @ w
    for i in range(4):
        for j in range(4):
            if i == j:
                continue
            G = d[i] @ d[j].T
            nj = (d[j] ** 2).sum(1)
            jstar = np.argmin(nj[None, :] - 2 * G, axis=1)
            s = np.abs(p[i] - p[j][jstar])
            want, idx, _ = o_sim.similarity_score(d[i], d[j], w, return_details=True)
            assert np.array_equal(jstar, idx)
            assert abs(np.sum(10 - 10 * np.log(s)) - want) < 1e-9 * abs(want)


def test_hamming_matches_reference(golden_dir):
    g = np.load(golden_dir + "/hamming.npz")
    assert np.array_equal(o_ham.distance_matrix(g["desc"]), g["D"])
    assert o_ham.distance(g["ex_y1"], g["ex_y2"]) == int(g["ex_d"])
    assert o_ham.SIGNED_LUT[0x80] == 1 and o_ham.SIGNED_LUT[0xFF] == 1 and o_ham.SIGNED_LUT[0x7F] == 7
    d = g["desc"]
    assert np.all(np.diag(o_ham.distance_matrix(d)) == 0)
    assert not np.array_equal(o_ham.distance_matrix(d, True), o_ham.distance_matrix(d, False))


def test_misc_golden(golden_dir):
    g = np.load(golden_dir + "/misc.npz")
    assert [o_cnn.compressed_size(int(v), 99.59) for v in g["vals"]] == list(g["compressed"])
    x, w = g["tw_x"], g["tw_w"]
    assert np.array_equal((x.reshape(-1, 2) @ w).reshape(3, 2, 2), g["tw_expected"])


def test_sda_oracle_shapes_and_saturation():
    dims = [1681, 64, 32]
    ws, bs = o_sda.make_weights(dims, seed=0)
    x = np.random.default_rng(0).uniform(0, 1, (3, 30, 1681))
    out = o_sda.sda_forward(x, ws, bs)
    assert out.shape == (90, 32) and out.min() >= 0 and out.max() <= 1
    assert np.allclose(o_sda.sigmoid(np.array([-800.0, 0.0, 800.0])), [0.0, 0.5, 1.0])


def test_cnnvtl_oracle_geometry():
    sizes = o_cnn.layer_sizes((192, 240))
    assert sizes == [256128, 157696, 49920, 49920, 33280] and sum(sizes) == 546944
    assert o_cnn.layer_sizes((224, 224)) == [279936, 173056, 55296, 55296, 36864]
    keep = o_cnn.make_keep_columns(sizes, seed=4)
    assert len(keep) <= 2243 and np.all(np.diff(keep) > 0)
    assert list(o_cnn.cast_int8_wrap(np.array([0.0, 127.9, 128.0, 255.0, 200.7]))) == [0, 127, -128, -1, -56]


def test_matcher_oracle_ties_and_padding():
    s = np.array([[1.0, 3.0, 3.0, 2.0], [np.nan, 0.5, 0.5, 0.5]])
    ts, ti = o_match.topk(s, 3)
    assert ti.tolist() == [[1, 2, 3], [1, 2, 3]]
    ts, ti = o_match.topk(s, 3, largest=False, exclude_band=0)
    assert ti.tolist() == [[3, 1, 2], [2, 3, -1]]
    c, ts, ti = o_match.threshold(s, 2.5, 3)
    assert c.tolist() == [2, 0] and ti.tolist() == [[1, 2, -1], [-1, -1, -1]]


# ------------------------------------------------------------------------------------------- independent cross-checks
def test_cnnvtl_oracle_matches_torch_conv2d():
    """The conv-head oracle (NumPy sliding windows, TF 'SAME' = extra pixel after) against an independent float64
    implementation: torch.nn.functional.conv2d / max_pool2d on CPU with the same explicit padding."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 255, (2, 43, 51, 3))
    params = o_cnn.make_weights(5)
    outs = o_cnn.conv_outputs(x, params)
    h = torch.from_numpy(x).permute(0, 3, 1, 2)
    for (name, kh, kw, cin, cout, stride, padding, relu), want in zip(o_cnn.LAYERS, outs):
        w = torch.from_numpy(params[name][0]).permute(3, 2, 0, 1)            # HWIO -> OIHW
        b = torch.from_numpy(params[name][1])
        if padding == "same":
            ph = max((-(-h.shape[2] // stride) - 1) * stride + kh - h.shape[2], 0)
            pw = max((-(-h.shape[3] // stride) - 1) * stride + kw - h.shape[3], 0)
            h = F.pad(h, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        h = F.conv2d(h, w, b, stride=stride)
        if relu:
            h = torch.relu(h)
        got = h.permute(0, 2, 3, 1).numpy()
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) <= 1e-9 * max(1.0, np.abs(want).max()), name
        if name in ("conv1", "conv2"):
            h = F.max_pool2d(h, 3, 2)


def test_sda_oracle_matches_torch():
    import torch
    from oracle import sda as o_sda
    dims = [37, 29, 23]
    ws, bs = o_sda.make_weights(dims, seed=4, scale="normal")
    x = np.random.default_rng(1).uniform(0, 1, (11, dims[0]))
    h = torch.from_numpy(x)
    for w, b in zip(ws, bs):
        h = torch.sigmoid(h @ torch.from_numpy(w) + torch.from_numpy(b))
    assert np.max(np.abs(h.numpy() - o_sda.sda_forward(x, ws, bs))) <= 1e-14


def test_image_oracle_matches_reference_lines_and_cv2(golden_dir):
    """oracle/images.py vs fixtures made by executing create_similarity_matrix.py:41-45 / create_distance_matrix.py:40
    verbatim and round-tripping through the real cv2.imwrite (tests/golden/make_golden.py::golden_images)."""
    from oracle import images as o_img
    g = np.load(golden_dir + "/images.npz")
    for name in "abr":
        m = o_img.to_reference_int(g["sim_scores_" + name])
        assert np.array_equal(m, g["sim_int_" + name])
        f = o_img.similarity_image_f64(m)
        assert np.array_equal(f, g["sim_img_f64_" + name])
        assert np.array_equal(o_img.imwrite_u8(f), g["sim_png_" + name])
    f = o_img.distance_image_f64(g["dist_D"])
    assert np.array_equal(f, g["dist_img_f64"])
    assert np.array_equal(o_img.imwrite_u8(f), g["dist_png"])
    assert np.array_equal(o_img.imwrite_u8(np.array([0.5, 1.5, 2.5, -3.0, 300.0, np.nan])), [0, 2, 2, 0, 255, 0])


def test_write_png_round_trip(tmp_path):
    """The pure-zlib PNG writer produces a file OpenCV reads back bit for bit (skipped without cv2)."""
    cv2 = pytest.importorskip("cv2")
    from deeploopcloser_b200.similarity import write_png
    img = np.random.default_rng(0).integers(0, 256, (37, 53), dtype=np.uint8)
    path = str(tmp_path / "m.png")
    write_png(path, img)
    assert np.array_equal(cv2.imread(path, cv2.IMREAD_GRAYSCALE), img)


def _blob_image(H, W, blobs, base=60.0):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.full((H, W), base)
    for bx, by, s, a in blobs:
        img += a * np.exp(-((xx - bx) ** 2 + (yy - by) ** 2) / (2 * s * s))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def test_surf_oracle_recovers_synthetic_blobs():
    """The detector oracle cannot be pinned against OpenCV's non-free SURF (absent here); what can be checked is the
    geometry it must satisfy: Gaussian blobs are found at their centres (x = column, y = row, sub-pixel), with a size
    that grows with the blob, strongest response first, bright and dark blobs alike, and nothing in a flat image."""
    from oracle import surf
    blobs = [(60.0, 50.0, 3.0, 150), (150.5, 100.25, 5.0, 120), (200.0, 150.0, 8.0, -50)]
    kp = surf.top_n(surf.detect(_blob_image(192, 240, blobs)), 3)
    assert len(kp) == 3 and np.all(np.diff(kp[:, 3]) <= 0)
    for (bx, by, s, _), row in zip(blobs, kp):
        assert abs(row[0] - bx) < 0.1 and abs(row[1] - by) < 0.1
        assert 5.0 * s <= row[2] <= 6.5 * s
    assert len(surf.detect(np.full((96, 128), 77, dtype=np.uint8))) == 0
    # a blob smaller than the first middle layer (15 x 15 filter) has no scale-space maximum there
    assert len(surf.detect(_blob_image(96, 128, [(64.0, 48.0, 2.0, 180)]))) == 0
    # layer geometry: filter sizes of SURF_create() defaults
    assert surf.layer_sizes() == [[9, 15, 21, 27, 33], [18, 30, 42, 54, 66], [36, 60, 84, 108, 132],
                                  [72, 120, 168, 216, 264]]


# ---------------------------------------------------------------------------------------------- config 1 (parity config)
def _config1_oracle(g, tag):
    """pixels -> patches -> float64 encoder for one weight set of the config-1 fixture."""
    seed, scale = {"normal": (1, "normal"), "xavier": (2, "xavier")}[tag]
    ws, bs = o_sda.make_weights([1681, 2500, 2500, 2500, 2500, 2500], seed=seed, scale=scale)
    x = np.stack([o_patch.extract_patches(g["frames"][i], g["xy"][i]) for i in range(len(g["frames"]))])
    return x, ws, bs, o_sda.sda_forward(x, ws, bs).reshape(len(x), 30, -1)


def test_config1_fixture_pins_the_oracle_end_to_end(golden_dir):
    """The 20 real datasets/test frames: oracle patches == the reference's own patch function (digest), and the oracle
    scores / matches on the oracle descriptors == the reference's own SimilarityCalculator (tests/golden/make_golden.py
    ::golden_config1), for every ordered pair, for N(0,1) and Xavier-scaled weights."""
    import hashlib
    g = np.load(golden_dir + "/config1_frames.npz")
    assert g["frames"].shape == (20, 192, 240) and g["frames"].dtype == np.uint8
    for tag in ("normal", "xavier"):
        x, ws, bs, desc = _config1_oracle(g, tag)
        assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == str(g["patches_sha256"])
        dig = np.array([desc.sum(), np.abs(desc).max(), desc[3, 7, 11], desc[19, 29, 2499]])
        assert np.allclose(dig, g["desc_digest_" + tag], rtol=1e-12, atol=0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            S, det = o_sim.similarity_matrix(desc, full_asymmetric=True, return_details=True)
        ref = g["S_" + tag]
        m = ~np.eye(20, dtype=bool)
        assert np.max(np.abs(S[m] - ref[m]) / np.maximum(1, np.abs(ref[m]))) < 1e-12
        for (i, j), (idx, _) in det.items():
            assert np.array_equal(idx, g["idx_" + tag][i, j])


def test_parity_report_classes(golden_dir):
    """The from-pixels report itself (tests/parity_report.py), driven on the CPU: a float32 evaluation of the encoder
    stands in for the device (its descriptor error is of the order the device's is), scores are the float64 scores of
    those descriptors. Everything must be explained; a corrupted score and a wrong match must not be."""
    import parity_report as pr   # tests/ is on sys.path (rootdir conftest)
    g = np.load(golden_dir + "/config1_frames.npz")
    x, ws, bs, desc = _config1_oracle(g, "normal")
    h = x.reshape(-1, x.shape[-1]).astype(np.float32)
    for w, b in zip(ws, bs):
        h = (1.0 / (1.0 + np.exp(-(h @ w.astype(np.float32) + b.astype(np.float32))))).astype(np.float32)
    dev = h.reshape(desc.shape).astype(np.float64)
    n = 8                                                   # sub-sequence: keeps the CPU suite short
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        S_dev = o_sim.similarity_matrix(dev[:n], full_asymmetric=True)
        S_ref = o_sim.similarity_matrix(desc[:n], full_asymmetric=True)
    rep = pr.report(S_dev, dev[:n], desc[:n])
    assert rep["pairs"] == n * (n - 1) and rep["unexplained"] == 0, rep
    assert rep["ok"] + rep["tie"] + rep["conditioning"] == rep["pairs"]
    assert rep["descriptor_max_rel_err"] < 1e-3
    # fixtures as the oracle side give the same classes
    rep2 = pr.report(S_dev, dev[:n], desc[:n], S_ref=S_ref)
    assert {k: rep2[k] for k in ("ok", "tie", "conditioning")} == {k: rep[k] for k in ("ok", "tie", "conditioning")}
    # a wrong score (2 % off) is caught ...
    bad = S_dev.copy()
    bad[1, 2] *= 1.02
    assert pr.report(bad, dev[:n], desc[:n])["unexplained"] == 1
    # ... and a device that matched a non-nearest patch in one row of one pair is, too
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        w_dev = o_sim.distinctive_weights(dev[:n])
        _, idx, _ = o_sim.similarity_score(dev[2], dev[5], w_dev, return_details=True)
        idx[4] = (idx[4] + 7) % 30
        wrong = S_dev.copy()
        wrong[2, 5] = np.sum(10.0 - 10.0 * np.log(np.abs((dev[2] - dev[5][idx]) @ w_dev)))
    rep3 = pr.report(wrong, dev[:n], desc[:n])
    assert rep3["unexplained"] == 1 and rep3["unexplained_detail"][0][0] == (2, 5)


def test_surf_compile_time_box_tables_match_the_oracle_patterns():
    """csrc/surf.cu surf_octave_fast_kernel hard-codes the box corners of the default pyramid as
    fast_cr(x, size) = (2 x size + 9) / 18 (integer division) and fast_margin = (size / 2) / step; dlc_surf_detect
    compares them with the runtime plan before using the kernel (surf_fast_ok). The same rule, restated here, must
    reproduce oracle/surf.py's float32 resize_pattern for every layer the fast kernel serves (octaves 0 and 1) - and,
    as it happens, for the whole default pyramid."""
    from oracle import surf

    def cr(x, size):
        return (2 * x * size + 9) // 18

    for octave, sizes in enumerate(surf.layer_sizes(4, 3)):
        for size in sizes:
            for name, src in (("dx", surf.DX), ("dy", surf.DY), ("dxy", surf.DXY)):
                want = [box[:4] for box in surf.resize_pattern(src, size)]
                got = [tuple(cr(v, size) for v in box[:4]) for box in src]
                assert got == want, (octave, size, name)
            lobes = [cr(x, size) for x in (0, 3, 6, 9)]
            assert lobes[0] == 0 and lobes[3] == size and len({b - a for a, b in zip(lobes, lobes[1:])}) == 1
    # region bookkeeping of the kernel: 34 halo samples, widest margin 16, plane dimension 67 for both steps
    for step in (1, 2):
        sizes = [(9 + 6 * l) * step for l in range(5)]
        margins = [(s // 2) // step for s in sizes]
        assert margins == [4, 7, 10, 13, 16]
        extent = max((33 - m + 16) * step + s for m, s in zip(margins, sizes)) + 1
        assert extent == 66 * step + 1 and -(-extent // step) == 67
