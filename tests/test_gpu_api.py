"""`-m gpu`: the reference-facing Python API (same import paths as the reference) running on the CUDA kernels,
checked against the oracle and the reference-made golden vectors."""
import warnings

import numpy as np
import pytest

from oracle import cnnvtl as o_cnn
from oracle import hamming as o_ham
from oracle import sda as o_sda
from oracle import similarity as o_sim

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel_err(a, b):
    return float(np.max(np.abs(np.asarray(a, float) - b) / np.maximum(1.0, np.abs(b))))


def test_example(cuda):
    """Port of the reference's only test (test/TensorflowWrapperTest.py:11-21), body kept line for line with
    `tf.Session` -> `tw.Session`."""
    import src.utils.TensorflowWrapper as tw
    x = tw.constant([[[1, 2], [3, 4]], [[5, 6], [7, 8]], [[9, 10], [11, 12]]])
    w = tw.constant([[2, 2], [2, 2]])
    expected = np.array([[[6, 6], [14, 14]], [[22, 22], [30, 30]], [[38, 38], [46, 46]]]).astype(np.float64)

    y = x.matmul(w.to_tf()).to_tf()

    with tw.Session() as sess:
        actual = sess.run(y)

    assert np.array_equal(expected, actual)
    assert actual.dtype == np.float64


def test_tensorwrapper_encoder_layer(cuda):
    """x.corrupt(0).matmul(W).add(b).sigmoid() - the exact expression of SDAV.py:129 - through placeholder + feed."""
    import src.utils.TensorflowWrapper as tw
    rng = np.random.default_rng(0)
    W, b = rng.standard_normal((96, 40)), rng.standard_normal(40)
    x0 = tw.placeholder(tw.float64, [None, 30, 96])
    level = tw.placeholder(tw.float64, [])
    h0 = x0.corrupt(level).matmul(tw.constant(W)).add(tw.constant(b)).sigmoid()
    x = rng.uniform(0, 1, (3, 30, 96))
    with tw.Session() as sess:
        out = sess.run(h0.to_tf(), feed_dict={x0.to_tf(): x, level.to_tf(): 0})
    assert out.shape == (3, 30, 40)
    assert rel_err(out, o_sda.sigmoid(x @ W + b)) < 1e-5
    with tw.Session(seed=1) as sess:  # corruption level 0.5 zeroes exactly half of each frame's mask
        m = sess.run(tw.random_mask(tw.constant([30, 96], dtype=tw.int32), tw.constant(0.5)).to_tf())
    assert m.shape == (30, 96) and m.sum() == 30 * 96 / 2


def test_sdav_transform(cuda):
    from src.sdav.network.SDAV import SDAV
    model = SDAV(seed=1)
    x = np.random.default_rng(0).integers(0, 256, (3, 30, 1681)).astype(np.float64) / 255.0
    out = model.transform(x)
    assert out.shape == (90, 2500) and out.dtype == np.float64          # flat, like SDAV.py:163
    ref = o_sda.sda_forward(x, model._weights, model._biases)
    err = rel_err(out, ref)
    print("SDAV.transform N(0,1) weights: max |a-b|/max(1,|b|) =", err)
    assert err <= TOL
    with pytest.raises(ValueError):
        model.transform(np.zeros((2, 30, 100)))


def test_da_and_sda_transform(cuda):
    from src.sdav.network.StackedDenoisingAutoencoderVariants import SDA
    sda = SDA([30, 200], [64, 48], seed=3)
    x = np.random.default_rng(1).uniform(0, 1, (30, 200))
    out = sda.transform(x)
    ref = o_sda.sda_forward(x, [l._w0 for l in sda._layers], [l._b0 for l in sda._layers])
    assert out.shape == (30, 48) and rel_err(out, ref) <= TOL


def test_cv_input_parser(cuda, golden_dir):
    from src.sdav.input.CvInputParser import CvInputParser, get_vectorized_patches_from_key_points
    g = np.load(golden_dir + "/patches.npz")
    parser = CvInputParser(30, 41)
    out = parser.parse(g["img"][0], key_points=g["xy"][0])
    assert out.dtype == np.float64 and np.array_equal(out, g["out"][0])       # bit-exact vs the reference

    class KP:
        def __init__(self, x, y):
            self.pt = (x, y)
    ints = get_vectorized_patches_from_key_points(g["img"][1], [KP(*p) for p in g["xy"][1]], 41)
    assert np.array_equal(ints, np.rint(g["out"][1] * 255).astype(int))


def test_similarity_calculator(cuda, golden_dir):
    from src.sdav.similarity.SimilarityCalculator import SimilarityCalculator
    g = np.load(golden_dir + "/similarity.npz")
    desc, ref = g["desc_a"], g["S_a"]
    calc = SimilarityCalculator(desc)
    assert abs(calc.similarity_score(desc[0], desc[1]) - ref[0, 1]) <= TOL * abs(ref[0, 1])
    assert abs(calc.similarity_score(desc[3], desc[2]) - ref[3, 2]) <= TOL * abs(ref[3, 2])
    S = calc.similarity_matrix()
    iu = np.triu_indices(len(ref), 1)
    assert np.max(np.abs(S[iu] - ref[iu]) / np.abs(ref[iu])) <= TOL and np.array_equal(S, S.T)
    Si = calc.similarity_matrix(reference_int=True)
    assert Si.dtype == np.int64 and Si[0, 1] == int(ref[0, 1])      # truncation toward zero (Appendix A.8)
    s, i = calc.loop_candidates(k=2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = o_sim.similarity_matrix(desc)
    np.fill_diagonal(want, -np.inf)
    assert np.array_equal(i[:, 0], want.argmax(1))


def test_distance_calculator(cuda, golden_dir):
    from src.cnn_vtl.similarity.DistanceCalculator import DistanceCalculator
    g = np.load(golden_dir + "/hamming.npz")
    assert DistanceCalculator.calculate_distance(g["ex_y1"], g["ex_y2"]) == int(g["ex_d"])   # basic_example.py recipe
    assert DistanceCalculator.calculate_distance(list(g["desc"][0]), list(g["desc"][1])) == g["D"][0, 1]
    assert np.array_equal(DistanceCalculator.distance_matrix(g["desc"]), g["D"])


@pytest.mark.parametrize("hw", [(192, 240), (67, 83)])
def test_cnnvtl_transform(cuda, hw):
    from src.cnn_vtl.network.cnn_vtl import CnnVtl
    H, W = hw
    n = 3
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, (n, H, W, 3)).astype(np.float64)
    params = o_cnn.make_weights(3)
    sizes = o_cnn.layer_sizes(hw)
    keep = o_cnn.make_keep_columns(sizes, compress_factor=99.0 if H < 100 else 99.59, seed=4)
    net = CnnVtl(input_shape=[n, H, W, 3], batch_size=2, weights=params, keep_cols=keep)
    net.DEVICE_CHUNK = 2                                  # two device passes (2 + 1 images)
    assert net.layer_sizes == sizes
    got = net.transform(x)
    outs = o_cnn.conv_outputs(x, params)
    want, scaled = o_cnn.descriptors_from_outputs(outs, keep)
    assert got.shape == want.shape and got.dtype == np.int8
    diff = (got.astype(np.int16) - want.astype(np.int16))
    diff = (diff + 128) % 256 - 128                      # wrap-aware difference
    mism = np.flatnonzero(diff.ravel())
    # int8 truncation flips by one step where the float64 value sits within float32 rounding of an integer
    frac = scaled.ravel()[mism] - np.floor(scaled.ravel()[mism])
    print("cnn_vtl %s: %d of %d descriptor bytes differ (all +-1 at an integer boundary)" % (hw, mism.size, diff.size))
    assert np.all(np.abs(diff) <= 1)
    assert mism.size <= max(2, 0.002 * diff.size)
    assert np.all(np.minimum(frac, 1 - frac) < 2e-3)
    # and the Hamming matrix of the GPU descriptors is exact for those descriptors
    from src.cnn_vtl.similarity.DistanceCalculator import DistanceCalculator
    assert np.array_equal(DistanceCalculator.distance_matrix(got), o_ham.distance_matrix(got))


def test_streaming_loop_closer(cuda):
    """Config-5 style streaming on one GPU: frames inserted by one step are found again (top-1, cosine ~ 1) when the
    same frames come back; the frame descriptor equals the oracle's mean of patch descriptors."""
    import torch
    from deeploopcloser_b200.streaming import StreamingLoopCloser
    from oracle import patches as o_patch
    dims = [1681, 96, 64]
    ws, bs = o_sda.make_weights(dims, seed=2, scale="normal")   # N(0,1): saturating, frame descriptors well separated
    sl = StreamingLoopCloser(capacity_per_rank=64, dims=dims, k=3)
    sl.set_weights(ws, bs)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (8, 120, 160), dtype=np.uint8)
    xy = np.stack([rng.uniform(0, 160, (8, 30)), rng.uniform(0, 120, (8, 30))], -1).astype(np.float32)
    f_d, x_d = torch.from_numpy(frames).cuda(), torch.from_numpy(xy).cuda()
    q = sl.frame_descriptors(f_d, x_d).cpu().numpy()
    x = np.concatenate([o_patch.extract_patches(frames[i], xy[i]) for i in range(8)])
    want = o_sda.sda_forward(x, ws, bs).reshape(8, 30, -1).mean(1)
    assert rel_err(q, want) <= TOL
    s0, i0 = sl.step(f_d, x_d)                       # empty database: all padding
    assert np.all(i0.cpu().numpy() == -1)
    s1, i1 = sl.step(f_d, x_d)                       # same frames again: each finds its own earlier copy
    assert np.array_equal(i1[:, 0].cpu().numpy(), np.arange(8))
    assert np.all(np.abs(s1[:, 0].cpu().numpy() - 1.0) < 2e-3)
    assert len(sl.db.local) == 16


def test_matrix_images_match_reference(cuda, golden_dir):
    """a6 / a10 tails on the device (dlc_matrix_image) vs fixtures made by executing the reference's normalisation
    lines verbatim + the real cv2.imwrite: pixels identical."""
    from deeploopcloser_b200.distance import distance_image
    from deeploopcloser_b200.similarity import similarity_image
    g = np.load(golden_dir + "/images.npz")
    for name in "abr":
        scores = g["sim_scores_" + name]
        # the product's scores are float32: use float32-representable inputs whose truncation is unambiguous
        s32 = scores.astype(np.float32)
        if np.array_equal(np.trunc(s32.astype(np.float64)), g["sim_int_" + name]):
            assert np.array_equal(similarity_image(s32), g["sim_png_" + name]), name
        assert np.array_equal(similarity_image(g["sim_int_" + name].astype(np.float32)), g["sim_png_" + name]), name
    assert np.array_equal(distance_image(g["dist_D"]), g["dist_png"])


def test_matrix_image_edge_cases(cuda):
    import torch

    from deeploopcloser_b200 import ops
    from oracle import images as o_img
    # +inf (a matched pair of identical patches) saturates and leaves the range of the finite scores alone
    s = np.array([[-1, -400.7, np.inf], [-400.7, -1, -20.2], [np.inf, -20.2, -1]], dtype=np.float32)
    img = ops.matrix_image(torch.from_numpy(s).cuda(), ops.IMG_SIMILARITY, truncate_int=True).cpu().numpy()
    finite = np.where(np.isfinite(s), s, -1).astype(np.float64)
    ref = o_img.imwrite_u8(o_img.similarity_image_f64(o_img.to_reference_int(finite)))
    assert img[0, 2] == 255 and img[2, 0] == 255
    m = np.isfinite(s)
    assert np.array_equal(img[m], ref[m])
    # without truncation, constant matrix (0 / 0 -> black), empty matrix
    r = np.random.default_rng(1).normal(0, 50, (33, 65)).astype(np.float32)
    img = ops.matrix_image(torch.from_numpy(r).cuda(), ops.IMG_SIMILARITY).cpu().numpy()
    assert np.array_equal(img, o_img.imwrite_u8(o_img.similarity_image_f64(r.astype(np.float64))))
    c = torch.full((4, 4), 7.0, device="cuda")
    assert int(ops.matrix_image(c, ops.IMG_SIMILARITY).max()) == 0
    assert ops.matrix_image(torch.empty((0, 5), dtype=torch.int32, device="cuda"), ops.IMG_DISTANCE).shape == (0, 5)
    d = np.random.default_rng(2).integers(0, 9000, (130, 130)).astype(np.int32)
    img = ops.matrix_image(torch.from_numpy(d).cuda(), ops.IMG_DISTANCE).cpu().numpy()
    assert np.array_equal(img, o_img.imwrite_u8(o_img.distance_image_f64(d.astype(np.int64))))


def test_create_matrix_scripts(cuda, tmp_path):
    """The two script drop-ins end to end on synthetic frames: dataset directory in, PNG out, image == oracle."""
    cv2 = pytest.importorskip("cv2")
    from oracle import images as o_img
    from src.cnn_vtl import create_distance_matrix
    from src.sdav import create_similarity_matrix
    rng = np.random.default_rng(5)
    data = tmp_path / "frames"
    data.mkdir()
    for i in range(5):
        cv2.imwrite(str(data / ("f%03d.png" % i)), rng.integers(0, 256, (96, 128, 3), dtype=np.uint8))
    S = create_similarity_matrix.main([str(data), str(tmp_path / "s.png"), "--keypoints", "seeded"])
    assert S.dtype == np.int64 and S.shape == (5, 5) and np.all(np.diag(S) == -1) and np.array_equal(S, S.T)
    finite = np.isfinite(S.astype(np.float64)).all()
    png = cv2.imread(str(tmp_path / "s.png"), cv2.IMREAD_GRAYSCALE)
    if finite:
        assert np.array_equal(png, o_img.imwrite_u8(o_img.similarity_image_f64(S)))
    D = create_distance_matrix.main([str(data), str(tmp_path / "d.png")])
    assert D.shape == (5, 5) and np.all(np.diag(D) == 0)
    png = cv2.imread(str(tmp_path / "d.png"), cv2.IMREAD_GRAYSCALE)
    assert np.array_equal(png, o_img.imwrite_u8(o_img.distance_image_f64(D)))


def test_pipeline_detects_keypoints_when_none_given(cuda):
    """LoopClosurePipeline.run(frames) with no keypoints: fast-Hessian detector -> patch gather -> encoder -> score
    matrix, all on the device; equals the same pipeline fed the oracle detector's keypoints."""
    import torch

    from deeploopcloser_b200.pipeline import LoopClosurePipeline
    from oracle import surf
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[0:120, 0:160].astype(np.float64)
    frames = []
    for _ in range(5):
        img = np.full((120, 160), 100.0)
        for _ in range(60):
            s = rng.uniform(2.6, 7.0)
            img += rng.choice([-1, 1]) * rng.uniform(30, 120) * np.exp(
                -((xx - rng.uniform(0, 160)) ** 2 + (yy - rng.uniform(0, 120)) ** 2) / (2 * s * s))
        frames.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
    frames = np.stack(frames)
    dims = [1681, 128, 64]
    ws, bs = o_sda.make_weights(dims, seed=1, scale="xavier")
    pipe = LoopClosurePipeline(dims)
    pipe.set_weights(ws, bs)
    f_d = torch.from_numpy(frames).cuda()
    got = pipe.run(f_d, k=3)
    xy_ref = []
    for f in frames:
        kp = surf.top_n(surf.detect(f), 30)
        assert len(kp) == 30                       # enough structure in the synthetic frames
        xy_ref.append(kp[:, :2].astype(np.float32))
    want = pipe.run(f_d, torch.from_numpy(np.stack(xy_ref)).cuda(), k=3)
    assert torch.equal(got["descriptors"], want["descriptors"]) and torch.equal(got["similarity"], want["similarity"])
