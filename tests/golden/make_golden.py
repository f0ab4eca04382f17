"""Generates tests/golden/*.npz by running the REFERENCE's own importable code (authoring container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [/root/reference]

Imports (unmodified, from the read-only mount): src.sdav.similarity.SimilarityCalculator,
src.cnn_vtl.similarity.DistanceCalculator, src.sdav.input.CvInputParser (patch functions; SURF detection is not
available, keypoints are seeded), src.utils.MathUtils. The TensorFlow classes cannot be imported here.
The fixtures travel with the repo; nothing at test time reads /root/reference.
"""
import os
import sys
import warnings

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))


class FakeKeyPoint:  # the only attribute the reference reads is .pt (CvInputParser.py:111)
    def __init__(self, x, y):
        self.pt = (float(x), float(y))


def golden_patches():
    import cv2
    from src.sdav.input.CvInputParser import get_vectorized_patches_from_key_points
    rng = np.random.default_rng(0)
    imgs, kps, outs = [], [], []
    for f in ("outdoor_kennedylong_000000.ppm", "outdoor_kennedylong_000007.ppm"):
        img = cv2.imread(os.path.join(REF, "datasets/test", f), cv2.IMREAD_GRAYSCALE)
        h, w = img.shape
        xy = np.stack([rng.uniform(0, w, 30), rng.uniform(0, h, 30)], axis=1).astype(np.float32)
        # edge cases: corners, half-integers (banker's rounding), x beyond H (the x/y swap quirk clamps it)
        xy[:6] = [[0, 0], [w - 1, h - 1], [230.5, 10.5], [20.5, 21.5], [239.49, 0.5], [100.5, 191.0]]
        pts = [FakeKeyPoint(x, y) for x, y in xy]
        out = get_vectorized_patches_from_key_points(img, pts, 41) / 255.0   # parse(): CvInputParser.py:26-27
        imgs.append(img); kps.append(xy); outs.append(out)
    np.savez_compressed(os.path.join(HERE, "patches.npz"), img=np.stack(imgs), xy=np.stack(kps),
                        out=np.stack(outs).astype(np.float64))


def golden_similarity():
    from src.sdav.similarity.SimilarityCalculator import SimilarityCalculator
    rng = np.random.default_rng(1)
    cases = {}
    # (a) sigmoid-like descriptors, small D;  (b) saturated 0/1-ish descriptors like N(0,1)-weight SDA outputs
    for name, n, p, d, sat in (("a", 5, 30, 96, False), ("b", 4, 30, 200, True), ("c", 3, 7, 64, False)):
        x = rng.standard_normal((n, p, d)) * (8.0 if sat else 1.0)
        desc = (1.0 / (1.0 + np.exp(-x))).astype(np.float32).astype(np.float64)  # float32-representable
        calc = SimilarityCalculator(desc)
        S = np.full((n, n), np.nan)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(n):
                for j in range(n):
                    if i != j:
                        S[i, j] = calc.similarity_score(desc[i], desc[j])
        cases["desc_" + name] = desc
        cases["S_" + name] = S
    np.savez_compressed(os.path.join(HERE, "similarity.npz"), **cases)


def golden_hamming():
    from src.cnn_vtl.similarity.DistanceCalculator import DistanceCalculator
    rng = np.random.default_rng(2)
    desc = rng.integers(-128, 128, size=(9, 203)).astype(np.int8)
    desc[0, :4] = [-128, 127, -1, 0]
    desc[1, :4] = [127, -128, 0, -1]
    D = np.array([[DistanceCalculator.calculate_distance(a, b) for b in desc] for a in desc], dtype=np.int64)
    # basic_example.py recipe: change one element to 23
    y1 = desc[2].copy(); y2 = y1.copy(); y2[0] = 23
    ex = DistanceCalculator.calculate_distance(y1, y2)
    np.savez_compressed(os.path.join(HERE, "hamming.npz"), desc=desc, D=D, ex_y1=y1, ex_y2=y2, ex_d=np.int64(ex))


def golden_misc():
    from src.utils.MathUtils import MathUtils
    vals = np.array([256128, 157696, 49920, 49920, 33280, 290400, 186624, 64896, 43264, 1, 1000], dtype=np.int64)
    comp = np.array([MathUtils.compressed_size(int(v), 99.59) for v in vals], dtype=np.int64)
    # the reference's only test vector (test/TensorflowWrapperTest.py:12-14)
    x = np.array([[[1, 2], [3, 4]], [[5, 6], [7, 8]], [[9, 10], [11, 12]]], dtype=np.float64)
    w = np.array([[2, 2], [2, 2]], dtype=np.float64)
    expected = np.array([[[6, 6], [14, 14]], [[22, 22], [30, 30]], [[38, 38], [46, 46]]], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "misc.npz"), vals=vals, compressed=comp, tw_x=x, tw_w=w,
                        tw_expected=expected)


def _exec_reference_lines(path, first, last, env):
    """Runs lines first..last (1-based, inclusive) of a reference script - the scripts themselves cannot be imported
    (hard-coded paths, TensorFlow) - in `env`, so the fixture holds what the reference's own statements compute."""
    with open(os.path.join(REF, path)) as f:
        lines = f.readlines()[first - 1:last]
    exec(compile("".join(lines), path, "exec"), env)
    return env


def golden_images():
    """create_similarity_matrix.py:31,41-45 + :48 and create_distance_matrix.py:31,40-41: the reference's int64
    matrix, its normalisation lines executed verbatim, and the real cv2.imwrite -> imread round trip."""
    import tempfile

    import cv2
    g = np.load(os.path.join(HERE, "similarity.npz"))
    h = np.load(os.path.join(HERE, "hamming.npz"))
    out = {}
    tmp = tempfile.mkdtemp()
    rng = np.random.default_rng(3)
    big = rng.normal(-300.0, 120.0, (37, 37))
    big = (big + big.T) / 2
    for name, S in (("a", g["S_a"]), ("b", g["S_b"]), ("r", big)):
        S = np.array(S, dtype=np.float64)
        n = len(S)
        similarity_matrix = np.full([n, n], -1)                       # :31 (int64)
        for i in range(n):
            for j in range(i + 1, n):
                similarity_matrix[i, j] = S[i, j]                     # :36-37, float -> int64 store
                similarity_matrix[j, i] = S[i, j]
        env = _exec_reference_lines("src/sdav/create_similarity_matrix.py", 41, 45,
                                    {"similarity_matrix": similarity_matrix, "np": np})
        path = os.path.join(tmp, "s.png")
        cv2.imwrite(path, env["similarity_img"])                      # :48
        out["sim_scores_" + name] = np.where(np.eye(n, dtype=bool), -1.0, np.triu(S, 1) + np.triu(S, 1).T)
        out["sim_int_" + name] = similarity_matrix
        out["sim_img_f64_" + name] = env["similarity_img"]
        out["sim_png_" + name] = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    distance_matrix = np.full([len(h["D"]), len(h["D"])], -1)         # :31
    distance_matrix[:, :] = h["D"]
    env = _exec_reference_lines("src/cnn_vtl/create_distance_matrix.py", 40, 40,
                                {"distance_matrix": distance_matrix, "np": np})
    path = os.path.join(tmp, "d.png")
    cv2.imwrite(path, env["distance_img"])                            # :41
    out["dist_D"] = h["D"]
    out["dist_img_f64"] = env["distance_img"]
    out["dist_png"] = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    np.savez_compressed(os.path.join(HERE, "images.npz"), **out)


def golden_config1():
    """SURVEY 8(d) config 1, the parity config: the 20 real frames of datasets/test read like CvInputParser.py:32
    (cv2.imread GRAYSCALE), seeded keypoints (np.random.default_rng(0), x in U[0,240), y in U[0,192), 30 per frame),
    the reference's OWN patch function (CvInputParser.py:100-123, 27), the float64 oracle encoder (the reference's
    TensorFlow graph cannot run here) with SDA weights seed 1 (N(0,1), the reference's initialisation) and seed 2
    (Xavier-scaled), and the reference's OWN SimilarityCalculator on those descriptors for every ordered pair i != j.
    Stored: the frames, keypoints, a digest of the reference's patch array, and per weight set the 20 x 20 scores,
    the matched patch index of every row and the gap to the runner-up (for the tie report)."""
    import hashlib

    import cv2
    from src.sdav.input.CvInputParser import get_vectorized_patches_from_key_points
    from src.sdav.similarity.SimilarityCalculator import SimilarityCalculator
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import sda as o_sda
    from oracle import similarity as o_sim
    names = sorted(f for f in os.listdir(os.path.join(REF, "datasets/test")) if f.endswith(".ppm"))
    assert len(names) == 20
    frames = np.stack([cv2.imread(os.path.join(REF, "datasets/test", f), cv2.IMREAD_GRAYSCALE) for f in names])
    n, h, w = frames.shape
    rng = np.random.default_rng(0)
    xy = np.stack([rng.uniform(0, w, (n, 30)), rng.uniform(0, h, (n, 30))], -1).astype(np.float32)
    x = np.stack([get_vectorized_patches_from_key_points(frames[i], [FakeKeyPoint(px, py) for px, py in xy[i]], 41)
                  / 255.0 for i in range(n)])                                     # [20, 30, 1681] float64
    out = {"frames": frames, "xy": xy, "names": np.array(names),
           "patches_sha256": np.array(hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest())}
    dims = [1681, 2500, 2500, 2500, 2500, 2500]
    for tag, seed, scale in (("normal", 1, "normal"), ("xavier", 2, "xavier")):
        ws, bs = o_sda.make_weights(dims, seed=seed, scale=scale)
        desc = o_sda.sda_forward(x, ws, bs).reshape(n, 30, -1)
        calc = SimilarityCalculator(desc)
        S = np.full((n, n), -1.0)
        idx = np.zeros((n, n, 30), dtype=np.int8)
        gap = np.zeros((n, n, 30))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(n):
                for j in range(n):
                    if i == j:
                        continue
                    S[i, j] = calc.similarity_score(desc[i], desc[j])             # the reference class, unmodified
                    idx[i, j] = o_sim.match_indices(desc[i], desc[j])
                    gap[i, j] = o_sim.nn_margin(desc[i], desc[j])
        out["S_" + tag] = S
        out["idx_" + tag] = idx
        out["gap_" + tag] = gap
        out["desc_digest_" + tag] = np.array([desc.sum(), np.abs(desc).max(), desc[3, 7, 11], desc[19, 29, 2499]])
    np.savez_compressed(os.path.join(HERE, "config1_frames.npz"), **out)


if __name__ == "__main__":
    golden_patches(); golden_similarity(); golden_hamming(); golden_misc(); golden_images(); golden_config1()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
