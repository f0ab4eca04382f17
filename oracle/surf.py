"""Keypoint detector oracle: the fast-Hessian ("SURF") detector behind get_top_n_key_points
(src/sdav/input/CvInputParser.py:36-46: cv2.xfeatures2d.SURF_create().detect, sort by -response, first n).
TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED. The arithmetic lives in a third-party dependency that is absent from /root/reference AND from this
image: opencv-contrib-python==3.4.2.17 (requirements.txt:7), modules/xfeatures2d/src/surf.cpp, non-free. This file
restates the published algorithm (Bay, Ess, Tuytelaars, Van Gool, "Speeded-Up Robust Features", CVIU 2008) with the
constants and conventions of that implementation as documented: SURF_create() defaults hessianThreshold = 100,
nOctaves = 4, nOctaveLayers = 3; 9x9 base filters growing by 6 per layer and doubling per octave, sampling step
2^octave; box-filter Dxx / Dyy / Dxy on the integral image, det = Dxx Dyy - 0.81 Dxy^2; strict 3x3x3 maximum in the
three middle layers of every octave; quadratic sub-sample interpolation (rejected when an offset exceeds one sample);
keypoints whose orientation window cannot be sampled are dropped. The detection ORDER of the reference is not
deterministic (parallel layers + mutex) and its sort is stable, so ties in response have no defined order there; here
ties break by (octave, layer, row, column).
What IS pinned: the CUDA detector equals this file bit for bit (tests/test_gpu_kernels.py), and property tests place
synthetic blobs whose position and scale the detector must recover (tests/test_oracle_golden.py)."""
import numpy as np

HAAR_SIZE0, HAAR_SIZE_INC = 9, 6
DX = ((0, 2, 3, 7, 1), (3, 2, 6, 7, -2), (6, 2, 9, 7, 1))
DY = ((2, 0, 7, 3, 1), (2, 3, 7, 6, -2), (2, 6, 7, 9, 1))
DXY = ((1, 1, 4, 4, 1), (5, 1, 8, 4, -1), (1, 5, 4, 8, -1), (5, 5, 8, 8, 1))
ORI_RADIUS = 6
f32 = np.float32


def integral(img):
    """int32 [H+1, W+1], first row and column zero."""
    s = np.zeros((img.shape[0] + 1, img.shape[1] + 1), dtype=np.int64)
    s[1:, 1:] = np.cumsum(np.cumsum(img.astype(np.int64), axis=0), axis=1)
    return s.astype(np.int32)


def resize_pattern(src, size):
    """Box corners (x1, y1, x2, y2) and weight of a 9x9 pattern scaled to `size` (float32 arithmetic)."""
    ratio = f32(size) / f32(9)
    out = []
    for x1, y1, x2, y2, w in src:
        a, b, c, d = (int(np.rint(np.float64(ratio * f32(v)))) for v in (x1, y1, x2, y2))
        out.append((a, b, c, d, f32(w) / (f32(c - a) * f32(d - b))))
    return out


def layer_sizes(n_octaves=4, n_layers=3):
    return [[(HAAR_SIZE0 + HAAR_SIZE_INC * l) << o for l in range(n_layers + 2)] for o in range(n_octaves)]


def haar(sum_, pattern, size, step, ni, nj):
    """Response of one box pattern at every sample: float32(sum over boxes of float32(box sum) * w, added in float64)."""
    acc = np.zeros((ni, nj), dtype=np.float64)
    ii = np.arange(ni) * step
    jj = np.arange(nj) * step
    for x1, y1, x2, y2, w in pattern:
        box = (sum_[np.ix_(ii + y1, jj + x1)].astype(np.int64) + sum_[np.ix_(ii + y2, jj + x2)]
               - sum_[np.ix_(ii + y2, jj + x1)] - sum_[np.ix_(ii + y1, jj + x2)]).astype(np.int32)
        acc += (box.astype(np.float32) * w).astype(np.float64)
    return acc.astype(np.float32)


def det_layer(sum_, size, step):
    """Hessian determinant of one layer: float32 [H // step, W // step], zero where the filter does not fit."""
    H, W = sum_.shape[0] - 1, sum_.shape[1] - 1
    det = np.zeros((H // step, W // step), dtype=np.float32)
    if size > H or size > W:
        return det
    ni, nj = 1 + (H - size) // step, 1 + (W - size) // step
    margin = (size // 2) // step
    dx = haar(sum_, resize_pattern(DX, size), size, step, ni, nj)
    dy = haar(sum_, resize_pattern(DY, size), size, step, ni, nj)
    dxy = haar(sum_, resize_pattern(DXY, size), size, step, ni, nj)
    d = (dx * dy).astype(np.float32) - ((f32(0.81) * dxy).astype(np.float32) * dxy).astype(np.float32)
    det[margin:margin + ni, margin:margin + nj] = d.astype(np.float32)
    return det


def solve3(A, b):
    """Gaussian elimination with partial pivoting in float64, fixed operation order (mirrored by the CUDA kernel)."""
    M = [[float(A[r][c]) for c in range(3)] + [float(b[r])] for r in range(3)]
    for k in range(3):
        p = k
        for r in range(k + 1, 3):
            if abs(M[r][k]) > abs(M[p][k]):
                p = r
        if M[p][k] == 0.0:
            return None
        M[k], M[p] = M[p], M[k]
        for r in range(k + 1, 3):
            f = M[r][k] / M[k][k]
            for c in range(k, 4):
                M[r][c] = M[r][c] - f * M[k][c]
    x = [0.0, 0.0, 0.0]
    for k in (2, 1, 0):
        s = M[k][3]
        for c in range(k + 1, 3):
            s = s - M[k][c] * x[c]
        x[k] = s / M[k][k]
    return x


def orientation_samplable(x, y, size, H, W):
    """The orientation stage drops a keypoint when no gradient sample of its radius-6s disc fits in the image."""
    s = f32(size) * f32(1.2) / f32(9.0)
    grad = 2 * int(np.rint(np.float64(f32(2) * s)))
    if H + 1 < grad or W + 1 < grad:
        return False
    half = f32(grad - 1) / f32(2)
    for i in range(-ORI_RADIUS, ORI_RADIUS + 1):
        for j in range(-ORI_RADIUS, ORI_RADIUS + 1):
            if i * i + j * j <= ORI_RADIUS * ORI_RADIUS:
                px = int(np.rint(np.float64(f32(x) + f32(j) * s - half)))
                py = int(np.rint(np.float64(f32(y) + f32(i) * s - half)))
                if 0 <= py < H + 1 - grad and 0 <= px < W + 1 - grad:
                    return True
    return False


def detect(img, hessian_threshold=100.0, n_octaves=4, n_layers=3):
    """All keypoints of a uint8 [H, W] image as rows (x, y, size, response, octave, layer, i, j), detection order
    (octave, layer, row, column)."""
    H, W = img.shape
    sum_ = integral(img)
    thr = f32(hessian_threshold)
    out = []
    for o, sizes in enumerate(layer_sizes(n_octaves, n_layers)):
        step = 1 << o
        dets = [det_layer(sum_, s, step) for s in sizes]
        rows, cols = H // step, W // step
        for l in range(1, n_layers + 1):
            size = sizes[l]
            margin = (sizes[l + 1] // 2) // step + 1
            if rows - 2 * margin <= 0 or cols - 2 * margin <= 0:
                continue
            c = dets[l][margin:rows - margin, margin:cols - margin]
            ok = c > thr
            for dl in (-1, 0, 1):
                for di in (-1, 0, 1):
                    for dj in (-1, 0, 1):
                        if dl == 0 and di == 0 and dj == 0:
                            continue
                        nb = dets[l + dl][margin + di:rows - margin + di, margin + dj:cols - margin + dj]
                        ok &= c > nb
            for i0, j0 in zip(*np.nonzero(ok)):
                i, j = int(i0) + margin, int(j0) + margin
                N9 = [dets[l + dl][i - 1:i + 2, j - 1:j + 2].reshape(9) for dl in (-1, 0, 1)]
                val0 = N9[1][4]
                sum_i = step * (i - (size // 2) // step)
                sum_j = step * (j - (size // 2) // step)
                cy = f32(sum_i) + f32(size - 1) * f32(0.5)
                cx = f32(sum_j) + f32(size - 1) * f32(0.5)
                ds = size - sizes[l - 1]
                two, four = f32(2), f32(4)
                b = [-(N9[1][5] - N9[1][3]) / two, -(N9[1][7] - N9[1][1]) / two, -(N9[2][4] - N9[0][4]) / two]
                axy = (N9[1][8] - N9[1][6] - N9[1][2] + N9[1][0]) / four
                axs = (N9[2][5] - N9[2][3] - N9[0][5] + N9[0][3]) / four
                ays = (N9[2][7] - N9[2][1] - N9[0][7] + N9[0][1]) / four
                A = [[N9[1][3] - two * N9[1][4] + N9[1][5], axy, axs],
                     [axy, N9[1][1] - two * N9[1][4] + N9[1][7], ays],
                     [axs, ays, N9[0][4] - two * N9[1][4] + N9[2][4]]]
                x = solve3(A, b)
                if x is None:
                    continue
                x = [f32(v) for v in x]
                if not ((x[0] != 0 or x[1] != 0 or x[2] != 0) and abs(x[0]) <= 1 and abs(x[1]) <= 1 and abs(x[2]) <= 1):
                    continue
                px = f32(cx + x[0] * f32(step))
                py = f32(cy + x[1] * f32(step))
                ksize = f32(np.rint(np.float64(f32(size) + x[2] * f32(ds))))
                if not orientation_samplable(px, py, ksize, H, W):
                    continue
                out.append((float(px), float(py), float(ksize), float(val0), o, l, i, j))
    return np.array(out, dtype=np.float64).reshape(-1, 8)


def top_n(keypoints, n):
    """Sort by response, descending (CvInputParser.py:44-45); ties keep detection order. -> [<= n, 8]"""
    order = np.argsort(-keypoints[:, 3], kind="stable")
    return keypoints[order[:n]]
