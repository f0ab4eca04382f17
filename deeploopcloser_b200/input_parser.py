"""Drop-in for `src.sdav.input.CvInputParser` (reference src/sdav/input/CvInputParser.py).
`parse(image)` = top-n keypoints -> 41x41 patches around them (window shifted inside the image, rows indexed by x -
the reference's quirk) -> / 255.0 -> float64 [n, 1681]. The gather + normalise runs on the B200
(dlc_patch_gather_f64, bit-identical to the reference's NumPy result). Keypoint detection (OpenCV SURF in the
reference, CvInputParser.py:36-46) runs on the B200 too (dlc_surf_detect, a restatement of the published fast-Hessian
detector - parity against OpenCV's non-free module cannot be pinned here); `key_points=` lets callers inject their
own keypoints."""
import math

import numpy as np


def _xy_array(key_points):
    """cv2.KeyPoint-like objects (with .pt) or an [n, 2] array of (x, y) -> float32 [n, 2]."""
    if isinstance(key_points, np.ndarray):
        arr = key_points
    elif len(key_points) and hasattr(key_points[0], "pt"):
        arr = np.array([[kp.pt[0], kp.pt[1]] for kp in key_points])
    else:
        arr = np.asarray(key_points)
    return np.ascontiguousarray(arr, dtype=np.float32).reshape(-1, 2)


class KeyPoint:
    """The attributes of cv2.KeyPoint that the reference reads (.pt, CvInputParser.py:111; .response, :44)."""
    __slots__ = ("pt", "size", "response")

    def __init__(self, x, y, size, response):
        self.pt = (float(x), float(y))
        self.size = float(size)
        self.response = float(response)

    def __repr__(self):
        return "KeyPoint(pt=(%.3f, %.3f), size=%g, response=%g)" % (self.pt[0], self.pt[1], self.size, self.response)


def get_top_n_key_points(img, n, detector="b200"):
    """Top n SURF keypoints in descending response order (CvInputParser.py:36-46).
    detector="b200" (default): the fast-Hessian detector on the GPU (dlc_surf_detect; SURF_create() defaults:
    threshold 100, 4 octaves, 3 layers). It restates the published algorithm - OpenCV's non-free SURF is not available
    to pin it against, so individual keypoints may differ from the reference's. detector="opencv": the reference's
    own call, for environments whose OpenCV has xfeatures2d (host code, not accelerated)."""
    if detector == "opencv":
        import cv2
        if not hasattr(cv2, "xfeatures2d") or not hasattr(cv2.xfeatures2d, "SURF_create"):
            raise RuntimeError("cv2.xfeatures2d.SURF_create is unavailable in this OpenCV build")
        key_points = list(cv2.xfeatures2d.SURF_create().detect(img, None))
        key_points.sort(key=lambda kp: -kp.response)
        return key_points[0:n]
    if detector != "b200":
        raise ValueError("detector must be 'b200' or 'opencv'")
    import torch

    from . import _cuda, ops
    _cuda.require_cuda()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim != 2:
        raise ValueError("expected a single-channel (grayscale) image")
    xy, info, found = ops.surf_detect(torch.from_numpy(img)[None].cuda(), top_n=int(n))
    m = min(int(found[0].item()), int(n))           # like the reference: fewer than n when the image has fewer
    xy, info = xy[0, :m].cpu().numpy(), info[0, :m].cpu().numpy()
    return [KeyPoint(xy[i, 0], xy[i, 1], info[i, 0], info[i, 1]) for i in range(m)]


def get_1d_boundaries(rect_shape, center_points, patch_size, axis):
    """Host restatement of the window rule (CvInputParser.py:49-89) for callers that use it directly."""
    if patch_size % 2 == 0:
        raise ValueError("Invalid patch size. Patch size must be an odd number")
    if len(rect_shape) != 2:
        raise ValueError("Invalid rect shape. It must be a list of two integers")
    center_points = np.asarray(center_points)
    if center_points.ndim != 2:
        raise ValueError("Invalid center points. center_points must be a numpy array of 2D coordinates")
    if center_points.shape[1] != 2:
        raise ValueError("Invalid center points. Coordinates must be in 2D")
    half = patch_size // 2
    c = center_points[:, axis]
    lo, hi = c - half, c + half
    fwd = np.where(lo < 0, -lo, 0)
    over = hi - rect_shape[axis] + 1
    back = np.where(over > 0, over, 0)
    return lo - back + fwd, hi - back + fwd


def get_2d_boundaries(rect_shape, coordinates, patch_size):
    x_lo, x_hi = get_1d_boundaries(rect_shape, coordinates, patch_size, 0)
    y_lo, y_hi = get_1d_boundaries(rect_shape, coordinates, patch_size, 1)
    return x_lo, x_hi, y_lo, y_hi


def _gather_f64(img, key_points, patch_size, swap_xy_quirk=True):
    import torch

    from . import _cuda, ops
    _cuda.require_cuda()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim != 2:
        raise ValueError("expected a single-channel (grayscale) image")
    xy = _xy_array(key_points)
    if len(xy) == 0:
        return np.zeros((0, patch_size ** 2))
    out = ops.patch_gather_f64(torch.from_numpy(img)[None].cuda(), torch.from_numpy(xy)[None].cuda(), patch_size,
                               swap_xy_quirk)
    return out.cpu().numpy()


def get_vectorized_patches_from_key_points(img, key_points, patch_size):
    """[n, patch_size^2] integer pixel patches (CvInputParser.py:100-123)."""
    return np.rint(_gather_f64(img, key_points, patch_size) * 255.0).astype(int)


class CvInputParser:
    def __init__(self, n_patches: int = 30, patch_size: int = 41, swap_xy_quirk: bool = True):
        self.n_patches = n_patches
        self.patch_size = patch_size
        self.swap_xy_quirk = swap_xy_quirk  # True = reference behaviour (rows indexed by x)

    def parse(self, image, key_points=None):
        """image: uint8 [H, W] -> float64 [n, patch_size^2] in [0, 1] (CvInputParser.py:19-28)."""
        if key_points is None:
            key_points = get_top_n_key_points(image, self.n_patches)
        else:
            key_points = _xy_array(key_points)[: self.n_patches]
        return _gather_f64(image, key_points, self.patch_size, self.swap_xy_quirk)

    def parse_from_path(self, image_path: str, key_points=None):
        import cv2
        image = cv2.imread(str(image_path), cv2.IMREAD_GRAYSCALE)
        if image is None:
            raise IOError("could not read image %s" % image_path)
        return self.parse(image, key_points)


def get_generator(file_pattern: str, shape: list):
    """src/sdav/input/InputGenerator.py:11-27 (sorted here; the reference globs unsorted)."""
    import glob
    n_patches = shape[0]
    patch_size = int(math.sqrt(shape[1]))

    def iteration():
        parser = CvInputParser(n_patches, patch_size)
        for f in sorted(glob.glob(file_pattern)):
            yield parser.parse_from_path(f)
    return iteration
